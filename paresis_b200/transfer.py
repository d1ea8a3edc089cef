"""Device -> host result copies through pinned staging memory.

``fetch`` returns a numpy array that is a zero-copy view of a pinned host tensor (torch's caching
pinned allocator recycles the block once the array is garbage collected), filled by an
asynchronous copy on a dedicated copy stream.  ``bytes_d2h`` / ``bytes_h2d`` count what crossed
PCIe through this module (bench.py reports them).
"""
import numpy as np
import torch

bytes_d2h = 0
bytes_h2d = 0
_copy_streams = {}


def _copy_stream(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=key)
    return _copy_streams[key]


class Pending:
    """A device->host copy in flight; ``wait()`` gives the numpy view."""

    def __init__(self, tensor, dtype=None):
        global bytes_d2h
        if dtype is not None and tensor.dtype != dtype:
            tensor = tensor.to(dtype)               # cast on the device, on the producing stream
        self._src = tensor                          # keep alive until the copy is done
        self.host = torch.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True)
        ready = torch.cuda.Event()
        ready.record()
        stream = _copy_stream(tensor.device)
        stream.wait_event(ready)
        with torch.cuda.stream(stream):
            self.host.copy_(tensor, non_blocking=True)
            self.done = torch.cuda.Event()
            self.done.record()
        bytes_d2h += tensor.numel() * tensor.element_size()

    def wait(self):
        self.done.synchronize()
        self._src = None
        return self.host.numpy()


def fetch(tensor, dtype=None):
    return Pending(tensor, dtype).wait()


def upload(array, dtype, device):
    """Host array -> device tensor (counted)."""
    global bytes_h2d
    a = np.ascontiguousarray(array)
    if not a.flags.writeable:
        a = a.copy()
    t = torch.as_tensor(a).to(device, dtype=dtype).contiguous()
    bytes_h2d += t.numel() * t.element_size()
    return t
