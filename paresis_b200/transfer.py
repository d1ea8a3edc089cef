"""Device -> host result copies through pinned staging memory.

``fetch`` returns a numpy array that is a zero-copy view of a pinned host tensor (torch's caching
pinned allocator recycles the block once the array is garbage collected), filled by an
asynchronous copy on a dedicated copy stream.  ``bytes_d2h`` / ``bytes_h2d`` count what crossed
PCIe through this module (bench.py reports them).
"""
import weakref

import numpy as np
import torch

bytes_d2h = 0
bytes_h2d = 0
_copy_streams = {}

# Pinned staging buffers are expensive to create (page locking: ~50 us per MiB), so they are
# recycled: a buffer goes back to its pool when the numpy array handed to the caller -- and every
# view of it -- has been garbage collected.
_pool = {}
pinned_allocs = 0


def _pinned(shape, dtype):
    global pinned_allocs
    key = (tuple(shape), dtype)
    free = _pool.setdefault(key, [])
    if free:
        return free.pop()
    pinned_allocs += 1
    return torch.empty(shape, dtype=dtype, pin_memory=True)


def _release(key, tensor):
    free = _pool.setdefault(key, [])
    if len(free) < 8:
        free.append(tensor)


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
        self.host = _pinned(tensor.shape, tensor.dtype)
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
        host, self.host = self.host, None
        arr = host.numpy()
        weakref.finalize(arr, _release, (tuple(host.shape), host.dtype), host)
        return arr


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
