"""Device -> host result copies through pinned staging memory.

``fetch`` returns a numpy array that is a zero-copy view of a pinned host tensor, filled by an
asynchronous copy on the library's copy stream (``paresis_transfer_d2h``: one C call per copy --
event record, stream wait, cudaMemcpyAsync, event record).  ``bytes_d2h`` / ``bytes_h2d`` count what
crossed PCIe through this module (bench.py reports them).
"""
import ctypes
import weakref

import numpy as np
import torch

from . import _cabi as abi

bytes_d2h = 0
bytes_h2d = 0

# Pinned staging buffers are expensive to create (page locking: ~50 us per MiB), so they are
# recycled together with their transfer lane (two CUDA events): a buffer goes back to its pool when
# the numpy array handed to the caller -- and every view of it -- has been garbage collected.
_pool = {}
pinned_allocs = 0


def _pinned(shape, dtype):
    global pinned_allocs
    key = (tuple(shape), dtype)
    free = _pool.setdefault(key, [])
    if free:
        return free.pop()
    pinned_allocs += 1
    lane = ctypes.c_void_p()
    abi._check(abi.lib.paresis_transfer_lane_create(ctypes.byref(lane)), "paresis_transfer_lane_create")
    return torch.empty(shape, dtype=dtype, pin_memory=True), lane


def _release(key, entry):
    free = _pool.setdefault(key, [])
    if len(free) < 8:
        free.append(entry)
    else:
        abi.lib.paresis_transfer_lane_destroy(entry[1])


class Pending:
    """A device->host copy in flight; ``wait()`` gives the numpy view."""

    def __init__(self, tensor, dtype=None):
        global bytes_d2h
        if dtype is not None and tensor.dtype != dtype:
            tensor = tensor.to(dtype)               # cast on the device, on the producing stream
        if not tensor.is_contiguous():
            tensor = tensor.contiguous()
        self._src = tensor                          # keep alive until the copy is done
        self._key = (tuple(tensor.shape), tensor.dtype)
        self.host, self.lane = _pinned(tensor.shape, tensor.dtype)
        nbytes = tensor.numel() * tensor.element_size()
        abi._check(abi.lib.paresis_transfer_d2h(self.lane, ctypes.c_void_p(self.host.data_ptr()),
                                                ctypes.c_void_p(tensor.data_ptr()), nbytes, abi._stream()),
                   "paresis_transfer_d2h")
        bytes_d2h += nbytes
        # A Pending that is dropped without wait() (a membrane map nobody reads) must not hand its pinned buffer and its
        # source tensor back to the allocators while the copy is in flight on the library's copy stream: wait, then recycle.
        self._finalizer = weakref.finalize(self, _abandon, self._key, self.host, self.lane, self._src)

    def wait(self):
        self._finalizer.detach()
        abi._check(abi.lib.paresis_transfer_wait(self.lane), "paresis_transfer_wait")
        self._src = None
        host, self.host = self.host, None
        arr = host.numpy()
        weakref.finalize(arr, _release, self._key, (host, self.lane))
        return arr


def _abandon(key, host, lane, src):
    abi.lib.paresis_transfer_wait(lane)
    del src
    _release(key, (host, lane))


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
