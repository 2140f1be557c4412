"""TEST INFRASTRUCTURE ONLY: numpy-backed stand-in for b200.runtime.CudaRuntime so the host logic of
the product (b200/tracking.py, b200/pipeline.py, the libs mirror) can run against the host-emulation
build of the kernels in a container without a GPU.  Injected with b200.runtime.set_runtime()."""
import ctypes

import numpy as np

from b200 import runtime as rt_mod


class HostsimRuntime:
    is_device = False

    def __init__(self, lib):
        self.lib = lib

    def stream_ptr(self):
        return None

    def empty(self, shape, dtype):
        return rt_mod.Buf(np.zeros(tuple(shape), dtype), shape, dtype, self)

    zeros = empty

    def from_host(self, arr):
        a = np.ascontiguousarray(arr).copy()
        return rt_mod.Buf(a, a.shape, a.dtype, self)

    def upload(self, buf, arr):
        buf.t[...] = np.asarray(arr, dtype=buf.dtype).reshape(buf.shape)

    def to_host(self, buf):
        return buf.t.copy()

    def view(self, buf, shape):
        n = int(np.prod(shape)) if len(shape) else 1
        return rt_mod.Buf(buf.t.reshape(-1)[:n].reshape(tuple(shape)), shape, buf.dtype, self)

    def clone(self, buf):
        return rt_mod.Buf(buf.t.copy(), buf.shape, buf.dtype, self)

    def ptr_of(self, t):
        return t.ctypes.data_as(ctypes.c_void_p)

    def sync(self):
        pass

    def nvtx(self, name):
        import contextlib
        return contextlib.nullcontext()

    # streams / events: the emulation is synchronous, so these are no-ops with the CudaRuntime signatures
    def new_stream(self, high_priority=False):
        return None

    def on_stream(self, stream):
        import contextlib
        return contextlib.nullcontext()

    def record_event(self):
        return None

    def wait_event(self, ev):
        pass
