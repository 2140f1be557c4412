"""Shared plumbing of the geometry-layer mirrors (backprojection / transformation3d / projection / reprojection /
rigid_flow): the reference's layers are ``torch.nn.Module``s called with batched CUDA tensors; here every ``forward`` is
one launch of csrc/geometry.cu through the C ABI.  Inputs may be torch tensors (any device), NumPy arrays or the
device-backed ``tracking.DevArray``; the result comes back as a torch tensor on the input's device when the input was a
torch tensor, else as a NumPy array.  Matrices ([N,4,4]) are read on the host -- they are 16 numbers."""
import ctypes

import numpy as np

from b200 import runtime, tracking


def is_torch(x):
    return type(x).__module__.startswith("torch")


def host(x):
    """ndarray view of a small matrix argument (torch tensor / ndarray)."""
    if is_torch(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def mat_ptr(a):
    a = np.ascontiguousarray(a, np.float64)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def to_dev(x, shape):
    """float32 device buffer of a map-like argument, reshaped to `shape`."""
    rt = runtime.get()
    if isinstance(x, tracking.DevArray):
        return x.dev
    if isinstance(x, runtime.Buf):
        return x
    if is_torch(x):
        t = x.detach()
        if getattr(rt, "is_device", False) and t.is_cuda and str(t.dtype) == "torch.float32" and t.is_contiguous():
            return runtime.Buf(t.reshape(shape), shape, np.float32, rt)
        x = t.float().cpu().numpy()
    return rt.from_host(np.ascontiguousarray(np.asarray(x, np.float32).reshape(shape)))


def wrap(buf, like, shape):
    """Result in the caller's currency: torch tensor on `like`'s device, or ndarray."""
    rt = runtime.get()
    if is_torch(like):
        import torch
        if getattr(rt, "is_device", False):
            t = buf.t.reshape(shape)
            return t if like.is_cuda else t.cpu()
        return torch.from_numpy(buf.numpy().reshape(shape)).to(like.device)
    return buf.numpy().reshape(shape)


def batch(x):
    s = tuple(x.shape)
    return s[0] if len(s) >= 3 else 1
