"""Helpers shared by the parity tests."""
import ctypes

import numpy as np


def hptr(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def dptr(t):
    assert t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def bf16_round(x):
    """Round a float32 numpy array to bf16 precision (round-to-nearest-even), returned as float32."""
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def img_to_tensor(img):
    """deep_models.py:160-163: uint8 HWC -> float32 [1,3,H,W] in [0,1] (division in float64)."""
    import torch
    return torch.from_numpy(np.transpose(img / 255, (2, 0, 1))).unsqueeze(0).float()
