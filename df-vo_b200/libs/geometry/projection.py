"""``Projection`` with the reference's interface (libs/geometry/projection.py:15-58) on csrc/geometry.cu."""
import numpy as np

from b200 import runtime
from . import _layers as L


class Projection:
    def __init__(self, height, width, eps=1e-7):
        self.height, self.width, self.eps = int(height), int(width), float(eps)

    def cuda(self):
        return self

    def to(self, *a, **k):
        return self

    def forward(self, points3d, K, normalized=True):
        """points3d [N,4,H*W], K [N,4,4] -> pixel coordinates [N,H,W,2] (normalised to [-1,1] if asked)."""
        rt = runtime.get()
        n, hw = tuple(points3d.shape)[0], self.height * self.width
        Kh = L.host(K).reshape(-1, 4, 4)
        src = L.to_dev(points3d, (n, 4, hw))
        out = rt.empty((n, self.height, self.width, 2), np.float32)
        for i in range(n):
            keep, p = L.mat_ptr(Kh[min(i, Kh.shape[0] - 1)][:3, :])
            rt.lib.check(rt.lib.dfvo_project(src.ptr.value + i * 4 * hw * 4, self.height, self.width, p, self.eps, int(bool(normalized)),
                                             out.ptr.value + i * hw * 2 * 4, rt.stream_ptr()))
        return L.wrap(out, points3d, (n, self.height, self.width, 2))

    __call__ = forward
