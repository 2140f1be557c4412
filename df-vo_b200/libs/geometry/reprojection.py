"""``Reprojection`` with the reference's interface (libs/geometry/reprojection.py:20-56): backprojection, rigid transform and
projection fused into one kernel (csrc/geometry.cu::k_reproject) -- 4 B read and 8 B written per pixel instead of the three
layers' 4 + 16 + 16 + 16 + 8."""
import numpy as np

from b200 import runtime
from . import _layers as L
from .backprojection import Backprojection
from .projection import Projection
from .transformation3d import Transformation3D


class Reprojection:
    def __init__(self, height, width):
        self.height, self.width = int(height), int(width)
        self.backproj = Backprojection(height, width)          # the sub-layers of the reference stay reachable
        self.transform = Transformation3D()
        self.project = Projection(height, width)

    def cuda(self):
        return self

    def to(self, *a, **k):
        return self

    def _run(self, depth, T, K, inv_K, normalized, flow):
        rt = runtime.get()
        n, hw = L.batch(depth), self.height * self.width
        Th, Kh, iKh = (L.host(m).reshape(-1, 4, 4) for m in (T, K, inv_K))
        d = L.to_dev(depth, (n, self.height, self.width))
        out = rt.empty((n, 2, self.height, self.width) if flow else (n, self.height, self.width, 2), np.float32)
        for i in range(n):
            k1, pT = L.mat_ptr(Th[min(i, Th.shape[0] - 1)])
            k2, pK = L.mat_ptr(Kh[min(i, Kh.shape[0] - 1)][:3, :])
            k3, pI = L.mat_ptr(iKh[min(i, iKh.shape[0] - 1)][:3, :3])
            dp, op = d.ptr.value + i * hw * 4, out.ptr.value + i * hw * 2 * 4
            if flow:
                rt.lib.check(rt.lib.dfvo_rigid_flow(dp, self.height, self.width, pT, pK, pI, op, rt.stream_ptr()))
            else:
                rt.lib.check(rt.lib.dfvo_reproject(dp, self.height, self.width, pT, pK, pI, self.project.eps, int(bool(normalized)), op,
                                                   rt.stream_ptr()))
        return L.wrap(out, depth, out.shape)

    def forward(self, depth, T, K, inv_K, normalized=True):
        """depth [N,1,H,W], T / K / inv_K [N,4,4] -> xy [N,H,W,2]."""
        return self._run(depth, T, K, inv_K, normalized, False)

    __call__ = forward
