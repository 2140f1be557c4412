"""``Backprojection`` with the reference's interface (libs/geometry/backprojection.py:17-67) on csrc/geometry.cu."""
import numpy as np

from b200 import runtime
from . import _layers as L


class Backprojection:
    def __init__(self, height, width):
        self.height, self.width = int(height), int(width)

    def cuda(self):
        return self

    def to(self, *a, **k):
        return self

    def forward(self, depth, inv_K, img_like_out=False):
        """depth [N,1,H,W], inv_K [N,4,4] -> homogeneous points [N,4,H*W] (or [N,4,H,W])."""
        rt = runtime.get()
        n, hw = L.batch(depth), self.height * self.width
        iK = L.host(inv_K).reshape(-1, 4, 4)
        out = rt.empty((n, 4, hw), np.float32)
        d = L.to_dev(depth, (n, self.height, self.width))
        for i in range(n):
            keep, p = L.mat_ptr(iK[min(i, iK.shape[0] - 1)][:3, :3])
            rt.lib.check(rt.lib.dfvo_backproject(d.ptr.value + i * hw * 4, self.height, self.width, p, out.ptr.value + i * 4 * hw * 4,
                                                 rt.stream_ptr()))
        shape = (n, 4, self.height, self.width) if img_like_out else (n, 4, hw)
        return L.wrap(out, depth, shape)

    __call__ = forward
