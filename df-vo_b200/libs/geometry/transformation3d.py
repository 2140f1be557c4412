"""``Transformation3D`` with the reference's interface (libs/geometry/transformation3d.py:13-31) on csrc/geometry.cu."""
import numpy as np

from b200 import runtime
from . import _layers as L


class Transformation3D:
    def cuda(self):
        return self

    def to(self, *a, **k):
        return self

    def forward(self, points, T):
        """points [N,4,M], T [N,4,4] -> T @ points."""
        rt = runtime.get()
        s = tuple(points.shape)
        n, m = s[0], int(np.prod(s[2:]))
        Th = L.host(T).reshape(-1, 4, 4)
        src = L.to_dev(points, (n, 4, m))
        out = rt.empty((n, 4, m), np.float32)
        for i in range(n):
            keep, p = L.mat_ptr(Th[min(i, Th.shape[0] - 1)])
            rt.lib.check(rt.lib.dfvo_transform3d(src.ptr.value + i * 4 * m * 4, m, p, out.ptr.value + i * 4 * m * 4, rt.stream_ptr()))
        return L.wrap(out, points, s)

    __call__ = forward
