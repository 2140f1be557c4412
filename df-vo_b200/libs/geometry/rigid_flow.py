"""``RigidFlow`` with the reference's interface (libs/geometry/rigid_flow.py:17-58): Reprojection + PixToFlow
(layers.py:232-266) in one kernel; returns the flow [N,2,H,W]."""
from .reprojection import Reprojection


class RigidFlow:
    def __init__(self, height, width):
        self.height, self.width = int(height), int(width)
        self.reprojection = Reprojection(height, width)

    def cuda(self):
        return self

    def to(self, *a, **k):
        return self

    def forward(self, depth, T, K, inv_K, normalized=True):
        """depth [N,1,H,W], T / K / inv_K [N,4,4] -> rigid flow [N,2,H,W] in pixels.  As in the reference, `normalized` only
        reaches the reprojection; the tracker calls it with normalized=False (E_tracker.py:676-683), which is the case the
        fused kernel implements -- normalized=True falls back on the two-step definition."""
        if normalized:
            xy = self.reprojection(depth, T, K, inv_K, True)
            import numpy as np
            from . import _layers as L
            g = np.stack(np.meshgrid(range(self.width), range(self.height), indexing="xy"), 0).astype(np.float32)[None]
            if L.is_torch(xy):
                import torch
                return xy.permute(0, 3, 1, 2) - torch.from_numpy(g).to(xy.device)
            return np.transpose(xy, (0, 3, 1, 2)) - g
        return self.reprojection._run(depth, T, K, inv_K, False, True)

    __call__ = forward
