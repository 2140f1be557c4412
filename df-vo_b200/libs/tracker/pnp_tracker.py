"""``PnpTracker`` with the reference's interface (libs/tracker/pnp_tracker.py:23-212).

The PnP fallback only runs when the essential-matrix tracker is rejected (dfvo.py:227).  Keypoint filtering is
vectorised NumPy on the arrays the driver passes in; the RANSAC repeats (OpenCV's subset stream replayed, EPnP minimal
solver, reprojection scoring) and the final least-squares refit run on the device (csrc/pnp.cu)."""
import numpy as np

from b200 import tracking
from libs.geometry.camera_modules import SE3


class PnpTracker:
    def __init__(self, cfg, cam_intrinsics):
        self.cfg = cfg
        self.cam_intrinsics = cam_intrinsics
        assert not cfg.kp_selection.rigid_flow_kp.enable, "rigid_flow_kp is a 'next' row (SURVEY.md 8f rank 1)"

    def compute_pose_3d2d(self, kp1, kp2, depth_1, is_iterative):
        """pnp_tracker.py:45-125 -> {'pose': SE3 (view-2 -> view-1), 'kp1', 'kp2'}."""
        depth_1 = np.asarray(depth_1)
        height, width = depth_1.shape
        keep = (kp2[:, 0] >= 0) & (kp2[:, 0] < width)
        kp1, kp2 = kp1[keep], kp2[keep]
        keep = (kp2[:, 1] >= 0) & (kp2[:, 1] < height)
        kp1, kp2 = kp1[keep], kp2[keep]
        ki = kp1.astype(int)
        d = depth_1[ki[:, 1], ki[:, 0]]
        keep = (d != 0) & (d < self.cfg.depth.max_depth) & (d > self.cfg.depth.min_depth)
        kp1, kp2 = kp1[keep], kp2[keep]
        repeat = self.cfg.pnp_tracker.ransac.repeat if is_iterative else 3
        K = [float(self.cam_intrinsics.cx), float(self.cam_intrinsics.cy), float(self.cam_intrinsics.fx), float(self.cam_intrinsics.fy)]
        T, _ = tracking.compute_pose_3d2d(tracking.default_engine(), np.asarray(kp1, np.float64), np.asarray(kp2, np.float64),
                                          d[keep], K, repeat=repeat, iters=self.cfg.pnp_tracker.ransac.iter,
                                          reproj_thre=self.cfg.pnp_tracker.ransac.reproj_thre)
        pose = SE3(T)                                                  # already inverted (:118: solvePnP gives ref -> cur)
        return {"pose": pose, "kp1": kp1, "kp2": kp2}

    def compute_rigid_flow_kp(self, cur_data, ref_data, pose):
        raise NotImplementedError("rigid-flow keypoints are a 'next' row (SURVEY.md 8f rank 1)")
