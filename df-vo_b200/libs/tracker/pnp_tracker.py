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
        self.K = [float(cam_intrinsics.cx), float(cam_intrinsics.cy), float(cam_intrinsics.fx), float(cam_intrinsics.fy)]

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
        """pnp_tracker.py:126-145 (kp_selection_good_depth, :148-212, is the same routine as the E-tracker's)."""
        import copy
        from libs.tracker.E_tracker import EssTracker
        rigid_pose = copy.deepcopy(pose)
        ref_data["rigid_flow_pose"] = SE3(rigid_pose.inv_pose)
        k = EssTracker.kp_selection_good_depth(self, cur_data, ref_data, self.cfg.pnp_tracker.iterative_kp.score_method)
        ref_data["kp_depth"], cur_data["kp_depth"] = k["kp1_depth"][0], k["kp2_depth"][0]
        ref_data["kp_depth_uniform"], cur_data["kp_depth_uniform"] = k["kp1_depth_uniform"][0], k["kp2_depth_uniform"][0]
        cur_data["rigid_flow_mask"] = k["rigid_flow_mask"]

    _dev = None          # bound below (shared helper of the two trackers)


from libs.tracker.E_tracker import EssTracker as _E          # noqa: E402
PnpTracker._dev = _E._dev
