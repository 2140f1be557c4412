"""``EssTracker`` with the reference's interface (libs/tracker/E_tracker.py:129-705) on the dfvo_b200
kernels: five repeated essential-matrix RANSACs (replaying OpenCV's sampling sequence), GRIC model
selection, pose recovery, and scale recovery from triangulated-vs-CNN depth."""
import numpy as np

from b200 import tracking
from libs.geometry.camera_modules import SE3


def get_E_from_pose(pose):
    """E_tracker.py:102-127: E = [t]x R with t normalised."""
    R = pose.R
    t = pose.t / np.linalg.norm(pose.t)
    tx = np.array([[0, -t[2, 0], t[1, 0]], [t[2, 0], 0, -t[0, 0]], [-t[1, 0], t[0, 0], 0]])
    return tx @ R


class EssTracker:
    def __init__(self, cfg, cam_intrinsics, timers):
        self.cfg = cfg
        self.prev_scale = 0
        self.prev_pose = SE3()
        self.cam_intrinsics = cam_intrinsics
        self.timers = timers
        assert not cfg.kp_selection.rigid_flow_kp.enable, "rigid_flow_kp is a 'next' row (SURVEY.md 8f rank 1)"
        assert cfg.e_tracker.validity.method == "GRIC", "dfvo_b200 implements e_tracker.validity.method GRIC (the default)"
        self.K = [float(cam_intrinsics.cx), float(cam_intrinsics.cy), float(cam_intrinsics.fx), float(cam_intrinsics.fy)]

    def compute_pose_2d2d(self, kp_ref, kp_cur, is_iterative):
        """E_tracker.py:154-307 -> {'pose': SE3 (cur -> ref, unit translation or identity), 'inliers': bool[N]}."""
        repeat = self.cfg.e_tracker.ransac.repeat if is_iterative else 3                 # :179
        r = tracking.compute_pose_2d2d(tracking.default_engine(), np.ascontiguousarray(kp_ref, np.float64),
                                       np.ascontiguousarray(kp_cur, np.float64), self.K, repeat=repeat,
                                       reproj_thre=self.cfg.e_tracker.ransac.reproj_thre)
        pose = SE3()
        pose.R = r["R"]
        pose.t = r["t"]
        return {"pose": pose, "inliers": r["inliers"]}

    def scale_recovery(self, cur_data, ref_data, E_pose, is_iterative):
        """E_tracker.py:442-474."""
        assert self.cfg.scale_recovery.method == "simple", "iterative scale recovery is a 'next' row (SURVEY.md 8f rank 1)"
        return {"scale": self.scale_recovery_simple(cur_data, ref_data, E_pose, is_iterative)}

    def scale_recovery_simple(self, cur_data, ref_data, E_pose, is_iterative):
        """E_tracker.py:476-507."""
        src = self.cfg.scale_recovery.iterative_kp.kp_src if is_iterative else self.cfg.scale_recovery.kp_src
        return self.find_scale_from_depth(ref_data[src], cur_data[src], E_pose.inv_pose, cur_data["depth"])

    def find_scale_from_depth(self, kp1, kp2, T_21, depth2):
        """E_tracker.py:571-643."""
        c = self.cfg.scale_recovery.ransac
        assert c.method == "depth_ratio", "dfvo_b200 implements scale_recovery.ransac.method depth_ratio (the default)"
        return tracking.find_scale_from_depth(tracking.default_engine(), np.asarray(kp1, np.float64), np.asarray(kp2, np.float64),
                                              np.asarray(T_21, np.float64), np.asarray(depth2), self.K, c.min_samples,
                                              c.max_trials, c.stop_prob, c.thre)

    def compute_rigid_flow_kp(self, cur_data, ref_data, pose):
        raise NotImplementedError("rigid-flow keypoints are a 'next' row (SURVEY.md 8f rank 1)")

    def scale_recovery_iterative(self, cur_data, ref_data, E_pose):
        raise NotImplementedError("iterative scale recovery is a 'next' row (SURVEY.md 8f rank 1)")
