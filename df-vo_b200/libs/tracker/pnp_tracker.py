"""``PnpTracker`` with the reference's interface (libs/tracker/pnp_tracker.py:23-212).

The PnP fallback only runs when the essential-matrix tracker is rejected (dfvo.py:227).  This round the
keypoint filtering / unprojection are vectorised NumPy and the RANSAC itself is still
``cv2.solvePnPRansac`` on the host -- the EPnP + LM device port is listed as open work in DESIGN.md; it is
not a fallback for a CUDA path, it is the one row of SURVEY 8(a) not yet moved to the device."""
import numpy as np

from libs.geometry.camera_modules import SE3
from libs.geometry.ops_3d import unprojection_kp


class PnpTracker:
    def __init__(self, cfg, cam_intrinsics):
        self.cfg = cfg
        self.cam_intrinsics = cam_intrinsics
        assert not cfg.kp_selection.rigid_flow_kp.enable, "rigid_flow_kp is a 'next' row (SURVEY.md 8f rank 1)"

    def compute_pose_3d2d(self, kp1, kp2, depth_1, is_iterative):
        """pnp_tracker.py:45-125 -> {'pose': SE3 (view-2 -> view-1), 'kp1', 'kp2'}."""
        import cv2
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
        XYZ = unprojection_kp(kp1, d[keep], self.cam_intrinsics)
        best_rt, best_inl = [], 0
        repeat = self.cfg.pnp_tracker.ransac.repeat if is_iterative else 3
        for _ in range(repeat):
            order = np.arange(0, kp2.shape[0], 1)
            np.random.shuffle(order)                                   # host RNG, as the reference (:92)
            nX, n2 = XYZ.copy()[order], kp2.copy()[order]
            if n2.shape[0] > 4:
                flag, r, t, inl = cv2.solvePnPRansac(objectPoints=nX, imagePoints=n2, cameraMatrix=self.cam_intrinsics.mat,
                                                     distCoeffs=None, iterationsCount=self.cfg.pnp_tracker.ransac.iter,
                                                     reprojectionError=self.cfg.pnp_tracker.ransac.reproj_thre)
                if flag and inl.shape[0] > best_inl:
                    best_rt, best_inl = [r, t], inl.shape[0]
        pose = SE3()
        if len(best_rt) != 0:
            pose.R = cv2.Rodrigues(best_rt[0])[0]
            pose.t = best_rt[1]
        pose.pose = pose.inv_pose                                      # :118 (solvePnP gives ref -> cur)
        return {"pose": pose, "kp1": kp1, "kp2": kp2}

    def compute_rigid_flow_kp(self, cur_data, ref_data, pose):
        raise NotImplementedError("rigid-flow keypoints are a 'next' row (SURVEY.md 8f rank 1)")
