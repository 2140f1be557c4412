"""Synthetic KITTI-odometry-shaped sequence on disk + analytic flow/depth injection, for running the
reference DRIVER (libs/dfvo.py, unmodified) end to end -- once against the reference's own hot-path
packages (golden generation) and once against the dfvo_b200 ``libs`` mirror (tests).  TEST INFRA ONLY."""
import os

import numpy as np

from . import synth


def write_sequence(root, n_frames, h, w, seq="00"):
    """dataset/kitti_odom layout the reference KittiOdom loader expects (kitti.py:63-160): calib.txt with
    a P2 row, image_2/%06d.png.  Returns the intrinsics [cx, cy, fx, fy] at (h, w)."""
    import cv2
    d = os.path.join(root, seq, "image_2")
    os.makedirs(d, exist_ok=True)
    for i in range(n_frames):
        cv2.imwrite(os.path.join(d, "%06d.png" % i), synth.value_noise_image(h, w, 100 + i)[:, :, ::-1])
    # utils.load_kitti_odom_intrinsics rescales from the raw 370x1226 KITTI size (utils.py:240-262)
    fx = fy = 718.856
    cx, cy = 607.1928, 185.2157
    with open(os.path.join(root, seq, "calib.txt"), "w") as f:
        for k in range(4):
            f.write("P%d: %.6f 0 %.6f 0 0 %.6f %.6f 0 0 0 1 0\n" % (k, fx, cx, fy, cy))
    return [cx / 1226.0 * w, cy / 370.0 * h, fx / 1226.0 * w, fy / 370.0 * h]


frame_inputs = synth.frame_inputs
MODES = synth.SEQUENCE_MODES


def patch_deep_model(DeepModel, h, w, K):
    """Replace the two network calls of the facade by the analytic frame inputs (identical for the
    reference run and the dfvo_b200 run); everything downstream of them is the code under test."""
    def initialize_models(self):
        class _D:
            feed_height, feed_width = h, w
        self.depth = _D()
        self.flow = None

    def forward_depth(self, imgs):
        return frame_inputs(self._t, h, w, K, MODES[self._t % len(MODES)])["depth"]

    def forward_flow(self, in_cur_data, in_ref_data, forward_backward):
        t = in_cur_data["id"]
        f = frame_inputs(t, h, w, K, MODES[t % len(MODES)])
        s, g = in_ref_data["id"], in_cur_data["id"]
        return {(s, g): f["fwd"], (g, s): f["bwd"], (s, g, "diff"): f["diff"]}

    DeepModel.initialize_models = initialize_models
    DeepModel.forward_depth = forward_depth
    DeepModel.forward_flow = forward_flow


def canonical_order(kp1, h, w, rows=10, cols=10):
    """Permutation that sorts selected keypoints cell-major, then by ascending pixel index -- the order
    the dfvo_b200 selection kernel emits.  The reference's order inside a cell is whatever
    ``np.argpartition`` left (implementation-defined); RANSAC draws depend on the order, so end-to-end
    pose comparisons hand both sides identically ordered arrays."""
    x, y = kp1[:, 0].astype(int), kp1[:, 1].astype(int)
    ye = np.array([int(h / rows * r) for r in range(rows + 1)])
    xe = np.array([int(w / cols * c) for c in range(cols + 1)])
    r = np.searchsorted(ye, y, side="right") - 1
    c = np.searchsorted(xe, x, side="right") - 1
    return np.lexsort((y * w + x, r * cols + c))


def patch_canonical_kp_order(KeypointSampler):
    orig = KeypointSampler.update_kp_data

    def update_kp_data(self, cur_data, ref_data, kp_sel_outputs):
        orig(self, cur_data, ref_data, kp_sel_outputs)
        if "kp_best" in ref_data and hasattr(ref_data["kp_best"], "shape"):
            h, w = cur_data["depth"].shape
            o = canonical_order(ref_data["kp_best"], h, w, self.cfg.kp_selection.local_bestN.num_row,
                                self.cfg.kp_selection.local_bestN.num_col)
            ref_data["kp_best"], cur_data["kp_best"] = ref_data["kp_best"][o], cur_data["kp_best"][o]
    KeypointSampler.update_kp_data = update_kp_data


def run_driver(dfvo_module, cfg, n_frames):
    """Run the reference driver class on ``n_frames`` and return the global poses [n,4,4]."""
    vo = dfvo_module.DFVO(cfg)
    orig = vo.deep_model_inference

    def wrapped():
        vo.deep_models._t = vo.cur_data["id"]
        return orig()
    vo.deep_model_inference = wrapped
    vo.main()
    return np.stack([vo.global_poses[i].pose for i in range(n_frames)])
