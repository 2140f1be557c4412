"""``KeypointSampler`` with the reference's interface (libs/matching/keypoint_sampler.py:18-163)."""
import numpy as np

from .kp_selection import bestN_flow_kp, local_bestN, sampled_kp


class KeypointSampler:
    def __init__(self, cfg):
        self.cfg = cfg
        self.kps = {}
        if cfg.kp_selection.sampled_kp.enable:
            self.kps["uniform"] = self.generate_kp_samples(cfg.image.height, cfg.image.width, cfg.crop.flow_crop,
                                                           cfg.kp_selection.sampled_kp.num_kp)

    def get_feat_track_methods(self, method_idx):
        return {1: "deep_flow"}[method_idx]

    def generate_kp_samples(self, img_h, img_w, crop, N):
        """keypoint_sampler.py:51-74."""
        y0, y1 = int(crop[0][0] * img_h), int(crop[0][1] * img_h)
        x0, x1 = int(crop[1][0] * img_w), int(crop[1][1] * img_w)
        return np.linspace(0, (x1 - x0) * (y1 - y0) - 1, N, dtype=int)

    def kp_selection(self, cur_data, ref_data):
        """keypoint_sampler.py:76-143.  The dense float64 grids the reference builds (2 x 7.5 MB per frame)
        are never materialised: the selection kernels work on the device-resident flow."""
        outputs = {"good_kp_found": True}
        sel = self.cfg.kp_selection
        if sel.local_bestN.enable:
            outputs.update(local_bestN(None, None, ref_data, self.cfg, outputs))
        elif sel.bestN.enable:
            outputs.update(bestN_flow_kp(None, None, ref_data, self.cfg, outputs))
        if sel.sampled_kp.enable:
            outputs.update(sampled_kp(None, None, ref_data, self.kps["uniform"], self.cfg, outputs))
        return outputs

    def update_kp_data(self, cur_data, ref_data, kp_sel_outputs):
        """keypoint_sampler.py:145-163."""
        sel = self.cfg.kp_selection
        if sel.local_bestN.enable or sel.bestN.enable:
            ref_data["kp_best"] = kp_sel_outputs["kp1_best"][0]
            cur_data["kp_best"] = kp_sel_outputs["kp2_best"][0]
            cur_data["fb_flow_mask"] = kp_sel_outputs["fb_flow_mask"]
        if sel.sampled_kp.enable:
            ref_data["kp_list"] = kp_sel_outputs["kp1_list"][0]
            cur_data["kp_list"] = kp_sel_outputs["kp2_list"][0]
