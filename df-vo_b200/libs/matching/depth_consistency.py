"""``DepthConsistency`` (libs/matching/depth_consistency.py) needs the experimental PoseNet and is disabled
in every shipped configuration (default_configuration.yml:124-125); it is outside the dfvo_b200 hot path
(SURVEY.md section 2 row 8).  The class exists so ``libs/dfvo.py`` imports unchanged."""


class DepthConsistency:
    def __init__(self, cfg, cam_intrinsics):
        raise NotImplementedError("kp_selection.depth_consistency needs deep_pose; outside the dfvo_b200 hot path")
