"""Hot-path configuration keys with the reference's default values
(options/examples/default_configuration.yml:7-168, SURVEY.md section 5 'Config / flags').  The libs
mirror reads the caller's cfg object (the reference's EasyDict) directly; this dictionary is what the
device pipeline / bench use when no DF-VO configuration file is around."""


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(d):
    return AttrDict({k: _wrap(v) if isinstance(v, dict) else v for k, v in d.items()})


def default_cfg(height=376, width=1241):
    return _wrap({
        "dataset": "kitti_odom", "seed": 4869, "frame_step": 1,
        "image": {"height": height, "width": width},
        "depth": {"max_depth": 50, "min_depth": 0, "depth_src": None, "deep_depth": {"network": "monodepth2", "pretrained_model": None}},
        "deep_flow": {"network": "liteflow", "flow_net_weight": None, "forward_backward": True},
        "deep_pose": {"enable": False},
        "online_finetune": {"enable": False, "lr": 0.00001, "num_frames": 200,
                            "flow": {"enable": False, "scales": [1, 2, 3, 4, 5], "loss": {"flow_consistency": 0.005, "flow_smoothness": 0.1}},
                            "depth": {"enable": False, "scales": [0, 1, 2, 3], "pose_src": "DF-VO",
                                      "loss": {"apperance_loss": 1, "disparity_smoothness": 0.001, "depth_consistency": 0.001}},
                            "pose": {"enable": False}},
        "crop": {"depth_crop": [[0.3, 1], [0, 1]], "flow_crop": [[0, 1], [0, 1]]},
        "kp_selection": {
            "local_bestN": {"enable": True, "num_bestN": 2000, "num_row": 10, "num_col": 10, "score_method": "flow", "thre": 0.1},
            "bestN": {"enable": False, "num_bestN": 2000},
            "sampled_kp": {"enable": False, "num_kp": 2000},
            "depth_consistency": {"enable": False, "thre": 0.05},
            "rigid_flow_kp": {"enable": False, "num_bestN": 2000, "num_row": 10, "num_col": 10, "score_method": "opt_flow",
                              "rigid_flow_thre": 5, "optical_flow_thre": 0.1},
        },
        "tracking_method": "hybrid",
        "e_tracker": {"ransac": {"reproj_thre": 0.2, "repeat": 5}, "validity": {"method": "GRIC"}, "kp_src": "kp_best",
                      "iterative_kp": {"enable": False, "kp_src": "kp_depth", "score_method": "opt_flow"}},
        "scale_recovery": {"method": "simple", "kp_src": "kp_best", "iterative_kp": {"enable": False, "kp_src": "kp_depth", "score_method": "rigid_flow"},
                           "ransac": {"method": "depth_ratio", "min_samples": 3, "max_trials": 100, "stop_prob": 0.99, "thre": 0.1}},
        "pnp_tracker": {"ransac": {"iter": 100, "reproj_thre": 1, "repeat": 5}, "kp_src": "kp_best",
                        "iterative_kp": {"enable": False, "kp_src": "kp_depth", "score_method": "rigid_flow"}},
    })
