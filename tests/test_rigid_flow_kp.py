"""CPU: the rigid-flow keypoint / iterative scale row (E_tracker.py:509-569,645-705, kp_selection.py:203-324)
run in the host-emulation build against the oracle and the reference golden."""
import os
import sys

import pytest

import rigid_cases


def _engine(hostsim_lib):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from runtime import HostsimRuntime
    from b200 import runtime as rt_mod, tracking
    rt_mod.set_runtime(HostsimRuntime(hostsim_lib))
    return tracking.Engine(376, 1241)


@pytest.mark.parametrize("name", ["outliers"])
def test_rigid_flow_map_and_selection(hostsim_lib, name):
    print("max |map - oracle| = %.2e px" % rigid_cases.check_maps_and_selection(_engine(hostsim_lib), name))


@pytest.mark.parametrize("name,kp_src", [("clean", "kp_depth"), ("outliers", "kp_best")])
def test_iterative_scale_vs_reference(hostsim_lib, name, kp_src):
    rigid_cases.check_iterative_scale(_engine(hostsim_lib), name, kp_src)


@pytest.mark.parametrize("kp_src", ["kp_best"])
def test_mirror_esstracker_iterative_scale(hostsim_lib, kp_src):
    """The reference-API mirror (df-vo_b200/libs/tracker/E_tracker.py) with kp_selection.rigid_flow_kp.enable and
    scale_recovery.method 'iterative' (kitti_*_extend.yml): scale_recovery() and compute_rigid_flow_kp() against the golden
    produced by the reference classes."""
    import numpy as np
    import synthdata
    from b200 import config, tracking
    eng = _engine(hostsim_lib)
    tracking._default_engine = eng
    from libs.geometry.camera_modules import SE3, Intrinsics
    from libs.tracker.E_tracker import EssTracker
    cfg = config.default_cfg(376, 1241)
    cfg.kp_selection.rigid_flow_kp.enable = True
    cfg.scale_recovery.method = "iterative"
    cfg.scale_recovery.kp_src = kp_src
    K = synthdata.kitti_intrinsics(376, 1241)
    ess = EssTracker(cfg, Intrinsics(K), None)
    g = np.load(os.path.join(rigid_cases.G, "rigid_flow_kp_376x1241.npz"))
    fr, depth_proc, kp1, kp2 = rigid_cases.frame("clean")
    key = "clean_" + kp_src
    cur = {"depth": depth_proc, "raw_depth": fr["depth"], "kp_best": kp2}
    ref = {"flow": fr["flow_fwd"], "flow_diff": fr["flow_diff"], "raw_depth": fr["depth"], "depth": depth_proc, "kp_best": kp1}
    np.random.seed(4869)
    r = ess.compute_pose_2d2d(kp1, kp2, True)
    so = ess.scale_recovery(cur, ref, r["pose"], False)
    want = float(g[key + "_scale"])
    assert abs(so["scale"] - want) < 1e-6 * want and abs(ess.prev_scale - want) < 1e-6 * want
    assert int(np.random.randint(0, 2 ** 31 - 1)) == int(g[key + "_rng_after"])
    assert so["ref_kp_depth"].shape == (2000, 2) and np.asarray(so["rigid_flow_mask"]).shape == (376, 1241)
    if kp_src == "kp_best":
        assert np.array_equal(so["ref_kp_depth"].astype(np.int32), g[key + "_kp1_uniform"])
        hyb = SE3(r["pose"].pose.copy())
        hyb.t = r["pose"].t * so["scale"]
        ess.compute_rigid_flow_kp(cur, ref, hyb)
        lin = (ref["kp_depth"][:, 1] * 1241 + ref["kp_depth"][:, 0]).astype(np.int64)
        assert np.array_equal(np.sort(lin), g[key + "_best_idx_sorted"])
        assert np.array_equal(ref["kp_depth_uniform"].astype(np.int32), g[key + "_kp1_uniform_hyb"])
