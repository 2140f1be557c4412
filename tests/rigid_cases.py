"""Shared checks of the rigid-flow keypoint / iterative scale row against the oracle and the reference
golden (tests/golden/rigid_flow_kp_376x1241.npz); run on the CPU emulation build by test_rigid_flow_kp.py and on the GPU by
test_gpu_selection.py."""
import os

import numpy as np

import synthdata
from b200 import hostmath, tracking
from oracle import vo

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
H, W = 376, 1241
CASES = {"clean": dict(seed=51), "outliers": dict(seed=52, outlier_frac=0.3, diff_sigma=0.12)}


def frame(name):
    fr = synthdata.analytic_frame(h=H, w=W, **CASES[name])
    depth_proc = (fr["depth"] * ((fr["depth"] < 50) & (fr["depth"] > 0))).astype(np.float64)
    good, cells = vo.local_bestn_indices(fr["flow_diff"])
    kp1, kp2 = vo.keypoints_from_indices([np.sort(np.concatenate(cells))], fr["flow_fwd"], W)
    return fr, depth_proc, kp1, kp2


def check_maps_and_selection(engine, name):
    """Device rigid-flow inconsistency map vs the oracle (float32 operation order differs: 2e-3 px on ~10^3 px coordinates),
    and -- fed the oracle's map so that no pixel sits on a threshold differently -- the 'uniform' and 'best' keypoint lists
    bit-equal to the oracle's and to the reference golden."""
    g = np.load(os.path.join(G, "rigid_flow_kp_376x1241.npz"))
    K = synthdata.kitti_intrinsics(H, W)
    rt = engine.rt
    fr, depth_proc, kp1, kp2 = frame(name)
    key = name + "_kp_best"
    T = g[key + "_rigid_flow_pose"]
    d_depth, d_flow = rt.from_host(fr["depth"]), rt.from_host(fr["flow_fwd"][None])
    d_diff = rt.from_host(np.ascontiguousarray(fr["flow_diff"][:, :, 0])[None])
    o = engine.rigid_flow_keypoints(d_depth, d_flow, d_diff, T, K, score_method="opt_flow")
    dev_map = o["rigid_flow_diff"].numpy()
    want_map = vo.rigid_flow_diff(fr["depth"], fr["flow_fwd"], T, K)
    err = np.abs(dev_map - want_map)
    assert err.max() < 2e-3 * max(1.0, float(np.abs(want_map).max()) / 30), err.max()
    assert np.abs(dev_map[::9, ::9] - g[key + "_rigid_flow_diff_s9"]).max() < 5e-3
    # selection on the oracle's map (uploaded over the device map) -> exact lists
    rf = engine._rf[(10, 10, 20)]
    rf["map"].upload(want_map)
    cells, quota = 100, 20
    u, b = rf["u"], rf["b"]
    lib, st = engine.lib, rt.stream_ptr()
    lib.check(lib.dfvo_uniform_cells(rf["map"].ptr, d_diff.ptr, H, W, 10, 10, 2000, 5.0, 0.1, u["idx"].ptr, u["cc"].ptr, st))
    lib.check(lib.dfvo_local_bestn(d_diff.ptr, rf["map"].ptr, H, W, 10, 10, 2000, 0.1, 5.0, b["idx"].ptr, b["cc"].ptr,
                                   rf["st"].ptr, st))
    best, uniform = vo.opt_rigid_flow_kp(want_map, fr["flow_diff"][:, :, 0], score_method="opt_flow")
    iu, cu = u["idx"].numpy().reshape(cells, quota), u["cc"].numpy()
    ib, cb = b["idx"].numpy().reshape(cells, quota), b["cc"].numpy()
    for c in range(cells):
        assert cu[c] == len(uniform[c]) and np.array_equal(iu[c, :cu[c]], uniform[c]), ("uniform", c)
        assert cb[c] == len(best[c]) and np.array_equal(ib[c, :cb[c]], best[c]), ("best", c)
    lin = np.concatenate([iu[c, :cu[c]] for c in range(cells)])
    assert np.array_equal(np.stack([lin % W, lin // W], 1).astype(np.int32), g[key + "_kp1_uniform"])
    return float(err.max())


def check_iterative_scale(engine, name, kp_src):
    """tracking.scale_recovery_iterative on the device port vs EssTracker.scale_recovery (method 'iterative') of the
    reference: same E-pose, same RNG consumption, scale to 1e-9 for kp_src kp_best (the selection does not enter) and 1e-6
    for kp_src kp_depth (it does, through a float32 map whose last bits differ from torch's)."""
    g = np.load(os.path.join(G, "rigid_flow_kp_376x1241.npz"))
    K = synthdata.kitti_intrinsics(H, W)
    rt = engine.rt
    fr, depth_proc, kp1, kp2 = frame(name)
    key = "%s_%s" % (name, kp_src)
    np.random.seed(4869)
    r = tracking.compute_pose_2d2d(engine, kp1, kp2, K)
    E = np.eye(4); E[:3, :3] = r["R"]; E[:3, 3:] = r["t"]
    assert np.abs(E - g[key + "_E_pose"]).max() < 1e-9
    d_depth, d_flow = rt.from_host(fr["depth"]), rt.from_host(fr["flow_fwd"][None])
    d_diff = rt.from_host(np.ascontiguousarray(fr["flow_diff"][:, :, 0])[None])
    select = lambda T, sm: engine.rigid_flow_keypoints(d_depth, d_flow, d_diff, T, K, score_method=sm, want_best=False)
    find = lambda a, b: tracking.find_scale_from_depth(engine, a, b, np.linalg.inv(E), depth_proc, K)
    o = tracking.scale_recovery_iterative(select, find, E, 0, (kp1, kp2), kp_src=kp_src)
    want = float(g[key + "_scale"])
    assert abs(o["scale"] - want) < (1e-9 if kp_src == "kp_best" else 1e-6) * abs(want), (o["scale"], want)
    assert int(np.random.randint(0, 2 ** 31 - 1)) == int(g[key + "_rng_after"])
    return o["scale"]
