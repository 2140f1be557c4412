"""CPU: the PnP RANSAC of csrc/pnp.cu (run in the host-emulation build) against cv2.solvePnPRansac and the reference
PnpTracker golden.  The same checks run on the GPU in test_gpu_depth_pose.py."""
import os
import sys

import numpy as np

import pnp_cases

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _engine(hostsim_lib):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from runtime import HostsimRuntime
    from b200 import runtime as rt_mod, tracking
    rt_mod.set_runtime(HostsimRuntime(hostsim_lib))
    return tracking.Engine(376, 1241)


def test_pnp_ransac_vs_cv2(hostsim_lib):
    exact, total = pnp_cases.check_vs_cv2(_engine(hostsim_lib))
    print("repeats reproduced exactly: %d / %d" % (exact, total))


def test_epnp_cooperative_vs_sequential_vs_cv2(hostsim_lib):
    worst = pnp_cases.check_epnp_minimal(_engine(hostsim_lib), samples=4)
    print("EPnP minimal solver vs cv2: %.2e rad, %.2e" % tuple(worst))


def test_scale_ransac_on_device_vs_sklearn(hostsim_lib):
    worst = pnp_cases.check_scale_ransac_vs_sklearn(_engine(hostsim_lib), cases=60)
    print("scale RANSAC vs sklearn: worst relative difference %.2e" % worst)


def test_cooperative_five_point_vs_sequential(hostsim_lib):
    iters = pnp_cases.check_coop_five_point_vs_sequential(_engine(hostsim_lib))
    print("RANSAC iterations:", iters)


def test_fused_tracker_tail_vs_stepwise(hostsim_lib):
    pnp_cases.check_fused_tail_vs_stepwise(_engine(hostsim_lib))


def test_pnp_tracker_vs_reference_golden(hostsim_lib):
    worst = pnp_cases.check_vs_reference_golden(_engine(hostsim_lib), np.load(os.path.join(G, "trackers_2000.npz")))
    print("worst rotation / relative translation difference: %.2e rad, %.2e" % worst)


def test_homography_ransac_gric_vs_cv2(hostsim_lib):
    worst = pnp_cases.check_homography_vs_cv2(_engine(hostsim_lib))
    print("worst relative GRIC-H difference: %.2e" % worst)


def test_pose_solver_edge_cases(hostsim_lib):
    """Sentinel behaviour at the degenerate ends (pnp_tracker.py:97,113-118; E_tracker.py:196,216-217,299-300):
    fewer than 5 PnP correspondences -> identity pose and no solver call; hopeless correspondences -> the solvers still
    return a model (RANSAC always keeps its best) but the essential-matrix tracker falls back to R = I, t = 0;
    no more than 10 keypoints -> E-tracker gives up before consuming any randomness."""
    import synthdata
    from b200 import tracking
    eng = _engine(hostsim_lib)
    K = synthdata.kitti_intrinsics()
    rs = np.random.RandomState(3)
    # PnP with 4 points: n > 4 is false -> identity, but the five shuffles are still drawn (pnp_tracker.py:90-97)
    k1, k2, d = rs.uniform(0, 300, (4, 2)), rs.uniform(0, 300, (4, 2)), rs.uniform(5, 20, 4)
    np.random.seed(1)
    T, ninl = tracking.compute_pose_3d2d(eng, k1, k2, d, K)
    a = np.random.randint(0, 2 ** 31 - 1)
    np.random.seed(1)
    for _ in range(5):
        np.random.shuffle(np.arange(4))
    assert np.array_equal(T, np.eye(4)) and ninl == 0 and a == np.random.randint(0, 2 ** 31 - 1)
    # E-tracker with 10 keypoints: early out, RNG untouched (E_tracker.py:196)
    kp = rs.uniform(0, 300, (10, 2))
    np.random.seed(2)
    r = tracking.compute_pose_2d2d(eng, kp, kp + 1.0, K)
    after = np.random.randint(0, 2 ** 31 - 1)
    assert not r["valid"] and np.array_equal(r["R"], np.eye(3)) and not r["t"].any() and r["inliers"].all()
    np.random.seed(2)
    assert after == np.random.randint(0, 2 ** 31 - 1)
    # pure-noise correspondences: finite outputs, identity pose or a rejected / low-support model
    kp_ref = np.stack([rs.uniform(0, 1241, 400), rs.uniform(0, 376, 400)], 1)
    kp_cur = np.stack([rs.uniform(0, 1241, 400), rs.uniform(0, 376, 400)], 1)
    np.random.seed(3)
    r = tracking.compute_pose_2d2d(eng, kp_ref, kp_cur, K)
    assert np.isfinite(r["R"]).all() and np.isfinite(r["t"]).all() and r["inliers"].sum() < 60
    T, ninl = tracking.compute_pose_3d2d(eng, kp_ref, kp_cur, rs.uniform(5, 40, 400), K)
    assert np.isfinite(T).all() and ninl < 60


def test_homography_edge_cases_vs_cv2(hostsim_lib):
    """Degenerate ends of findHomography: all points collinear (getSubset never passes checkSubset: cv2 returns None, the
    device reports found = 0), an exact 12-point homography, the 5-point minimum above the 4-point model, pure noise
    (RANSAC runs all 2000 iterations and keeps a 5-inlier model) -- inlier counts / masks equal, H to 1e-8."""
    import cv2
    eng = _engine(hostsim_lib)
    rt = eng.rt
    rs = np.random.RandomState(0)

    def run(p1, p2):
        H, mask = cv2.findHomography(p1, p2, method=cv2.RANSAC, confidence=0.99, ransacReprojThreshold=1)
        h = eng.homography_launch(rt.from_host(p1), rt.from_host(p2), p1.shape[0])
        rt.wait_event(h["done"])
        return H, mask, h["H"].numpy().reshape(3, 3), h["mask"].numpy(), h["info"].numpy()

    x = np.linspace(10, 500, 50)
    p1 = np.stack([x, 2 * x + 5], 1)
    H, mask, Hd, md, info = run(p1, p1 + np.array([3.0, 1.0]))
    assert H is None and info[0] == 0 and not md.any()
    p1 = rs.uniform(0, 500, (12, 2))
    Ht = np.array([[1.01, 0.02, 3], [-0.01, 0.99, -2], [1e-5, 2e-5, 1]])
    q = (Ht @ np.c_[p1, np.ones(12)].T).T
    p2 = q[:, :2] / q[:, 2:]
    for n in (12, 5):
        H, mask, Hd, md, info = run(p1[:n].copy(), p2[:n].copy())
        assert info[0] == 1 and info[1] == n == int(mask.sum()) and np.abs(H - Hd).max() < 1e-8
    p1, p2 = rs.uniform(0, 1000, (300, 2)), rs.uniform(0, 1000, (300, 2))
    H, mask, Hd, md, info = run(p1, p2)
    assert info[2] == 2000 and np.array_equal(mask.ravel(), md) and np.abs(H - Hd).max() / np.abs(H).max() < 1e-8
