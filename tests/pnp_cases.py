"""Shared checks of the device PnP RANSAC (csrc/pnp.cu) against cv2.solvePnPRansac and the reference PnpTracker golden;
run on the CPU emulation build by test_pnp.py and on the GPU by test_gpu_depth_pose.py.

Tolerance: OpenCV's EPnP reads its null-space vectors from the LEFT singular vectors of a rank-deficient 12x12 matrix
(5 points = 10 equations), i.e. from normalised round-off, so its minimal-sample poses are only reproducible to ~1e-6;
the RANSAC trajectory usually still coincides (same winning iteration, same inlier count, refit pose equal to 1e-12).
When it does not, the inlier sets differ by a few borderline points and the refit poses by < 1e-4 rad / 1e-3 |t|
(BASELINE north-star tolerance), which is what is asserted; the fraction of exactly reproduced repeats is asserted too."""
import numpy as np

import synthdata
from b200 import hostmath, tracking


def scene(seed, outlier_frac, noise, n=1500, h=376, w=1241):
    K = synthdata.kitti_intrinsics(h, w)
    cx, cy, fx, fy = K
    Kmat = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    rs = np.random.RandomState(seed)
    kp1 = np.stack([rs.uniform(0, w, n), rs.uniform(0.3 * h, h, n)], 1)
    d = rs.uniform(4, 45, n)
    XYZ = (np.linalg.inv(Kmat) @ np.concatenate([kp1, np.ones((n, 1))], 1).T).T * d[:, None]
    rvec = np.array([0.002, 0.015, -0.001]) * (1 + rs.rand())
    tvec = np.array([0.03, -0.01, -0.8])
    Xc = (hostmath.rodrigues(rvec) @ XYZ.T).T + tvec
    kp2 = np.stack([fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy], 1) + rs.standard_normal((n, 2)) * noise
    bad = rs.rand(n) < outlier_frac
    kp2[bad] += rs.uniform(-40, 40, (int(bad.sum()), 2))
    return K, Kmat, kp1, d, XYZ, kp2


def pose_delta(Ra, ta, Rb, tb):
    dR = Ra.T @ Rb
    ang = float(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)))
    return ang, float(np.linalg.norm(ta - tb) / max(np.linalg.norm(tb), 1e-12))


def check_vs_cv2(engine):
    """Per repeat: the device RANSAC either reproduces cv2.solvePnPRansac exactly (inlier count equal, pose to 1e-9) or
    lands on a neighbouring consensus set (see the module docstring); the best-of-5 result the tracker returns
    (pnp_tracker.py:108-110) must agree within the north-star tolerance in every scene."""
    import cv2
    exact = total = 0
    for seed, outl, noise in [(1, 0.0, 0.05), (2, 0.3, 0.05), (3, 0.5, 0.2), (4, 0.1, 0.5), (5, 0.3, 0.3), (6, 0.6, 0.1)]:
        K, Kmat, kp1, d, XYZ, kp2 = scene(seed, outl, noise)
        n = XYZ.shape[0]
        np.random.seed(7)
        perms = []
        for _ in range(5):
            o = np.arange(n)
            np.random.shuffle(o)
            perms.append(o)
        rt, info = engine.pnp_ransac(XYZ, kp2, perms, K, 100, 1.0)
        ref = []
        for r in range(5):
            flag, rv, tv, inl = cv2.solvePnPRansac(objectPoints=XYZ[perms[r]].copy(), imagePoints=kp2[perms[r]].copy(), cameraMatrix=Kmat,
                                                   distCoeffs=None, iterationsCount=100, reprojectionError=1)
            assert bool(info[r, 0]) == bool(flag)
            ref.append((inl.shape[0], cv2.Rodrigues(rv)[0], tv.ravel()))
            ang, dt = pose_delta(hostmath.rodrigues(rt[r, :3]), rt[r, 3:], ref[r][1], ref[r][2])
            assert abs(int(info[r, 1]) - inl.shape[0]) <= 0.08 * inl.shape[0], (seed, r, info[r], inl.shape[0])
            assert ang < 1e-3 and dt < 5e-2, (seed, r, ang, dt)
            total += 1
            exact += int(info[r, 1] == inl.shape[0] and ang < 1e-9 and dt < 1e-9)
        # what the tracker returns: the first repeat with the largest inlier count
        bc = max(range(5), key=lambda r: (ref[r][0], -r))
        bd = max(range(5), key=lambda r: (int(info[r, 1]), -r))
        ang, dt = pose_delta(hostmath.rodrigues(rt[bd, :3]), rt[bd, 3:], ref[bc][1], ref[bc][2])
        assert ang < 1e-4 and dt < 1e-3, ("best of 5", seed, ang, dt)
    assert exact >= 0.8 * total, "only %d of %d repeats reproduce cv2.solvePnPRansac exactly" % (exact, total)
    return exact, total


def check_vs_reference_golden(engine, g):
    """tracking.compute_pose_3d2d after the same RNG consumption as the reference run (E-tracker shuffles, scale RANSAC)
    against PnpTracker.compute_pose_3d2d of the unmodified reference (tests/golden/trackers_2000.npz)."""
    K = synthdata.kitti_intrinsics()
    cases = {"out00": dict(seed=31, outlier_frac=0.0), "out30": dict(seed=32, outlier_frac=0.3),
             "out60": dict(seed=33, outlier_frac=0.6), "still": dict(seed=34, outlier_frac=0.1, zero_motion=True)}
    worst = (0.0, 0.0)
    for name, kw in cases.items():
        kp_ref, kp_cur, info = synthdata.correspondences(n=2000, **kw)
        np.random.seed(4869)
        r = tracking.compute_pose_2d2d(engine, kp_ref, kp_cur, K)
        depth = info["depth"].astype(np.float32)
        dp = (depth * ((depth < 50) & (depth > 0))).astype(np.float64)
        if np.linalg.norm(r["t"]) != 0:
            pose = np.eye(4); pose[:3, :3] = r["R"]; pose[:3, 3:] = r["t"]
            tracking.find_scale_from_depth(engine, kp_ref, kp_cur, np.linalg.inv(pose), dp, K)
        # PnpTracker's keypoint filter (pnp_tracker.py:64-78)
        keep = (kp_cur[:, 0] >= 0) & (kp_cur[:, 0] < 1241)
        k1, k2 = kp_ref[keep], kp_cur[keep]
        keep = (k2[:, 1] >= 0) & (k2[:, 1] < 376)
        k1, k2 = k1[keep], k2[keep]
        ki = k1.astype(int)
        d = dp[ki[:, 1], ki[:, 0]]
        keep = (d != 0) & (d < 50) & (d > 0)
        k1, k2, d = k1[keep], k2[keep], d[keep]
        assert k1.shape[0] == int(g[name + "_pnp_nkp"])
        T, ninl = tracking.compute_pose_3d2d(engine, k1, k2, d, K)
        want = g[name + "_pnp_pose"]
        ang, dt = pose_delta(T[:3, :3], T[:3, 3], want[:3, :3], want[:3, 3])
        assert ang < 1e-4 and dt < 1e-3, (name, ang, dt)
        assert int(np.random.randint(0, 2 ** 31 - 1)) == int(g[name + "_rng_after"]), "host RNG position differs from the reference run"
        worst = (max(worst[0], ang), max(worst[1], dt))
    return worst


def check_homography_vs_cv2(engine):
    """csrc/homog.cu against cv2.findHomography(RANSAC, 1 px, 0.99) + the reference GRIC-H (E_tracker.py:199-215):
    RANSAC inlier masks bit-equal, H to 1e-8 relative (measured 2e-10), GRIC-H to 1e-9 relative (measured 7e-12)."""
    import cv2
    from oracle import vo
    cases = {"out00": dict(seed=31, outlier_frac=0.0), "out30": dict(seed=32, outlier_frac=0.3), "out60": dict(seed=33, outlier_frac=0.6),
             "still": dict(seed=34, outlier_frac=0.1, zero_motion=True), "still60": dict(seed=35, outlier_frac=0.6, zero_motion=True)}
    worst = 0.0
    for name, kw in cases.items():
        kp_ref, kp_cur, _ = synthdata.correspondences(n=2000, **kw)
        n = kp_ref.shape[0]
        H, mask = cv2.findHomography(kp_cur, kp_ref, method=cv2.RANSAC, confidence=0.99, ransacReprojThreshold=1)
        want = vo.calc_gric(vo.homography_residual(H, kp_cur, kp_ref), 0.8, n, "HMat")
        h = engine.homography_launch(engine.rt.from_host(kp_cur), engine.rt.from_host(kp_ref), n)
        engine.rt.wait_event(h["done"])
        Hd, md, info, g = h["H"].numpy().reshape(3, 3), h["mask"].numpy(), h["info"].numpy(), float(h["gric"].numpy()[0])
        assert info[0] == 1 and np.array_equal(md, mask.ravel()), (name, info, int(mask.sum()))
        assert np.abs(Hd - H).max() / np.abs(H).max() < 1e-8, name
        assert abs(g - want) < 1e-9 * abs(want), (name, g, want)
        worst = max(worst, abs(g - want) / abs(want))
    return worst


def check_epnp_minimal(engine, samples=6):
    """csrc/pnp.cu's minimal solver through the stage entry dfvo_epnp_minimal: the lane-cooperative kernel (one warp per sample, one
    column of the 12x12 Jacobi SVD per lane) against the one-thread-per-sample kernel, and both against cv2.solvePnP(SOLVEPNP_EPNP)
    on the same 5 points.  Tolerances: cooperative vs sequential differ only in the summation order of the 12-term dot products, but
    EPnP's null-space basis is normalised round-off (pnp_cases docstring), so poses agree to ~1e-6, not to the last bit; vs OpenCV
    the same bound applies (rotation angle < 1e-5 rad, relative translation < 1e-5) on well-conditioned samples."""
    import ctypes
    import cv2
    K, Kmat, kp1, d, XYZ, kp2 = scene(11, 0.0, 0.0, n=400)
    cx, cy, fx, fy = K
    rs = np.random.RandomState(3)
    idx = np.stack([rs.choice(400, 5, replace=False) for _ in range(samples)])
    obj = np.ascontiguousarray(XYZ[idx].astype(np.float32).astype(np.float64).reshape(-1, 3))
    img = np.ascontiguousarray(kp2[idx].astype(np.float32).astype(np.float64).reshape(-1, 2))
    rt = engine.rt
    d_obj, d_img = rt.from_host(obj), rt.from_host(img)
    out = {}
    for coop in (0, 1):
        d_rt, d_ok = rt.zeros((samples, 12), np.float64), rt.zeros((samples,), np.int32)
        engine.lib.check(engine.lib.dfvo_epnp_minimal(d_obj.ptr, d_img.ptr, samples, fx, fy, cx, cy, coop, d_rt.ptr, d_ok.ptr, rt.stream_ptr()))
        rt.sync()
        out[coop] = (d_rt.numpy().copy(), d_ok.numpy().copy())
    assert out[0][1].all() and out[1][1].all()
    worst = [0.0, 0.0]
    for s in range(samples):
        Ra, ta = out[0][0][s, :9].reshape(3, 3), out[0][0][s, 9:]
        Rb, tb = out[1][0][s, :9].reshape(3, 3), out[1][0][s, 9:]
        ang, dt = pose_delta(Ra, ta, Rb, tb)
        assert ang < 1e-5 and dt < 1e-5, ("coop vs sequential", s, ang, dt)
        flag, rv, tv = cv2.solvePnP(obj[5 * s:5 * s + 5], img[5 * s:5 * s + 5], Kmat, None, flags=cv2.SOLVEPNP_EPNP)
        ang2, dt2 = pose_delta(Rb, tb, cv2.Rodrigues(rv)[0], tv.ravel())
        worst = [max(worst[0], ang2), max(worst[1], dt2)]
    assert worst[0] < 1e-4 and worst[1] < 1e-4, ("vs cv2 EPNP", worst)
    return worst


def check_scale_ransac_vs_sklearn(engine, cases=40):
    """Engine.ransac_scale (csrc/ransac.cu::k_scale_ransac: the RANSAC trial loop on the device, fed with NumPy's MT19937 state) against
    sklearn.linear_model.RANSACRegressor as E_tracker.py:626-636 calls it, from the same generator state: same scale (1e-12 relative;
    the final refit's long dot products are summed in a different order than BLAS does) and the SAME generator position afterwards
    (checked by drawing from both) -- over sample counts that exercise scikit-learn's three sampling branches' two reachable ones
    (permutation for n < 300, tracking selection above) and inlier ratios from 20 % to 100 %."""
    from sklearn import linear_model
    from b200 import hostmath
    rs = np.random.RandomState(99)
    worst = 0.0
    for case in range(cases):
        n = int(rs.choice([12, 40, 150, 299, 301, 700, 1500, 2000]))
        w = rs.uniform(0.2, 1.0)
        s_true = rs.uniform(0.5, 20.0)
        x = (1.0 / s_true) * (1.0 + rs.standard_normal(n) * 0.01)
        bad = rs.rand(n) > w
        x[bad] = rs.uniform(0.01, 3.0, int(bad.sum()))
        seed = 1000 + case
        np.random.seed(seed)
        np.random.rand(case % 7 * 89)                       # leave the generator at an arbitrary position (incl. near the 624 wrap)
        st0 = np.random.get_state()
        ransac = linear_model.RANSACRegressor(estimator=linear_model.LinearRegression(fit_intercept=False), min_samples=3, max_trials=100,
                                              stop_probability=0.99, residual_threshold=0.1)
        ransac.fit(x.reshape(-1, 1), np.ones((n, 1)))
        want = float(ransac.estimator_.coef_[0, 0])
        after_ref = np.random.randint(0, 2 ** 31 - 1, 4)
        np.random.set_state(st0)
        got = engine.ransac_scale(x, 3, 100, 0.99, 0.1, np.random)
        after_dev = np.random.randint(0, 2 ** 31 - 1, 4)
        np.random.set_state(st0)
        host = hostmath.ransac_scale(x, 3, 100, 0.99, 0.1, np.random)
        assert np.array_equal(after_ref, after_dev), ("generator position differs from sklearn's", case, n)
        assert abs(got - want) <= 1e-12 * abs(want), (case, n, got, want)
        assert abs(host - want) <= 1e-12 * abs(want)
        worst = max(worst, abs(got - want) / abs(want))
    return worst


def check_fused_tail_vs_stepwise(engine, only=None):
    """Engine.essential_tail (best repeat -> recoverPose -> GRIC vote -> cheirality gate -> depth ratios -> scale regressor on the
    device, one read) against the step-by-step host orchestration (tracking.compute_pose_2d2d + find_scale_from_depth) from the same
    generator state: same pose (bit-equal: same kernels on the same E), same vote, same scale to 1e-12 (T_21 is inverted
    analytically on the device, by LAPACK on the host) and the same generator position afterwards -- moving and still camera,
    0 / 30 / 60 % outliers."""
    K = synthdata.kitti_intrinsics()
    cases = {"out00": dict(seed=31, outlier_frac=0.0), "out30": dict(seed=32, outlier_frac=0.3), "out60": dict(seed=33, outlier_frac=0.6),
             "still": dict(seed=34, outlier_frac=0.1, zero_motion=True)}
    rt = engine.rt
    seen_scale = seen_reject = 0
    if only:                                       # the CPU emulation run keeps the two cheap cases; the GPU run all four
        cases = {k: v for k, v in cases.items() if k in only}
    for name, kw in cases.items():
        kp_ref, kp_cur, info = synthdata.correspondences(n=2000, **kw)
        n = kp_ref.shape[0]
        depth = info["depth"].astype(np.float32)
        dp32 = (depth * ((depth < 50) & (depth > 0))).astype(np.float32)
        # ---- step by step
        np.random.seed(4869)
        r = tracking.compute_pose_2d2d(engine, kp_ref, kp_cur, K)
        scale_a = None
        if np.linalg.norm(r["t"]) != 0:
            pose = np.eye(4); pose[:3, :3] = r["R"]; pose[:3, 3:] = r["t"]
            scale_a = tracking.find_scale_from_depth(engine, kp_ref, kp_cur, np.linalg.inv(pose), dp32.astype(np.float64), K)
        after_a = np.random.randint(0, 2 ** 31 - 1, 4)
        # ---- fused
        np.random.seed(4869)
        perms = []
        for _ in range(5):
            order = np.arange(0, n, 1)
            np.random.shuffle(order)
            perms.append(order)
        b_ref, b_cur, b_depth = rt.from_host(kp_ref), rt.from_host(kp_cur), rt.from_host(dp32)
        h = engine.homography_launch(b_cur, b_ref, n)
        w = engine.essential_launch(b_cur, b_ref, n, perms, K, threshold=0.2)
        o = engine.essential_tail(w, h, b_cur, b_ref, n, K, b_depth, np.random)
        after_b = np.random.randint(0, 2 ** 31 - 1, 4)
        assert o["valid"] == r["valid"] and np.array_equal(o["R"], r["R"]) and np.array_equal(o["t"], r["t"]), name
        assert o["cheirality"] == r["cheirality"] and np.array_equal(o["ransac_info"], r["ransac_info"]) and np.array_equal(o["E_gric"], r["E_gric"])
        assert np.array_equal(after_a, after_b), ("generator position", name)
        if scale_a is None:
            assert o["scale"] == -1 and o["scale_status"] == -3
            seen_reject += 1
        else:
            assert abs(o["scale"] - scale_a) <= 1e-12 * abs(scale_a), (name, o["scale"], scale_a)
            seen_scale += 1
    assert seen_scale >= (1 if only else 2) and seen_reject >= 1
    return seen_scale, seen_reject


def check_coop_five_point_vs_sequential(engine):
    """csrc/fivept.cuh::solve_coop (10 lanes per minimal sample) against the one-thread-per-sample solver inside the full RANSAC
    (DFVO_HYP_COOP is read per call): same inlier masks, iteration counts and GRIC, E to 1e-10."""
    import os
    K = synthdata.kitti_intrinsics()
    kp_ref, kp_cur, _ = synthdata.correspondences(seed=32, n=600, outlier_frac=0.3)
    n = kp_ref.shape[0]
    rt = engine.rt
    b_ref, b_cur = rt.from_host(kp_ref), rt.from_host(kp_cur)
    np.random.seed(3)
    perms = [np.random.permutation(n) for _ in range(2)]
    res = {}
    prev = os.environ.get("DFVO_HYP_COOP")
    try:
        for coop in ("0", "1"):
            os.environ["DFVO_HYP_COOP"] = coop
            w = engine.essential_launch(b_cur, b_ref, n, perms, K, threshold=0.2)
            res[coop] = (w["E"].numpy().copy(), w["mask"].numpy().copy(), w["info"].numpy().copy(), w["gric"].numpy().copy())
    finally:
        if prev is None:
            os.environ.pop("DFVO_HYP_COOP", None)
        else:
            os.environ["DFVO_HYP_COOP"] = prev
    a, b = res["0"], res["1"]
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.abs(a[0] - b[0]).max() < 1e-10 and np.abs(a[3] - b[3]).max() <= 1e-9 * np.abs(a[3]).max()
    return int(a[2][0, 1])
