"""GPU parity: monodepth2 (DeepModel.forward_depth, deep_models.py:184-206) against the golden produced
by the reference facade, depth post-processing (dfvo.py:314-319) against cv2 / the oracle, and the pose
solvers (E_tracker.py:223-296) against cv2 goldens and the oracle replay (oracle/cvreplay.py)."""
import os

import numpy as np
import pytest
import torch

from b200 import native
from oracle import cvreplay, synth, vo
from util import dptr

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_monodepth2_vs_reference_golden(dev_lib):
    g = np.load(os.path.join(G, "deep_models_70x150.npz"))
    fh, fw = [int(x) for x in g["feed_hw"]]
    enc, dec = synth.monodepth2_weights(4869, fh, fw)
    feed = cu(g["depth_feed"][None])
    for prec, tol in ((native.PREC_FP32, 2e-5), (native.PREC_BF16, 4e-2)):
        ctx = native.Context(dev_lib)
        ctx.load_weights(native.NET_MONODEPTH2, enc)
        ctx.load_weights(native.NET_MONODEPTH2, dec)
        ctx.monodepth2_build(fh, fw, prec)
        out = torch.zeros((fh, fw), dtype=torch.float32, device="cuda")
        ctx.monodepth2_forward(dptr(feed), dptr(out))
        torch.cuda.synchronize()
        rel = np.abs(out.cpu().numpy() - g["depth"]) / g["depth"]
        assert rel.max() < tol, (prec, rel.max())
        ctx.close()


def test_monodepth2_full_feed_runs(dev_lib):
    """192x640 (the model-zoo feed size): finite, positive, fp32 vs bf16 within a few percent."""
    fh, fw = 192, 640
    enc, dec = synth.monodepth2_weights(4869, fh, fw)
    rs = np.random.RandomState(2)
    feed = cu(rs.uniform(0, 1, (1, 3, fh, fw)).astype(np.float32))
    outs = []
    for prec in (native.PREC_FP32, native.PREC_BF16):
        ctx = native.Context(dev_lib)
        ctx.load_weights(native.NET_MONODEPTH2, enc); ctx.load_weights(native.NET_MONODEPTH2, dec)
        ctx.monodepth2_build(fh, fw, prec)
        out = torch.zeros((fh, fw), dtype=torch.float32, device="cuda")
        ctx.monodepth2_forward(dptr(feed), dptr(out))
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
        ctx.close()
    assert np.isfinite(outs[0]).all() and (outs[0] > 0).all()
    assert (np.abs(outs[0] - outs[1]) / outs[0]).mean() < 0.03


def test_depth_post(dev_lib):
    import cv2
    rs = np.random.RandomState(4)
    d = rs.uniform(0.5, 80, (192, 640)).astype(np.float32)
    H, W = 376, 1241
    raw = torch.zeros((H, W), dtype=torch.float32, device="cuda")
    out = torch.zeros_like(raw)
    dd = cu(d)
    dev_lib.check(dev_lib.dfvo_depth_post(dptr(dd), 192, 640, H, W, 0.3, 1.0, 0.0, 1.0, 0.0, 50.0, dptr(raw), dptr(out), None))
    torch.cuda.synchronize()
    r2 = cv2.resize(d, (W, H), interpolation=cv2.INTER_NEAREST)
    assert np.array_equal(raw.cpu().numpy(), r2)
    assert np.array_equal(out.cpu().numpy().astype(np.float64), vo.preprocess_depth(r2, [[0.3, 1], [0, 1]], [0, 50]))


def test_five_point_vs_cv2(dev_lib):
    g = np.load(os.path.join(G, "cv_solvers.npz"))
    cx, cy, fx, fy = synth.kitti_intrinsics()
    kp_ref, kp_cur, _ = synth.correspondences(seed=44, n=400, outlier_frac=0.0)
    x1 = (kp_cur - np.array([cx, cy])) / fx
    x2 = (kp_ref - np.array([cx, cy])) / fx
    subs, sols = g["five_point_subsets"], g["five_point_solutions"]
    M = subs.shape[0]
    E = torch.zeros((M, 10, 9), dtype=torch.float64, device="cuda")
    n = torch.zeros(M, dtype=torch.int32, device="cuda")
    d1, d2 = cu(x1[subs]), cu(x2[subs])            # keep the device tensors alive across the call
    dev_lib.check(dev_lib.dfvo_five_point(dptr(d1), dptr(d2), M, dptr(E), dptr(n), None))
    torch.cuda.synchronize()
    E, n = E.cpu().numpy(), n.cpu().numpy()
    dists, count_mismatch = [], 0
    for i in range(M):
        ref = sols[i][~np.isnan(sols[i]).any(1)].reshape(-1, 9)
        count_mismatch += int(len(ref) != n[i])       # near-double roots may be classified differently
        if n[i] == 0:
            continue
        for r in ref:
            dists.append(min(min(np.abs(m - r).max(), np.abs(m + r).max()) for m in E[i, :n[i]]))
    dists = np.sort(dists)
    assert count_mismatch <= 2
    assert np.median(dists) < 1e-10 and dists[int(0.95 * len(dists))] < 1e-6


@pytest.mark.parametrize("name,kw", [("out00", dict(seed=41, outlier_frac=0.0)), ("out30", dict(seed=42, outlier_frac=0.3)),
                                     ("out60", dict(seed=43, outlier_frac=0.6))])
def test_essential_ransac_and_recover_pose_vs_cv2(dev_lib, name, kw):
    g = np.load(os.path.join(G, "cv_solvers.npz"))
    cx, cy, fx, fy = synth.kitti_intrinsics()
    N, MI = 2000, 1000
    kp_ref, kp_cur, _ = synth.correspondences(n=N, **kw)
    subsets = cu(cvreplay.subset_stream(N, 5, MI))
    p1, p2 = cu(kp_cur), cu(kp_ref)
    ws = torch.zeros(dev_lib.dfvo_essential_workspace_bytes(N, 1, MI), dtype=torch.uint8, device="cuda")
    E = torch.zeros((1, 9), dtype=torch.float64, device="cuda")
    mask = torch.zeros((1, N), dtype=torch.uint8, device="cuda")
    info = torch.zeros((1, 4), dtype=torch.int32, device="cuda")
    gric = torch.zeros(1, dtype=torch.float64, device="cuda")
    dev_lib.check(dev_lib.dfvo_essential_ransac(dptr(p1), dptr(p2), N, None, 1, dptr(subsets), MI, fx, fy, cx, cy, 0.2, 0.99,
                                                dptr(ws), ws.numel(), dptr(E), dptr(mask), dptr(info), dptr(gric), None))
    torch.cuda.synchronize()
    Eref = g[name + "_E"]
    Eg = E.cpu().numpy()[0].reshape(3, 3)
    assert min(np.abs(Eg - Eref).max(), np.abs(Eg + Eref).max()) < 1e-9
    assert np.array_equal(mask.cpu().numpy()[0], g[name + "_mask"])          # bit-exact inlier mask vs cv2
    # GRIC-E of the winner vs the oracle's restatement of gric.py
    Kmat = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    F = np.linalg.inv(Kmat.T) @ Eref @ np.linalg.inv(Kmat)
    want = vo.calc_gric(vo.fundamental_residual(F, kp_cur, kp_ref), 0.8, N, "EMat")
    assert abs(gric.item() - want) < 1e-6 * abs(want)
    Rt = torch.zeros(12, dtype=torch.float64, device="cuda")
    pm = torch.zeros(N, dtype=torch.uint8, device="cuda")
    pi = torch.zeros(5, dtype=torch.int32, device="cuda")
    dEref = cu(Eref)
    dev_lib.check(dev_lib.dfvo_recover_pose(dptr(dEref), dptr(p1), dptr(p2), N, fx, cx, cy, dptr(Rt), dptr(pm), dptr(pi), None))
    torch.cuda.synchronize()
    Rt = Rt.cpu().numpy()
    assert int(pi[0].item()) == int(g[name + "_cheir"])
    assert np.abs(Rt[:9].reshape(3, 3) - g[name + "_R"]).max() < 1e-12 and np.abs(Rt[9:] - g[name + "_t"][:, 0]).max() < 1e-12
    assert np.array_equal(pm.cpu().numpy(), (g[name + "_pmask"] > 0).astype(np.uint8))


def test_score_hypotheses_config4(dev_lib):
    """BASELINE config #4 shape (10k hypotheses x 2048 correspondences) against the oracle formula."""
    rs = np.random.RandomState(6)
    kp_ref, kp_cur, _ = synth.correspondences(seed=45, n=2048, outlier_frac=0.3)
    cx, cy, fx, fy = synth.kitti_intrinsics()
    x1 = (kp_cur - np.array([cx, cy])) / fx
    x2 = (kp_ref - np.array([cx, cy])) / fx
    M = 10000
    E = rs.standard_normal((M, 9))
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    thr2 = (0.2 / fx) ** 2 * 400          # looser so random models have non-trivial counts
    counts = torch.zeros(M, dtype=torch.int32, device="cuda")
    dE, d1, d2 = cu(E), cu(x1), cu(x2)             # keep the device tensors alive across the call
    dev_lib.check(dev_lib.dfvo_score_hypotheses(dptr(dE), M, dptr(d1), dptr(d2), 2048, thr2, dptr(counts), None))
    torch.cuda.synchronize()
    got = counts.cpu().numpy()
    idx = rs.choice(M, 256, replace=False)
    for i in idx:
        want = int((cvreplay.sampson_errors(E[i].reshape(3, 3), x1, x2) <= thr2).sum())
        assert int(got[i]) == want


def test_lanczos_feed_bit_exact_with_pil(dev_lib):
    import PIL.Image as pil
    from b200 import lanczos
    H, W, oh, ow = 376, 1241, 192, 640
    img = synth.value_noise_image(H, W, 9)
    ref = np.asarray(pil.fromarray(img).resize((ow, oh), pil.LANCZOS))
    bh, kh, ksh = lanczos.coeffs(W, ow)
    bv, kv, ksv = lanczos.coeffs(H, oh)
    d = [cu(a) for a in (img, bh, kh, bv, kv)]
    tmp = torch.zeros((H, ow, 3), dtype=torch.uint8, device="cuda")
    out = torch.zeros((oh, ow, 3), dtype=torch.uint8, device="cuda")
    f = torch.zeros((3, oh, ow), dtype=torch.float32, device="cuda")
    dev_lib.check(dev_lib.dfvo_lanczos_resize_u8(dptr(d[0]), H, W, dptr(d[1]), dptr(d[2]), ksh, dptr(d[3]), dptr(d[4]), ksv, oh, ow,
                                                 dptr(tmp), dptr(out), dptr(f), None))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(f.cpu().numpy(), np.transpose(ref, (2, 0, 1)).astype(np.float32) / np.float32(255))


def _gpu_engine():
    from b200 import runtime as rt_mod, tracking
    rt = rt_mod.CudaRuntime(0)
    rt_mod.set_runtime(rt)
    return tracking.Engine(376, 1241, rt)


def test_pnp_ransac_vs_cv2(dev_lib):
    """csrc/pnp.cu on the device against cv2.solvePnPRansac (pnp_tracker.py:98): see tests/pnp_cases.py for the bar."""
    import pnp_cases
    exact, total = pnp_cases.check_vs_cv2(_gpu_engine())
    assert exact >= 0.8 * total


def test_epnp_cooperative_vs_sequential_vs_cv2(dev_lib):
    import pnp_cases
    pnp_cases.check_epnp_minimal(_gpu_engine(), samples=64)


def test_scale_ransac_on_device_vs_sklearn(dev_lib):
    import pnp_cases
    pnp_cases.check_scale_ransac_vs_sklearn(_gpu_engine(), cases=60)


def test_fused_tracker_tail_vs_stepwise(dev_lib):
    import pnp_cases
    pnp_cases.check_fused_tail_vs_stepwise(_gpu_engine())


def test_pnp_tracker_vs_reference_golden(dev_lib):
    import pnp_cases
    ang, dt = pnp_cases.check_vs_reference_golden(_gpu_engine(), np.load(os.path.join(G, "trackers_2000.npz")))
    assert ang < 1e-4 and dt < 1e-3


def test_homography_ransac_gric_vs_cv2(dev_lib):
    """csrc/homog.cu on the device against cv2.findHomography + GRIC-H (E_tracker.py:199-215)."""
    import pnp_cases
    assert pnp_cases.check_homography_vs_cv2(_gpu_engine()) < 1e-9
