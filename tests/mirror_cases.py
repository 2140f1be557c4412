"""Shared cases for the `libs.*` mirror additions of round 2 (geometry layers, opt_rigid_flow_kp, triangulation, DevArray
snapshots, capacity-allocated workspaces): run on the host-emulation build by the CPU suite and on the device by the GPU
suite.  The checker side is torch / cv2 / the oracle; the code under test is df-vo_b200/libs + b200 over the C ABI."""
import numpy as np

import synthdata


def geometry_inputs(h=47, w=83, seed=3):
    rs = np.random.RandomState(seed)
    K = synthdata.kitti_intrinsics(h, w)
    cx, cy, fx, fy = K
    Km = np.eye(4); Km[:3, :3] = [[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]]
    iKm = np.eye(4); iKm[:3, :3] = np.linalg.inv(Km[:3, :3])
    T = np.eye(4); T[:3, :3] = synthdata.rodrigues(np.array([0.01, -0.02, 0.005])); T[:3, 3] = [0.1, -0.05, 0.8]
    depth = rs.uniform(2.0, 40.0, (1, 1, h, w)).astype(np.float32)
    return depth, T, Km, iKm


def torch_layers(depth, T, Km, iKm, normalized):
    """The reference layers restated with torch on the CPU (backprojection.py:45-63, transformation3d.py:21-31,
    projection.py:31-52, layers.py:252-266), float32."""
    import torch
    _, _, h, w = depth.shape
    d = torch.from_numpy(depth)
    Kt, iKt, Tt = (torch.from_numpy(m).float().unsqueeze(0) for m in (Km, iKm, T))
    mesh = np.meshgrid(range(w), range(h), indexing="xy")
    idc = torch.tensor(np.stack(mesh, axis=0).astype(np.float32))
    ones = torch.ones(1, 1, h * w)
    xy = torch.cat([torch.unsqueeze(torch.stack([idc[0].view(-1), idc[1].view(-1)], 0), 0), ones], 1)
    pts = torch.matmul(iKt[:, :3, :3], xy)
    pts = d.view(1, 1, -1) * pts
    pts = torch.cat([pts, ones], 1)
    tr = torch.matmul(Tt, pts)
    p2 = torch.matmul(Kt[:, :3, :], tr)
    pix = p2[:, :2, :] / (p2[:, 2:3, :] + 1e-7)
    pix = pix.view(1, 2, h, w).permute(0, 2, 3, 1).clone()
    flow = pix.permute(0, 3, 1, 2) - idc.unsqueeze(0)
    if normalized:
        pix[..., 0] /= w - 1
        pix[..., 1] /= h - 1
        pix = (pix - 0.5) * 2
    return pts.numpy(), tr.numpy(), pix.numpy(), flow.numpy()


def check_geometry_layers(as_torch):
    """Backprojection / Transformation3D / Projection / Reprojection / RigidFlow mirrors vs the torch layers."""
    import torch
    from libs.geometry.backprojection import Backprojection
    from libs.geometry.projection import Projection
    from libs.geometry.reprojection import Reprojection
    from libs.geometry.rigid_flow import RigidFlow
    from libs.geometry.transformation3d import Transformation3D
    depth, T, Km, iKm = geometry_inputs()
    h, w = depth.shape[2:]
    conv = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()) if as_torch else (lambda a: a)
    back = lambda a: a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    d, Tt, Kt, iKt = conv(depth), conv(T[None]), conv(Km[None]), conv(iKm[None])
    for normalized in (True, False):
        pts, tr, pix, flow = torch_layers(depth, T, Km, iKm, normalized)
        p = Backprojection(h, w)(d, iKt)
        assert tuple(p.shape) == (1, 4, h * w)
        assert np.allclose(back(p), pts, rtol=2e-6, atol=1e-5)
        q = Transformation3D()(p, Tt)
        assert np.allclose(back(q), tr, rtol=2e-6, atol=2e-5)
        xy = Projection(h, w)(q, Kt, normalized)
        assert tuple(xy.shape) == (1, h, w, 2)
        tol = 2e-5 if normalized else 2e-3          # pixel units when not normalised
        assert np.abs(back(xy) - pix).max() < tol
        xy2 = Reprojection(h, w)(d, Tt, Kt, iKt, normalized)
        assert np.abs(back(xy2) - pix).max() < tol
        if not normalized:
            f = RigidFlow(h, w)(d, Tt, Kt, iKt, normalized=False)
            assert tuple(f.shape) == (1, 2, h, w)
            assert np.abs(back(f) - flow).max() < 2e-3
    assert Backprojection(h, w)(d, iKt, img_like_out=True).shape[2:] == (h, w)


def check_triangulation():
    """libs.geometry.ops_3d.triangulation vs cv2.triangulatePoints for identity and general first views."""
    import cv2
    from libs.geometry import ops_3d
    kp_ref, kp_cur, info = synthdata.correspondences(seed=5, n=300, outlier_frac=0.0)
    cx, cy, fx, fy = info["K"]
    k1 = (kp_ref - [cx, cy]) / [fx, fy]
    k2 = (kp_cur - [cx, cy]) / [fx, fy]
    T2 = np.eye(4); T2[:3, :3] = info["R"]; T2[:3, 3] = info["t"]
    G = np.eye(4); G[:3, :3] = synthdata.rodrigues(np.array([0.02, 0.01, -0.03])); G[:3, 3] = [0.3, -0.2, 0.1]
    for T1w, T2w in ((np.eye(4), T2), (G, T2 @ G)):
        X, X1, X2 = ops_3d.triangulation(k1, k2, T1w, T2w)
        Xc = cv2.triangulatePoints(T1w[:3], T2w[:3], np.ascontiguousarray(k1.T), np.ascontiguousarray(k2.T))
        Xc = Xc / Xc[3]
        assert X.shape == (3, 300) and X1.shape == (3, 300) and X2.shape == (3, 300)
        scale = np.abs(Xc[:3]).max(0) + 1.0
        assert (np.abs(X - Xc[:3]) / scale).max() < 1e-6
        assert (np.abs(X1 - T1w[:3] @ Xc) / scale).max() < 1e-6 and (np.abs(X2 - T2w[:3] @ Xc) / scale).max() < 1e-6


def check_opt_rigid_flow_kp(engine):
    """libs.matching.kp_selection.opt_rigid_flow_kp (the free function) == the tracker-internal path == the oracle."""
    import rigid_cases
    from b200 import config, tracking
    from oracle import vo
    from libs.matching import kp_selection
    tracking._default_engine = engine
    fr, depth_proc, kp1, kp2 = rigid_cases.frame("outliers")
    cfg = config.default_cfg(376, 1241)
    cfg.kp_selection.rigid_flow_kp.enable = True
    T = np.eye(4); T[:3, :3] = fr["R"]; T[:3, 3] = np.asarray(fr["t"]).reshape(3)
    rmap = vo.rigid_flow_diff(fr["depth"], fr["flow_fwd"], T, fr["K"])
    ref = {"rigid_flow_diff": rmap[..., None], "flow_diff": fr["flow_diff"], "flow": fr["flow_fwd"]}
    for method in ("opt_flow", "rigid_flow"):
        out = kp_selection.opt_rigid_flow_kp(None, None, ref, cfg, {}, method)
        best, uniform = vo.opt_rigid_flow_kp(rmap, fr["flow_diff"][:, :, 0], score_method=method)
        lin = lambda k: (k[0][:, 1] * 1241 + k[0][:, 0]).astype(np.int64)
        assert np.array_equal(np.sort(lin(out["kp1_depth"])), np.sort(np.concatenate(best))), method
        assert np.array_equal(lin(out["kp1_depth_uniform"]), np.concatenate(uniform)), method
        assert np.asarray(out["rigid_flow_mask"]).shape == (376, 1241)
        f = fr["flow_fwd"]
        k1 = out["kp1_depth"][0].astype(int)
        assert np.allclose(out["kp2_depth"][0], out["kp1_depth"][0] + np.stack([f[0, k1[:, 1], k1[:, 0]], f[1, k1[:, 1], k1[:, 0]]], 1))


def check_devarray_copy(rt):
    """DevArray.copy() is a snapshot: later writes into the engine's buffer do not reach it (dfvo.py:329-332)."""
    from b200 import tracking
    buf = rt.from_host(np.arange(12, dtype=np.float32).reshape(3, 4))
    a = tracking.DevArray(buf, (3, 4))
    snap = a.copy()
    buf.upload(np.zeros((3, 4), np.float32))
    assert np.array_equal(np.asarray(snap), np.arange(12, dtype=np.float32).reshape(3, 4))
    assert np.array_equal(np.asarray(a), np.zeros((3, 4), np.float32))
    b = a.copy()                                   # host copy exists now: snapshot of the host data
    assert np.array_equal(np.asarray(b), np.zeros((3, 4)))


def check_varying_keypoint_counts(engine):
    """The capacity-allocated RANSAC workspaces give the same result for a keypoint count whatever counts came before, and
    the caches do not grow with the number of distinct counts (per-n workspaces)."""
    from b200 import tracking
    K = synthdata.kitti_intrinsics()
    kp_ref, kp_cur, _ = synthdata.correspondences(seed=32, n=2600, outlier_frac=0.3)

    def run(n):
        np.random.seed(4869)
        return tracking.compute_pose_2d2d(engine, kp_ref[:n].copy(), kp_cur[:n].copy(), K, repeat=3)
    first = run(700)
    for n in (1999, 350, 1200):
        run(n)
    again = run(700)
    assert np.array_equal(first["inliers"], again["inliers"]) and np.array_equal(first["R"], again["R"]) and np.array_equal(first["t"], again["t"])
    assert len(engine._ess_ws) == 1 and len(engine._h_ws) == 1
    cap = next(iter(engine._ess_ws.values()))["cap"]
    assert cap >= 2000
    run(cap + 100)                                 # beyond the capacity: grows once, still one entry
    assert len(engine._ess_ws) == 1 and next(iter(engine._ess_ws.values()))["cap"] >= cap + 100
    engine._subsets_cap = 3                        # LRU bound of the subset tables
    for n in (100, 200, 300, 400, 500):
        run(n)
    assert len(engine._subsets) <= 3
