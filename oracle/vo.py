"""CPU restatement (numpy + the reference's own third-party solvers cv2 / sklearn) of the host
side of the tracking hot path: depth post-processing, correspondence selection, E-tracker, GRIC,
scale recovery, PnP tracker.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Each function cites the reference lines it restates.  Where the reference calls OpenCV /
scikit-learn (``opencv-python==3.4.3.18``, ``scikit-learn==0.20.3`` pinned in
``envs/requirement.yml:296,233``; 4.13.0 / 1.9.0 in this image) the oracle calls the same entry
point with the same arguments -- those libraries *are* the reference's algorithm for that step.
White-box replays of the OpenCV solvers live in ``oracle/cvreplay.py``.
"""
import math

import numpy as np


# ----------------------------------------------------------------------------------------
# depth post-processing  (dfvo.py:314-319, utils.py:89-114)
# ----------------------------------------------------------------------------------------
def resize_nearest(depth, W, H):
    """``cv2.resize(depth, (W, H), interpolation=cv2.INTER_NEAREST)`` (OpenCV resizeNN): src index =
    min(floor(dst * ifx), in - 1) with ``ifx = 1 / (out / in)`` in float64 (not bit-identical to in/out)."""
    h, w = depth.shape
    ys = np.minimum(np.floor(np.arange(H) * (1.0 / (H / h))).astype(np.int64), h - 1)
    xs = np.minimum(np.floor(np.arange(W) * (1.0 / (W / w))).astype(np.int64), w - 1)
    return depth[ys][:, xs]


def preprocess_depth(depth, crop, depth_range):
    """``utils.preprocess_depth`` (utils.py:89-114): zero outside the crop box and outside the
    (min, max) *exclusive* range; result float64 (mask multiply promotes)."""
    min_d, max_d = depth_range
    h, w = depth.shape
    y0, y1 = int(h * crop[0][0]), int(h * crop[0][1])
    x0, x1 = int(w * crop[1][0]), int(w * crop[1][1])
    mask = np.zeros((h, w))
    mask[y0:y1, x0:x1] = 1
    rng = (depth < max_d) * (depth > min_d)
    return depth * (mask * rng)


# ----------------------------------------------------------------------------------------
# correspondence selection  (kp_selection.py:33-200, keypoint_sampler.py:76-163)
# ----------------------------------------------------------------------------------------
def cell_bounds(h, w, rows, cols):
    """Cell slices of ``local_bestN`` (kp_selection.py:129-133): ``x1 = int(h/rows*(r+1)) - 1`` is
    an inclusive corner used as an exclusive slice end."""
    out = []
    for r in range(rows):
        for c in range(cols):
            y0, x0 = int(h / rows * r), int(w / cols * c)
            y1, x1 = int(h / rows * (r + 1)) - 1, int(w / cols * (c + 1)) - 1
            out.append((y0, y1, x0, x1))
    return out


def local_bestn_indices(flow_diff, rows=10, cols=10, N=2000, thre=0.1, depth_diff=None, depth_thre=0.05):
    """``local_bestN`` with score_method 'flow' (kp_selection.py:74-200) reduced to what it
    decides: returns ``(good, idx)`` where ``idx`` is a list (cell-major) of sorted linear pixel
    indices (y*w+x) per cell -- the *set* the reference's ``argpartition`` selects (its order inside
    the first k is implementation-defined).  Ties at the k-th value are resolved by
    the smaller linear index (the CUDA kernel uses the same rule)."""
    h, w = flow_diff.shape[:2]
    fd = flow_diff.reshape(h, w)
    if (fd < thre).sum() < N * 0.1:                                  # kp_selection.py:121-125
        return False, []
    n_best = math.floor(N / (rows * cols))
    good_regions = 0
    sel = []
    for (y0, y1, x0, x1) in cell_bounds(h, w, rows, cols):
        tile = fd[y0:y1, x0:x1]
        mask = tile < thre
        if depth_diff is not None:
            mask = mask & (depth_diff.reshape(h, w)[y0:y1, x0:x1] < depth_thre)
        ys, xs = np.where(mask)
        k = min(n_best, len(ys))
        if k != 0:
            good_regions += 1
        score = tile[ys, xs]
        lin = (ys + y0) * w + (xs + x0)
        order = np.lexsort((lin, score))[:k]                         # k smallest, ties -> smaller index
        sel.append(np.sort(lin[order]))
    if good_regions < (rows * cols) * 0.1:                           # kp_selection.py:175-179
        return False, []
    return True, sel


def bestn_indices(flow_diff, N=2000):
    """``bestN_flow_kp`` (kp_selection.py:33-71): the N smallest of the whole map (as a sorted set;
    ties -> smaller linear index).  NB the reference uses ``argpartition(.., N)[:N]``."""
    fd = flow_diff.reshape(-1)
    lin = np.arange(fd.size)
    return np.sort(np.lexsort((lin, fd))[:N])


def keypoints_from_indices(idx_list, flow_fwd, w):
    """``KeypointSampler.kp_selection`` (keypoint_sampler.py:101-104): kp1 = pixel grid (x,y) float64,
    kp2 = kp1 + forward flow (float32 promoted)."""
    lin = np.concatenate(idx_list) if len(idx_list) else np.zeros((0,), np.int64)
    ys, xs = lin // w, lin % w
    kp1 = np.stack([xs, ys], 1).astype(np.float64)
    kp2 = kp1 + np.stack([flow_fwd[0, ys, xs], flow_fwd[1, ys, xs]], 1).astype(np.float64)
    return kp1, kp2


# ----------------------------------------------------------------------------------------
# GRIC  (gric.py:14-132)
# ----------------------------------------------------------------------------------------
def fundamental_residual(F, kp1, kp2):
    """``compute_fundamental_residual`` (gric.py:14-37); the O(N^2) ``diagonal()`` is the per-point
    bilinear form ``m1_i^T F m0_i``."""
    m0 = np.ones((3, kp1.shape[0])); m0[:2] = kp1.T
    m1 = np.ones((3, kp2.shape[0])); m1[:2] = kp2.T
    Fm0 = F @ m0
    Ftm1 = F.T @ m1
    m1Fm0 = np.einsum("ij,ij->j", Fm0, m1)
    return m1Fm0 ** 2 / (np.sum(Fm0[:2] ** 2, axis=0) + np.sum(Ftm1[:2] ** 2, axis=0))


def homography_residual(H_in, kp1, kp2):
    """``compute_homography_residual`` (gric.py:40-91)."""
    H = H_in.flatten()
    m0 = np.ones((3, kp1.shape[0])); m0[:2] = kp1.T
    m1 = np.ones((3, kp2.shape[0])); m1[:2] = kp2.T
    G0 = np.stack([H[0] - m1[0] * H[6], H[1] - m1[0] * H[7], -m0[0] * H[6] - m0[1] * H[7] - H[8]])
    G1 = np.stack([H[3] - m1[1] * H[6], H[4] - m1[1] * H[7], -m0[0] * H[6] - m0[1] * H[7] - H[8]])
    magG0 = np.sqrt((G0 * G0).sum(0))
    magG1 = np.sqrt((G1 * G1).sum(0))
    magG0G1 = G0[0] * G1[0] + G0[1] * G1[1]
    alpha = np.arccos(magG0G1 / (magG0 * magG1))
    wgt = m0[0] * H[6] + m0[1] * H[7] + H[8]
    alg0 = m0[0] * H[0] + m0[1] * H[1] + H[2] - m1[0] * wgt
    alg1 = m0[0] * H[3] + m0[1] * H[4] + H[5] - m1[1] * wgt
    D1, D2 = alg0 / magG0, alg1 / magG1
    return (D1 * D1 + D2 * D2 - 2.0 * D1 * D2 * np.cos(alpha)) / np.sin(alpha)


def calc_gric(res, sigma, n, model):
    """``calc_GRIC`` (gric.py:94-132).  The reference accumulates in a Python loop (sequential
    float64 adds); ``math.fsum``-free sequential order is kept with ``np.cumsum``'s last element."""
    R = 4
    K = {"FMat": 7, "EMat": 5, "HMat": 8}[model]
    D = {"FMat": 3, "EMat": 3, "HMat": 2}[model]
    lam3RD = 2.0 * (R - D)
    tmp = res[:n] * (1.0 / sigma ** 2)
    terms = np.where(tmp <= lam3RD, tmp, lam3RD)
    s = 0.0
    for t in terms:                      # same left-to-right accumulation as the reference loop
        s += t
    return s + n * D * np.log(R) + K * np.log(R * n)


# ----------------------------------------------------------------------------------------
# E-tracker  (E_tracker.py:154-307), default config: validity.method == 'GRIC'
# ----------------------------------------------------------------------------------------
def compute_pose_2d2d(kp_ref, kp_cur, K, repeat=5, reproj_thre=0.2, rng=np.random, trace=None):
    """``EssTracker.compute_pose_2d2d(is_iterative=True)`` with ``validity.method: GRIC``.
    K = [cx, cy, fx, fy].  ``rng`` must expose ``shuffle`` (the reference uses the *global*
    ``np.random``).  Returns dict(R, t, inliers, valid, best_E, cheirality)."""
    import cv2
    cx, cy, fx, fy = K
    Kmat = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    pp = (cx, cy)
    R, t = np.eye(3), np.zeros((3, 1))
    best_inlier_cnt = 0
    best_inliers = np.ones((kp_ref.shape[0], 1)) == 1
    best_E = None
    out = {"R": R, "t": t, "valid": False, "best_E": None, "cheirality": 0}
    if kp_cur.shape[0] <= 10:                                        # E_tracker.py:196,217
        out["inliers"] = best_inliers[:, 0]
        return out
    H, H_inl = cv2.findHomography(kp_cur, kp_ref, method=cv2.RANSAC, confidence=0.99,
                                  ransacReprojThreshold=1)            # :199-205
    H_res = homography_residual(H, kp_cur, kp_ref)
    H_gric = calc_gric(H_res, 0.8, kp_cur.shape[0], "HMat")
    num_valid = 0
    for _ in range(repeat):                                          # :223-286
        order = np.arange(0, kp_cur.shape[0], 1)
        rng.shuffle(order)
        nc, nr = kp_cur.copy()[order], kp_ref.copy()[order]
        E, inl = cv2.findEssentialMat(nc, nr, focal=fx, pp=pp, method=cv2.RANSAC, prob=0.99,
                                      threshold=reproj_thre)
        F = np.linalg.inv(Kmat.T) @ E @ np.linalg.inv(Kmat)
        E_res = fundamental_residual(F, nc, nr)
        E_gric = calc_gric(E_res, 0.8, kp_cur.shape[0], "EMat")
        valid_case = H_gric > E_gric
        if trace is not None:
            trace.append(dict(order=order, E=E, inl=inl, E_gric=E_gric, H_gric=H_gric))
        if inl.sum() > best_inlier_cnt:
            best_E = E
            best_inlier_cnt = inl.sum()
            revert = np.zeros_like(order)
            revert[order] = np.arange(order.shape[0])
            best_inliers = inl[list(revert)]
        num_valid += int(valid_case)
    out["valid"] = num_valid > repeat / 2
    out["best_E"] = best_E
    if out["valid"]:
        cnt, R_, t_, _ = cv2.recoverPose(best_E, kp_cur, kp_ref, focal=fx, pp=pp)
        out["cheirality"] = cnt
        if cnt > kp_cur.shape[0] * 0.1:
            R, t = R_, t_
    out["R"], out["t"] = R, t
    out["inliers"] = best_inliers[:, 0] == 1
    out["H_gric"] = H_gric
    return out


# ----------------------------------------------------------------------------------------
# scale recovery  (E_tracker.py:476-507, 571-643; ops_3d.py:15-67)
# ----------------------------------------------------------------------------------------
def triangulate_x2(kp1n, kp2n, T_21):
    """``ops_3d.triangulation(kp1_norm, kp2_norm, eye(4), T_21)`` -> X2 [3,N] (ops_3d.py:44-67)."""
    import cv2
    X = cv2.triangulatePoints(np.eye(4)[:3], T_21[:3], np.ascontiguousarray(kp1n.T), np.ascontiguousarray(kp2n.T))
    X = X / X[3]
    return T_21[:3] @ X


def sparse_depth(kp, XYZ, height, width):
    """``convert_sparse3D_to_depth`` (ops_3d.py:15-41): truncation toward zero, last writer wins."""
    depth = np.zeros((height, width))
    kp_int = kp.astype(int)
    m1 = (kp_int[:, 0] >= 0) * (kp_int[:, 0] < width)
    kp_int = kp_int[m1]
    m2 = (kp_int[:, 1] >= 0) * (kp_int[:, 1] < height)
    kp_int = kp_int[m2]
    Z = XYZ[:, m1][:, m2]
    depth[kp_int[:, 1], kp_int[:, 0]] = Z[2]
    return depth


def depth_ratios(kp1, kp2, T_21, depth2, K):
    """First half of ``find_scale_from_depth`` (E_tracker.py:571-616): returns the vector
    ``depth_ratio`` in the (row-major mask) order the reference hands to the RANSAC regressor."""
    cx, cy, fx, fy = K
    kp1n, kp2n = kp1.copy(), kp2.copy()
    kp1n[:, 0] = (kp1[:, 0] - cx) / fx; kp1n[:, 1] = (kp1[:, 1] - cy) / fy
    kp2n[:, 0] = (kp2[:, 0] - cx) / fx; kp2n[:, 1] = (kp2[:, 1] - cy) / fy
    X2 = triangulate_x2(kp1n, kp2n, T_21)
    h, w = depth2.shape
    d_tri = sparse_depth(kp2, X2, h, w)
    d_tri[d_tri < 0] = 0
    valid = (depth2 > 0) * (d_tri > 0)
    return d_tri[valid] / depth2[valid], valid


def find_scale_from_depth(kp1, kp2, T_21, depth2, K, min_samples=3, max_trials=100, stop_prob=0.99, thre=0.1):
    """``EssTracker.find_scale_from_depth`` (E_tracker.py:571-643), ransac.method 'depth_ratio'.
    Consumes the global ``np.random`` stream exactly like the reference (``random_state=None``)."""
    from sklearn import linear_model
    ratio, valid = depth_ratios(kp1, kp2, T_21, depth2, K)
    if valid.sum() > 10:
        ransac = linear_model.RANSACRegressor(
            estimator=linear_model.LinearRegression(fit_intercept=False), min_samples=min_samples,
            max_trials=max_trials, stop_probability=stop_prob, residual_threshold=thre)
        ransac.fit(ratio.reshape(-1, 1), np.ones((ratio.shape[0], 1)))
        return ransac.estimator_.coef_[0, 0]
    return -1


# ----------------------------------------------------------------------------------------
# PnP tracker  (pnp_tracker.py:45-125, ops_3d.py:70-94)
# ----------------------------------------------------------------------------------------
def compute_pose_3d2d(kp1, kp2, depth_1, K, repeat=5, iters=100, reproj_thre=1, min_depth=0, max_depth=50,
                      rng=np.random):
    """``PnpTracker.compute_pose_3d2d(is_iterative=True)``; returns the 4x4 pose (cur->ref, i.e.
    the inverse of solvePnP's, pnp_tracker.py:113-118) and the filtered keypoints."""
    import cv2
    cx, cy, fx, fy = K
    Kmat = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    height, width = depth_1.shape
    m = (kp2[:, 0] >= 0) * (kp2[:, 0] < width)
    kp1, kp2 = kp1[m], kp2[m]
    m = (kp2[:, 1] >= 0) * (kp2[:, 1] < height)
    kp1, kp2 = kp1[m], kp2[m]
    ki = kp1.astype(int)
    d = depth_1[ki[:, 1], ki[:, 0]]
    m = (d != 0) * ((d < max_depth) * (d > min_depth))
    kp1, kp2, d = kp1[m], kp2[m], d[m]
    XYZ = np.stack([(kp1[:, 0] - cx) / fx, (kp1[:, 1] - cy) / fy, np.ones(len(d))], 1)
    invK = np.linalg.inv(Kmat)
    XYZ = (invK @ np.concatenate([kp1, np.ones((len(d), 1))], 1).T).T       # ops_3d.py:83-89
    XYZ = XYZ * d[:, None]
    best_rt, best_inl = [], 0
    for _ in range(repeat):
        order = np.arange(0, kp2.shape[0], 1)
        rng.shuffle(order)
        nX, n2 = XYZ.copy()[order], kp2.copy()[order]
        if n2.shape[0] > 4:
            flag, r, t, inl = cv2.solvePnPRansac(objectPoints=nX, imagePoints=n2, cameraMatrix=Kmat, distCoeffs=None,
                                                 iterationsCount=iters, reprojectionError=reproj_thre)
            if flag and inl.shape[0] > best_inl:
                best_rt, best_inl = [r, t], inl.shape[0]
    pose = np.eye(4)
    if len(best_rt) != 0:
        pose[:3, :3] = cv2.Rodrigues(best_rt[0])[0]
        pose[:3, 3:] = best_rt[1]
    return np.linalg.inv(pose), kp1, kp2, best_inl


# ----------------------------------------------------------------------------------------
# rigid-flow keypoints + iterative scale recovery  
# ----------------------------------------------------------------------------------------
def rigid_flow_diff(raw_depth, flow, T, K):
    """``EssTracker.kp_selection_good_depth`` up to ``rigid_flow_diff`` (E_tracker.py:666-691): the RigidFlow layer
    (rigid_flow.py:38-60 = Backprojection backprojection.py:45-63, Transformation3D, Projection projection.py:31-52 with
    normalized=False, PixToFlow layers.py:252-266) in the same torch float32 operations, then
    ``np.linalg.norm(rigid_flow - flow, axis=0)``.  raw_depth [h,w] f32, flow [2,h,w] f32, T 4x4 (float64, cast to
    float32 like ``torch.from_numpy(pose).float()``), K = [cx, cy, fx, fy].  Returns float32 [h,w]."""
    import torch
    h, w = raw_depth.shape
    cx, cy, fx, fy = K
    Km = np.eye(4); Km[:3, :3] = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    iKm = np.eye(4); iKm[:3, :3] = np.linalg.inv(Km[:3, :3])
    Kt = torch.from_numpy(Km).float().unsqueeze(0)
    iKt = torch.from_numpy(iKm).float().unsqueeze(0)
    Tt = torch.from_numpy(np.asarray(T, np.float64)).float().unsqueeze(0)
    depth = torch.from_numpy(np.ascontiguousarray(raw_depth)).float().unsqueeze(0).unsqueeze(0)
    mesh = np.meshgrid(range(w), range(h), indexing="xy")
    idc = torch.tensor(np.stack(mesh, axis=0).astype(np.float32))
    ones = torch.ones(1, 1, h * w)
    xy = torch.cat([torch.unsqueeze(torch.stack([idc[0].view(-1), idc[1].view(-1)], 0), 0), ones], 1)
    points = torch.matmul(iKt[:, :3, :3], xy)
    points = depth.view(1, 1, -1) * points
    points = torch.cat([points, ones], 1)
    points = torch.matmul(Tt, points)
    p2 = torch.matmul(Kt[:, :3, :], points)
    pix = p2[:, :2, :] / (p2[:, 2:3, :] + 1e-7)
    pix = pix.view(1, 2, h, w).permute(0, 2, 3, 1)
    rigid = (pix.permute(0, 3, 1, 2) - idc.unsqueeze(0)).numpy()[0]
    return np.linalg.norm(rigid - flow, axis=0)


def opt_rigid_flow_kp(rigid_diff, flow_diff, rows=10, cols=10, N=2000, rigid_thre=5, flow_thre=0.1, score_method="opt_flow"):
    """``opt_rigid_flow_kp`` (kp_selection.py:203-324) reduced to what it decides: per cell the 'best' set (sorted linear
    indices; ties at the k-th score by smaller index, like :func:`local_bestn_indices`) and the 'uniform' list (row-major
    order of ``np.where(valid_mask)``, every ``step``-th, kp_selection.py:277-282).  Returns (best, uniform): lists of
    int64 arrays of linear pixel indices per cell."""
    h, w = rigid_diff.shape
    n_best = math.floor(N / (rows * cols))
    best, uniform = [], []
    for (y0, y1, x0, x1) in cell_bounds(h, w, rows, cols):
        rd, fd = rigid_diff[y0:y1, x0:x1], flow_diff[y0:y1, x0:x1]
        valid = (rd < rigid_thre) & (fd < flow_thre)
        ys, xs = np.where(valid)
        lin = (ys + y0) * w + (xs + x0)
        n = len(lin)
        k = min(n_best, n)
        if k > 0:
            step = int(n / k)
            uniform.append(lin[np.arange(0, n, step)[:k]].astype(np.int64))
        else:
            uniform.append(np.zeros(0, np.int64))
        score = (rd if score_method == "rigid_flow" else fd)[valid]
        order = np.lexsort((lin, score))[:k]
        best.append(np.sort(lin[order]).astype(np.int64))
    return best, uniform


def scale_recovery_iterative(kp_ref_best, kp_cur_best, E_pose, depth2, raw_depth_ref, flow, flow_diff, K, prev_scale,
                             kp_src="kp_best", score_method="rigid_flow", ransac=None):
    """``EssTracker.scale_recovery_iterative`` (E_tracker.py:509-569): up to five rounds of [pose scaled by the current
    scale -> inverse -> rigid-flow keypoint selection (kp_selection_good_depth, :645-705) -> find_scale_from_depth on
    ``cfg.scale_recovery.kp_src`` keypoints]; stops when the scale moves by < 0.001.  E_pose: 4x4 (cur -> ref).
    Returns dict(scale, kp1_uniform, kp2_uniform, rigid_flow_diff, rigid_flow_pose, rounds)."""
    h, w = raw_depth_ref.shape
    fd = np.asarray(flow_diff).reshape(h, w)
    scale, delta = prev_scale, 0.001
    out = {}
    for it in range(5):
        P = np.array(E_pose, np.float64)
        P[:3, 3] = P[:3, 3] * scale
        T = np.linalg.inv(P)                                       # SE3(rigid_flow_pose.inv_pose)
        rd = rigid_flow_diff(raw_depth_ref, flow, T, K)
        best, uniform = opt_rigid_flow_kp(rd, fd, score_method=score_method)
        kp1u, kp2u = keypoints_from_indices(uniform, flow, w)
        ref_kp, cur_kp = (kp1u, kp2u) if kp_src == "kp_depth" else (kp_ref_best, kp_cur_best)
        new_scale = find_scale_from_depth(ref_kp, cur_kp, np.linalg.inv(np.array(E_pose, np.float64)), depth2, K, **(ransac or {}))
        d = abs(new_scale - scale)
        scale = new_scale
        out = dict(scale=scale, kp1_uniform=kp1u, kp2_uniform=kp2u, rigid_flow_diff=rd, rigid_flow_pose=T, rounds=it + 1,
                   best=best, uniform=uniform)
        if d < delta:
            break
    return out
