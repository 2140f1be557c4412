"""Small host-side numerics of the tracking path that are inherently sequential or consume the host
RNG stream and therefore stay on the CPU, written from scratch in NumPy:

* GRIC model-selection scores (libs/tracker/gric.py:14-132) -- O(N) instead of the reference's O(N^2)
  ``diagonal()`` and Python loop;
* the scale RANSAC of ``EssTracker.find_scale_from_depth`` (E_tracker.py:618-641): a re-implementation of
  ``sklearn.linear_model.RANSACRegressor(LinearRegression(fit_intercept=False))`` for the 1-parameter
  model ``y = s * x`` that draws from ``np.random`` exactly like scikit-learn does, so the global RNG
  stream stays aligned with the reference (the reference's ~27 ms per frame drops to ~1 ms);
* image_grid / preprocess_depth equivalents used by the device pipeline's host mirror.
"""
import math

import numpy as np


def rodrigues(rvec):
    """cv2.Rodrigues(rvec)[0]: rotation vector -> matrix."""
    r = np.asarray(rvec, np.float64).reshape(3)
    th = float(np.linalg.norm(r))
    if th < 2.220446049250313e-16:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return math.cos(th) * np.eye(3) + (1 - math.cos(th)) * np.outer(k, k) + math.sin(th) * Kx


# ---------------------------------------------------------------------------------------------
# GRIC (gric.py)
# ---------------------------------------------------------------------------------------------
def fundamental_residual(F, kp1, kp2):
    """gric.py:14-37 with the per-point bilinear form instead of the N x N product."""
    x1, y1, x2, y2 = kp1[:, 0], kp1[:, 1], kp2[:, 0], kp2[:, 1]
    f0 = F[0, 0] * x1 + F[0, 1] * y1 + F[0, 2]
    f1 = F[1, 0] * x1 + F[1, 1] * y1 + F[1, 2]
    f2 = F[2, 0] * x1 + F[2, 1] * y1 + F[2, 2]
    g0 = F[0, 0] * x2 + F[1, 0] * y2 + F[2, 0]
    g1 = F[0, 1] * x2 + F[1, 1] * y2 + F[2, 1]
    m = x2 * f0 + y2 * f1 + f2
    return m * m / (f0 * f0 + f1 * f1 + g0 * g0 + g1 * g1)


def homography_residual(H_in, kp1, kp2):
    """gric.py:40-91."""
    H = np.asarray(H_in, np.float64).reshape(-1)
    x0, y0, x1, y1 = kp1[:, 0], kp1[:, 1], kp2[:, 0], kp2[:, 1]
    w = x0 * H[6] + y0 * H[7] + H[8]
    G0 = (H[0] - x1 * H[6], H[1] - x1 * H[7], -w)
    G1 = (H[3] - y1 * H[6], H[4] - y1 * H[7], -w)
    magG0 = np.sqrt(G0[0] * G0[0] + G0[1] * G0[1] + G0[2] * G0[2])
    magG1 = np.sqrt(G1[0] * G1[0] + G1[1] * G1[1] + G1[2] * G1[2])
    alpha = np.arccos((G0[0] * G1[0] + G0[1] * G1[1]) / (magG0 * magG1))
    alg0 = x0 * H[0] + y0 * H[1] + H[2] - x1 * w
    alg1 = x0 * H[3] + y0 * H[4] + H[5] - y1 * w
    D1, D2 = alg0 / magG0, alg1 / magG1
    return (D1 * D1 + D2 * D2 - 2.0 * D1 * D2 * np.cos(alpha)) / np.sin(alpha)


def calc_gric(res, sigma, n, model):
    """gric.py:94-132: sum_i min(res_i / sigma^2, 2(R-D)) + n D log R + K log(R n), R = 4."""
    R = 4
    K = {"FMat": 7, "EMat": 5, "HMat": 8}[model]
    D = {"FMat": 3, "EMat": 3, "HMat": 2}[model]
    lam = 2.0 * (R - D)
    t = np.asarray(res[:n], np.float64) * (1.0 / sigma ** 2)
    return float(np.minimum(t, lam).sum()) + n * D * math.log(R) + K * math.log(R * n)


# ---------------------------------------------------------------------------------------------
# scale RANSAC (E_tracker.py:618-641  ==  sklearn RANSACRegressor on y = s*x through the origin)
# ---------------------------------------------------------------------------------------------
def _sample_without_replacement(n_population, n_samples, rng):
    """sklearn.utils.random.sample_without_replacement(method='auto') as implemented by the scikit-learn
    in this image (1.9.0, sklearn/utils/_random.pyx:235-255): ``rng.permutation(n)[:k]`` when
    0.01 < k/n < 0.99, tracking selection when k/n <= 0.01, reservoir sampling when k/n >= 0.99 -- with
    the same draws from ``rng``.  (The reference's pinned 0.20.3 used tracking / reservoir / pool; the RNG
    stream therefore depends on the installed scikit-learn, exactly as it does for the reference.)"""
    ratio = n_samples / n_population if n_population else 1.0
    if 0.01 < ratio < 0.99:
        return rng.permutation(n_population)[:n_samples]
    out = np.empty(n_samples, dtype=np.int64)
    if ratio < 0.2:
        selected = set()
        for i in range(n_samples):
            j = int(rng.randint(n_population))
            while j in selected:
                j = int(rng.randint(n_population))
            selected.add(j)
            out[i] = j
    else:
        out[:] = np.arange(n_samples)
        for i in range(n_samples, n_population):
            j = int(rng.randint(0, i + 1))
            if j < n_samples:
                out[j] = i
    return out


def _dynamic_max_trials(n_inliers, n_samples, min_samples, probability):
    eps = np.spacing(1)
    inlier_ratio = n_inliers / float(n_samples)
    nom = max(eps, 1 - probability)
    denom = max(eps, 1 - inlier_ratio ** min_samples)
    if nom == 1:
        return 0
    if denom == 1:
        return float("inf")
    return abs(float(np.ceil(np.log(nom) / np.log(denom))))


def _fit_through_origin(x, y):
    """LinearRegression(fit_intercept=False) on one feature: minimum-norm least squares."""
    den = float(np.dot(x, x))
    return float(np.dot(x, y)) / den if den != 0.0 else 0.0


def ransac_scale(x, min_samples=3, max_trials=100, stop_probability=0.99, residual_threshold=0.1, rng=np.random):
    """Host statement of the scale fit that ``tracking.Engine.ransac_scale`` runs on the device (csrc/ransac.cu::k_scale_ransac);
    nothing on the product path calls it any more -- the tests use it as the readable twin of the kernel.
    Fit ``1 ~= s * x`` like ``RANSACRegressor(...).fit(x[:,None], ones)`` (E_tracker.py:626-636) and
    return ``estimator_.coef_[0,0]``.  Same trial loop, same acceptance rule (more inliers, or equal
    inliers and not-worse R^2 score), same dynamic max_trials, same final refit on the best inlier set."""
    x = np.asarray(x, np.float64).reshape(-1)
    n = x.shape[0]
    y = np.ones(n)
    n_inliers_best, score_best = 1, -np.inf
    inlier_best = None
    n_trials = 0
    while n_trials < max_trials:
        n_trials += 1
        idx = _sample_without_replacement(n, min_samples, rng)
        s = _fit_through_origin(x[idx], y[idx])
        inl = np.abs(y - s * x) <= residual_threshold
        n_inl = int(inl.sum())
        if n_inl < n_inliers_best:
            continue
        # estimator.score on the inlier subset = r2_score with constant y_true: 1.0 if exact else 0.0
        resid = y[inl] - s * x[inl]
        score = 1.0 if float(np.dot(resid, resid)) == 0.0 else 0.0
        if n_inl == n_inliers_best and score < score_best:
            continue
        n_inliers_best, score_best, inlier_best = n_inl, score, inl
        max_trials = min(max_trials, _dynamic_max_trials(n_inliers_best, n, min_samples, stop_probability))
        if n_inliers_best >= np.inf or score_best >= np.inf:
            break
    if inlier_best is None:
        raise ValueError("RANSAC could not find a valid consensus set")
    return _fit_through_origin(x[inlier_best], y[inlier_best])


# ---------------------------------------------------------------------------------------------
# misc
# ---------------------------------------------------------------------------------------------
def last_writer_depth_ratio_sparse(kp2, z_tri, depth_at_kp, h, w):
    """Same as :func:`last_writer_depth_ratio` when only the CNN depth at the keypoints' pixels is on the
    host (``depth_at_kp[i] = depth2[int(kp2_y), int(kp2_x)]``, gathered on the device)."""
    ki = kp2.astype(int)
    ok = (ki[:, 0] >= 0) & (ki[:, 0] < w) & (ki[:, 1] >= 0) & (ki[:, 1] < h)
    lin = ki[ok, 1].astype(np.int64) * w + ki[ok, 0]
    z = np.asarray(z_tri, np.float64)[ok]
    d = np.asarray(depth_at_kp, np.float64)[ok]
    uniq, first = np.unique(lin[::-1], return_index=True)
    zt = z[::-1][first]
    zt = np.where(zt < 0, 0.0, zt)
    dp = d[::-1][first]
    valid = (dp > 0) & (zt > 0)
    return zt[valid] / dp[valid], int(valid.sum())


def last_writer_depth_ratio(kp2, z_tri, depth2):
    """The part of find_scale_from_depth between triangulation and RANSAC (E_tracker.py:598-616,
    ops_3d.py:15-41) without materialising the dense map: keypoints truncate toward zero to pixels,
    out-of-image ones are dropped, duplicates are last-writer-wins, negative depths become zero; the ratio
    vector is ordered by row-major pixel index like ``depth_tri[valid_mask]``."""
    h, w = depth2.shape
    ki = kp2.astype(int)
    ok = (ki[:, 0] >= 0) & (ki[:, 0] < w) & (ki[:, 1] >= 0) & (ki[:, 1] < h)
    lin = ki[ok, 1].astype(np.int64) * w + ki[ok, 0]
    z = np.asarray(z_tri, np.float64)[ok]
    # last occurrence of each pixel wins; np.unique on the reversed array returns first occurrences there
    rl = lin[::-1]
    uniq, first = np.unique(rl, return_index=True)
    zt = z[::-1][first]
    zt = np.where(zt < 0, 0.0, zt)
    dp = depth2.reshape(-1)[uniq]
    valid = (dp > 0) & (zt > 0)
    return zt[valid] / dp[valid], int(valid.sum())
