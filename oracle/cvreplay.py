"""White-box replays of the OpenCV calib3d solvers the reference calls, validated
black-box against ``cv2`` (4.13.0 here; the reference pins opencv-python 3.4.3.18 whose source is not
vendored).  TEST INFRASTRUCTURE ONLY.

They exist so the CUDA RANSAC kernels can be checked stage by stage (subset stream, per-hypothesis
inlier counts, sequential acceptance rule, adaptive iteration count) and not only on the final pose.
"""
import math

import numpy as np

CV_RNG_COEFF = 4164903690


class CvRNG:
    """``cv::RNG`` multiply-with-carry generator; ``RNG((uint64)-1)`` is what
    ``RANSACPointSetRegistrator::run`` constructs on every call (ptsetreg.cpp)."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * CV_RNG_COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else a + self.next() % (b - a)


def subset_stream(count, model_points, n_subsets, rng=None):
    """``getSubset``: ``model_points`` distinct indices per subset; a draw equal to an earlier index of
    the same subset is redrawn.  Returns int32 [n_subsets, model_points]."""
    rng = rng or CvRNG()
    out = np.zeros((n_subsets, model_points), np.int32)
    for s in range(n_subsets):
        i = 0
        while i < model_points:
            v = rng.uniform(0, count)
            if v in out[s, :i]:
                continue
            out[s, i] = v
            i += 1
    return out


def ransac_update_num_iters(p, ep, model_points, max_iters):
    """``RANSACUpdateNumIters`` (ptsetreg.cpp)."""
    p = max(p, 0.0); p = min(p, 1.0)
    ep = max(ep, 0.0); ep = min(ep, 1.0)
    num = max(1.0 - p, np.finfo(np.float64).tiny)        # DBL_MIN
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < np.finfo(np.float64).tiny:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else int(round(num / denom))


def sampson_errors(E, x1, x2):
    """``EMEstimatorCallback::computeError`` (five-point.cpp): x1, x2 normalised [N,2]."""
    X1 = np.concatenate([x1, np.ones((x1.shape[0], 1))], 1)
    X2 = np.concatenate([x2, np.ones((x2.shape[0], 1))], 1)
    Ex1 = X1 @ E.T
    Etx2 = X2 @ E
    x2tEx1 = (X2 * Ex1).sum(1)
    return x2tEx1 ** 2 / (Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2)


def five_point_cv(x1s, x2s):
    """OpenCV's own minimal solver: ``findEssentialMat`` with exactly 5 points returns every real
    solution stacked as [3k,3] in the order RANSAC iterates them."""
    import cv2
    E, _ = cv2.findEssentialMat(x1s, x2s, focal=1.0, pp=(0.0, 0.0), method=cv2.RANSAC, prob=0.99, threshold=1.0)
    if E is None:
        return np.zeros((0, 3, 3))
    return E.reshape(-1, 3, 3)


def find_essential_mat_replay(p1, p2, focal, pp, prob=0.99, threshold=1.0, max_iters=1000, solver=five_point_cv,
                              trace=None):
    """Replay of ``cv2.findEssentialMat(p1, p2, focal, pp, RANSAC, prob, threshold)``.
    Returns (E [3,3], mask uint8 [N], iterations).  ``trace`` (list) receives one dict per iteration:
    subset, counts per candidate, whether the best changed."""
    N = p1.shape[0]
    x1 = (np.asarray(p1, np.float64) - np.array(pp)) / focal
    x2 = (np.asarray(p2, np.float64) - np.array(pp)) / focal
    thr = threshold / focal
    thr2 = thr * thr
    rng = CvRNG()
    niters = max_iters
    best_good = -1
    best = None
    it = 0
    while it < niters:
        S = subset_stream(N, 5, 1, rng)[0]
        it += 1
        cands = solver(x1[S], x2[S])
        counts = []
        for Ek in cands:
            err = sampson_errors(Ek, x1, x2)
            mask = err <= thr2
            good = int(mask.sum())
            counts.append(good)
            if good > max(best_good, 4):
                best = (Ek.copy(), mask.copy())
                best_good = good
                niters = ransac_update_num_iters(prob, (N - good) / N, 5, niters)
        if trace is not None:
            trace.append(dict(subset=S, counts=counts, niters=niters, best_good=best_good))
    if best is None:
        return None, np.zeros(N, np.uint8), it
    return best[0], best[1].astype(np.uint8), it


# ---------------------------------------------------------------------------------------------
# recoverPose
# ---------------------------------------------------------------------------------------------
def decompose_essential(E):
    """``decomposeEssentialMat`` (five-point.cpp): SVD, det sign fix, W = [[0,1,0],[-1,0,0],[0,0,1]],
    R1 = U W Vt, R2 = U Wt Vt, t = U[:,2]."""
    import cv2
    _, U, Vt = cv2.SVDecomp(np.asarray(E, np.float64), flags=cv2.SVD_MODIFY_A)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0, 1, 0], [-1, 0, 0], [0, 0, 1.0]])
    return U @ W @ Vt, U @ W.T @ Vt, U[:, 2:3].copy()


def triangulate_cv(P0, P1, x0, x1):
    import cv2
    return cv2.triangulatePoints(P0, P1, np.ascontiguousarray(x0.T), np.ascontiguousarray(x1.T))


def recover_pose_replay(E, p1, p2, focal, pp, dist_thresh=50.0, triangulate=triangulate_cv):
    """Replay of ``cv2.recoverPose(E, p1, p2, focal, pp)`` (five-point.cpp): candidates in the order
    (R1,t), (R2,t), (R1,-t), (R2,-t); masks ``Q_z*Q_w > 0``, ``z1 < dist``, ``z2 > 0``, ``z2 < dist``; the
    first candidate with the maximum count wins."""
    x1 = (np.asarray(p1, np.float64) - np.array(pp)) / focal
    x2 = (np.asarray(p2, np.float64) - np.array(pp)) / focal
    R1, R2, t = decompose_essential(E)
    P0 = np.eye(3, 4)
    cands = [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    masks = []
    for R, tt in cands:
        P = np.concatenate([R, tt], 1)
        Q = triangulate(P0, P, x1, x2)
        mask = (Q[2] * Q[3]) > 0
        Q = Q / Q[3]
        mask = (Q[2] < dist_thresh) & mask
        Q2 = P @ Q
        mask = (Q2[2] > 0) & mask
        mask = (Q2[2] < dist_thresh) & mask
        masks.append(mask)
    goods = [int(m.sum()) for m in masks]
    # cv: if good1 >= good2 && good1 >= good3 && good1 >= good4 -> 1; else if good2 >= ... (first max wins)
    best = int(np.argmax(goods))
    return goods[best], cands[best][0], cands[best][1], masks[best]
