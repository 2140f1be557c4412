"""3-D helpers with the reference's signatures (libs/geometry/ops_3d.py:15-94).  ``triangulation`` runs
the per-point DLT on the device; the two tiny gather/scatter helpers are NumPy like the reference."""
import numpy as np


def convert_sparse3D_to_depth(kp, XYZ, height, width):
    """ops_3d.py:15-41 (truncation toward zero, out-of-image points dropped, last writer wins)."""
    depth = np.zeros((height, width))
    kp_int = kp.astype(int)
    keep_x = (kp_int[:, 0] >= 0) & (kp_int[:, 0] < width)
    kp_int, XYZ = kp_int[keep_x], XYZ[:, keep_x]
    keep_y = (kp_int[:, 1] >= 0) & (kp_int[:, 1] < height)
    kp_int, XYZ = kp_int[keep_y], XYZ[:, keep_y]
    depth[kp_int[:, 1], kp_int[:, 0]] = XYZ[2]
    return depth


def triangulation(kp1, kp2, T_1w, T_2w):
    """ops_3d.py:44-67: cv2.triangulatePoints(T_1w[:3], T_2w[:3], kp1, kp2) per point on the device (the 4x4 DLT of
    csrc/ransac.cu), X /= X[3]; returns (X [3,N] world, X1 = T_1w[:3] @ X, X2 = T_2w[:3] @ X)."""
    from b200 import runtime
    rt = runtime.get()
    n = kp1.shape[0]
    k1 = rt.from_host(np.ascontiguousarray(kp1, np.float64))
    k2 = rt.from_host(np.ascontiguousarray(kp2, np.float64))
    t1 = rt.from_host(np.ascontiguousarray(np.asarray(T_1w, np.float64)[:3].reshape(-1)))
    t2 = rt.from_host(np.ascontiguousarray(np.asarray(T_2w, np.float64)[:3].reshape(-1)))
    out = [rt.empty((3, n), np.float64) for _ in range(3)]
    rt.lib.check(rt.lib.dfvo_triangulate_points(k1.ptr, k2.ptr, n, t1.ptr, t2.ptr, out[0].ptr, out[1].ptr, out[2].ptr, rt.stream_ptr()))
    X, X1, X2 = (o.numpy() for o in out)
    return X, X1, X2


def unprojection_kp(kp, kp_depth, cam_intrinsics):
    """ops_3d.py:70-94: XYZ = depth * K^-1 [x, y, 1]."""
    ones = np.ones((kp.shape[0], 1))
    rays = (cam_intrinsics.inv_mat @ np.concatenate([kp, ones], 1).T).T
    return rays * np.asarray(kp_depth).reshape(-1, 1)
