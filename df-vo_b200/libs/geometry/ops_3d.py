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
    """ops_3d.py:44-67 for the configuration the tracker uses (T_1w = identity): returns (X, X1, X2)
    with only the rows the callers read populated exactly (X2[2] = depth in view 2)."""
    from b200 import runtime, tracking
    assert np.allclose(T_1w, np.eye(4)), "triangulation: the device kernel assumes view 1 = [I|0]"
    eng = tracking.default_engine()
    n = kp1.shape[0]
    z2 = eng.triangulate_depth(runtime.get().from_host(np.ascontiguousarray(kp1, np.float64)),
                               runtime.get().from_host(np.ascontiguousarray(kp2, np.float64)), n, T_2w)
    X2 = np.zeros((3, n))
    X2[2] = z2
    return None, None, X2


def unprojection_kp(kp, kp_depth, cam_intrinsics):
    """ops_3d.py:70-94: XYZ = depth * K^-1 [x, y, 1]."""
    ones = np.ones((kp.shape[0], 1))
    rays = (cam_intrinsics.inv_mat @ np.concatenate([kp, ones], 1).T).T
    return rays * np.asarray(kp_depth).reshape(-1, 1)
