"""One compute_pose_2d2d (5 repeats + homography vote + recoverPose) and one PnP call per outlier fraction, for an ncu launch
list of the tracker kernels (development aid):  ncu --metrics gpu__time_duration.sum ... python scripts/prof_tracker.py 0.6"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "df-vo_b200"))
import numpy as np, torch
import synthdata as synth
from b200 import runtime as rt_mod, tracking

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.6
rt = rt_mod.CudaRuntime(0); rt_mod.set_runtime(rt)
eng = tracking.Engine(376, 1241, rt)
K = synth.kitti_intrinsics(376, 1241)
a, b, info = synth.correspondences(seed=32, n=2000, outlier_frac=frac)
for i in range(3):
    np.random.seed(4869)
    r = tracking.compute_pose_2d2d(eng, a, b, K)
    torch.cuda.synchronize()
d = info["depth"][a[:, 1].astype(int), a[:, 0].astype(int)]
for i in range(2):
    np.random.seed(4869)
    pose, inl = tracking.compute_pose_3d2d(eng, a, b, d, K)
    torch.cuda.synchronize()
print("iterations", r["ransac_info"][:, 1], "inliers", int(r["inliers"].sum()), "pnp inliers", inl)
