"""One fused E-tracker frame (selection excluded: keypoints are given) -- homography model, five essential-matrix repeats, the
device-side tail (recoverPose, vote, depth ratios, scale RANSAC) -- and one PnP call per outlier fraction, for an ncu launch list of
the tracker kernels (development aid):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/prof_tracker.py 0.6"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "df-vo_b200"))
import numpy as np
import torch

import synthdata as synth
from b200 import runtime as rt_mod, tracking

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.6
rt = rt_mod.CudaRuntime(0); rt_mod.set_runtime(rt)
eng = tracking.Engine(376, 1241, rt)
K = synth.kitti_intrinsics(376, 1241)
a, b, info = synth.correspondences(seed=32, n=2000, outlier_frac=frac)
n = a.shape[0]
depth = info["depth"].astype(np.float32)
dp32 = (depth * ((depth < 50) & (depth > 0))).astype(np.float32)
b_ref, b_cur, b_depth = rt.from_host(a), rt.from_host(b), rt.from_host(dp32)
for i in range(3):
    np.random.seed(4869)
    perms = []
    for _ in range(5):
        order = np.arange(0, n, 1)
        np.random.shuffle(order)
        perms.append(order)
    h = eng.homography_launch(b_cur, b_ref, n)
    w = eng.essential_launch(b_cur, b_ref, n, perms, K, threshold=0.2)
    o = eng.essential_tail(w, h, b_cur, b_ref, n, K, b_depth, np.random)
    torch.cuda.synchronize()
d = info["depth"][a[:, 1].astype(int), a[:, 0].astype(int)]
for i in range(2):
    np.random.seed(4869)
    pose, inl = tracking.compute_pose_3d2d(eng, a, b, d, K)
    torch.cuda.synchronize()
print("iterations", o["ransac_info"][:, 1], "valid", o["valid"], "scale", o["scale"], "ratios", o["n_ratios"], "pnp inliers", inl)
