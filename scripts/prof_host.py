"""Where does a pipeline step spend its time?  Wall-clock sections with syncs in between (dev aid)."""
import sys, os, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "df-vo_b200"))
import numpy as np, torch
import bench
from b200 import native, pipeline, tracking, runtime as rt_mod
import synthdata as synth
rt = rt_mod.CudaRuntime(0); rt_mod.set_runtime(rt)
K, frames, analytic = bench.make_inputs(0)
enc, dec = synth.monodepth2_weights(4869, 192, 640)
pipe = pipeline.FramePipeline(K, 376, 1241, runtime=rt)
pipe.load_weights(synth.liteflownet_weights(), enc, dec)
H, W = 376, 1241
d_fwd = [rt.from_host(a["fwd"][None]) for a in analytic]; d_bwd = [rt.from_host(a["bwd"][None]) for a in analytic]
d_diff = [rt.from_host(a["diff"][None, :, :, 0]) for a in analytic]; d_depth = [rt.from_host(a["depth"]) for a in analytic]
d_frames = [rt.from_host(f) for f in frames]; d_feeds = [pipe.eng.depth_feed(f) for f in d_frames]
T = {}
def tic(): torch.cuda.synchronize(); return time.perf_counter()
def add(k, t0): torch.cuda.synchronize(); T[k] = T.get(k, 0) + time.perf_counter() - t0
np.random.seed(1)
import cv2
from b200 import hostmath
for it in range(24):
    slot = it % 8
    t0 = tic(); d = pipe.eng.depth(d_feeds[slot]); add("depth_net", t0)
    t0 = tic(); pipe.eng.flow([d_frames[slot - 1], d_frames[slot]]); add("flow_net", t0)
    pipe.eng.flow_fwd.t.copy_(d_fwd[slot].t); pipe.eng.flow_bwd.t.copy_(d_bwd[slot].t); pipe.eng.flow_diff.t.copy_(d_diff[slot].t)
    raw = pipe._buf("raw0", (H, W), np.float32); dep = pipe._buf("dep0", (H, W), np.float32)
    t0 = tic(); pipe.eng.depth_post(d_depth[slot], pipe.cfg.crop.depth_crop, 0.0, 50.0, raw, dep); add("depth_post", t0)
    t0 = tic(); good, n, k1, k2 = pipe.eng.select_local_bestn(pipe.eng.flow_diff, pipe.eng.flow_fwd, 10, 10, 2000, 0.1); add("select", t0)
    t0 = tic(); kp_ref = k1.numpy()[:n]; kp_cur = k2.numpy()[:n]; add("kp_d2h", t0)
    t0 = tic(); perms = [np.random.permutation(n) for _ in range(5)]; add("perms", t0)
    t0 = tic(); w = pipe.eng.essential_launch(k2, k1, n, perms, K); add("ess_gpu", t0)
    t0 = tic(); Hm, _ = cv2.findHomography(kp_cur, kp_ref, method=cv2.RANSAC, confidence=0.99, ransacReprojThreshold=1); add("cv2_homography", t0)
    t0 = tic(); hostmath.calc_gric(hostmath.homography_residual(Hm, kp_cur, kp_ref), 0.8, n, "HMat"); add("gric_h", t0)
    t0 = tic(); info = w["info"].numpy(); g = w["gric"].numpy(); m = w["mask"].numpy(); add("ess_d2h", t0)
    t0 = tic(); Rt, ch = pipe.eng.recover_pose(w, 0, k2, k1, n, K); add("recover_pose", t0)
    pose = np.eye(4); pose[:3, :3] = Rt[:9].reshape(3, 3); pose[:3, 3] = Rt[9:]
    t0 = tic(); s = pipe.scale_recovery(kp_ref, kp_cur, k2, np.linalg.inv(pose), dep, n); add("scale", t0)
    if it == 7:
        T.clear()       # discard warm-up
for k, v in T.items():
    print("%-16s %8.3f ms/frame" % (k, v / 16 * 1e3))
print("iters info", info[:, 1], "n", n, "scale", s)
# cProfile of full steps
pipe2 = pipe
pr = cProfile.Profile()
def inject_infer(img, fid):
    slot = fid % 8
    st = pipeline.FrameState(); st.id = fid; s2 = fid & 1
    st.img = d_frames[slot]
    d = pipe.eng.depth(d_feeds[slot])
    st.raw_depth = pipe._buf("raw%d" % s2, (H, W), np.float32); st.depth = pipe._buf("dep%d" % s2, (H, W), np.float32)
    if pipe.ref is not None: pipe.eng.flow([pipe.ref.img, st.img])
    pipe.eng.flow_fwd.t.copy_(d_fwd[slot].t); pipe.eng.flow_bwd.t.copy_(d_bwd[slot].t); pipe.eng.flow_diff.t.copy_(d_diff[slot].t)
    pipe.eng.depth_post(d_depth[slot], pipe.cfg.crop.depth_crop, 0.0, 50.0, st.raw_depth, st.depth)
    return st
pipe.infer = inject_infer
for _ in range(4): pipe.step(None)
torch.cuda.synchronize()
pr.enable()
t0 = time.perf_counter()
for _ in range(20): pipe.step(None)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pr.disable()
print("full step %.3f ms" % (dt / 20 * 1e3))
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
