"""cProfile of FramePipeline.step in the bench's default mode (3 engines, pipelined tracker): where the HOST spends a step.
Development aid:  python scripts/prof_host_step.py"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "df-vo_b200"))
import numpy as np
import torch

import bench
import synthdata as synth
from b200 import native, pipeline, runtime as rt_mod

rt = rt_mod.CudaRuntime(0); rt_mod.set_runtime(rt)
H, W = bench.H, bench.W
K, frames, analytic = bench.make_inputs(0)
enc, dec = synth.monodepth2_weights(4869, bench.FEED_H, bench.FEED_W)
d_frames = [rt.from_host(f) for f in frames]
d_fwd = [rt.from_host(a["fwd"][None]) for a in analytic]; d_bwd = [rt.from_host(a["bwd"][None]) for a in analytic]
d_diff = [rt.from_host(a["diff"][None, :, :, 0]) for a in analytic]; d_depth = [rt.from_host(a["depth"]) for a in analytic]


def inject(pipe, st):
    slot = st.id % bench.N_DISTINCT
    if st.fwd is not None:
        st.fwd.t.copy_(d_fwd[slot].t); st.bwd.t.copy_(d_bwd[slot].t); st.diff.t.copy_(d_diff[slot].t)
    with pipe.depth_stream(st.id):
        tmp = pipe._buf("dsrc%d" % pipe.slot(st.id), (H, W), np.float32)
        tmp.t.copy_(d_depth[slot].t)
        pipe.eng.depth_post(tmp, pipe.cfg.crop.depth_crop, 0.0, 50.0, st.raw_depth, st.depth)


np.random.seed(4869)
pipe = pipeline.FramePipeline(K, H, W, precision=native.PREC_BF16, runtime=rt, overlap=True, inflight=3, inject=inject, pipelined=True)
pipe.load_weights(synth.liteflownet_weights(), enc, dec)
for _ in range(12):
    pipe.step(d_frames[pipe.stage % bench.N_DISTINCT])
torch.cuda.synchronize()
n = 64
t0 = time.perf_counter()
for _ in range(n):
    pipe.step(d_frames[pipe.stage % bench.N_DISTINCT])
torch.cuda.synchronize()
print("unprofiled: %.3f ms / step" % ((time.perf_counter() - t0) / n * 1e3))
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
for _ in range(n):
    pipe.step(d_frames[pipe.stage % bench.N_DISTINCT])
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pr.disable()
print("profiled: %.3f ms / step" % (dt / n * 1e3))
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(30)
    print(s.getvalue()[:6500])
