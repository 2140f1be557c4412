"""Host-side profile of FramePipeline.track with an idle GPU (the networks of the frame are finished before the tracker starts):
per-frame wall time by (branch, outlier fraction), a cProfile over the tracked frames, and the number of device->host reads.
Development aid:  python scripts/prof_host_tracker.py"""
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
    tmp = pipe._buf("dsrc%d" % pipe.slot(st.id), (H, W), np.float32)
    tmp.t.copy_(d_depth[slot].t)
    pipe.eng.depth_post(tmp, pipe.cfg.crop.depth_crop, 0.0, 50.0, st.raw_depth, st.depth)


np.random.seed(4869)
pipe = pipeline.FramePipeline(K, H, W, precision=native.PREC_BF16, runtime=rt, overlap=False, inject=inject)
pipe.load_weights(synth.liteflownet_weights(), enc, dec)
ref = None
ms = {}
pr = cProfile.Profile()
reads = [0]
n0 = rt_mod.Buf.numpy


def counted(self):
    reads[0] += 1
    return n0(self)


rt_mod.Buf.numpy = counted
nframes = 0
for fid in range(8 + 32):
    cur = pipe.infer(d_frames[fid % bench.N_DISTINCT], fid)
    torch.cuda.synchronize()
    if ref is not None:
        prof = fid >= 8
        r0 = reads[0]
        if prof:
            pr.enable()
        t0 = time.perf_counter()
        rel = pipe.track(cur, ref)
        dt = (time.perf_counter() - t0) * 1e3
        if prof:
            pr.disable()
            key = "%s@%.1f" % (pipe.last.get("mode"), bench.FRAME_OUTLIERS[fid % bench.N_DISTINCT])
            ms.setdefault(key, []).append((dt, reads[0] - r0))
            nframes += 1
        pipe.motion = rel.copy()
    pipe.ref = ref = cur
for k, v in sorted(ms.items()):
    print("%-10s %.3f ms   %d device->host reads" % (k, float(np.mean([a for a, _ in v])), int(np.mean([b for _, b in v]))))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(32)
print("profiled frames:", nframes)
print(s.getvalue()[:7000])
