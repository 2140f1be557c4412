"""GPU time of one network forward as the product runs it (CUDA-graph replay on a side stream), timed with one event pair
around the call -- development aid: compare with the per-kernel sums of the ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "df-vo_b200"))
import numpy as np, torch
import synthdata as synth
from b200 import native, pipeline, runtime as rt_mod

H, W = 376, 1241
rt = rt_mod.CudaRuntime(0); rt_mod.set_runtime(rt)
K = synth.kitti_intrinsics(H, W)
pipe = pipeline.FramePipeline(K, H, W, precision=native.PREC_BF16, runtime=rt)
enc, dec = synth.monodepth2_weights(4869, 192, 640)
pipe.load_weights(synth.liteflownet_weights(), enc, dec)
frames = [rt.from_host(synth.value_noise_image(H, W, i)) for i in range(2)]
s = torch.cuda.Stream()
feed = pipe.eng.depth_feed(frames[0])
torch.cuda.synchronize()
for name, fn in (("flow", lambda: pipe.eng.flow(frames)), ("depth", lambda: pipe.eng.depth(feed))):
    with torch.cuda.stream(s):
        for _ in range(4):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20):
            fn()
        e1.record(s)
    torch.cuda.synchronize()
    print("%s forward: %.3f ms (graphs=%s pdl=%s)" % (name, e0.elapsed_time(e1) / 20, os.environ.get("DFVO_GRAPHS", "1"), os.environ.get("DFVO_PDL", "1")))
