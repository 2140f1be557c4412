"""Ad-hoc device timing of the LiteFlowNet path (development aid; bench.py is the contract)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "df-vo_b200"))
import numpy as np, torch
from b200 import native
from oracle import synth
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import dptr

H, W = 376, 1241
lib = native.load()
w = synth.liteflownet_weights()
ref, cur = synth.value_noise_image(H, W, 1), synth.value_noise_image(H, W, 2)
d_ref, d_cur = torch.from_numpy(ref).cuda(), torch.from_numpy(cur).cuda()
for prec, name in [pn for pn in ((1, "bf16"), (0, "fp32")) if pn[1] in os.environ.get("QB_PRECS", "bf16,fp32")]:
    ctx = native.Context(lib); ctx.load_weights(0, w); ctx.liteflow_build(H, W, 1, prec)
    fwd = torch.zeros((2, H, W), device="cuda"); bwd = torch.zeros_like(fwd); diff = torch.zeros((H, W), device="cuda")
    args = ([d_ref.data_ptr(), d_cur.data_ptr()], dptr(fwd), dptr(bwd), dptr(diff))
    for _ in range(int(os.environ.get("QB_WARM", "3"))): ctx.liteflow_forward(*args)
    torch.cuda.synchronize()
    n = int(os.environ.get("QB_ITERS", "10"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): ctx.liteflow_forward(*args)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("%s: device %.3f ms/pair, wall %.3f ms/pair" % (name, e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n), flush=True)
    ctx.close()
