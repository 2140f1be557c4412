#!/usr/bin/env python
"""bench.py -- VO frames/s of the DF-VO tracking hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): KITTI-odometry-shaped stream of 376x1241 RGB frames, full hybrid
tracker per frame: monodepth2 depth + LiteFlowNet forward/backward flow + forward-backward consistency +
local best-N selection + 5x essential-matrix RANSAC with GRIC + pose recovery + scale recovery (PnP
fallback when the E-model is rejected).  One *step* = one VO frame.  With N GPUs every rank tracks its
own sequence (weak scaling); weights are generated on rank 0 and broadcast with NCCL.

There are no trained weights or KITTI frames offline: frames are seeded synthetic textures and both
networks run with seeded random-init weights (their cost is the real cost).  Random-weight flow has no
consistent correspondences, so the tracker stages (selection, RANSAC, scale, PnP) are fed analytic
rigid-scene flow / depth of the same shapes, copied over the network outputs on the device inside the
timed region; every kernel of the frame runs every step.

The pipeline is `b200.pipeline.FramePipeline(overlap=True)`: LiteFlowNet and monodepth2 of frame t are enqueued on
their own streams (their bodies replayed as CUDA graphs) while frame t-1 is tracked on a third stream, so K steps =
K frames inferred AND K frames tracked (`DFVO_OVERLAP=0` times the in-order variant; both give identical poses,
tests/test_gpu_pipeline.py).

Output: ONE JSON line (rank 0).  `value` = frames/s with frames already in HBM; `e2e` = the same metric
through the public API (host uint8 frames in pinned memory -> pose on the host) with H2D/D2H inside the
timed region; `roofline` = the tcgen05 convolution kernels against the measured bf16 peak (per-launch CUDA events
on an in-order stream; `traffic` = their DRAM bytes per frame from ncu, profiles/conv_tc_traffic.json);
`cpu_baseline` = the CPU oracle port of the same frame timed on this box's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "df-vo_b200"))

import numpy as np

H, W = 376, 1241
FEED_H, FEED_W = 192, 640
METRIC = "VO frames/sec on 376x1241 pairs"
WORKLOAD = "KITTI-odom seq-shape stream (376x1241), full hybrid tracker (depth+flow+E-RANSAC/PnP), 1 frame per step"
N_DISTINCT = 8                     # distinct synthetic frames cycled through the run


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=3, help="frames of the CPU baseline sample")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ------------------------------------------------------------------------------------------------
# synthetic inputs (synthdata.py: seeded generators only, no arithmetic of the path)
# ------------------------------------------------------------------------------------------------
def make_inputs(seed):
    import synthdata as synth
    K = synth.kitti_intrinsics(H, W)
    frames = [synth.value_noise_image(H, W, seed * 100 + i) for i in range(N_DISTINCT)]
    modes = ["normal"] * N_DISTINCT
    modes[5] = "still"             # one PnP-fallback frame per cycle (GRIC prefers the homography)
    analytic = [synth.frame_inputs(i, H, W, K, modes[i]) for i in range(N_DISTINCT)]
    return K, frames, analytic


# ------------------------------------------------------------------------------------------------
# CPU baseline / --impl reference: the oracle port of one frame on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_frame(oracle_state, ref_img, cur_img, analytic, K):
    """One reference frame on the CPU (oracle restatement of dfvo.py:299-345 + 121-262): both networks in
    torch fp32, consistency map, then selection / E-tracker / scale (or PnP) on the analytic flow."""
    import torch
    from oracle import nets, vo
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    p_flow, p_enc, p_dec = oracle_state
    to_t = lambda im: torch.from_numpy(np.transpose(im / 255, (2, 0, 1))).unsqueeze(0).float()
    with torch.no_grad():
        nets.liteflow_inference_flow(p_flow, to_t(ref_img), to_t(cur_img))
        import PIL.Image as pil
        feed = np.transpose(np.asarray(pil.fromarray(cur_img).resize((FEED_W, FEED_H), pil.LANCZOS), np.float32) / 255, (2, 0, 1))[None]
        d = nets.monodepth2_inference_depth(p_enc, p_dec, torch.from_numpy(np.ascontiguousarray(feed)))[0, 0].numpy()
    vo.preprocess_depth(vo.resize_nearest(d, W, H), [[0.3, 1], [0, 1]], [0, 50])
    depth = vo.preprocess_depth(analytic["depth"], [[0.3, 1], [0, 1]], [0, 50])
    good, cells = vo.local_bestn_indices(analytic["diff"])
    pose = np.eye(4)
    if not good:
        return pose
    kp1, kp2 = vo.keypoints_from_indices(cells, analytic["fwd"], W)
    r = vo.compute_pose_2d2d(kp1, kp2, K)
    pose[:3, :3], pose[:3, 3:] = r["R"], r["t"]
    scale = -1
    if np.linalg.norm(r["t"]) != 0:
        scale = vo.find_scale_from_depth(kp1, kp2, np.linalg.inv(pose), depth, K)
        if scale != -1:
            pose[:3, 3] *= scale
    if np.linalg.norm(r["t"]) == 0 or scale == -1:
        pose = vo.compute_pose_3d2d(kp1, kp2, depth, K)[0]
    return pose


def cpu_baseline(n_frames, K, frames, analytic):
    import torch
    from oracle import nets, synth
    # all host threads it can usefully use: oneDNN/OpenMP convolutions stop scaling (and then collapse) well
    # before a 100+-core host is saturated, so the pool is capped; `cores` reports what was used
    cores = min(os.cpu_count() or 1, int(os.environ.get("DFVO_CPU_THREADS", "32")))
    torch.set_num_threads(cores)
    enc, dec = synth.monodepth2_weights(4869, FEED_H, FEED_W)
    state = (nets.to_torch(synth.liteflownet_weights()),
             {k: v for k, v in nets.to_torch(enc).items() if not isinstance(v, int)}, nets.to_torch(dec))
    np.random.seed(4869)
    poses = []
    cpu_frame(state, frames[0], frames[1], analytic[1], K)          # warm-up (thread pools, allocator)
    np.random.seed(4869)
    t0 = time.perf_counter()
    for i in range(1, n_frames + 1):
        poses.append(cpu_frame(state, frames[(i - 1) % N_DISTINCT], frames[i % N_DISTINCT], analytic[i % N_DISTINCT], K))
    dt = time.perf_counter() - t0
    return dict(value=n_frames / dt, unit="frames/s", cores=cores, kind="port",
                sample="%d frames of the same workload through oracle/ (torch-fp32 networks + cv2/sklearn solvers), %.1f s" % (n_frames, dt)), poses


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) > 8 for i in range(4) if r[5 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    K, frames, analytic = make_inputs(0)
    # `steps` bounded samples; each step = one frame of the same workload on all host threads
    n = max(1, min(args.steps, 6))
    base, _ = cpu_baseline(n, K, frames, analytic)
    line = dict(impl="reference", metric=METRIC, value=base["value"], unit="frames/s", n_gpus=args.gpus, steps=n, warmup=1,
                ms_per_step=1e3 / base["value"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=WORKLOAD, note="reference's CPU path = oracle port (the reference is pure "
                "Python + torch/cv2/sklearn; /root/reference does not travel to the GPU box)"),
                cpu_baseline=base, e2e=dict(value=base["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_b200(args):
    import torch
    rank, local_rank, world = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for --impl b200)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from b200 import native, pipeline, runtime as rt_mod
    import synthdata as synth
    rt = rt_mod.CudaRuntime(local_rank)
    rt_mod.set_runtime(rt)
    lib = rt.lib

    # ---- weights: rank 0 generates, NCCL broadcast to the others (the only collective of the path)
    enc, dec = synth.monodepth2_weights(4869, FEED_H, FEED_W)
    flow_w = synth.liteflownet_weights()
    if world > 1:
        from b200 import multi
        flow_w, enc, dec = multi.broadcast_weights([flow_w, enc, dec], src=0, device=torch.device("cuda", local_rank))

    K, frames, analytic = make_inputs(rank)
    np.random.seed(4869 + rank)
    overlap = os.environ.get("DFVO_OVERLAP", "1") != "0"
    inflight = int(os.environ.get("DFVO_INFLIGHT", "2")) if overlap else 1       # network engines in flight (measured: 2 > 1 by ~5 %)
    pipe = pipeline.FramePipeline(K, H, W, precision=native.PREC_BF16, runtime=rt, overlap=overlap, inflight=inflight)
    pipe.load_weights(flow_w, enc, dec)

    # device-resident copies of everything a step consumes
    d_frames = [rt.from_host(f) for f in frames]
    d_fwd = [rt.from_host(a["fwd"][None]) for a in analytic]
    d_bwd = [rt.from_host(a["bwd"][None]) for a in analytic]
    d_diff = [rt.from_host(a["diff"][None, :, :, 0]) for a in analytic]
    d_depth = [rt.from_host(a["depth"]) for a in analytic]
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    state = dict(resident=True, i=0, h2d=0, d2h=0)

    def inject(pipe, slot, st):
        # analytic flow / depth over the (random-weight) network outputs: D2D, inside the timed region
        if st.fwd is not None:
            st.fwd.t.copy_(d_fwd[slot].t); st.bwd.t.copy_(d_bwd[slot].t); st.diff.t.copy_(d_diff[slot].t)
        with pipe.depth_stream(st.id):                  # ordered after the depth network's own post-processing
            tmp = pipe._buf("dsrc%d" % pipe.slot(st.id), (H, W), np.float32)
            tmp.t.copy_(d_depth[slot].t)
            pipe.eng.depth_post(tmp, pipe.cfg.crop.depth_crop, 0.0, 50.0, st.raw_depth, st.depth)

    def make_infer(pipe):
        def infer(img, fid):
            slot = fid % N_DISTINCT
            st = pipeline.FrameState()
            st.id = fid
            s2 = pipe.slot(fid)
            eng = pipe.engine_for(fid)
            if state["resident"]:
                st.img = d_frames[slot]
                feed = eng.depth_feed(st.img)
            else:
                st.img = pipe._buf("img%d" % s2, (H, W, 3), np.uint8)
                st.img.t.copy_(pinned[slot], non_blocking=True)                       # H2D from pinned memory
                pipe.mark_image_ready(st)
                feed = eng.depth_feed(st.img)                                          # PIL-exact LANCZOS + ToTensor on the device
                state["h2d"] += frames[slot].nbytes
            st.raw_depth = pipe._buf("raw%d" % s2, (H, W), np.float32)
            st.depth = pipe._buf("dep%d" % s2, (H, W), np.float32)
            with pipe.depth_stream(fid):                    # monodepth2 on its side stream (overlap mode), as FramePipeline.infer does
                d = eng.depth(feed)
                eng.depth_post(d, pipe.cfg.crop.depth_crop, 0.0, 50.0, st.raw_depth, st.depth)
            if pipe.ref is not None:
                pipe.wait_reference_image()
                st.fwd, st.bwd, st.diff = pipe.flow_slot(s2)
                eng.flow([pipe.ref.img, st.img], out=(st.fwd, st.bwd, st.diff))
            inject(pipe, slot, st)
            return st
        return infer

    pipe.infer = make_infer(pipe)
    # count the bytes the host side moves per step
    up0, dn0 = rt_mod.Buf.upload, rt_mod.Buf.numpy

    def up(self, arr):
        state["h2d"] += int(np.asarray(arr).nbytes)
        return up0(self, arr)

    def dn(self):
        a = dn0(self)
        state["d2h"] += int(a.nbytes)
        return a
    rt_mod.Buf.upload, rt_mod.Buf.numpy = up, dn

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, resident):
        state["resident"] = resident
        state["h2d"] = state["d2h"] = 0
        l0 = lib.dfvo_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n_steps):
            pipe.step(None)               # overlap mode: networks of frame t on one stream while frame t-1 is tracked on the other
        if overlap:                       # the closing event waits for both of the pipeline's streams
            cs = torch.cuda.current_stream()
            for sx in pipe.s_nets + pipe.s_depths + [pipe.s_trk]:
                cs.wait_stream(sx)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            from b200 import multi
            ms = multi.max_over_ranks(ms, device=torch.device("cuda", local_rank))
        return ms, lib.dfvo_launch_count() - l0

    # warm-up: every network engine needs three forwards before it replays its CUDA graph (eager, capture, replay)
    warmup = max(3 * inflight + 1, args.warmup)
    pipe.step(None)                                   # frame 0 (no flow yet)
    for _ in range(warmup):
        pipe.step(None)
    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
    ms, launches = timed(args.steps, True)
    clk = clocks.stop() if rank == 0 else None
    modes = dict(last=pipe.last.get("mode"))
    ms_e2e, _ = timed(args.steps, False)
    h2d, d2h = state["h2d"] / args.steps, state["d2h"] / args.steps

    # ---- roofline of the dominant kernels (tcgen05 convs): CUDA-event timing of every launch over a few steps, on an
    # in-order single-stream pipeline sharing the built networks (in the two-stream pipeline the tracker's and the depth
    # network's kernels run beside the convolutions, which would be charged to whichever launch they overlap)
    pipe.flush()
    torch.cuda.synchronize()
    prof = pipeline.FramePipeline(K, H, W, precision=native.PREC_BF16, runtime=rt, overlap=False, engine=pipe.eng)
    prof.infer = make_infer(prof)
    state["resident"] = True
    prof.step(None)
    prof.step(None)
    torch.cuda.synchronize()
    lib.dfvo_profile_enable(1)
    prof_steps = 5
    for _ in range(prof_steps):
        prof.step(None)
    torch.cuda.synchronize()
    import ctypes
    tc_ms, tc_n, tc_fl = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
    lib.dfvo_profile_read(ctypes.byref(tc_ms), ctypes.byref(tc_n), ctypes.byref(tc_fl))
    lib.dfvo_profile_enable(0)

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    achieved = (tc_fl.value / 1e12) / (tc_ms.value / 1e3) if tc_ms.value > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "conv_tc_traffic.json")))["dram_bytes_per_frame"]
    except Exception:
        pass
    if args.cpu_frames > 0 and world == 1:
        base, cpu_poses = cpu_baseline(args.cpu_frames, K, frames, analytic)
    else:
        base = None          # the CPU baseline is reported at N = 1 only (and skipped in ncu profiling runs)
    value = world * args.steps / (ms / 1e3)
    line = dict(
        metric=METRIC, value=value, unit="frames/s", n_gpus=world, steps=args.steps, warmup=warmup,
        ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="bf16 tensor-core convs (fp32 accumulate) + fp32 flow/selection + fp64 pose solvers",
        data="synthetic frames + seeded random-init weights; tracker stages fed analytic rigid-scene flow/depth (see bench.py docstring)",
        config=dict(workload=WORKLOAD, image=[H, W], flow_net_input=[352, 1216], depth_feed=[FEED_H, FEED_W], keypoints=2000,
                    ransac_repeats=5, sequences_per_gpu=1, parallelism="1 sequence per GPU, NCCL weight broadcast only",
                    streams=("%d network engine(s) on their own streams, tracker %d frame(s) behind; K steps = K frames inferred and K tracked"
                             % (inflight, inflight)) if overlap else "1 (in order)",
                    l2="per-frame activation working set (>1 GB written/read per frame) exceeds the 126 MB L2; no explicit flush",
                    last_frame_branch=modes["last"]),
        clocks=clk,
        e2e=dict(value=world * args.steps / (ms_e2e / 1e3), unit="frames/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                 ms_per_step=ms_e2e / args.steps),
        gpu_launches=int(launches),
        roofline=dict(kernel="k_conv_tc (tcgen05 implicit-GEMM conv, %d launches/frame)" % (tc_n.value // prof_steps), bound="tensor",
                      achieved=achieved, peak=peak_tf, unit="TFLOP/s", frac=achieved / peak_tf if peak_tf else None, traffic=traffic,
                      peak_source=peak_src, algorithmic_gflop_per_frame=tc_fl.value / prof_steps / 1e9,
                      kernel_ms_per_frame=tc_ms.value / prof_steps, share_of_step=(tc_ms.value / prof_steps) / (ms / args.steps)),
        cpu_baseline=base,
    )
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
