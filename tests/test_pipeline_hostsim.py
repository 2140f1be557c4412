"""CPU: the device-resident pipeline (b200/pipeline.py, what bench.py times) reproduces the trajectory of
the unmodified reference driver (golden from oracle/gen_golden.py::gen_dfvo_driver) when fed the same
analytic network outputs -- kernels run in the host-emulation build."""
import os
import sys

import numpy as np
import pytest

from oracle import seqdata

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("overlap,iterative,threaded", [(False, False, False), (True, True, False), (True, False, True), (True, False, "pipelined")])
def test_pipeline_matches_reference_driver(hostsim_lib, overlap, iterative, threaded):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from runtime import HostsimRuntime
    from b200 import pipeline, runtime as rt_mod
    rt_mod.set_runtime(HostsimRuntime(hostsim_lib))
    # iterative: kp_selection.rigid_flow_kp + scale_recovery.method 'iterative' (SURVEY 8f rank 1), golden from the same
    # unmodified driver under ablation_scale_iterative.yml's settings
    g = np.load(os.path.join(G, "dfvo_driver_iter_188x620.npz" if iterative else "dfvo_driver_188x620.npz"))
    h, w = [int(v) for v in g["hw"]]
    K = list(g["K"])
    n = g["poses"].shape[0]

    class Injected(pipeline.FramePipeline):
        def infer(self, img, fid):
            f = seqdata.frame_inputs(fid, h, w, K, seqdata.MODES[fid % len(seqdata.MODES)])
            st = pipeline.FrameState()
            st.id = fid
            slot = self.slot(fid)
            st.raw_depth = self._buf("raw%d" % slot, (h, w), np.float32)
            st.depth = self._buf("dep%d" % slot, (h, w), np.float32)
            d = self._buf("dsrc", (h, w), np.float32).upload(f["depth"])
            self.eng.depth_post(d, self.cfg.crop.depth_crop, 0.0, 50.0, st.raw_depth, st.depth)
            if not self.eng.flow_ready:
                self.eng.flow_fwd = self.rt.empty((1, 2, h, w), np.float32)
                self.eng.flow_bwd = self.rt.empty((1, 2, h, w), np.float32)
                self.eng.flow_diff = self.rt.empty((1, h, w), np.float32)
                self.eng.flow_ready = True
            st.fwd, st.bwd, st.diff = self.flow_slot(slot)
            st.fwd.upload(f["fwd"][None]); st.bwd.upload(f["bwd"][None]); st.diff.upload(f["diff"][None, :, :, 0])
            return st

    np.random.seed(4869)
    from b200 import config
    cfg = config.default_cfg(h, w)
    if iterative:
        cfg.kp_selection.rigid_flow_kp.enable = True
        cfg.scale_recovery.method = "iterative"
    inflight = 2 if (iterative or threaded) else 1     # two frames in flight (tracker two frames behind); threaded: tracker on its own host thread
    p = Injected(K, h, w, cfg=cfg, overlap=overlap, inflight=inflight, tracker_thread=(threaded is True), pipelined=(threaded == "pipelined"))
    lag = p.lag
    if overlap:                          # two-stream mode: step(t) returns the pose of frame t-inflight, flush() the rest
        for _ in range(lag):
            assert p.step(None) is None
    tail = []
    for t in range(n):
        if not overlap:
            pose = p.step(None)
        elif t + lag < n:
            pose = p.step(None)
        else:
            if not tail:
                f = p.flush()
                tail = f if isinstance(f, list) else [f]
            pose = tail.pop(0)
        dR = pose[:3, :3].T @ g["poses"][t][:3, :3]
        ang = np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))
        dt = np.linalg.norm(pose[:3, 3] - g["poses"][t][:3, 3])
        assert ang < 1e-6 and dt < 1e-6 * max(1.0, np.linalg.norm(g["poses"][t][:3, 3])), (t, ang, dt)
    p.close()
    modes = list(p.modes.values())
    assert "PnP" in modes and "const" in modes and "E" in modes      # all three branches of dfvo.py:121-262 exercised


def test_pipeline_real_infer_plumbing(hostsim_lib):
    """The un-injected FramePipeline.infer/step (upload -> device LANCZOS feed -> depth net -> depth post ->
    flow net -> tracking) runs end to end on a small frame pair; the depth it produces equals the stage-level
    path (Engine.depth on the PIL feed), i.e. the wiring bench.py times is the tested one."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from runtime import HostsimRuntime
    import synthdata
    from b200 import native, pipeline, runtime as rt_mod
    rt = HostsimRuntime(hostsim_lib)
    rt_mod.set_runtime(rt)
    h, w = 64, 96
    K = synthdata.kitti_intrinsics(h, w)
    enc, dec = synthdata.monodepth2_weights(4869, 64, 96)      # feed >= 64: the decoder reflection-pads the 1/32 map
    p = pipeline.FramePipeline(K, h, w, precision=native.PREC_FP32, runtime=rt)
    p.load_weights(synthdata.liteflownet_weights(), enc, dec)
    frames = [synthdata.value_noise_image(h, w, 11), synthdata.value_noise_image(h, w, 12)]
    np.random.seed(1)
    for f in frames:
        pose = p.step(f)
        assert pose.shape == (4, 4) and np.all(np.isfinite(pose))
    assert p.eng.flow_ready and np.all(np.isfinite(p.ref.fwd.numpy())) and np.all(np.isfinite(p.ref.diff.numpy()))
    assert np.abs(p.ref.fwd.numpy()).max() > 0
    # depth of the last frame through the host (PIL) feed == through the device feed
    feed_host = rt.from_host(p.depth_feed_host(frames[1]))
    d_host = p.eng.depth(feed_host, out=rt.empty((p.eng.feed_h, p.eng.feed_w), np.float32)).numpy()
    d_dev = p.eng.depth(p.eng.depth_feed(rt.from_host(frames[1]))).numpy()
    assert np.array_equal(d_host, d_dev)
    raw = p.ref.raw_depth.numpy()
    assert raw.shape == (h, w) and np.all(np.isfinite(raw)) and raw.max() > 0


def test_pipelined_three_engines_real_infer_equals_in_order(hostsim_lib):
    """The bench's default mode -- three network engines in flight, tracker split into enqueue / read halves -- with the real
    (un-injected) infer path on small frames gives the poses of the in-order pipeline, frame by frame, and hands them out
    `lag` = 4 steps late; flush() returns the rest."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from runtime import HostsimRuntime
    import synthdata
    from b200 import native, pipeline, runtime as rt_mod
    rt = HostsimRuntime(hostsim_lib)
    rt_mod.set_runtime(rt)
    h, w = 64, 96
    K = synthdata.kitti_intrinsics(h, w)
    enc, dec = synthdata.monodepth2_weights(4869, 64, 96)
    flow_w = synthdata.liteflownet_weights()
    frames = [synthdata.value_noise_image(h, w, 40 + i) for i in range(7)]

    def run(**kw):
        np.random.seed(5)
        p = pipeline.FramePipeline(K, h, w, precision=native.PREC_FP32, runtime=rt, **kw)
        p.load_weights(flow_w, enc, dec)
        out = []
        for f in frames:
            r = p.step(f)
            if r is not None:
                out.append(r.copy())
        tail = p.flush() if kw.get("overlap") else None
        for q in (tail if isinstance(tail, list) else [tail]):
            if q is not None:
                out.append(q.copy())
        return p, out, [p.poses[i] for i in sorted(p.poses)]

    pa, oa, alla = run()
    pb, ob, allb = run(overlap=True, inflight=3, pipelined=True)
    assert pb.lag == 4 and pb.nslots == 6
    assert len(oa) == len(frames) and len(ob) == len(frames) and len(allb) == len(frames)
    for i in range(len(frames)):
        assert np.array_equal(alla[i], allb[i]), "pose of frame %d differs between the in-order and the pipelined three-engine mode" % i
        assert np.array_equal(oa[i], ob[i])
