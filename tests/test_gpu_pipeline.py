"""GPU: the two-stream FramePipeline (networks of frame t on one stream -- replayed as CUDA graphs -- while frame t-1
is tracked on the other, homography vote on a host thread) is the same computation as the in-order pipeline:
identical flows, depths and poses frame by frame (dfvo.py:299-345 + 121-262 once per frame either way)."""
import numpy as np
import pytest

import synthdata
from b200 import native, pipeline, runtime as rt_mod

pytestmark = pytest.mark.gpu


def run(overlap, frames, analytic, K, h, w, enc, dec, flow_w, inflight=1, tracker_thread=False, pipelined=False):
    rt = rt_mod.CudaRuntime(0)
    rt_mod.set_runtime(rt)
    np.random.seed(4869)
    p = pipeline.FramePipeline(K, h, w, precision=native.PREC_BF16, runtime=rt, overlap=overlap, inflight=inflight, tracker_thread=tracker_thread, pipelined=pipelined)
    p.load_weights(flow_w, enc, dec)
    base_infer = p.infer
    net_flows = {}

    def infer(img, fid):
        # real networks on the real frame; the tracker then consumes analytic rigid-scene flow / depth (random-init
        # networks give no consistent correspondences) copied over the network outputs on the same stream
        st = base_infer(img, fid)
        a = analytic[fid % len(analytic)]
        if st.fwd is not None:
            net_flows[fid] = (st.fwd.t.clone(), st.diff.t.clone())
            st.fwd.upload(a["fwd"][None]); st.bwd.upload(a["bwd"][None]); st.diff.upload(a["diff"][None, :, :, 0])
        with p.depth_stream(fid):                 # same stream as (hence ordered after) the depth network
            net_flows.setdefault("depth", {})[fid] = st.raw_depth.t.clone()
            d = p._buf("dsrc%d" % p.slot(fid), (h, w), np.float32).upload(a["depth"])
            p.eng.depth_post(d, p.cfg.crop.depth_crop, 0.0, 50.0, st.raw_depth, st.depth)
        return st

    p.infer = infer
    poses = []
    for f in frames:
        r = p.step(f)
        if r is not None:
            poses.append(r.copy())
    last = p.flush()
    for q in (last if isinstance(last, list) else [last]):
        if q is not None:
            poses.append(q.copy())
    rt.torch.cuda.synchronize()
    p.close()
    return poses, net_flows, [p.poses[i] for i in sorted(p.poses)]


def test_overlap_pipeline_equals_in_order(dev_lib):
    h, w = 128, 416
    K = synthdata.kitti_intrinsics(h, w)
    enc, dec = synthdata.monodepth2_weights(4869, 64, 96)
    flow_w = synthdata.liteflownet_weights()
    n = 7                                         # > 2 uses of every buffer slot: eager, captured and replayed forwards
    frames = [synthdata.value_noise_image(h, w, 30 + i) for i in range(n)]
    analytic = [synthdata.frame_inputs(i, h, w, K, "normal") for i in range(n)]
    pa, fa, all_a = run(False, frames, analytic, K, h, w, enc, dec, flow_w)
    pb, fb, all_b = run(True, frames, analytic, K, h, w, enc, dec, flow_w)
    pc, fc, all_c = run(True, frames, analytic, K, h, w, enc, dec, flow_w, inflight=2)       # two network engines in flight
    pd, fd, all_d = run(True, frames, analytic, K, h, w, enc, dec, flow_w, inflight=2, tracker_thread=True)   # + tracker on its own host thread
    pe, fe, all_e = run(True, frames, analytic, K, h, w, enc, dec, flow_w, inflight=2, pipelined=True)        # tracker split into enqueue / read halves
    assert len(pa) == n and len(pb) == n and len(pc) == n and len(pd) == n and len(pe) == n
    for i in range(1, n):
        assert (fa[i][0] == fb[i][0]).all() and (fa[i][1] == fb[i][1]).all(), "network flow of frame %d differs" % i
        assert (fa["depth"][i] == fb["depth"][i]).all(), "network depth of frame %d differs" % i
    for i in range(1, n):
        assert (fa[i][0] == fc[i][0]).all() and (fa["depth"][i] == fc["depth"][i]).all(), "two-engine networks of frame %d differ" % i
    for i in range(n):
        assert np.array_equal(all_a[i], all_b[i]), "pose of frame %d differs between in-order and overlapped pipeline" % i
        assert np.array_equal(all_a[i], all_c[i]), "pose of frame %d differs with two frames in flight" % i
        assert np.array_equal(all_a[i], all_d[i]), "pose of frame %d differs with the tracker thread" % i
        assert np.array_equal(pc[i], pd[i]), "step() / flush() hand out different poses with the tracker thread (frame %d)" % i
        assert np.array_equal(all_a[i], all_e[i]) and np.array_equal(pc[i], pe[i]), "pose of frame %d differs in the pipelined-tracker mode" % i
    assert not np.allclose(all_a[-1], np.eye(4))
