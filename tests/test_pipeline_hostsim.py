"""CPU: the device-resident pipeline (b200/pipeline.py, what bench.py times) reproduces the trajectory of
the unmodified reference driver (golden from oracle/gen_golden.py::gen_dfvo_driver) when fed the same
analytic network outputs -- kernels run in the host-emulation build."""
import os
import sys

import numpy as np

from oracle import seqdata

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pipeline_matches_reference_driver(hostsim_lib):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from runtime import HostsimRuntime
    from b200 import pipeline, runtime as rt_mod
    rt_mod.set_runtime(HostsimRuntime(hostsim_lib))
    g = np.load(os.path.join(G, "dfvo_driver_188x620.npz"))
    h, w = [int(v) for v in g["hw"]]
    K = list(g["K"])
    n = g["poses"].shape[0]

    class Injected(pipeline.FramePipeline):
        def infer(self, img, fid):
            f = seqdata.frame_inputs(fid, h, w, K, seqdata.MODES[fid % len(seqdata.MODES)])
            st = pipeline.FrameState()
            st.id = fid
            slot = fid & 1
            st.raw_depth = self._buf("raw%d" % slot, (h, w), np.float32)
            st.depth = self._buf("dep%d" % slot, (h, w), np.float32)
            d = self._buf("dsrc", (h, w), np.float32).upload(f["depth"])
            self.eng.depth_post(d, self.cfg.crop.depth_crop, 0.0, 50.0, st.raw_depth, st.depth)
            if not self.eng.flow_ready:
                self.eng.flow_fwd = self.rt.empty((1, 2, h, w), np.float32)
                self.eng.flow_bwd = self.rt.empty((1, 2, h, w), np.float32)
                self.eng.flow_diff = self.rt.empty((1, h, w), np.float32)
                self.eng.flow_ready = True
            self.eng.flow_fwd.upload(f["fwd"][None]); self.eng.flow_bwd.upload(f["bwd"][None]); self.eng.flow_diff.upload(f["diff"][None, :, :, 0])
            return st

    np.random.seed(4869)
    p = Injected(K, h, w)
    modes = []
    for t in range(n):
        pose = p.step(None)
        modes.append(p.last.get("mode"))
        dR = pose[:3, :3].T @ g["poses"][t][:3, :3]
        ang = np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))
        dt = np.linalg.norm(pose[:3, 3] - g["poses"][t][:3, 3])
        assert ang < 1e-6 and dt < 1e-6 * max(1.0, np.linalg.norm(g["poses"][t][:3, 3])), (t, ang, dt)
    assert "PnP" in modes and "const" in modes and "E" in modes      # all three branches of dfvo.py:121-262 exercised
