"""Shared body of the drop-in sequence tests: tests/vo_driver.py (the reference driver's call sequence) against the
`libs.*` mirror, compared with the golden trajectories the UNMODIFIED reference driver produced with the reference's own
packages (tests/golden/dfvo_driver_*.npz, oracle/gen_golden.py::gen_dfvo_driver)."""
import os
import sys

import numpy as np

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ITERATIVE = {"kp_selection.rigid_flow_kp.enable": True, "scale_recovery.method": "iterative"}      # ablation_scale_iterative.yml


def make_cfg(h, w, extra=None, **paths):
    from b200 import config
    cfg = config.default_cfg(h, w)
    for k, v in {**(extra or {}), **paths}.items():
        node = cfg
        parts = k.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    return cfg


def fresh_libs():
    """(Re-)import the mirror package as `libs` (another test may have imported the reference's)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mine = os.path.join(root, "df-vo_b200")
    sys.path[:] = [mine] + [p for p in sys.path if p != mine]
    for k in [k for k in sys.modules if k == "libs" or k.startswith("libs.")]:
        del sys.modules[k]
    import libs
    assert libs.__file__.startswith(mine)


def check_poses(poses, gold):
    assert poses.shape == gold.shape
    for i in range(poses.shape[0]):
        dR = poses[i][:3, :3].T @ gold[i][:3, :3]
        ang = np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))
        dt = np.linalg.norm(poses[i][:3, 3] - gold[i][:3, 3])
        scale = max(1.0, np.linalg.norm(gold[i][:3, 3]))
        assert ang < 1e-6 and dt / scale < 1e-6, (i, ang, dt)       # BASELINE tolerance 1e-4 rad / 1e-3; observed ~1e-12
    assert np.linalg.norm(gold[-1][:3, 3]) > 1.0


def analytic_hooks(h, w, K, seen=None):
    """Network outputs replaced by the analytic frame inputs AFTER the real calls ran (random-weight networks give no
    usable correspondences); `seen` collects what the networks really produced."""
    from oracle import seqdata

    def depth(fid, raw):
        if seen is not None:
            seen.setdefault("depth", {})[fid] = np.array(raw)
        return seqdata.frame_inputs(fid, h, w, K, seqdata.MODES[fid % len(seqdata.MODES)])["depth"]

    def flow(cur_id, ref_id, flows):
        f = seqdata.frame_inputs(cur_id, h, w, K, seqdata.MODES[cur_id % len(seqdata.MODES)])
        for key, val in (((ref_id, cur_id), f["fwd"]), ((cur_id, ref_id), f["bwd"]), ((ref_id, cur_id, "diff"), f["diff"])):
            a = flows[key]
            if seen is not None and key[0] == ref_id:          # the forward flow and the inconsistency map
                seen.setdefault("flow", {})[(cur_id,) + key[2:]] = np.array(np.asarray(a))
            if hasattr(a, "dev"):                              # device-backed array of the mirror: overwrite in place
                a.dev.upload(np.ascontiguousarray(val, np.float32).reshape(a.dev.shape))
                a._host = None
            else:
                flows[key] = val
    return dict(depth=depth, flow=flow)
