"""GPU parity of correspondence selection (kp_selection.py:33-200, keypoint_sampler.py:76-143):
bit-exact index sets against golden vectors produced by the reference, and keypoints (float64,
exact) against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import synth, vo
from util import dptr

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"easy": dict(seed=21), "outliers": dict(seed=22, outlier_frac=0.3, diff_sigma=0.12),
         "sparse": dict(seed=23, diff_sigma=2.0), "toofew": dict(seed=24, diff_sigma=60.0)}


@pytest.mark.parametrize("name", list(CASES))
def test_local_bestn_and_bestn(dev_lib, name):
    g = np.load(os.path.join(G, "selection_376x1241.npz"))
    H, W = 376, 1241
    fr = synth.analytic_frame(h=H, w=W, **CASES[name])
    diff = torch.from_numpy(np.ascontiguousarray(fr["flow_diff"][..., 0])).cuda()
    flow = torch.from_numpy(fr["flow_fwd"]).cuda()
    idx = torch.zeros(2000, dtype=torch.int32, device="cuda")
    cc = torch.zeros(100, dtype=torch.int32, device="cuda")
    st = torch.zeros(4, dtype=torch.int32, device="cuda")
    dev_lib.check(dev_lib.dfvo_local_bestn(dptr(diff), None, H, W, 10, 10, 2000, 0.1, 0.05, dptr(idx), dptr(cc), dptr(st), None))
    kp1 = torch.zeros((2000, 2), dtype=torch.float64, device="cuda")
    kp2 = torch.zeros_like(kp1)
    n = torch.zeros(1, dtype=torch.int32, device="cuda")
    dev_lib.check(dev_lib.dfvo_gather_keypoints(dptr(idx), dptr(cc), 100, 20, dptr(flow), H, W, dptr(kp1), dptr(kp2), dptr(n), None))
    torch.cuda.synchronize()
    idx_h, st_h, n_h = idx.cpu().numpy(), st.cpu().numpy(), int(n.item())
    assert bool(st_h[0]) == bool(g[name + "_local_bestN_good"])
    sel = idx_h[idx_h >= 0]
    assert np.array_equal(np.sort(sel), g[name + "_local_bestN_idx_sorted"])          # bit-exact set vs reference
    good, cells = vo.local_bestn_indices(fr["flow_diff"])
    o1, o2 = vo.keypoints_from_indices(cells, fr["flow_fwd"], W)
    assert n_h == o1.shape[0] == st_h[1]
    assert np.array_equal(kp1.cpu().numpy()[:n_h], o1) and np.array_equal(kp2.cpu().numpy()[:n_h], o2)
    ws = torch.zeros(dev_lib.dfvo_bestn_workspace_bytes(H, W), dtype=torch.uint8, device="cuda")
    bi = torch.zeros(2000, dtype=torch.int32, device="cuda")
    dev_lib.check(dev_lib.dfvo_bestn(dptr(diff), H, W, 2000, dptr(bi), dptr(ws), ws.numel(), None))
    torch.cuda.synchronize()
    assert np.array_equal(bi.cpu().numpy(), g[name + "_bestN_idx_sorted"])


def test_local_bestn_degenerate(dev_lib):
    """Early-outs (kp_selection.py:121-125,175-179) and ties at the k-th value."""
    H, W = 376, 1241
    rs = np.random.RandomState(9)
    # (a) almost nothing below threshold -> good_kp_found False
    d = np.full((H, W), 5.0, np.float32); d[:3, :40] = 0.01
    # (b) enough pixels but concentrated in < 10 cells -> False
    e = np.full((H, W), 5.0, np.float32); e[:36, :600] = 0.01
    # (c) massive ties: constant map -> smallest indices of every cell
    c = np.full((H, W), 0.05, np.float32)
    for arr, want_good in ((d, False), (e, False), (c, True)):
        diff = torch.from_numpy(arr).cuda()
        idx = torch.zeros(2000, dtype=torch.int32, device="cuda")
        cc = torch.zeros(100, dtype=torch.int32, device="cuda")
        st = torch.zeros(4, dtype=torch.int32, device="cuda")
        dev_lib.check(dev_lib.dfvo_local_bestn(dptr(diff), None, H, W, 10, 10, 2000, 0.1, 0.05, dptr(idx), dptr(cc), dptr(st), None))
        torch.cuda.synchronize()
        good, cells = vo.local_bestn_indices(arr[..., None])
        assert bool(st.cpu().numpy()[0]) == good == want_good
        if good:
            sel = idx.cpu().numpy()
            assert np.array_equal(sel, np.concatenate(cells))


def _gpu_engine():
    from b200 import runtime as rt_mod, tracking
    rt = rt_mod.CudaRuntime(0)
    rt_mod.set_runtime(rt)
    return tracking.Engine(376, 1241, rt)


@pytest.mark.parametrize("name", ["clean", "outliers"])
def test_rigid_flow_map_and_selection(dev_lib, name):
    """SURVEY 8f rank 1 on the device: rigid-flow inconsistency map vs the oracle, 'uniform' / 'best' keypoint lists of
    opt_rigid_flow_kp bit-equal to the oracle and the reference golden (tests/rigid_cases.py)."""
    import rigid_cases
    assert rigid_cases.check_maps_and_selection(_gpu_engine(), name) < 5e-3


@pytest.mark.parametrize("kp_src", ["kp_best", "kp_depth"])
def test_iterative_scale_vs_reference(dev_lib, kp_src):
    import rigid_cases
    rigid_cases.check_iterative_scale(_gpu_engine(), "outliers", kp_src)
