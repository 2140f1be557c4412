"""GPU: the round-2 `libs.*` mirror additions on the device (same cases as the CPU test_mirror_api.py): geometry layers with
CUDA torch tensors in / out, opt_rigid_flow_kp, triangulation (X, X1, X2), DevArray snapshots, capacity workspaces."""
import pytest

import mirror_cases

pytestmark = pytest.mark.gpu


@pytest.fixture()
def cuda_rt(dev_lib):
    from b200 import runtime as rt_mod
    rt = rt_mod.CudaRuntime(0)
    rt_mod.set_runtime(rt)
    return rt


@pytest.mark.parametrize("as_torch", [False, True])
def test_geometry_layers(cuda_rt, as_torch):
    mirror_cases.check_geometry_layers(as_torch)


def test_geometry_layers_cuda_tensors(cuda_rt):
    """The reference calls the layers with CUDA tensors and reads `.detach().cpu().numpy()` (E_tracker.py:676-685)."""
    import numpy as np
    import torch
    from libs.geometry.rigid_flow import RigidFlow
    depth, T, Km, iKm = mirror_cases.geometry_inputs()
    h, w = depth.shape[2:]
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()
    f = RigidFlow(h, w).cuda()(cu(depth), cu(T[None]), cu(Km[None]), cu(iKm[None]), normalized=False)
    assert f.is_cuda and tuple(f.shape) == (1, 2, h, w)
    want = mirror_cases.torch_layers(depth, T, Km, iKm, False)[3]
    assert np.abs(f.detach().cpu().numpy() - want).max() < 2e-3


def test_triangulation_returns_all_views(cuda_rt):
    mirror_cases.check_triangulation()


def test_opt_rigid_flow_kp_free_function(cuda_rt):
    from b200 import tracking
    mirror_cases.check_opt_rigid_flow_kp(tracking.Engine(376, 1241, cuda_rt))


def test_devarray_copy_is_a_snapshot(cuda_rt):
    mirror_cases.check_devarray_copy(cuda_rt)


def test_workspaces_do_not_grow_with_keypoint_count(cuda_rt):
    from b200 import tracking
    mirror_cases.check_varying_keypoint_counts(tracking.Engine(376, 1241, cuda_rt))
