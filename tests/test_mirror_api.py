"""CPU (host-emulation build): round-2 additions to the `libs.*` mirror and the host runtime -- geometry layers, the free
function opt_rigid_flow_kp, ops_3d.triangulation returning (X, X1, X2), DevArray snapshots, capacity-allocated workspaces.
The same cases run on the device in test_gpu_mirror.py."""
import os
import sys

import pytest

import mirror_cases


@pytest.fixture()
def hostsim_rt(hostsim_lib):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from runtime import HostsimRuntime
    from b200 import runtime as rt_mod
    rt = HostsimRuntime(hostsim_lib)
    rt_mod.set_runtime(rt)
    return rt


@pytest.mark.parametrize("as_torch", [False, True])
def test_geometry_layers(hostsim_rt, as_torch):
    mirror_cases.check_geometry_layers(as_torch)


def test_triangulation_returns_all_views(hostsim_rt):
    mirror_cases.check_triangulation()


def test_opt_rigid_flow_kp_free_function(hostsim_rt):
    from b200 import tracking
    mirror_cases.check_opt_rigid_flow_kp(tracking.Engine(376, 1241, hostsim_rt))


def test_devarray_copy_is_a_snapshot(hostsim_rt):
    mirror_cases.check_devarray_copy(hostsim_rt)


def test_workspaces_do_not_grow_with_keypoint_count(hostsim_rt):
    from b200 import tracking
    mirror_cases.check_varying_keypoint_counts(tracking.Engine(376, 1241, hostsim_rt))
