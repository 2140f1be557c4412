"""CPU: the PnP RANSAC of csrc/pnp.cu (run in the host-emulation build) against cv2.solvePnPRansac and the reference
PnpTracker golden.  The same checks run on the GPU in test_gpu_depth_pose.py."""
import os
import sys

import numpy as np

import pnp_cases

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _engine(hostsim_lib):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from runtime import HostsimRuntime
    from b200 import runtime as rt_mod, tracking
    rt_mod.set_runtime(HostsimRuntime(hostsim_lib))
    return tracking.Engine(376, 1241)


def test_pnp_ransac_vs_cv2(hostsim_lib):
    exact, total = pnp_cases.check_vs_cv2(_engine(hostsim_lib))
    print("repeats reproduced exactly: %d / %d" % (exact, total))


def test_pnp_tracker_vs_reference_golden(hostsim_lib):
    worst = pnp_cases.check_vs_reference_golden(_engine(hostsim_lib), np.load(os.path.join(G, "trackers_2000.npz")))
    print("worst rotation / relative translation difference: %.2e rad, %.2e" % worst)


def test_homography_ransac_gric_vs_cv2(hostsim_lib):
    worst = pnp_cases.check_homography_vs_cv2(_engine(hostsim_lib))
    print("worst relative GRIC-H difference: %.2e" % worst)
