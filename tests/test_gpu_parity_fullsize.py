"""GPU: precision-mode parity at the BASELINE sizes (VERDICT r1 "What's weak" 1): the product networks in every precision
mode (fp32 CUDA-core, tf32 tcgen05, bf16 tcgen05) against the CPU oracle at 376x1241 (flow) and 192x640 (depth), and the
downstream effect of the difference on mask bits, selected keypoints and the pose (tests/parity_cases.py explains the
end-to-end construction).  The table is written to gpurun_out/parity_fullsize.json; DESIGN.md section 5 quotes it.

Tolerances (floating point; the tolerance is the assertion):
  fp32 mode : flow EPE max < 1e-3 px, |d flow_diff| max < 2e-3 px, depth rel < 1e-4           (summation order only)
  tf32 mode : flow EPE mean < 0.01 px, depth rel max < 1e-2                                   (10-bit mantissa operands)
  bf16 mode : flow EPE mean < 0.06 px, depth rel max < 6e-2                                   (8-bit mantissa operands)
  chain on IDENTICAL inputs ('exact_inputs'): mask flips < 1e-4, keypoint set and RANSAC inlier count equal up to that,
             pose within 1e-4 rad / 1e-3 of the cv2 arm  (the north-star tolerance, met where it can be met)
  chain with a mode's error field: the selection keeps the 20 SMALLEST of ~4500 candidates per cell, i.e. extremes of a
             noisy field, so any perturbation (even 1e-2 px of white noise on the oracle's own input changes 83 % of the
             set) re-draws the set; what must hold is that the pose stays as accurate against the scene's TRUE motion as the
             reference arm's: rot err < 5e-4 rad, translation-direction err < 5e-3 for every mode.
"""
import numpy as np
import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def table(dev_lib):
    from b200 import runtime as rt_mod, tracking
    rt = rt_mod.CudaRuntime(0)
    rt_mod.set_runtime(rt)
    eng = tracking.Engine(pc.H, pc.W, rt)
    t = pc.measure_all(dev_lib, eng)
    pc.write_report(t)
    for mode, row in t.items():
        print(mode, {k: (round(v, 6) if isinstance(v, float) else v) for k, v in row.get("flow", {}).items()})
        print(mode, "chain", {k: (round(v, 6) if isinstance(v, float) else v) for k, v in row["chain"].items()})
    return t


def test_fp32_mode_matches_oracle_at_full_size(table):
    f, d = table["fp32"]["flow"], table["fp32"]["depth"]
    assert f["epe_max"] < 1e-3 and f["diff_abs_err_max"] < 2e-3, f
    assert d["rel_err_max"] < 1e-4, d


def test_tf32_mode_vs_oracle_at_full_size(table):
    f, d = table["tf32"]["flow"], table["tf32"]["depth"]
    assert f["epe_mean"] < 0.01 and f["epe_bwd_mean"] < 0.01, f
    assert d["rel_err_max"] < 1e-2, d


def test_bf16_mode_vs_oracle_at_full_size(table):
    f, d = table["bf16"]["flow"], table["bf16"]["depth"]
    assert f["epe_mean"] < 0.06 and f["epe_bwd_mean"] < 0.06, f
    assert d["rel_err_max"] < 6e-2, d
    # ordering of the modes: each step of operand precision buys accuracy
    assert table["fp32"]["flow"]["epe_mean"] < table["tf32"]["flow"]["epe_mean"] < f["epe_mean"]


def test_chain_on_identical_inputs_meets_the_north_star_tolerance(table):
    c = table["exact_inputs"]["chain"]
    assert c["mask_flip_frac"] < 1e-4 and c["kp_changed_frac"] < 5e-3, c
    assert c["pose_rot_rad"] < 1e-4 and c["pose_tdir"] < 1e-3, c


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_chain_pose_accuracy_is_kept_in_every_mode(table, mode):
    for c in table[mode]["chain_scenes"]:
        assert c["gt_rot_err_dev"] < 5e-4 and c["gt_tdir_err_dev"] < 5e-3, (mode, c)
        assert c["inliers_dev"] > 0.8 * c["kp_dev"], (mode, c)
