#!/bin/bash
# first GPU bring-up: every group in its own process under its own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 15 gpurun_out/$name.log; }
run stage_fp32 python -m pytest tests/test_gpu_stage_ops.py -q --timeout 120 -k "not tcgen05" -p no:cacheprovider
run stage_tc python -m pytest tests/test_gpu_stage_ops.py -q --timeout 120 -k "tcgen05" -p no:cacheprovider
run liteflow python -m pytest tests/test_gpu_liteflow.py -q --timeout 300 -p no:cacheprovider
run quick_bench python scripts/quick_bench.py
