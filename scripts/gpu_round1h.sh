#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 45 gpurun_out/$name.log; }
run liteflow python -m pytest tests/test_gpu_liteflow.py tests/test_gpu_stage_ops.py -q --timeout 300 -p no:cacheprovider
run prof_host python scripts/prof_host.py
