#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-400; }
run tests_conv python -m pytest tests/test_gpu_stage_ops.py tests/test_gpu_liteflow.py -x -q --timeout 300 -p no:cacheprovider -W ignore
DFVO_TC_TRACE=1 run trace_tc python scripts/trace_tc.py
run bench python bench.py --steps 40 --warmup 5 --cpu-frames 0
