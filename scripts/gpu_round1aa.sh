#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run tests_conv python -m pytest tests/test_gpu_stage_ops.py tests/test_gpu_liteflow.py tests/test_gpu_depth_pose.py -x -q --timeout 600 -p no:cacheprovider -W ignore
run net_time python scripts/net_time.py
run bench1 python bench.py --cpu-frames 0
DFVO_TC_TRACE=1 run trace_tc python scripts/trace_tc.py
