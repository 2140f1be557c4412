#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 1 gpurun_out/$name.log | cut -c1-160; }
run tests_pipe python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_liteflow.py -x -q --timeout 400 -p no:cacheprovider -W ignore
DFVO_INFLIGHT=2 run b_in2 python bench.py --cpu-frames 0
run b_in1 python bench.py --cpu-frames 0
DFVO_INFLIGHT=2 run b_in2b python bench.py --cpu-frames 0
