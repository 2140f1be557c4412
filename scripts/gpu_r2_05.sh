#!/bin/bash
# round 2, GPU call 5: layer-chain kernels (cooperative, grid barriers), warp replay -- network parity tests first, then suite + bench + trace
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run tests_nets 600 python -m pytest tests/test_gpu_liteflow.py tests/test_gpu_depth_pose.py -q --timeout 300 -p no:cacheprovider -W ignore -x
run tests_gpu 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -W ignore
run bench 600 python bench.py --warmup 3 --steps 60
DFVO_TC_TRACE=1 run trace 300 python scripts/trace_tc.py
DFVO_CONV_CHAIN=0 run bench_nochain 600 python bench.py --warmup 3 --steps 60 --no-extras --cpu-frames 0
