#!/bin/bash
# halo conv + two-stream pipeline: suite, bench (overlap on/off), per-layer trace, host profile
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 5 gpurun_out/$name.log | cut -c1-700; }
run tests_gpu python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider -W ignore
run bench python bench.py --steps 40 --warmup 5
DFVO_OVERLAP=0 run bench_inorder python bench.py --steps 40 --warmup 5 --cpu-frames 0
DFVO_CONV_HALO=0 run bench_nohalo python bench.py --steps 40 --warmup 5 --cpu-frames 0
DFVO_TC_TRACE=1 run trace_tc python scripts/trace_tc.py
run prof_host python scripts/prof_host.py
