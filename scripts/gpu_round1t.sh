#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 5 gpurun_out/$name.log | cut -c1-400; }
run tests_gpu python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider -W ignore
run bench python bench.py
DFVO_PDL=0 run bench_nopdl python bench.py --cpu-frames 0
DFVO_HEAD_TC=0 run bench_nohead python bench.py --cpu-frames 0
DFVO_TC_TRACE=1 run trace_tc python scripts/trace_tc.py
