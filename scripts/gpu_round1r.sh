#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 5 gpurun_out/$name.log | cut -c1-400; }
run tests_gpu python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider -W ignore
run bench python bench.py --steps 60 --warmup 8
DFVO_GRAPHS=0 run bench_nograph python bench.py --steps 60 --warmup 8 --cpu-frames 0
