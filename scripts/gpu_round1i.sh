#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 12 gpurun_out/$name.log | cut -c1-400; }
run tests_gpu python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -W ignore
DFVO_CPU_THREADS=32 run bench python bench.py --steps 40 --warmup 5
