#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 5 gpurun_out/$name.log | cut -c1-400; }
run tests_gpu python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider -W ignore
run bench python bench.py
DFVO_STEM_WINDOW=0 run bench_nowin python bench.py --cpu-frames 0
run ncu_frame ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_frame.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0
