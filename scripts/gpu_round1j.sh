#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 6 gpurun_out/$name.log | cut -c1-300; }
run tests_lf python -m pytest tests/test_gpu_liteflow.py -q --timeout 600 -p no:cacheprovider -W ignore
run bench python bench.py --steps 40 --warmup 5 --cpu-frames 0
run ncu_frame ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_frame.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0
