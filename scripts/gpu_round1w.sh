#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-400; }
run tests_gpu python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider -W ignore
run bench1 python bench.py
run bench2 python bench.py --cpu-frames 0
run prof_host python scripts/prof_host.py
