#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 2 gpurun_out/$name.log | cut -c1-200; }
run tests_gpu python -m pytest tests -m gpu -x -q --timeout 400 -p no:cacheprovider -W ignore
run bench python bench.py --warmup 3 --steps 60
