#!/bin/bash
# end-of-round validation: GPU suite, smoke, both bench arms
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-400; }
run tests_gpu python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider -W ignore
run smoke python __graft_entry__.py smoke
run bench_ref python bench.py --impl reference
run bench python bench.py
