#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 6 gpurun_out/$name.log; }
run smoke python __graft_entry__.py smoke
run depth_pose python -m pytest tests/test_gpu_depth_pose.py -q --timeout 300 -p no:cacheprovider
run bench python bench.py --steps 30 --warmup 5
