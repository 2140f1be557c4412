#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 12 gpurun_out/$name.log; }
run depth_pose python -m pytest tests/test_gpu_depth_pose.py -q --timeout 300 -p no:cacheprovider
run selection python -m pytest tests/test_gpu_selection.py -q --timeout 120 -p no:cacheprovider
