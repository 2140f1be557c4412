#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 6 gpurun_out/$name.log; }
run ncu_convtc ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -c 6 -o gpurun_out/prof_convtc -f python scripts/prof_conv.py
run sel python -m pytest tests/test_gpu_selection.py -q --timeout 120 -p no:cacheprovider
