#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run ncu_halo6 ncu --set full --import-source on --clock-control none -k regex:k_conv_halo --launch-skip 354 --launch-count 6 -f -o gpurun_out/prof_halo_L6 python scripts/trace_tc.py
