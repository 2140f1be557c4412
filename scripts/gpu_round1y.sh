#!/bin/bash
# final evidence of the round: launch list of a frame, ncu --set full of the level-2 halo convs
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run ncu_frame ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_frame.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0
run ncu_halo ncu --set full --import-source on --clock-control none -k regex:k_conv_halo --launch-skip 427 --launch-count 17 -f -o gpurun_out/prof_halo_L2 python scripts/trace_tc.py
