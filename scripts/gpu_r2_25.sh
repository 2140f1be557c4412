#!/bin/bash
# round 2, GPU call 25: ncu launch list of whole frames (in-order pipeline so that one frame's kernels are contiguous)
mkdir -p gpurun_out
DFVO_OVERLAP=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2600 -c 800 --csv --log-file gpurun_out/launches_frame.csv python bench.py --steps 6 --warmup 4 --no-extras --cpu-frames 0 > gpurun_out/ncu_frame.log 2>&1; echo rc=$?
tail -2 gpurun_out/ncu_frame.log | cut -c1-200; wc -l gpurun_out/launches_frame.csv
