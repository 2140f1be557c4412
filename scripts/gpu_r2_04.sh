#!/bin/bash
# round 2, GPU call 4: lean conv epilogue (FADD2/FMUL2, no per-chunk overhead), warp-cooperative 5-point hypotheses, parallel homography
# subset stream -- GPU suite, bench, conv trace, tracker launch list at 60 % outliers
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run tests_gpu 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -W ignore
run bench 600 python bench.py --warmup 3 --steps 60
DFVO_TC_TRACE=1 run trace 300 python scripts/trace_tc.py
run ncu_trk 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tracker_60.csv python scripts/prof_tracker.py 0.6
run ncu_trk0 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tracker_00.csv python scripts/prof_tracker.py 0.0
