#!/bin/bash
# round 2, GPU call 19: ncu launch lists of the fused tracker at 0 / 30 / 60 % outliers
mkdir -p gpurun_out
for f in 0.0 0.3 0.6; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tracker_$f.csv python scripts/prof_tracker.py $f > gpurun_out/ncu_trk_$f.log 2>&1; echo "rc=$? ($f)"; tail -1 gpurun_out/ncu_trk_$f.log
done
