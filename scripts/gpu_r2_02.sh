#!/bin/bash
# round 2, GPU call 2: GPU suite with the TMA-store epilogue (UTMASTG) + in-kernel phase stamps of the conv kernel, bench,
# launch list of the tracker kernels at 60 % outliers
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run tests_gpu 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -W ignore
run bench 600 python bench.py --warmup 3 --steps 60
DFVO_TC_TRACE=1 run trace 300 python scripts/trace_tc.py
DFVO_HALO_DBG=1 DFVO_GRAPHS=0 run halo_dbg 300 python scripts/trace_tc.py
DFVO_TMA_STORE=0 DFVO_TC_TRACE=1 run trace_notma 300 python scripts/trace_tc.py
run ncu_trk 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tracker_60.csv python scripts/prof_tracker.py 0.6
