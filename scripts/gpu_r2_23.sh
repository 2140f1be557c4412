#!/bin/bash
# round 2, GPU call 23: racecheck after giving every lane its own shared-memory column in the EPnP SVD; pnp tests; bench line
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc ($name)"; tail -n 2 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_pnp 300 python -m pytest tests/test_gpu_depth_pose.py -q -p no:cacheprovider -W ignore -x || exit 1
run san_racecheck 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_run.py
run bench_vo 900 python bench.py
