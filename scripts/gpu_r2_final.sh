#!/bin/bash
# round 2, final GPU call: full GPU suite, smoke, the default bench line (all extras), reference arm, the other configs, conv traces,
# tracker launch lists, compute-sanitizer.  Guarded: stops if the suite fails.
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc ($name)"; tail -n 2 gpurun_out/$name.log | cut -c1-300; return $rc; }
run tests_gpu 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -W ignore -x || exit 1
run smoke 300 python __graft_entry__.py smoke || exit 1
run bench_vo 900 python bench.py || exit 1
run bench_ref 400 python bench.py --impl reference --steps 3 --warmup 1
run bench_corr64 200 python bench.py --config corr64
run bench_ransac10k 200 python bench.py --config ransac10k
run bench_pairs64 300 python bench.py --config pairs64
DFVO_TC_TRACE=1 run trace_convs 200 python scripts/trace_tc.py
DFVO_TC_TRACE=1 DFVO_CONV_CHAIN=1 run trace_convs_chains 200 python scripts/trace_tc.py
for f in 0.0 0.6; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tracker_$f.csv python scripts/prof_tracker.py $f > gpurun_out/ncu_trk_$f.log 2>&1; echo "rc=$? (ncu tracker $f)"
done
run san_memcheck 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_run.py
run san_racecheck 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_run.py
run san_synccheck 400 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/sanitize_run.py
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/san_*.log
