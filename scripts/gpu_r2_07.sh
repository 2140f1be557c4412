#!/bin/bash
# round 2, GPU call 7: full GPU suite (chains now opt-in + bit-equality test, sliced corr kernel, batched consistency), corr64 bench,
# compute-sanitizer memcheck / racecheck / synccheck over scripts/sanitize_run.py
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run tests_gpu 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -W ignore -x
run corr64 300 python bench.py --config corr64
run san_memcheck 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_run.py
run san_racecheck 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_run.py
run san_synccheck 600 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/sanitize_run.py
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|Error:" gpurun_out/san_*.log | sort | uniq -c | head -20
