#!/bin/bash
# round 2, GPU call 1: whole GPU suite (new: parity at full size, drop-in on the device, mirror, tf32 stage tests), default
# bench line with extras, per-launch conv trace (baseline of this round)
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run tests_gpu 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -W ignore -x --deselect tests/test_gpu_parity_fullsize.py
run tests_parity 600 python -m pytest tests/test_gpu_parity_fullsize.py -q --timeout 600 -p no:cacheprovider -W ignore -s
run bench 600 python bench.py --warmup 3 --steps 60
DFVO_TC_TRACE=1 run trace 300 python scripts/trace_tc.py
