#!/bin/bash
# round 2, GPU call 11: diagnostic -- networks-only ceiling of the frame pipeline (tracker replaced by an identity pose), 2 and 1 engines
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc ($name)"; tail -n 2 gpurun_out/$name.log | cut -c1-400; return $rc; }
B="python bench.py --warmup 3 --steps 80 --no-extras --cpu-frames 0"
DFVO_INFLIGHT=2 run d_notrack2 240 $B --diag-no-track
DFVO_INFLIGHT=1 run d_notrack1 240 $B --diag-no-track
DFVO_INFLIGHT=3 run d_notrack3 240 $B --diag-no-track
