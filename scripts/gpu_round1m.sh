#!/bin/bash
# halo-resident conv kernel bring-up: descriptor variants x sub-tile counts on the conv parity tests, then suite/bench/trace
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-300; }
for bo in 0 1; do for S in 1 2 4; do
  DFVO_HALO_BO=$bo DFVO_HALO_S=$S run conv_bo${bo}_S${S} python -m pytest tests/test_gpu_stage_ops.py -k "tcgen05 and not stride2" -q --timeout 200 -p no:cacheprovider -W ignore
done; done
