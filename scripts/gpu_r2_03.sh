#!/bin/bash
# round 2, GPU call 3: ncu --set full of epilogue-bound conv launches (direct-store and TMA-store epilogues), source-level stalls
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
DFVO_TMA_STORE=0 run ncu_direct 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 1 -c 1 -o gpurun_out/prof_halo_direct_1x1 -f python scripts/prof_conv.py 3
DFVO_TMA_STORE=0 run ncu_direct2 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 1 -c 1 -o gpurun_out/prof_halo_direct_3x3 -f python scripts/prof_conv.py 4
DFVO_TMA_STORE=1 run ncu_tma 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 1 -c 1 -o gpurun_out/prof_halo_tma_1x1 -f python scripts/prof_conv.py 3
ls -la gpurun_out/*.ncu-rep
