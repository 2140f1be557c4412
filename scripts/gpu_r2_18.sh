#!/bin/bash
# round 2, GPU call 18: source-level ncu profile of the EPnP hypotheses kernel
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_pnp_hypotheses_coop -c 1 -o gpurun_out/prof_pnp_coop -f python scripts/prof_tracker.py 0.0 > gpurun_out/ncu_pnp.log 2>&1; echo rc=$?
tail -3 gpurun_out/ncu_pnp.log; ls -la gpurun_out/prof_pnp_coop.ncu-rep
