#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_host_tracker.py > gpurun_out/prof_host_tracker.log 2>&1; echo rc=$?
head -70 gpurun_out/prof_host_tracker.log
