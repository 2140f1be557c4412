#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_host_step.py > gpurun_out/prof_host_step.log 2>&1; echo rc=$?
head -120 gpurun_out/prof_host_step.log
