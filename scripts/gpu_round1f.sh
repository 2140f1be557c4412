#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 8 gpurun_out/$name.log; }
run dbg_score python scripts/dbg_score.py
run dropin python -m pytest tests/test_dropin_driver.py -q --timeout 600 -m gpu -p no:cacheprovider -W ignore
