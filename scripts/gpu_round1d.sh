#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 6 gpurun_out/$name.log; }
run stage_tc python -m pytest tests/test_gpu_stage_ops.py -q --timeout 120 -k "tcgen05" -p no:cacheprovider
run liteflow python -m pytest tests/test_gpu_liteflow.py -q --timeout 300 -p no:cacheprovider
QB_PRECS=bf16 QB_ITERS=20 run quick_bench python scripts/quick_bench.py
QB_PRECS=bf16 QB_ITERS=1 QB_WARM=1 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16_b.csv python scripts/quick_bench.py
