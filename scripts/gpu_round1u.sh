#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run tests_pipe python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_liteflow.py -x -q --timeout 600 -p no:cacheprovider -W ignore
run bench1 python bench.py --cpu-frames 0
run bench2 python bench.py --cpu-frames 0
run bench3 python bench.py
run ncu_dram ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv_ --csv --log-file gpurun_out/conv_dram.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0
