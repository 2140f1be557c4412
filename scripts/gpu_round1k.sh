#!/bin/bash
# full GPU suite + bench + launch list + conv_tc DRAM traffic + source-level profile of the 5-point kernel
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
run tests_gpu python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider -W ignore
run bench python bench.py --steps 40 --warmup 5
run ncu_frame ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_frame.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0
run ncu_dram ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv_tc --csv --log-file gpurun_out/convtc_dram.csv python bench.py --steps 1 --warmup 3 --cpu-frames 0
run ncu_hyp ncu --set full --import-source on --clock-control none -k regex:k_hypotheses -c 1 -f -o gpurun_out/prof_hyp python bench.py --steps 1 --warmup 3 --cpu-frames 0
echo "=== trace_tc"; DFVO_TC_TRACE=1 timeout 600 python scripts/trace_tc.py > gpurun_out/trace_tc.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/trace_tc.log
