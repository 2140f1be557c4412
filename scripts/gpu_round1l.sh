#!/bin/bash
# re-entry baseline: GPU suite + bench + per-frame launch list + ncu --set full of the biggest k_conv_tc launches
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 6 gpurun_out/$name.log | cut -c1-600; }
run tests_gpu python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider -W ignore
run bench python bench.py --steps 40 --warmup 5
run ncu_frame ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_frame.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0
run ncu_convtc ncu --set full --import-source on --clock-control none -k regex:k_conv_tc -s 330 -c 12 -f -o gpurun_out/prof_convtc python bench.py --steps 1 --warmup 3 --cpu-frames 0
run smoke python __graft_entry__.py smoke
