#!/bin/bash
# round 2, GPU call 8: lane-cooperative EPnP, monodepth2 stem on the tensor cores (two-source window conv), single-slice correlation
# variant; full GPU suite, bench (all extras), corr64, tracker launch list, conv trace
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run tests_gpu 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -W ignore -x
run bench 900 python bench.py --warmup 3 --steps 60
DFVO_MONO_STEM_TC=0 run bench_nostemtc 300 python bench.py --warmup 3 --steps 60 --no-extras --cpu-frames 0
DFVO_PNP_COOP=0 run bench_nopnpcoop 300 python bench.py --warmup 3 --steps 60 --no-extras --cpu-frames 0
run corr64 300 python bench.py --config corr64
echo "=== ncu tracker"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tracker_60.csv python scripts/prof_tracker.py 0.6 > gpurun_out/ncu_trk.log 2>&1; echo rc=$?
DFVO_TC_TRACE=1 run trace 300 python scripts/trace_tc.py
python - <<'PY'
import json
for f in ['bench','bench_nostemtc','bench_nopnpcoop']:
    for l in open('gpurun_out/%s.log'%f):
        if l.startswith('{"metric"'):
            d=json.loads(l); r=d['roofline']
            print('%-18s value %.1f e2e %.1f  launches/frame %d  kernel_ms %.3f frac %.3f trk %s'%(f,d['value'],d['e2e']['value'],d['gpu_launches']/d['steps'],r['kernel_ms_per_frame'],r['frac'],json.dumps(d['config'].get('tracker_ms_by_branch_and_outliers'))))
PY
grep -h "k_pnp\|k_h_" gpurun_out/launches_tracker_60.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' | sort | uniq -c | sort -k3 -n -r | head -20
