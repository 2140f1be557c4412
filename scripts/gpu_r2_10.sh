#!/bin/bash
# round 2, GPU call 10: one-warp-per-sample EPnP with the rsqrt rotation -- guarded steps, stop at the first failure; then 3 engines in flight
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_pnp 200 python -m pytest tests/test_gpu_depth_pose.py -q -p no:cacheprovider -W ignore -x -k "pnp" || exit 1
echo "=== ncu tracker"; timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tracker_60.csv python scripts/prof_tracker.py 0.6 > gpurun_out/ncu_trk.log 2>&1; echo rc=$?
grep -h "k_pnp" gpurun_out/launches_tracker_60.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' | sort | uniq -c | head
B="python bench.py --warmup 3 --steps 80 --no-extras --cpu-frames 0"
DFVO_INFLIGHT=2 run b_in2 300 $B || exit 1
DFVO_INFLIGHT=3 run b_in3 300 $B
DFVO_INFLIGHT=2 run b_in2b 300 $B
python - <<'PY'
import json
for f in ['b_in2','b_in3','b_in2b']:
    try:
        for l in open('gpurun_out/%s.log'%f):
            if l.startswith('{"metric"'):
                d=json.loads(l); r=d['roofline']
                print('%-10s value %.1f e2e %.1f lat %.2f launches/frame %d  kernel_ms %.3f frac %.3f trk %s'%(f,d['value'],d['e2e']['value'],d['e2e'].get('latency_ms',0),d['gpu_launches']/d['steps'],r['kernel_ms_per_frame'],r['frac'],json.dumps(d['config'].get('tracker_ms_by_branch_and_outliers'))))
    except Exception as e: print(f, e)
PY
