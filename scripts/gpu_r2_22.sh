#!/bin/bash
# round 2, GPU call 22: pinned staging for the small host->device uploads -- guarded tests, bench
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc ($name)"; tail -n 2 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_trk 400 python -m pytest tests/test_gpu_depth_pose.py tests/test_gpu_pipeline.py tests/test_gpu_selection.py -q -p no:cacheprovider -W ignore -x || exit 1
B="python bench.py --warmup 3 --steps 96 --no-extras --cpu-frames 0"
run b_def 240 $B || exit 1
DFVO_PIPELINED=0 DFVO_INFLIGHT=2 run b_p0_i2 240 $B
run b_def2 240 $B
python - <<'PY'
import json
for f in ['b_def','b_p0_i2','b_def2']:
    try:
        for l in open('gpurun_out/%s.log'%f):
            if l.startswith('{"metric"'):
                d=json.loads(l); r=d['roofline']
                print('%-10s value %.1f e2e %.1f lat %.2f launches/frame %d  kernel_ms %.3f trk %s'%(f,d['value'],d['e2e']['value'],d['e2e'].get('latency_ms',0),d['gpu_launches']/d['steps'],r['kernel_ms_per_frame'],json.dumps(d['config'].get('tracker_ms_by_branch_and_outliers'))))
    except Exception as e: print(f, e)
PY
timeout 200 python scripts/prof_host_step.py 2>&1 | head -45
