#!/bin/bash
# round 2, GPU call 15: fused E-tracker tail -- guarded tests, then bench {fused on/off} x {tracker thread on/off}
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc ($name)"; tail -n 2 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_tail 200 python -m pytest tests/test_gpu_depth_pose.py -q -p no:cacheprovider -W ignore -x -k "fused or scale_ransac" || exit 1
run t_pipe 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dropin.py -q -p no:cacheprovider -W ignore -x || exit 1
B="python bench.py --warmup 3 --steps 80 --no-extras --cpu-frames 0"
DFVO_TRACKER_THREAD=0 DFVO_FUSED_TAIL=1 run b_f1_t0 240 $B || exit 1
DFVO_TRACKER_THREAD=1 DFVO_FUSED_TAIL=1 run b_f1_t1 240 $B
DFVO_TRACKER_THREAD=0 DFVO_FUSED_TAIL=0 run b_f0_t0 240 $B
DFVO_TRACKER_THREAD=1 DFVO_FUSED_TAIL=1 DFVO_INFLIGHT=3 run b_f1_t1_i3 240 $B
python - <<'PY'
import json
for f in ['b_f1_t0','b_f1_t1','b_f0_t0','b_f1_t1_i3']:
    try:
        for l in open('gpurun_out/%s.log'%f):
            if l.startswith('{"metric"'):
                d=json.loads(l); r=d['roofline']
                print('%-12s value %.1f e2e %.1f lat %.2f launches/frame %d  kernel_ms %.3f trk %s'%(f,d['value'],d['e2e']['value'],d['e2e'].get('latency_ms',0),d['gpu_launches']/d['steps'],r['kernel_ms_per_frame'],json.dumps(d['config'].get('tracker_ms_by_branch_and_outliers'))))
    except Exception as e: print(f, e)
PY
timeout 200 python scripts/prof_host_tracker.py > gpurun_out/prof_host_tracker.log 2>&1; head -8 gpurun_out/prof_host_tracker.log
