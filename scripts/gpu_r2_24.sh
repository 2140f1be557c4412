#!/bin/bash
# round 2, GPU call 24: layer chains again, now that the tracker no longer bounds the pipeline
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc ($name)"; return $rc; }
B="python bench.py --warmup 3 --steps 96 --no-extras --cpu-frames 0"
DFVO_CONV_CHAIN=1 run c_all 240 $B
DFVO_CONV_CHAIN=1 DFVO_CHAIN_MAX_PIXELS=14000 run c_coarse 240 $B
DFVO_CONV_CHAIN=0 run c_off 240 $B
python - <<'PY'
import json
for f in ['c_all','c_coarse','c_off']:
    try:
        for l in open('gpurun_out/%s.log'%f):
            if l.startswith('{"metric"'):
                d=json.loads(l); r=d['roofline']
                print('%-10s value %.1f e2e %.1f lat %.2f launches/frame %d  kernel_ms %.3f frac %.3f trk %s'%(f,d['value'],d['e2e']['value'],d['e2e'].get('latency_ms',0),d['gpu_launches']/d['steps'],r['kernel_ms_per_frame'],r['frac'],json.dumps(d['config'].get('tracker_ms_by_branch_and_outliers'))))
    except Exception as e: print(f, e)
PY
