#!/bin/bash
# round 2, GPU call 6: fused warp + correlation on mma.sync; fps matrix over {chains: all / coarse levels only / off} x {1, 2 engines in flight}
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 2 gpurun_out/$name.log | cut -c1-200; }
run tests_nets 600 python -m pytest tests/test_gpu_liteflow.py tests/test_gpu_stage_ops.py -q --timeout 300 -p no:cacheprovider -W ignore -x -k "liteflow or correlation"
B="python bench.py --warmup 3 --steps 60 --no-extras --cpu-frames 0"
DFVO_INFLIGHT=2 run m_all_2 300 $B
DFVO_INFLIGHT=1 run m_all_1 300 $B
DFVO_CHAIN_MAX_PIXELS=14000 DFVO_INFLIGHT=2 run m_coarse_2 300 $B
DFVO_CHAIN_MAX_PIXELS=14000 DFVO_INFLIGHT=1 run m_coarse_1 300 $B
DFVO_CONV_CHAIN=0 DFVO_INFLIGHT=2 run m_off_2 300 $B
DFVO_CONV_CHAIN=0 DFVO_INFLIGHT=1 run m_off_1 300 $B
DFVO_CONV_CHAIN=0 DFVO_CORR_MMA=0 DFVO_INFLIGHT=2 run m_off_2_oldcorr 300 $B
run corr64 300 python bench.py --config corr64
for f in m_all_2 m_all_1 m_coarse_2 m_coarse_1 m_off_2 m_off_1 m_off_2_oldcorr; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
for l in open('gpurun_out/%s.log'%f):
    if l.startswith('{"metric"'):
        d=json.loads(l); r=d['roofline']
        print('%-18s value %.1f e2e %.1f  launches/frame %d  conv %s kernel_ms %.3f frac %.3f'%(f,d['value'],d['e2e']['value'],d['gpu_launches']/d['steps'],r['kernel'][-24:-1],r['kernel_ms_per_frame'],r['frac']))
PY
done
