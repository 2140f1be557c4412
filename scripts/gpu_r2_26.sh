#!/bin/bash
# round 2, GPU call 26: register-blocked flow-head kernel -- network parity tests at full size, bench with / without it
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc ($name)"; tail -n 2 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_nets 400 python -m pytest tests/test_gpu_liteflow.py tests/test_gpu_parity_fullsize.py -q -p no:cacheprovider -W ignore -x || exit 1
B="python bench.py --warmup 3 --steps 96 --no-extras --cpu-frames 0"
run b_head8 240 $B || exit 1
DFVO_FLOW_HEAD8=0 run b_head1 240 $B
run b_head8b 240 $B
python - <<'PY'
import json
for f in ['b_head8','b_head1','b_head8b']:
    try:
        for l in open('gpurun_out/%s.log'%f):
            if l.startswith('{"metric"'):
                d=json.loads(l); r=d['roofline']
                print('%-10s value %.1f e2e %.1f lat %.2f launches/frame %d  kernel_ms %.3f trk %s'%(f,d['value'],d['e2e']['value'],d['e2e'].get('latency_ms',0),d['gpu_launches']/d['steps'],r['kernel_ms_per_frame'],json.dumps(d['config'].get('tracker_ms_by_branch_and_outliers'))))
    except Exception as e: print(f, e)
PY
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_fullsize.json'))
for m in ['fp32','tf32','bf16']: print(m, d[m]['flow']['epe_mean'], d[m]['flow']['epe_max'])
PY
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_flow_head -c 30 --csv --log-file gpurun_out/launches_heads.csv python scripts/trace_tc.py > /dev/null 2>&1; grep -h "k_flow_head" gpurun_out/launches_heads.csv | awk -F'","' '{print $5, $9, $NF}' | sed 's/(const.*)//' | tail -12
