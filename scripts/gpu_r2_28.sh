#!/bin/bash
# round 2, GPU call 28: speculative flow network in the libs mirror -- drop-in / mirror tests, the default bench line (e2e_libs)
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc ($name)"; tail -n 1 gpurun_out/$name.log | cut -c1-200; return $rc; }
run t_dropin 400 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_mirror.py -q -p no:cacheprovider -W ignore -x || exit 1
run bench_vo 900 python bench.py || exit 1
python - <<'PY'
import json
for l in open('gpurun_out/bench_vo.log'):
    if l.startswith('{"metric"'):
        d=json.loads(l); print('value %.1f e2e %.1f lat %.2f libs %.1f (%.2f ms)'%(d['value'],d['e2e']['value'],d['e2e']['latency_ms'],d['e2e_libs']['value'],d['e2e_libs']['ms_per_step']))
PY
