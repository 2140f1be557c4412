#!/bin/bash
# round 2, GPU call 9: lane-cooperative EPnP with group-masked shuffles -- guarded steps, stop at the first failure
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_epnp 150 python -m pytest tests/test_gpu_depth_pose.py -q -p no:cacheprovider -W ignore -x -k "epnp" || exit 1
run t_pnp 300 python -m pytest tests/test_gpu_depth_pose.py -q -p no:cacheprovider -W ignore -x -k "pnp or monodepth" || exit 1
run tests_gpu 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -W ignore -x || exit 1
run bench 420 python bench.py --warmup 3 --steps 60 --no-extras --cpu-frames 0 || exit 1
echo "=== ncu tracker"; timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tracker_60.csv python scripts/prof_tracker.py 0.6 > gpurun_out/ncu_trk.log 2>&1; echo rc=$?
python - <<'PY'
import json
for f in ['bench']:
    for l in open('gpurun_out/%s.log'%f):
        if l.startswith('{"metric"'):
            d=json.loads(l); r=d['roofline']
            print('%-18s value %.1f e2e %.1f  launches/frame %d  kernel_ms %.3f frac %.3f trk %s'%(f,d['value'],d['e2e']['value'],d['gpu_launches']/d['steps'],r['kernel_ms_per_frame'],r['frac'],json.dumps(d['config'].get('tracker_ms_by_branch_and_outliers'))))
PY
grep -h "k_pnp" gpurun_out/launches_tracker_60.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' | sort | uniq -c | head
