#!/bin/bash
# round 2, GPU call 29: DRAM traffic of the tensor-core convolution launches (roofline.traffic), then the full GPU suite once more
mkdir -p gpurun_out
DFVO_OVERLAP=0 timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv_ --csv --log-file gpurun_out/conv_dram.csv python bench.py --steps 2 --warmup 3 --no-extras --cpu-frames 0 > gpurun_out/ncu_dram.log 2>&1; echo "rc=$? (ncu dram)"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/conv_dram.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
ids=set(r[0] for r in rows[hdr+1:] if len(r)>5)
print('conv launches profiled:', len(ids))
PY
timeout 600 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -W ignore -x > gpurun_out/tests_gpu.log 2>&1; echo "rc=$? (tests)"; tail -1 gpurun_out/tests_gpu.log
