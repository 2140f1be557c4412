#!/bin/bash
# round 2: two ranks under torchrun -- default bench line and the pair-sharded batched mode
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$? (vo x2)"; tail -n 1 gpurun_out/bench_2gpu.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --config pairs64 > gpurun_out/bench_pairs64_2gpu.log 2>&1; echo "rc=$? (pairs64 x2)"; tail -n 1 gpurun_out/bench_pairs64_2gpu.log | cut -c1-400
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.log 2>&1; echo "rc=$? (reference arm x2)"; tail -n 1 gpurun_out/bench_ref_2gpu.log | cut -c1-300
