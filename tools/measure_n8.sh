#!/bin/bash
# 8-GPU bench lines (run under `gpurun --gpus 8`): headline, BASELINE config 4 (64 problems over 8 GPUs) and config 5
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
timeout 400 $R bench.py --gpus 8 --skip-cpu 2>/dev/null | tail -1 > gpurun_out/bench_r02_headline_n8.json
timeout 400 $R bench.py --gpus 8 --workload cfg4 --skip-cpu 2>/dev/null | tail -1 > gpurun_out/bench_r02_cfg4_n8.json
timeout 400 $R bench.py --gpus 8 --workload cfg5 --skip-cpu 2>/dev/null | tail -1 > gpurun_out/bench_r02_cfg5_n8.json
for f in headline cfg4 cfg5; do cut -c1-220 gpurun_out/bench_r02_${f}_n8.json; echo; done
