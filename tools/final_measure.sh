#!/bin/bash
# round-end evidence (run under gpurun): GPU test suite, both bench arms, the ncu launch list of the default bench command
# and one --set full capture of the two dominant kernels.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r01_final.json; cat gpurun_out/bench_r01_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r01_reference.json
cut -c1-300 gpurun_out/bench_r01_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_final_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_col|k_rowz_mid" -s 6 -c 2 -o gpurun_out/r01_v8_full \
    python bench.py --batch 8 --steps 1 --warmup 1 --iters 6 --skip-cpu --skip-e2e > /dev/null 2>&1
ls -la gpurun_out | tail -4
