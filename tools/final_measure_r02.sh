#!/bin/bash
# Round-2 evidence (run under gpurun, one B200): GPU test suite, both bench arms of the headline, one bench line per BASELINE
# configuration, the ncu launch list of the default bench command and --set full captures of the dominant kernels.
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/r02_gputest_final.log; cat $O/r02_gputest_final.log
timeout 400 python bench.py 2>/dev/null | tail -1 > $O/bench_r02_headline_n1.json; cut -c1-400 $O/bench_r02_headline_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/bench_r02_reference.json; cut -c1-300 $O/bench_r02_reference.json
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $w 2>/dev/null | tail -1 > $O/bench_r02_${w}_n1.json; cut -c1-260 $O/bench_r02_${w}_n1.json; echo
done
timeout 400 python bench.py --workload cfg2 --denoiser bf16 --skip-cpu 2>/dev/null | tail -1 > $O/bench_r02_cfg2_bf16_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_final_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_col|k_rowz_mid" -s 6 -c 2 -o $O/r02_pairs_full \
    python bench.py --batch 8 --steps 1 --warmup 1 --iters 6 --skip-cpu --skip-e2e > /dev/null 2>&1
DPX_TRACE=$O/r02_trace.bin python tools/exp_colvar.py --vars=- --reps 1 2>&1 | tail -1
ls -la $O | tail -12
