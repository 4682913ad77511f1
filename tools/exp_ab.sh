#!/bin/bash
# A/B helper (run under gpurun): fused-engine parity tests, then the resident-input number of the headline workload
mkdir -p gpurun_out
for v in "DPX_PAIRS=1"; do
env $v timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "fused or fft_backend or engine or headline" 2>&1 | tail -2
echo "== $v" | tee -a gpurun_out/exp_ab.log
env $v timeout 200 python bench.py --batch 8 --steps 3 --warmup 3 --skip-cpu --skip-e2e 2>&1 | tail -1 | tee -a gpurun_out/exp_ab.log
done
DPX_PAIRS=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_col|k_row" -s 4 -c 2 --csv --log-file gpurun_out/exp_ab_ncu.csv python bench.py --batch 4 --steps 1 --warmup 1 --iters 10 --skip-cpu --skip-e2e > /dev/null 2>&1
python - <<'P'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/exp_ab_ncu.csv')) if len(r)>5]
h=rows[0]; d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[h.index('ID')],r[h.index('Kernel Name')][:24]),{})[r[h.index('Metric Name')].split('.')[0][-28:]]=r[h.index('Metric Value')]
for k,v in d.items(): print(k,v)
P
