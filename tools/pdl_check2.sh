#!/bin/bash
# Dependent launches on every row kernel (half-spectrum engine, non-persistent pair kernels): GPU suite + A/B on odd batches and cfg1.
O=gpurun_out/r02_pdl_ab_rows.txt
: > $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2 >> $O
for s in "256 1" "512 1" "1024 3"; do set -- $s
  python tools/exp_colvar.py --vars=DPX_PDL=0,-,DPX_PDL=0,- --size $1 --batch $2 --reps 8 2>&1 | tail -4 | cut -c1-200 | sed "s/^/admm $1 x$2 /" >> $O
done
for p in 0 1; do DPX_PDL=$p python bench.py --workload cfg1 --skip-cpu 2>/dev/null | tail -1 | grep -o '"value": [0-9.]*' | head -1 | sed "s/^/cfg1 DPX_PDL=$p /" >> $O; done
cat $O
python bench.py --workload cfg1 2>/dev/null | tail -1 > gpurun_out/bench_r02_cfg1_n1.json
