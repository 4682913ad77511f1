#!/bin/bash
# Programmatic dependent launches on by default: full GPU suite, then A/B per size (DPX_PDL=0 = plain launches) and the small bench lines.
O=gpurun_out/r02_pdl_ab.txt
: > $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2 >> $O
for s in "256 2" "256 8" "512 8" "1024 8" "2048 8"; do set -- $s
  python tools/exp_colvar.py --vars=DPX_PDL=0,-,DPX_PDL=0,- --size $1 --batch $2 --reps 8 2>&1 | tail -4 | sed "s/^/admm $1 x$2 /" >> $O
done
python tools/exp_colvar.py --vars=DPX_PDL=0,-,DPX_PDL=0,- --size 1024 --method hqs --iters 24 --reps 8 2>&1 | tail -4 | sed "s/^/hqs 1024 x8 /" >> $O
cat $O
for w in cfg1 cfg4; do python bench.py --workload $w 2>/dev/null | tail -1 > gpurun_out/bench_r02_${w}_n1.json; cut -c1-150 gpurun_out/bench_r02_${w}_n1.json; echo; done
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_r02_headline_n1.json; cut -c1-150 gpurun_out/bench_r02_headline_n1.json
