#!/bin/bash
# Sizes added to the fused engine's table: parity against the cuFFT engine, then fused vs cuFFT engine time on the headline objective.
set -o pipefail
out=gpurun_out/r02_new_sizes.txt
: > $out
python -m pytest tests/test_parity_gpu.py -q -x -k "headline_size_vs_cufft" 2>&1 | tail -3 >> $out
for hw in "480 640 32" "600 800 32" "864 1152 16" "2400 3200 4"; do
  set -- $hw
  for be in 2 1; do
    python tools/exp_colvar.py --height $1 --width $2 --batch $3 --iters 20 --reps 4 --backend $be 2>&1 | tail -1 | sed "s/^/$1x$2 batch $3 backend $be /" >> $out
  done
done
cat $out
