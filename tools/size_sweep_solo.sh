#!/bin/bash
# Sizes whose tiles fit one per SM (512-thread CTAs): parity tests, then the headline objective at those sizes (not bench lines).
set -o pipefail
out=gpurun_out/r02_solo_tiles.txt
: > $out
python -m pytest tests/test_parity_gpu.py -q -x -k "4096 or headline_size_vs_cufft" 2>&1 | tail -3 >> $out
for hw in "4096 4096 8" "3072 3072 8" "2560 2560 8" "2160 3840 8" "1440 2560 8" "2048 2048 8"; do
  set -- $hw
  python tools/exp_colvar.py --height $1 --width $2 --batch $3 --iters 20 --reps 4 2>&1 | tail -1 | sed "s/^/$1x$2 /" >> $out
done
cat $out
