#!/bin/bash
# Round-2 closing refresh (run under gpurun, one B200) after the 512-thread tiles, the added sizes and the pcg graph-cache fix:
# GPU test suite, smoke, both bench arms of the headline, one bench line per BASELINE configuration.  No profiler in this script.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/r02_gputest_final.log; cat $O/r02_gputest_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/bench_r02_reference.json; cut -c1-200 $O/bench_r02_reference.json; echo
timeout 400 python bench.py 2>/dev/null | tail -1 > $O/bench_r02_headline_n1.json; cut -c1-300 $O/bench_r02_headline_n1.json; echo
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $w 2>/dev/null | tail -1 > $O/bench_r02_${w}_n1.json; cut -c1-200 $O/bench_r02_${w}_n1.json; echo
done
timeout 400 python bench.py --workload cfg2 --denoiser bf16 --skip-cpu 2>/dev/null | tail -1 > $O/bench_r02_cfg2_bf16_n1.json; cut -c1-200 $O/bench_r02_cfg2_bf16_n1.json; echo
