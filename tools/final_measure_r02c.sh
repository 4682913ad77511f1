#!/bin/bash
# Last refresh with the committed build: smoke, both arms of the headline, cfg2 / cfg4 / cfg5 lines (cfg1, cfg3 were taken with this build already).
O=gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/bench_r02_reference.json; cut -c1-120 $O/bench_r02_reference.json; echo
timeout 300 python bench.py 2>/dev/null | tail -1 > $O/bench_r02_headline_n1.json; cut -c1-160 $O/bench_r02_headline_n1.json; echo
for w in cfg4 cfg5 cfg2; do
  timeout 400 python bench.py --workload $w 2>/dev/null | tail -1 > $O/bench_r02_${w}_n1.json; cut -c1-160 $O/bench_r02_${w}_n1.json; echo
done
