#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --steps 6 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
for w in 6 12; do
timeout 600 python bench.py --steps 4 --warmup 3 --workload cfg4 --workers $w > $O/bench_cfg4_w$w.json 2> $O/bench_cfg4_w$w.err; echo "bench cfg4 w=$w rc=$?"
done
for w in 6 10; do
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg2 --workers $w > $O/bench_cfg2_w$w.json 2> $O/bench_cfg2_w$w.err; echo "bench cfg2 w=$w rc=$?"
done
