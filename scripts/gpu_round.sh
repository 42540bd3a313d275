#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --steps 4 --warmup 3 --workload cfg4 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg2 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg5 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench cfg5 rc=$?"
PLFEM_TIMING=1 PLFEM_HOST_THREADS=1 timeout 300 python scripts/gpu_forest_once.py 12 3 > $O/forest_timing.log 2>&1
PLFEM_SWEEP=levels timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_forest12_levels.csv python scripts/gpu_forest_once.py 12 2 > $O/ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
gzip -f $O/launches_forest12_levels.csv
