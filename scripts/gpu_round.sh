#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --durations=3 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_forest12.csv python scripts/gpu_forest_once.py 12 2 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
python scripts/summarize_launches.py $O/launches_forest12.csv $O/launches_forest12.txt > /dev/null 2>&1
rm -f $O/launches_forest12.csv
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
