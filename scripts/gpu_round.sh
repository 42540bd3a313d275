#!/bin/bash
# One GPU-box session: tests, bench lines, launch list, full ncu capture of the sweeps.  Outputs under gpurun_out/.
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_forest12.csv python scripts/gpu_forest_once.py 12 2 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
NCU_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'forward|backward' -o $O/sweeps_full -f python scripts/gpu_sweep_profile.py cfg1 12 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O
