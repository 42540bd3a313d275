#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x --durations=3 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 300 python scripts/gpu_sweep_profile.py cfg1 12 2>&1 | tail -4 > $O/p_cfg1.log
timeout 300 python scripts/gpu_sweep_profile.py cfg1 1 2>&1 | tail -2 > $O/p_cfg1s.log
timeout 300 python scripts/gpu_sweep_profile.py cfg5 1 2>&1 | tail -2 > $O/p_cfg5.log
timeout 300 python scripts/gpu_sweep_profile.py cfg2 12 2>&1 | tail -2 > $O/p_cfg2.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
