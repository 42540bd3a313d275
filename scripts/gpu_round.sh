#!/bin/bash
O=gpurun_out; mkdir -p $O
PLFEM_TIMING=1 PLFEM_HOST_THREADS=1 timeout 300 python scripts/gpu_forest_once.py 12 3 > $O/forest_timing.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
