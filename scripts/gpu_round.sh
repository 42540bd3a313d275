#!/bin/bash
O=gpurun_out; mkdir -p $O
grep -m1 "model name" /proc/cpuinfo > $O/cpu.txt; nproc >> $O/cpu.txt
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
PLFEM_TIMING=1 PLFEM_HOST_THREADS=1 timeout 300 python scripts/gpu_forest_once.py 12 3 > $O/forest_timing.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
# four host cores per GPU (what a rank gets on an 8-GPU node with 32 cores), forests in flight swept
for w in 6 9 12; do
  taskset -c 0-3 timeout 600 python bench.py --steps 6 --warmup 3 --workers $w > $O/bench_4cores_w$w.json 2> $O/bench_4cores_w$w.err; echo "bench 4 cores w=$w rc=$?"
done
