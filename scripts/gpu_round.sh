#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
# four host cores per GPU (what a rank gets on an 8-GPU node with 32 cores): forests in flight chosen by batch.default_workers()
taskset -c 0-3 timeout 600 python bench.py --steps 6 --warmup 3 > $O/bench_4cores_auto.json 2> $O/bench_4cores_auto.err; echo "bench 4 cores auto rc=$?"
taskset -c 0-3 timeout 600 python bench.py --steps 6 --warmup 3 --workers 12 > $O/bench_4cores_w12.json 2> $O/bench_4cores_w12.err; echo "bench 4 cores w12 rc=$?"
timeout 600 python bench.py --steps 6 --warmup 3 --workload cfg4 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?"
timeout 600 python bench.py --steps 4 --warmup 3 --workload cfg2 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg5 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench cfg5 rc=$?"
PLFEM_TIMING=1 PLFEM_HOST_THREADS=1 timeout 300 python scripts/gpu_forest_once.py 12 3 > $O/forest_timing.log 2>&1
