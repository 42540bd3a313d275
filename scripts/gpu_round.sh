#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
taskset -c 0-3 timeout 600 python bench.py --steps 6 --warmup 3 > $O/bench_4cores_auto.json 2> $O/bench_4cores_auto.err; echo "bench 4 cores auto rc=$?"
taskset -c 0-3 timeout 600 python bench.py --steps 6 --warmup 3 > $O/bench_4cores_auto_b.json 2> $O/bench_4cores_auto_b.err; echo "bench 4 cores auto (repeat) rc=$?"
