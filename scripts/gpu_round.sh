#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 600 python bench.py --steps 6 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
PLFEM_CGS=full timeout 600 python bench.py --steps 6 --warmup 3 > $O/bench_cfg1_cgsfull.json 2> $O/bench_cfg1_cgsfull.err; echo "bench cfg1 cgs full rc=$?"
