#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
for c in 5 3 2 1; do
PLFEM_SWEEP_CTAS_PER_SM=$c timeout 300 python scripts/gpu_sweep_profile.py cfg1 12 2>&1 | tail -2 > $O/pers${c}_cfg1.log
PLFEM_SWEEP_CTAS_PER_SM=$c timeout 600 python bench.py --steps 8 --warmup 3 > $O/bench_pers${c}.json 2> /dev/null; echo "bench $c rc=$?"
done
PLFEM_SWEEP_CTAS_PER_SM=5 timeout 300 python scripts/gpu_sweep_profile.py cfg5 1 2>&1 | tail -2 > $O/pers5_cfg5.log
PLFEM_SWEEP_CTAS_PER_SM=5 timeout 300 python scripts/gpu_sweep_profile.py cfg1 1 2>&1 | tail -2 > $O/pers5_cfg1s.log
