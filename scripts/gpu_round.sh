#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
PLFEM_TRACE_FILE=$O/trace_cfg1.bin timeout 300 python scripts/gpu_sweep_profile.py cfg1 12 > $O/trace_cfg1.log 2>&1; echo "trace cfg1 rc=$?"
python scripts/sweep_trace.py $O/trace_cfg1.bin > $O/trace_cfg1.txt 2>&1
PLFEM_TRACE_FILE=$O/trace_cfg1_single.bin timeout 300 python scripts/gpu_sweep_profile.py cfg1 1 > $O/trace_cfg1_single.log 2>&1; echo "trace cfg1 single rc=$?"
python scripts/sweep_trace.py $O/trace_cfg1_single.bin > $O/trace_cfg1_single.txt 2>&1
PLFEM_TRACE_FILE=$O/trace_cfg5.bin timeout 600 python scripts/gpu_sweep_profile.py cfg5 1 > $O/trace_cfg5.log 2>&1; echo "trace cfg5 rc=$?"
python scripts/sweep_trace.py $O/trace_cfg5.bin > $O/trace_cfg5.txt 2>&1
PLFEM_SWEEP=levels timeout 300 python scripts/gpu_sweep_profile.py cfg5 1 > $O/profile_cfg5_levels.log 2>&1
PLFEM_SWEEP=levels timeout 300 python scripts/gpu_sweep_profile.py cfg1 12 > $O/profile_cfg1_levels.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
PLFEM_SWEEP=levels timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1_levels.json 2> /dev/null; echo "bench cfg1 levels rc=$?"
du -sh $O
