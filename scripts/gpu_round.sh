#!/bin/bash
# One GPU-box session: tests, bench lines, launch lists, full ncu captures.  Outputs under gpurun_out/ (ncu reports stay in /tmp
# on the box: gpurun copies back at most 64 MiB; their CSV / text exports are what travels and what profiles/ keeps).
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_forest12.csv python scripts/gpu_forest_once.py 12 2 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
python scripts/summarize_launches.py $O/launches_forest12.csv $O/launches_forest12.txt > /dev/null 2>&1
gzip -f $O/launches_forest12.csv
NCU_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
  -k regex:'(forward|backward)_kernel<\(int\)4' -o /tmp/sweeps_full -f python scripts/gpu_sweep_profile.py cfg1 12 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
DESIGNS=12 python scripts/ncu_summary.py /tmp/sweeps_full.ncu-rep $O/ncu_full_forest12_sweeps.csv $O/traffic.json > $O/ncu_summary.log 2>&1
ncu -i /tmp/sweeps_full.ncu-rep --page details > $O/ncu_full_forest12_sweeps_details.txt 2>&1; gzip -f $O/ncu_full_forest12_sweeps_details.txt
NCU_RANGE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'assemble_rows_kernel|element_setup|spmm_b|resid_k|gemm_schur|gemm_w' -c 10 -o /tmp/asm_full -f python scripts/gpu_sweep_profile.py cfg1 12 > $O/ncu_asm.log 2>&1; echo "ncu asm rc=$?"
python scripts/ncu_summary.py /tmp/asm_full.ncu-rep $O/ncu_full_forest12_assembly_spmv_gemm.csv > /dev/null 2>&1
ncu -i /tmp/asm_full.ncu-rep --page details > $O/ncu_full_forest12_assembly_spmv_gemm_details.txt 2>&1; gzip -f $O/ncu_full_forest12_assembly_spmv_gemm_details.txt
NCU_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_cfg5_profile.csv \
  python scripts/gpu_sweep_profile.py cfg5 1 > $O/ncu_cfg5.log 2>&1; echo "cfg5 launch list rc=$?"
python scripts/summarize_launches.py $O/launches_cfg5_profile.csv $O/launches_cfg5_profile.txt > /dev/null 2>&1; gzip -f $O/launches_cfg5_profile.csv
timeout 120 scripts/micro/fp64_peak > $O/fp64_peak.txt 2>&1; echo "fp64 peak rc=$?"
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?"
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench cfg5 rc=$?"
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?"
PLFEM_TRACE_FILE=$O/trace_cfg1.bin timeout 300 python scripts/gpu_sweep_profile.py cfg1 12 > $O/trace_cfg1.log 2>&1
python scripts/sweep_trace.py $O/trace_cfg1.bin > $O/trace_cfg1.txt 2>&1
PLFEM_TRACE_FILE=$O/trace_cfg5.bin timeout 600 python scripts/gpu_sweep_profile.py cfg5 1 > $O/trace_cfg5.log 2>&1
python scripts/sweep_trace.py $O/trace_cfg5.bin > $O/trace_cfg5.txt 2>&1
rm -f $O/trace_cfg5.bin $O/trace_cfg1.bin
du -sh $O; ls $O
