#!/bin/bash
# One GPU-box session: tests, bench lines, launch list, full ncu capture of the sweeps.  Outputs under gpurun_out/
# (ncu reports stay in /tmp on the box: gpurun copies back at most 64 MiB; their CSV exports are what travels).
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
fi
PLFEM_TIMING=1 timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"
timeout 600 python bench.py --steps 6 --warmup 3 --inflight 24 --workers 4 > $O/bench_cfg1_b24w4.json 2> /dev/null; echo "bench cfg1 24x4 rc=$?"
timeout 600 python bench.py --steps 6 --warmup 3 --inflight 16 --workers 6 > $O/bench_cfg1_b16w6.json 2> /dev/null; echo "bench cfg1 16x6 rc=$?"
timeout 600 python bench.py --steps 6 --warmup 3 --inflight 12 --workers 8 > $O/bench_cfg1_b12w8.json 2> /dev/null; echo "bench cfg1 12x8 rc=$?"
if [ -z "$SKIP_NCU" ]; then
NCU_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
  -k regex:'(forward|backward)_kernel<\(int\)4' -o /tmp/sweeps_full -f python scripts/gpu_sweep_profile.py cfg1 12 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
DESIGNS=12 python scripts/ncu_summary.py /tmp/sweeps_full.ncu-rep $O/ncu_full_forest12_sweeps.csv $O/traffic.json > $O/ncu_summary.log 2>&1
ncu -i /tmp/sweeps_full.ncu-rep --page details > $O/ncu_full_forest12_sweeps_details.txt 2>&1; gzip -f $O/ncu_full_forest12_sweeps_details.txt
NCU_RANGE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'assemble_rows_kernel|element_setup|spmm_b|resid_k' -c 6 -o /tmp/asm_full -f python scripts/gpu_sweep_profile.py cfg1 12 > $O/ncu_asm.log 2>&1; echo "ncu asm rc=$?"
python scripts/ncu_summary.py /tmp/asm_full.ncu-rep $O/ncu_full_forest12_assembly.csv > /dev/null 2>&1
ncu -i /tmp/asm_full.ncu-rep --page details > $O/ncu_full_forest12_assembly_details.txt 2>&1; gzip -f $O/ncu_full_forest12_assembly_details.txt
NCU_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_cfg5_profile.csv \
  python scripts/gpu_sweep_profile.py cfg5 1 > $O/ncu_cfg5.log 2>&1; echo "cfg5 launch list rc=$?"
python scripts/summarize_launches.py $O/launches_cfg5_profile.csv $O/launches_cfg5_profile.txt > /dev/null 2>&1; gzip -f $O/launches_cfg5_profile.csv
fi
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench cfg5 rc=$?"
du -sh $O; ls -la $O
