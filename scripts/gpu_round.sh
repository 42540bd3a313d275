#!/bin/bash
O=gpurun_out; mkdir -p $O
PLFEM_TIMING=1 PLFEM_HOST_THREADS=1 timeout 300 python scripts/gpu_forest_once.py 12 3 > $O/forest_timing.log 2>&1
