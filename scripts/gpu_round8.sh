#!/bin/bash
O=gpurun_out; mkdir -p $O
nproc > $O/n8_nproc.txt; free -g | head -2 >> $O/n8_nproc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"
