#!/bin/bash
# N = 8 at the driver's setting (K=20, W=5) with the round cadence floor (sys.sync.min_clocks=8, the new default)
O=gpurun_out/run17; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29721 bench.py --gpus 8 --steps 20 --warmup 5 > $O/n8_k20_floor8.log 2>&1
python scripts/summarize_bench_logs.py $O | grep -v "^    \[rank"
