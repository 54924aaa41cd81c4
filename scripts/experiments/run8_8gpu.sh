#!/bin/bash
# 8-GPU scaling point(s): the headline bench as the driver runs it (K=20, W=5) and a longer window
O=gpurun_out/run8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; extra=$2; shift; shift; env "$@" timeout 600 $TR --nproc-per-node 8 --master-port $((29520 + RANDOM % 200)) bench.py --gpus 8 $extra > $O/$name.log 2>&1; }
run n8_k20 "--steps 20 --warmup 5" X=1
run n8_k100 "--steps 100 --warmup 10" X=1
run n8_k100_sps300 "--steps 100 --warmup 10 --sync-per-sec 300" X=1
timeout 500 $TR --nproc-per-node 8 --master-port 29619 bench.py --impl nccl --gpus 8 --steps 20 --warmup 5 > $O/nccl8.log 2>&1
python scripts/summarize_bench_logs.py $O
