#!/bin/bash
# final N = 8 confirmation of the committed build: the driver's setting (K=20, W=5) and a 100-step window
O=gpurun_out/run14; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0 timeout 200 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "sgns_step_matches or native_step" > $O/pytest_sgns.log 2>&1; echo "rc=$?" >> $O/pytest_sgns.log
timeout 300 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 5 > $O/n8_k20.log 2>&1
timeout 300 $TR --nproc-per-node 8 --master-port 29712 bench.py --gpus 8 --steps 100 --warmup 5 > $O/n8_k100.log 2>&1
python scripts/summarize_bench_logs.py $O | grep -v "^    \[rank"
