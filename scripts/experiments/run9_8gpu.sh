#!/bin/bash
# scaling points as the driver runs them (K=20, W=5) at N=4 and N=8 + the multi-GPU contract tests on 4 GPUs
O=gpurun_out/run9; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi4.log 2>&1; echo "rc=$?" >> $O/pytest_multi4.log
timeout 400 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > $O/n8_k20.log 2>&1
timeout 400 $TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --steps 20 --warmup 5 > $O/n4_k20.log 2>&1
timeout 400 $TR --nproc-per-node 4 --master-port 29613 bench.py --impl nccl --gpus 4 --steps 20 --warmup 5 > $O/nccl4.log 2>&1
python scripts/summarize_bench_logs.py $O | grep -v "^    \[rank"
