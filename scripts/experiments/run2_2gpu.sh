#!/bin/bash
# 2-GPU validation of the device-resident round: gpu tests, 1- and 2-GPU bench, NCCL arm, host-round A/B
O=gpurun_out/run2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/bench1.log 2>&1
timeout 400 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench2_k20.log 2>&1
ADAPM_SYNC_TRACE=1 timeout 500 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 10 --profile > $O/bench2_k200_profile.log 2>&1
python scripts/analyze_kernel_trace.py gpurun_out/kernel_trace.rank0.tsv > $O/trace_rank0.analysis.txt 2>&1
ADAPM_HOST_ROUND=1 timeout 400 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 200 --warmup 10 > $O/bench2_k200_hostround.log 2>&1
timeout 300 python bench.py --impl nccl --gpus 1 --steps 20 --warmup 5 > $O/nccl1.log 2>&1
timeout 400 $TR --nproc-per-node 2 --master-port 29515 bench.py --impl nccl --gpus 2 --steps 20 --warmup 5 > $O/nccl2.log 2>&1
tail -n 3 $O/pytest_gpu.log; for f in bench1 bench2_k20 bench2_k200_profile bench2_k200_hostround nccl1 nccl2; do echo "== $f"; tail -c 600 $O/$f.log | head -c 600; echo; done
