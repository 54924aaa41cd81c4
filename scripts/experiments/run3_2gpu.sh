#!/bin/bash
O=gpurun_out/run3; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
CUDA_VISIBLE_DEVICES=0 timeout 400 python -m pytest tests/experimental_one_gpu_procs.py -x -q > $O/pytest_one_gpu_procs.log 2>&1; echo "rc=$?" >> $O/pytest_one_gpu_procs.log
ADAPM_VERBOSE=1 ADAPM_SYNC_TRACE=1 timeout 500 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 10 --profile > $O/b2_default.log 2>&1
python scripts/analyze_kernel_trace.py gpurun_out/kernel_trace.rank0.tsv > $O/trace_rank0.analysis.txt 2>&1
timeout 400 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 200 --warmup 10 --sync-per-sec 300 > $O/b2_sps300.log 2>&1
ADAPM_SYNC_WORK_BLOCKS=2 timeout 400 $TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 200 --warmup 10 > $O/b2_wb2.log 2>&1
ADAPM_SGNS_REGS=128 timeout 400 $TR --nproc-per-node 2 --master-port 29516 bench.py --gpus 2 --steps 200 --warmup 10 > $O/b2_r128.log 2>&1
ADAPM_MULTICAST=0 ADAPM_VMM=0 timeout 400 $TR --nproc-per-node 2 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 10 > $O/b2_ipc.log 2>&1
timeout 400 $TR --nproc-per-node 2 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 > $O/b2_k20.log 2>&1
timeout 400 $TR --nproc-per-node 2 --master-port 29519 bench.py --impl nccl --gpus 2 --steps 20 --warmup 5 > $O/nccl2.log 2>&1
python scripts/summarize_bench_logs.py $O
