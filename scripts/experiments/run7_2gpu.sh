#!/bin/bash
O=gpurun_out/run7; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; extra=$2; shift; shift; env "$@" timeout 500 $TR --nproc-per-node 2 --master-port $((29520 + RANDOM % 200)) bench.py --gpus 2 $extra > $O/$name.log 2>&1; }
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "native_step or word2vec" > $O/pytest_native.log 2>&1; echo "rc=$?" >> $O/pytest_native.log
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/n1_native.log 2>&1
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --loop python > $O/n1_python.log 2>&1
run n2_native_k200 "--steps 200 --warmup 10" X=1
run n2_python_k200 "--steps 200 --warmup 10 --loop python" X=1
run n2_native_k20 "--steps 20 --warmup 5" X=1
python scripts/summarize_bench_logs.py $O | grep -v "^    \[rank"
tail -5 $O/pytest_native.log
