#!/bin/bash
O=gpurun_out/run5; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" ADAPM_SYNC_TRACE=1 timeout 500 $TR --nproc-per-node 2 --master-port $((29520 + RANDOM % 200)) bench.py --gpus 2 --steps 150 --warmup 10 --profile > $O/$name.log 2>&1; python scripts/analyze_kernel_trace.py gpurun_out/kernel_trace.rank0.tsv > $O/$name.trace.txt 2>&1; }
run default X=1
run inflight ADAPM_SGNS_INFLIGHT=1
run burst ADAPM_SGNS_INFLIGHT=1 ADAPM_SYNC_SCAN_BLOCKS=2 ADAPM_SYNC_META_BLOCKS=4 ADAPM_SYNC_WORK_BLOCKS=4
run burst8 ADAPM_SGNS_INFLIGHT=1 ADAPM_SYNC_SCAN_BLOCKS=4 ADAPM_SYNC_META_BLOCKS=4 ADAPM_SYNC_WORK_BLOCKS=8
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
python scripts/summarize_bench_logs.py $O
for n in default inflight burst burst8; do echo "== $n"; head -22 $O/$n.trace.txt | cut -c1-150; done
