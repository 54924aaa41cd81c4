#!/bin/bash
O=gpurun_out/run6; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
runp() { name=$1; shift; env "$@" ADAPM_SYNC_TRACE=1 timeout 500 $TR --nproc-per-node 2 --master-port $((29520 + RANDOM % 200)) bench.py --gpus 2 --steps 150 --warmup 10 --profile > $O/$name.log 2>&1; python scripts/analyze_kernel_trace.py gpurun_out/kernel_trace.rank0.tsv > $O/$name.trace.txt 2>&1; }
run() { name=$1; extra=$2; shift; shift; env "$@" timeout 500 $TR --nproc-per-node 2 --master-port $((29520 + RANDOM % 200)) bench.py --gpus 2 --steps 200 --warmup 10 $extra > $O/$name.log 2>&1; }
runp default X=1
runp burst ADAPM_SYNC_SCAN_BLOCKS=2 ADAPM_SYNC_META_BLOCKS=4 ADAPM_SYNC_WORK_BLOCKS=4
run burst_inflight "" ADAPM_SGNS_INFLIGHT=1 ADAPM_SYNC_SCAN_BLOCKS=2 ADAPM_SYNC_META_BLOCKS=4 ADAPM_SYNC_WORK_BLOCKS=4
run burst_sps300 "--sync-per-sec 300" ADAPM_SGNS_INFLIGHT=1 ADAPM_SYNC_SCAN_BLOCKS=2 ADAPM_SYNC_META_BLOCKS=4 ADAPM_SYNC_WORK_BLOCKS=4
run default_sps300 "--sync-per-sec 300" ADAPM_SGNS_INFLIGHT=1
run burst_k20 "" ADAPM_SGNS_INFLIGHT=1 ADAPM_SYNC_SCAN_BLOCKS=2 ADAPM_SYNC_META_BLOCKS=4 ADAPM_SYNC_WORK_BLOCKS=4
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
python scripts/summarize_bench_logs.py $O | grep -v "^    \[rank"
for n in default burst; do echo "== $n"; head -24 $O/$n.trace.txt | cut -c1-150 | grep -v "commit\|resolve"; done
