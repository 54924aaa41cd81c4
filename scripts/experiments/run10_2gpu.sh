#!/bin/bash
# A/B of the round's row pass: TMA-engine variant (default) vs register variant, N=2, with the kernel timeline
O=gpurun_out/run10; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
runp() { name=$1; shift; env "$@" ADAPM_SYNC_TRACE=1 timeout 400 $TR --nproc-per-node 2 --master-port $((29520 + RANDOM % 200)) bench.py --gpus 2 --steps 150 --warmup 10 --profile > $O/$name.log 2>&1; python scripts/analyze_kernel_trace.py gpurun_out/kernel_trace.rank0.tsv > $O/$name.trace.txt 2>&1; }
run() { name=$1; extra=$2; shift; shift; env "$@" timeout 400 $TR --nproc-per-node 2 --master-port $((29520 + RANDOM % 200)) bench.py --gpus 2 $extra > $O/$name.log 2>&1; }
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
runp tma X=1
runp reg ADAPM_ROW_TMA=0
run tma_k20 "--steps 20 --warmup 5" X=1
run reg_k20 "--steps 20 --warmup 5" ADAPM_ROW_TMA=0
run tma_k200 "--steps 200 --warmup 10" X=1
CUDA_VISIBLE_DEVICES=0 timeout 300 python benchmarks/sgns_shared_bench.py > $O/sgns_shared_bench.log 2>&1
python scripts/summarize_bench_logs.py $O | grep -v "^    \[rank"
tail -4 $O/pytest_gpu.log | cut -c1-300; tail -2 $O/sgns_shared_bench.log | cut -c1-1500
for n in tma reg; do echo "== $n"; head -24 $O/$n.trace.txt | cut -c1-150 | grep -v "commit\|resolve"; done
