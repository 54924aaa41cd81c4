#!/bin/bash
O=gpurun_out/run4; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" ADAPM_SYNC_TRACE=1 timeout 500 $TR --nproc-per-node 2 --master-port $((29520 + RANDOM % 200)) bench.py --gpus 2 --steps 150 --warmup 10 --profile > $O/$name.log 2>&1; python scripts/analyze_kernel_trace.py gpurun_out/kernel_trace.rank0.tsv > $O/$name.trace.txt 2>&1; cp gpurun_out/kernel_trace.rank0.tsv.counts $O/$name.counts 2>/dev/null; }
run default X=1
run inflight ADAPM_SGNS_INFLIGHT=1
run wb2mb2 ADAPM_SYNC_WORK_BLOCKS=2 ADAPM_SYNC_META_BLOCKS=2
CUDA_VISIBLE_DEVICES=0 timeout 400 python -m pytest tests/test_gpu_device_round.py tests/test_gpu_ops.py -x -q > $O/pytest_part.log 2>&1; echo "rc=$?" >> $O/pytest_part.log
python scripts/summarize_bench_logs.py $O
for n in default inflight wb2mb2; do echo "== $n"; head -24 $O/$n.trace.txt | cut -c1-200; python - $O/$n.counts <<'PY'
import sys
try:
    rows=[list(map(int,l.split())) for l in open(sys.argv[1]).read().split('\n')[1:] if l.strip()]
    import statistics as st
    print("rounds",len(rows),"mean recs %.0f workA %.0f workC %.0f" % tuple(st.mean(r[i] for r in rows) for i in range(3)))
except Exception as e: print("no counts",e)
PY
done
