#!/bin/bash
O=gpurun_out/sgns_bulk; mkdir -p $O
ADAPM_SGNS_BULKRED=1 timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "sgns or word2vec or native_step" > $O/pytest_bulk.log 2>&1; echo "rc=$?" >> $O/pytest_bulk.log
tail -4 $O/pytest_bulk.log
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/n1_default.log 2>&1
ADAPM_SGNS_BULKRED=1 timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/n1_bulk.log 2>&1
ADAPM_SGNS_BULKRED=1 ADAPM_SGNS_REGS=104 timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/n1_bulk_lean.log 2>&1
ADAPM_SGNS_REGS=104 timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/n1_lean.log 2>&1
python scripts/summarize_bench_logs.py $O | cut -c1-120
