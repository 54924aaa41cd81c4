#!/bin/bash
O=gpurun_out/sgns_pipe; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "sgns or word2vec or native_step" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
ADAPM_SGNS_BULKRED=0 timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "sgns or word2vec" > $O/pytest_nobulk.log 2>&1; echo "rc=$?" >> $O/pytest_nobulk.log
tail -3 $O/pytest.log; tail -3 $O/pytest_nobulk.log
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/n1_bulk_pipe.log 2>&1
ADAPM_SGNS_REGS=104 timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/n1_bulk_pipe_lean.log 2>&1
ADAPM_SGNS_BULKRED=0 timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/n1_red_pipe.log 2>&1
python scripts/summarize_bench_logs.py $O | cut -c1-110
