#!/bin/bash
# 1-GPU A/B of the SGNS step: block shape (warps per SM) and L2 prefetches; plus the new GPU tests
O=gpurun_out/run11; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 100 --warmup 10 > $O/$name.log 2>&1; }
run base X=1
run pf3 ADAPM_SGNS_PREFETCH=3
run t192_r112 ADAPM_SGNS_THREADS=192 ADAPM_SGNS_REGS=112
run t192_r112_pf3 ADAPM_SGNS_THREADS=192 ADAPM_SGNS_REGS=112 ADAPM_SGNS_PREFETCH=3
run t192_r104_pf3 ADAPM_SGNS_THREADS=192 ADAPM_SGNS_REGS=104 ADAPM_SGNS_PREFETCH=3
run pf1 ADAPM_SGNS_PREFETCH=1
run pf2 ADAPM_SGNS_PREFETCH=2
ADAPM_SGNS_THREADS=192 ADAPM_SGNS_REGS=112 ADAPM_SGNS_PREFETCH=3 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "sgns" > $O/pytest_sgns_variant.log 2>&1; echo "rc=$?" >> $O/pytest_sgns_variant.log
timeout 300 python -m pytest tests/test_gpu_contract.py -q -m gpu -k "prepass or many_key" > $O/pytest_prepass.log 2>&1; echo "rc=$?" >> $O/pytest_prepass.log
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "shared_negatives" > $O/pytest_shared.log 2>&1; echo "rc=$?" >> $O/pytest_shared.log
timeout 300 python benchmarks/sgns_shared_bench.py > $O/sgns_shared_bench.log 2>&1
python scripts/summarize_bench_logs.py $O | grep -v "^    "
tail -3 $O/pytest_sgns_variant.log $O/pytest_prepass.log; tail -25 $O/pytest_shared.log | cut -c1-200; tail -3 $O/sgns_shared_bench.log | cut -c1-1200
