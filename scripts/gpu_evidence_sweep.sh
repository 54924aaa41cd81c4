#!/bin/bash
# One-GPU evidence sweep: tests, headline bench (+ per-kernel profile), micro-benchmarks and ONE `ncu --set full`
# capture per hot kernel, all copied into profiles/ (tracked). Run on a B200 box from the repository root:
#
#     gpurun --timeout 1500 -- 'bash scripts/gpu_evidence_sweep.sh'          # ~12 GPU-minutes
#
# Numbers printed by a run under ncu are never used as benchmark values (ncu replays kernels).
set -u
OUT=gpurun_out/sweep
mkdir -p $OUT profiles
NCU="ncu --set full --clock-control none --import-source on"

echo "== tests"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/tests.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench"; timeout 300 python bench.py --profile > $OUT/bench_1gpu.log 2>&1; grep '^{' $OUT/bench_1gpu.log > profiles/bench_1gpu_latest.json
ADAPM_SGNS_REGS=104 timeout 300 python bench.py --steps 150 --profile > $OUT/bench_1gpu_lean.log 2>&1
grep '^{' $OUT/bench_1gpu_lean.log > profiles/bench_1gpu_lean_latest.json
echo "== micro benchmarks"
timeout 300 python benchmarks/gemm_bench.py > profiles/gemm_bench_latest.jsonl 2> $OUT/gemm_bench.err
timeout 300 python benchmarks/app_bench.py > profiles/app_bench_latest.jsonl 2> $OUT/app_bench.err

capture() {   # name, kernel regex, launch-skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 $NCU -k "regex:$regex" -s $skip -c 1 -o $OUT/$name -f "$@" > $OUT/$name.ncu.log 2>&1
  if [ -f $OUT/$name.ncu-rep ]; then
    ncu -i $OUT/$name.ncu-rep --page raw --csv > profiles/prof_${name}_ncu_raw.csv 2>/dev/null
    ncu -i $OUT/$name.ncu-rep --page details > profiles/prof_${name}_ncu_details.txt 2>/dev/null
  else
    echo "ncu capture of $name failed (see $OUT/$name.ncu.log)"
  fi
}
echo "== ncu captures"
capture sgns_tma        sgns_step_tma_kernel 20 python bench.py --steps 30 --warmup 5
ADAPM_SGNS_REGS=104 capture sgns_tma_lean sgns_step_tma_kernel 20 python bench.py --steps 30 --warmup 5
capture kge_step        kge_step_kernel 5 python benchmarks/app_bench.py
capture mf_step         mf_step_kernel 5 python benchmarks/app_bench.py
capture gemm_persistent gemm_nt_tcgen05_persistent_kernel 30 python benchmarks/gemm_bench.py
capture gather_gemm     gather_gemm_kernel 2 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k gather
capture sampler         sample_kernel 20 python bench.py --steps 30 --warmup 5
ls -la profiles | tail -30
