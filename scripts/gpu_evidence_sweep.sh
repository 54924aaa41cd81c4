#!/bin/bash
# One-GPU evidence sweep: ONE `ncu --set full` capture per hot kernel (headline SGNS step at the 1M-vocab config, the
# sync-round kernels, KGE / MF / gather-GEMM / GEMM), compute-sanitizer passes over the multi-rank contract tests, and
# the micro-benchmarks; summaries are copied into profiles/ (tracked). Run on a B200 box from the repository root:
#
#     gpurun --timeout 2400 -- 'bash scripts/gpu_evidence_sweep.sh'
#
# Numbers printed by a run under ncu / compute-sanitizer are never used as benchmark values.
set -u
OUT=gpurun_out/sweep
mkdir -p $OUT profiles
NCU="ncu --set full --clock-control none --import-source on"

capture() {   # name, kernel regex, launch-skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 900 $NCU -k "regex:$regex" -s $skip -c 1 -o $OUT/$name -f "$@" > $OUT/$name.ncu.log 2>&1
  if [ -f $OUT/$name.ncu-rep ]; then
    ncu -i $OUT/$name.ncu-rep --page raw --csv > profiles/prof_${name}_ncu_raw.csv 2>/dev/null
    ncu -i $OUT/$name.ncu-rep --page details > profiles/prof_${name}_ncu_details.txt 2>/dev/null
    ncu -i $OUT/$name.ncu-rep --page source --csv 2>/dev/null | head -400 > profiles/prof_${name}_ncu_source_head.csv
  else
    echo "ncu capture of $name failed (see $OUT/$name.ncu.log)"; tail -5 $OUT/$name.ncu.log
  fi
}
echo "== ncu captures"
# headline config (1M vocab, d=300, neg 25): the fused SGNS step, default (128 regs, 1 GPU) and lean (104 regs, multi GPU)
capture sgns_tma_1m      sgns_step_tma_kernel 12 python bench.py --steps 6 --warmup 3 --loop python
ADAPM_SGNS_REGS=104 capture sgns_tma_lean_1m sgns_step_tma_kernel 12 python bench.py --steps 6 --warmup 3 --loop python
# sync-round kernels: three logical ranks on one GPU (host-sequenced round, so ncu's kernel serialisation is harmless)
capture round_row        phase_row_kernel 40 python -m pytest tests/test_gpu_contract.py -q -m gpu -k "many_key_operations_cuda and all"
capture round_resolve    "phase_meta_kernel.*1E.*0E" 40 python -m pytest tests/test_gpu_contract.py -q -m gpu -k "many_key_operations_cuda and all"
capture gemm_pair        gemm_nt_tcgen05_pair_kernel 12 python benchmarks/gemm_bench.py
if [ "${ADAPM_SWEEP_ALL:-0}" = "1" ]; then
capture round_scan       phase_scan_kernel 40 python -m pytest tests/test_gpu_contract.py -q -m gpu -k "many_key_operations_cuda and all"
capture round_b          phase_b_kernel 40 python -m pytest tests/test_gpu_contract.py -q -m gpu -k "many_key_operations_cuda and all"
capture kge_step         kge_step_kernel 5 python benchmarks/app_bench.py
capture mf_step          mf_step_kernel 5 python benchmarks/app_bench.py
capture gemm_persistent  gemm_nt_tcgen05_persistent_kernel 30 python benchmarks/gemm_bench.py
fi
capture gather_gemm      gather_gemm_kernel 0 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k gather
capture rescal           kge_rescal_kernel 0 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "rescal_and_dropout and RESCAL-64"

echo "== compute-sanitizer (multi-rank protocol on one GPU + the fused ops)"
# (racecheck loses track of the launches of the multi-threaded multi-rank harness - profiles/sanitizer_racecheck.txt; it is
#  opt-in here: ADAPM_SWEEP_RACECHECK=1)
for tool in memcheck ${ADAPM_SWEEP_RACECHECK:+racecheck} ${ADAPM_SWEEP_SYNCCHECK:+synccheck}; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --print-limit 20 \
      python -m pytest tests/test_gpu_contract.py tests/test_gpu_ops.py -q -m gpu -x \
      -k "locality_api_cuda or set_operation_cuda or set_under_relocation_cuda or sgns_step_matches or kge_complex or rescal_and_dropout or mf_step" \
      > $OUT/sanitizer_$tool.log 2>&1
  echo "rc=$?" >> $OUT/sanitizer_$tool.log
  { echo "compute-sanitizer --tool $tool over tests/test_gpu_contract.py (locality, set, set-under-relocation: 3-4 logical ranks on one GPU)"; \
    echo "and tests/test_gpu_ops.py (fused SGNS / ComplEx / RESCAL+dropout / MF kernels):"; \
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $OUT/sanitizer_$tool.log | tail -6; } > profiles/sanitizer_$tool.txt
  cat profiles/sanitizer_$tool.txt
done

echo "== micro benchmarks"
timeout 300 python benchmarks/gemm_bench.py > profiles/gemm_bench_latest.jsonl 2> $OUT/gemm_bench.err
timeout 300 python benchmarks/app_bench.py > profiles/app_bench_latest.jsonl 2> $OUT/app_bench.err
for c in kge mf; do timeout 400 python bench.py --config $c --steps 50 --warmup 5 > $OUT/bench_$c.log 2>&1; grep '^{' $OUT/bench_$c.log > profiles/bench_${c}_1gpu.json; done
ADAPM_CTR_KEYS=100000000 timeout 600 python bench.py --config ctr --steps 20 --warmup 3 > $OUT/bench_ctr.log 2>&1; grep '^{' $OUT/bench_ctr.log > profiles/bench_ctr_1gpu.json
timeout 300 python bench.py --impl nccl --steps 20 --warmup 5 > $OUT/nccl1.log 2>&1; grep '^{' $OUT/nccl1.log > profiles/nccl_arm_1gpu.json
ls -la profiles | tail -40
for f in profiles/bench_kge_1gpu.json profiles/bench_mf_1gpu.json profiles/bench_ctr_1gpu.json; do python - $f <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], "value %.3fG e2e %.3fG ms %.3f" % (j["value"]/1e9, j["e2e"]["value"]/1e9, j["ms_per_step"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
# gpurun merges only gpurun_out/ back: mirror the summaries written to profiles/ on the box
mkdir -p $OUT/profiles
cp profiles/prof_*_ncu_details.txt profiles/prof_*_ncu_raw.csv profiles/prof_*_ncu_source_head.csv profiles/sanitizer_*.txt \
   profiles/*_latest.jsonl profiles/bench_*_1gpu.json profiles/nccl_arm_1gpu.json $OUT/profiles/ 2>/dev/null
