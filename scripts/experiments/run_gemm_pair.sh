#!/bin/bash
# 1-GPU check of the cta_group::2 pair GEMM: correctness tests, then throughput vs cuBLAS. Every step under timeout.
O=gpurun_out/gemm_pair; mkdir -p $O
ADAPM_GEMM_IMPL=c timeout 180 python -m pytest tests/test_gpu_gemm.py -x -q -k "gemm_nt_bf16_matches or rank_count_epilogue or fp8_matches" > $O/pytest_pair.log 2>&1; echo "rc=$?" >> $O/pytest_pair.log
tail -6 $O/pytest_pair.log
ADAPM_GEMM_IMPL=c timeout 300 python benchmarks/gemm_bench.py > $O/gemm_pair.jsonl 2> $O/gemm_pair.err; echo "bench rc=$?"
timeout 300 python benchmarks/gemm_bench.py > $O/gemm_default.jsonl 2> $O/gemm_default.err
python - <<'PY'
import json
for f in ("gpurun_out/gemm_pair/gemm_pair.jsonl","gpurun_out/gemm_pair/gemm_default.jsonl"):
    print(f)
    for l in open(f):
        try: j=json.loads(l)
        except Exception: continue
        print("  %-28s mine %.4f ms  cublas %.4f ms  ratio %.2f  TF %.0f  fp8 %s" % (j["name"], j["tcgen05_ms"], j["cublas_bf16_out_ms"], j["cublas_bf16_out_ms"]/j["tcgen05_ms"], j["tcgen05_tflops"], j.get("tcgen05_fp8_tflops")))
PY
tail -3 $O/gemm_pair.err
