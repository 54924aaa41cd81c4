#!/bin/bash
# final 1-GPU check of the build: the whole GPU test tier + smoke + the bench line as the driver runs it
O=gpurun_out/run12; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/n1_k20.log 2>&1
tail -4 $O/pytest_gpu.log | cut -c1-300; tail -2 $O/smoke.log; python scripts/summarize_bench_logs.py $O | grep -v "^    "
