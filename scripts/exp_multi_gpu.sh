#!/bin/bash
# A/B matrix for the multi-GPU step time (run on an N-GPU box from the repository root):
#
#     gpurun --gpus 2 --timeout 900 -- 'bash scripts/exp_multi_gpu.sh 2'        # ~1 min per variant
#
# Every variant is one `bench.py --profile` run; the summary line has the headline value, the per-step time
# distribution with intent signalling on, the locality counters and the round statistics. Logs: gpurun_out/exp_*.log
N=${1:-2}
STEPS=${STEPS:-200}
mkdir -p gpurun_out
run() {
  local tag=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps $STEPS --warmup 10 --profile "$@" > gpurun_out/exp_${N}_$tag.log 2>&1
  python - gpurun_out/exp_${N}_$tag.log $tag <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); p = d['profile']; s = p.get('steps_with_intent', {})
        print("%-14s value %.4g  e2e %.4g  ms/step %.3f | per-step mean %.3f p50 %.3f p90 %.3f max %.2f | remote/slow rows per step %d/%d | "
              "reloc %d repl %d refresh %d rounds %d" % (
                  sys.argv[2], d['value'], d['e2e']['value'], d['ms_per_step'], s.get('mean', 0), s.get('p50', 0), s.get('p90', 0),
                  s.get('max', 0), s.get('rows_local_remote_slow', [0, 0, 0])[1] // 192, s.get('rows_local_remote_slow', [0, 0, 0])[2] // 192,
                  d['pm']['relocations'], d['pm']['replica_setups'], d['pm']['refreshes'], d['pm']['sync_rounds']))
        break
else:
    print(sys.argv[2], "FAILED - see", sys.argv[1])
PY
}
run default
ADAPM_SGNS_REGS=128 run regs128
run prepass --intent-prepass
ADAPM_SYNC_WORK_BLOCKS=1 run work1
ADAPM_SYNC_WORK_BLOCKS=4 run work4
run ra16 --read-ahead 16
run ra64 --read-ahead 48
run inflight1 --max-inflight 1
run idle1 --opt sys.sync.idle_period=1
run idle8 --opt sys.sync.idle_period=8
ADAPM_SYNC_TRACE=1 run trace
[ -f gpurun_out/kernel_trace.rank0.tsv ] && python scripts/analyze_kernel_trace.py gpurun_out/kernel_trace.rank0.tsv | head -14
