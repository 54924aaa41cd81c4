#!/bin/bash
# 2-GPU probe (run under gpurun --gpus 2): NVLS / multicast availability + the multi-GPU bench as the driver runs it.
mkdir -p gpurun_out/probe
O=gpurun_out/probe
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/smi.txt 2>&1
nvidia-smi topo -m >> $O/smi.txt 2>&1
cat /proc/sys/kernel/yama/ptrace_scope > $O/sys.txt 2>&1
grep -i cap /proc/self/status >> $O/sys.txt 2>&1
nproc >> $O/sys.txt
timeout 120 benchmarks/probes/mc_probe > $O/mc_probe.txt 2>&1
echo "rc=$?" >> $O/mc_probe.txt
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,NVLS timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 benchmarks/probes/symm_probe.py > $O/symm_probe.txt 2>&1
echo "rc=$?" >> $O/symm_probe.txt
grep -i -E "nvls|multicast|symm_mem|multimem" $O/symm_probe.txt | head -40 > $O/symm_probe_summary.txt
for cfg in "--steps 20 --warmup 5" "--steps 200 --warmup 10"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $cfg > "$O/bench2_$(echo $cfg | tr -d ' -').log" 2>&1
done
tail -c 3000 $O/mc_probe.txt; cat $O/symm_probe_summary.txt; tail -n 3 $O/sys.txt
