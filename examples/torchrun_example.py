"""One rank per process under torchrun (the way bench.py and the multi-GPU apps are launched):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
        examples/torchrun_example.py [--backend cpu|cuda]

`adapm_b200.Server` takes rank / world size / job name from the environment torchrun sets (RANK, WORLD_SIZE, LOCAL_RANK,
MASTER_PORT); `torch.distributed` (gloo on CPU, nccl on GPUs) is only used here to cross-check the parameter manager's
result with an all-reduce."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adapm_b200 as adapm  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="cpu", choices=["cpu", "cuda"])
    args = ap.parse_args()
    if args.backend == "cuda":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("gloo" if args.backend == "cpu" else "nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    num_keys, vpk = 64, 4
    adapm.setup(num_keys, 1)
    server = adapm.Server(vpk, backend=args.backend)            # rank / world / job from the torchrun environment
    assert server.my_rank() == rank and server.num_servers() == world
    kv = adapm.Worker(0, server)
    keys = torch.arange(num_keys, dtype=torch.int64)
    mine = keys[(keys % world) == ((rank + 1) % world)]          # keys whose home is the NEXT rank
    kv.intent(mine, kv.current_clock() + 1, kv.current_clock() + 100)   # ... will be used here: relocate / replicate them
    kv.advance_clock()
    kv.wait_sync(); kv.barrier()
    local = sum(1 for k in mine.tolist() if server.is_local(k))
    kv.wait(kv.push(keys, torch.full((num_keys * vpk,), float(rank + 1))))
    kv.barrier(); kv.wait_sync(); kv.barrier(); kv.wait_sync(); kv.barrier()
    out = torch.zeros(num_keys * vpk)
    kv.wait(kv.pull(keys, out))
    want = torch.tensor([float(rank + 1)])
    if args.backend == "cuda":
        want = want.cuda()
    dist.all_reduce(want)                                        # sum over ranks of (rank + 1)
    ok = bool(torch.all(out == float(want.item()))) and local == mine.numel()
    print(f"[rank {rank}] {local}/{mine.numel()} intended keys local, every value == {float(want.item())}: {'PASSED' if ok else 'FAILED'}",
          flush=True)
    kv.finalize()
    server.shutdown()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
