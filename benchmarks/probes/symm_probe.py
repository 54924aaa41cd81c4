"""Probe: torch symmetric memory + NVLS multicast between the ranks of one box (torchrun, one rank per GPU)."""
import os, sys, time
import torch, torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
x = torch.ones(1 << 20, device=dev)
dist.all_reduce(x)
torch.cuda.synchronize()
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(rank + 1)
    dist.barrier()
    print(f"[rank {rank}] symm_mem ok: multicast_ptr={getattr(hdl, 'multicast_ptr', None)} buffer_ptrs={len(hdl.buffer_ptrs)}", flush=True)
    if getattr(hdl, "multicast_ptr", 0):
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print(f"[rank {rank}] multimem_all_reduce_ -> {t[0].item()} (expect {world * (world + 1) / 2})", flush=True)
except Exception as e:  # noqa
    print(f"[rank {rank}] symm_mem FAILED: {type(e).__name__}: {e}", flush=True)
dist.barrier()
dist.destroy_process_group()
