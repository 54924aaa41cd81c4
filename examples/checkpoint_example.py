"""System-level checkpoint of a parameter store and restore (also into a job of a different size).

    python -m adapm_b200.launch -s 3 --backend cpu examples/checkpoint_example.py save /tmp/adapm_ck
    python -m adapm_b200.launch -s 2 --backend cpu examples/checkpoint_example.py load /tmp/adapm_ck
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # noqa: E402
import torch

import adapm_b200 as ad
from adapm_b200.utils.checkpoint import load_store, save_store

mode, prefix = sys.argv[1], sys.argv[2]
ad.setup(num_keys=1000, num_threads=1)
server = ad.Server(4)
kv = ad.Worker(0, server)
rank, world = server.my_rank(), server.num_servers()
keys = torch.arange(1000)
if mode == "save":
    mine = keys[rank::world]
    kv.wait(kv.set(mine, mine.float().repeat_interleave(4)))
    kv.barrier()
    kv.intent(torch.arange(0, 50), kv.current_clock() + 1)       # move some keys around before saving
    kv.advance_clock(); kv.wait_sync()
    n = save_store(kv, prefix)
    print(f"[rank {rank}] wrote {n} keys")
else:
    n = load_store(kv, prefix)
    kv.barrier()
    out = torch.zeros(4000)
    kv.wait(kv.pull(keys, out))
    assert torch.equal(out.view(-1, 4)[:, 0], keys.float())
    print(f"[rank {rank}] restored {n} keys; all 1000 rows verified")
kv.barrier()
kv.finalize()
server.shutdown()
