"""Legacy PS-Lite API on this framework: SimpleApp RPC and KVWorker / KVServer (range-sliced key-value store).

    python -m adapm_b200.launch -s 2 --backend cpu examples/legacy_kv_example.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # noqa: E402
import numpy as np

import adapm_b200 as ad
from adapm_b200.legacy import KVServer, KVServerDefaultHandle, KVWorker, SimpleApp, kAllNodes

ad.setup(num_keys=100, num_threads=1)
server = ad.Server(1)                       # the parameter manager also provides rendezvous + barriers
kv = ad.Worker(0, server)
rank, world = server.my_rank(), server.num_servers()

# ---- SimpleApp: (int head, byte-string body) requests with responses
app = SimpleApp(app_id=7, customer_id=0, server=server)
app.set_request_handle(lambda req, a: a.response(req, b"pong from %d" % rank))
got = []
app.set_response_handle(lambda res, a: got.append(bytes(res.body)))
kv.barrier()
app.wait(app.request(1, b"ping", kAllNodes))
print(f"[rank {rank}] SimpleApp responses: {sorted(got)}")

# ---- KVWorker / KVServer: every server holds one static key range and sums what is pushed
srv = KVServer(app_id=8, server=server)
srv.set_request_handle(KVServerDefaultHandle())
w = KVWorker(app_id=8, customer_id=1, server=server, num_keys=100)
kv.barrier()
keys = np.array([3, 40, 77, 99], dtype=np.int64)
w.wait(w.push(keys, np.ones(keys.size * 2, dtype=np.float32)))         # 2 values per key
kv.barrier()
out = np.zeros(keys.size * 2, dtype=np.float32)
w.wait(w.pull(keys, out))
print(f"[rank {rank}] pulled {out.tolist()} (every rank pushed 1.0)")
assert np.all(out == world)
kv.barrier()
kv.finalize()
server.shutdown()
