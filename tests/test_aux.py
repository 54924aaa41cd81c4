"""Auxiliary subsystems: key tracing, locality statistics, PS all-reduce, options, ActionTimer, launcher."""
import glob
import os
import subprocess
import sys

import pytest
import torch

from harness import run_cluster

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _trace_worker(kv, server, wid):
    kv.barrier()
    if wid == 0:
        kv.intent(torch.tensor([3, 4]), 1)      # key 3 is home at rank 1, key 4 at rank 0
    v = torch.zeros(4)
    kv.pull(torch.tensor([3, 5]), v)
    kv.advance_clock(); kv.wait_sync(); kv.barrier()
    kv.advance_clock(); kv.advance_clock(); kv.wait_sync(); kv.barrier()
    kv.pull(torch.tensor([3, 5]), v)
    kv.finalize()
    return None


def test_key_tracing_and_locality_stats(tmp_path):
    run_cluster(_trace_worker, world=2, workers=1, mode="threads", value_lengths=2, num_keys=12,
                options={"sys.trace.keys": "3,4", "sys.stats.out": str(tmp_path), "sys.stats.locality": 1})
    t0 = open(tmp_path / "traces.0.tsv").read().splitlines()
    t1 = open(tmp_path / "traces.1.tsv").read().splitlines()
    ev0 = [(l.split("\t")[1], l.split("\t")[3]) for l in t0]
    ev1 = [(l.split("\t")[1], l.split("\t")[3]) for l in t1]
    assert ("3", "INTENT_START") in ev0 and ("3", "ALLOC") in ev0     # key 3 relocated to rank 0 ...
    assert ("3", "DEALLOC") in ev1                                     # ... and left rank 1
    loc = open(tmp_path / "locality_stats.rank.0.tsv").read().splitlines()
    assert loc[0] == "Param\tAccesses\tLocalAccesses" and any(l.startswith("3\t2\t") for l in loc[1:])


def _allreduce_worker(kv, server, wid):
    from adapm_b200.utils.allreduce import ps_allreduce

    out = ps_allreduce(kv, 7, torch.tensor([float(wid + 1), 10.0 * (wid + 1)]))
    kv.barrier()
    kv.finalize()
    return out.tolist()


def test_ps_allreduce():
    res = run_cluster(_allreduce_worker, world=3, workers=1, mode="threads", value_lengths=4, num_keys=10)
    for r in res.values():
        assert r[0] == [6.0, 60.0]


def test_options_and_action_timer():
    from adapm_b200 import _C
    import adapm_b200 as ad

    # Poisson quantile used by the ActionTimer (values from scipy.stats.poisson.ppf): the reference's start-up
    # window is quantile_0.9999(Poisson(2 * 10)) = 39 clocks
    assert _C.poisson_quantile(20.0, 0.9999) == 39
    assert _C.poisson_quantile(399.0, 0.9999) == 475
    assert abs(_C.poisson_quantile(2e6, 0.9999) - 2005262) <= 2
    with pytest.raises(RuntimeError):
        ad.Server(2, num_keys=10, rank=0, world=1, backend="cpu", options={"sys.unknown_flag": 1})
    with pytest.raises(RuntimeError):
        ad.Server(2, num_keys=10, rank=0, world=1, backend="cpu", options={"sys.channels": 3})
    s = ad.Server(2, num_keys=10, rank=0, world=1, backend="cpu", job="opts",
                  options={"sys.techniques": "replication_only", "sys.sync.threshold": "inf", "sys.sync.max_per_sec": 50,
                           "sys.zmq_threads": 5, "sys.location_caches": 0, "sampling.scheme": "naive"})
    w = ad.Worker(0, s)
    with pytest.raises(IndexError):
        w.pull(torch.tensor([10]), torch.zeros(2))
    with pytest.raises(ValueError):
        w.pull(torch.tensor([1]), torch.zeros(3))
    w.finalize(); s.shutdown()


def test_launcher_runs_simple_app():
    out = subprocess.run([sys.executable, "-m", "adapm_b200.launch", "-s", "2", "--backend", "cpu", "-m",
                          "adapm_b200.apps.simple", "--", "-k", "10", "-t", "2", "-i", "2", "-v", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("iteration") == 2 * 2 * 2


def test_bindings_example_runs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "bindings_example.py")], capture_output=True,
                         text=True, timeout=180, cwd=ROOT, env={**os.environ, "ADAPM_BACKEND": "cpu"})
    assert out.returncode == 0 and out.stdout.count("done") == 4, out.stdout[-2000:] + out.stderr[-2000:]


def test_kv_match_and_sort():
    from adapm_b200.utils import kvmatch as km

    g = torch.Generator().manual_seed(3)
    keys = torch.randperm(1000, generator=g)[:300]
    srt = km.parallel_sort(keys.clone())
    assert torch.equal(srt, torch.sort(keys).values)
    sk, sv = km.sort_by_key(keys, torch.arange(600.), k=2)
    assert torch.equal(sk, srt) and sv.view(-1, 2)[0].tolist() == [2.0 * int(torch.argmin(keys)), 2.0 * int(torch.argmin(keys)) + 1]
    src_k = torch.tensor([1, 3, 5, 9]); src_v = torch.arange(8.)
    dst_k = torch.tensor([0, 1, 2, 5, 9, 11]); dst_v = torch.ones(12)
    n = km.parallel_ordered_match(src_k, src_v, dst_k, dst_v, k=2, op=km.PLUS)
    assert n == 6 and dst_v.tolist() == [1, 1, 1, 2, 1, 1, 5, 6, 7, 8, 1, 1]
    n = km.parallel_ordered_match(src_k, src_v, dst_k, dst_v, k=2, op=km.ASSIGN)
    assert n == 6 and dst_v.view(-1, 2)[1].tolist() == [0, 1] and dst_v.view(-1, 2)[4].tolist() == [6, 7]
    assert km.parallel_ordered_match(torch.tensor([100]), torch.zeros(2), dst_k, dst_v, k=2) == 0


def test_network_utils():
    import socket

    from adapm_b200.utils import net

    ifs = net.list_interfaces()
    assert ifs and all(len(p) == 2 for p in ifs)
    name, ip = net.get_available_interface_and_ip()
    assert net.get_ip(name) == ip and net.get_ip("no-such-if0") is None
    port = net.get_available_port()
    with socket.socket() as s:
        s.bind(("127.0.0.1", port))          # the port really is free


def test_lint_clean():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "lint.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]


def test_launcher_dry_run_mpi_and_ssh(tmp_path):
    base = [sys.executable, "-m", "adapm_b200.launch", "-s", "2", "--dry-run"]
    r = subprocess.run(base + ["--launcher", "mpi", "-m", "adapm_b200.apps.simple", "--", "-k", "10"], cwd=ROOT,
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("mpirun -n 2") and "-x ADAPM_JOB=" in r.stdout
    hf = tmp_path / "hosts"
    hf.write_text("localhost\n")
    r = subprocess.run(base + ["--launcher", "ssh", "-H", str(hf), "-m", "adapm_b200.apps.simple"], cwd=ROOT,
                       capture_output=True, text=True)
    lines = r.stdout.strip().splitlines()
    assert r.returncode == 0 and len(lines) == 2 and "RANK=1" in lines[1] and lines[0].startswith("ssh ")
    hf.write_text("hostA\nhostB\n")     # the fabric is single-node: two hosts are refused
    r = subprocess.run(base + ["--launcher", "ssh", "-H", str(hf), "-m", "x"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 2 and "single-node" in r.stderr


def _pool_worker(kv, server, wid):
    h = server._impl.slot_histogram()
    kv.barrier()
    kv.finalize()
    return h[100]          # free slots of length class 0 on this rank


def test_pool_sizing_auto_and_explicit():
    # auto: as many slots as the memory budget buys, capped at "every key + slack" (100k keys -> 125k slots)
    res = run_cluster(_pool_worker, world=2, workers=1, mode="threads", value_lengths=2, num_keys=100000)
    assert [r[0] for r in res.values()] == [125000 - 50000] * 2
    # a tight memory budget falls back to twice the home share (+1024)
    res = run_cluster(_pool_worker, world=2, workers=1, mode="threads", value_lengths=2, num_keys=100000,
                      options={"pool_bytes": 1 << 20})
    assert [r[0] for r in res.values()] == [2 * 50000 + 1024 - 50000] * 2
    # explicit factor
    res = run_cluster(_pool_worker, world=2, workers=1, mode="threads", value_lengths=2, num_keys=100000,
                      options={"pool_factor": 1.5})
    assert [r[0] for r in res.values()] == [75000 + 1024 - 50000] * 2


def _ckpt_save_worker(kv, server, wid, prefix=None):
    from adapm_b200.utils.checkpoint import save_store

    nk = server.num_keys()
    keys = torch.arange(wid, nk, server.num_servers())
    kv.wait(kv.set(keys, (keys.float().repeat_interleave(3) + 0.5)))
    kv.barrier()
    kv.intent(torch.tensor([1, 2, 3, 4, 5]), kv.current_clock() + 1)      # move / replicate a few keys first
    kv.advance_clock(); kv.wait_sync(); kv.barrier()
    kv.wait(kv.push(torch.tensor([1, 2, 3]), torch.ones(9)))
    n = save_store(kv, prefix)
    kv.finalize()
    return n


def _ckpt_load_worker(kv, server, wid, prefix=None):
    from adapm_b200.utils.checkpoint import load_store

    n = load_store(kv, prefix)
    kv.barrier()
    nk = server.num_keys()
    out = torch.zeros(nk * 3)
    kv.wait(kv.pull(torch.arange(nk), out))
    kv.barrier()
    kv.finalize()
    return n, out.view(nk, 3)[:, 0].tolist()


def test_store_checkpoint_roundtrip_across_world_sizes(tmp_path):
    import functools

    prefix = str(tmp_path / "ck")
    nk = 50
    res = run_cluster(functools.partial(_ckpt_save_worker, prefix=prefix), world=3, workers=1, mode="threads",
                      value_lengths=3, num_keys=nk)
    assert sum(r[0] for r in res.values()) == nk                      # every key written exactly once
    res = run_cluster(functools.partial(_ckpt_load_worker, prefix=prefix), world=2, workers=1, mode="threads",
                      value_lengths=3, num_keys=nk)
    assert sum(r[0][0] for r in res.values()) == nk
    want = [k + 0.5 + (3.0 if k in (1, 2, 3) else 0.0) for k in range(nk)]   # 3 ranks pushed +1 each
    for r in res.values():
        assert r[0][1] == want


def test_examples_run(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    launch = [sys.executable, "-m", "adapm_b200.launch", "--backend", "cpu"]
    r = subprocess.run(launch + ["-s", "2", os.path.join(ROOT, "examples", "legacy_kv_example.py")], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "pulled [2.0" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    ck = str(tmp_path / "ck")
    r = subprocess.run(launch + ["-s", "3", os.path.join(ROOT, "examples", "checkpoint_example.py"), "save", ck], cwd=ROOT,
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    r = subprocess.run(launch + ["-s", "2", os.path.join(ROOT, "examples", "checkpoint_example.py"), "load", ck], cwd=ROOT,
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all 1000 rows verified" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_torchrun_launch_with_gloo():
    """One rank per process under `python -m torch.distributed.run` (how bench.py and the multi-GPU apps are launched):
    rank / world / job come from the environment torchrun sets; gloo cross-checks the result with an all-reduce."""
    import socket

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "3", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "examples", "torchrun_example.py")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2500:]
    assert out.count("PASSED") == 3 and "FAILED" not in out, out[-2500:]


def _allreduce_sum_worker(kv, server, wid):
    out = server.allreduce_sum([float(wid + 1), 0.5, -2.0 * wid])
    out2 = server.allreduce_sum([1.0])           # reusable
    kv.barrier()
    kv.finalize()
    return out, out2


@pytest.mark.parametrize("mode", ["threads", "procs"])
def test_control_block_allreduce(mode):
    res = run_cluster(_allreduce_sum_worker, world=3, workers=1, mode=mode, value_lengths=1, num_keys=8)
    for r in res.values():
        assert r[0][0] == [6.0, 1.5, -6.0] and r[0][1] == [3.0]


def test_fd_channel_between_ranks():
    """The unix-socket channel that carries VMM heap handles / the multicast object between the ranks of a box."""
    from adapm_b200 import _C

    assert _C._fdpass_selftest()


def _cadence_worker(kv, server, wid):
    import time

    errors = []
    keys = torch.arange(8, dtype=torch.int64)
    kv.wait(kv.push(keys, torch.ones(8 * 2, dtype=server.dtype)))
    kv.barrier()
    # 1) no clock progress: a round only every min_clocks_wait_ms (200 ms), not every millisecond
    r0 = server.counters()["sync_rounds"]
    t0 = time.perf_counter()
    time.sleep(1.0)
    idle_rounds = server.counters()["sync_rounds"] - r0
    el = time.perf_counter() - t0
    if not (1 <= idle_rounds <= el / 0.2 + 3):
        errors.append(f"{idle_rounds} rounds in {el:.2f} s without clock progress (one per 200 ms expected)")
    # 2) requests that wait for rounds are served at once, also when only ONE rank asks (a round is collective: the
    #    peers see the request in the control block; two floor-delayed rounds would take >= 0.4 s)
    if server.my_rank() == 0:
        t0 = time.perf_counter()
        kv.wait_sync()
        if time.perf_counter() - t0 > 0.3:
            errors.append(f"WaitSync took {time.perf_counter() - t0:.3f} s under the cadence floor")
    kv.barrier()
    # 3) clock progress releases rounds: 4 clocks per round
    r1 = server.counters()["sync_rounds"]
    t0 = time.perf_counter()
    for _ in range(40):
        kv.advance_clock()
        time.sleep(0.002)
    busy_rounds = server.counters()["sync_rounds"] - r1
    if busy_rounds < 2:       # (a loaded CI box runs few rounds in 0.1 s; without clock progress it would be at most one)
        errors.append(f"only {busy_rounds} rounds while the clock advanced by 40 in {time.perf_counter() - t0:.2f} s")
    kv.barrier()
    kv.finalize()
    return errors


def test_round_cadence_floor_in_worker_clocks():
    """sys.sync.min_clocks: a new round starts only after the workers advanced that many clocks (default for the
    device-resident round on GPUs; opt-in here), but never later than min_clocks_wait_ms, and at once for WaitSync."""
    res = run_cluster(_cadence_worker, world=2, workers=1, mode="threads", value_lengths=2, num_keys=16, dtype="float32",
                      options={"sys.sync.min_clocks": 4, "sys.sync.min_clocks_wait_ms": 200})
    errs = [e for r in res.values() for k, v in r.items() if k != "counters" for e in v]
    assert not errs, "\n".join(errs)


def _metrics_worker(kv, server, wid):
    import urllib.request

    from adapm_b200.utils.metrics import start_metrics_server

    keys = torch.arange(10, dtype=torch.int64)
    kv.wait(kv.push(keys, torch.ones(10 * 2, dtype=server.dtype)))
    kv.intent(keys, kv.current_clock(), kv.current_clock() + 50)
    kv.wait_sync(); kv.barrier()
    _, port = start_metrics_server(server, workers=[kv])
    body = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=10).read().decode()
    kv.barrier()
    kv.finalize()
    return body


def test_prometheus_metrics_endpoint():
    """utils/metrics.py: the node's counters and the workers' locality counters are scraped over HTTP while the job runs."""
    pytest.importorskip("prometheus_client")
    res = run_cluster(_metrics_worker, world=2, workers=1, mode="threads", value_lengths=2, num_keys=16, dtype="float32")
    for rank, r in res.items():
        body = r[0]
        assert f'adapm_sync_rounds_total{{rank="{rank}"}}' in body, body[:600]
        assert "adapm_protocol_errors_total" in body and "adapm_worker_push_params_total" in body
        rounds = [float(l.split()[-1]) for l in body.splitlines() if l.startswith("adapm_sync_rounds_total")]
        assert rounds and rounds[0] >= 2          # the WaitSync before the scrape ran at least two rounds
    moved = sum(float(l.split()[-1]) for r in res.values() for l in r[0].splitlines()
                if l.startswith(("adapm_relocations_total", "adapm_replica_setups_total")))
    assert moved > 0                              # both ranks wanted all ten keys: relocated or replicated


def _save_for_serving(kv, server, wid):
    from adapm_b200.utils.checkpoint import save_store

    d = 8
    keys = torch.arange(server.num_keys(), dtype=torch.int64)
    mine = keys[(keys % server.num_servers()) == server.my_rank()]
    g = torch.Generator().manual_seed(1)
    table = torch.randn(server.num_keys(), 2 * d, generator=g)
    table[2] = table[0] * 3.0                      # key 2 is parallel to key 0 (syn0 keys are the even ones)
    kv.wait(kv.set(mine, table[mine].reshape(-1)))
    kv.barrier()
    save_store(kv, server._ck)
    kv.barrier()
    kv.finalize()
    return table[:4].tolist()


def test_serving_from_a_store_checkpoint(tmp_path):
    """adapm_b200.serve: a 2-rank job saves its store, a single-rank server loads it and answers /pull, /topk, /metrics."""
    pytest.importorskip("fastapi")
    pytest.importorskip("uvicorn")
    import json
    import threading
    import time
    import urllib.request

    import uvicorn

    from adapm_b200.serve import load_service, make_app

    ck = str(tmp_path / "model")

    def setup(server):
        server._ck = ck

    res = run_cluster(_save_for_serving, world=2, workers=1, mode="threads", setup_fn=setup, value_lengths=16, num_keys=40)
    table4 = res[0][0]
    svc = load_service(ck, embed_dim=8, backend="cpu", stride=2, offset=0)
    import socket

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    srv = uvicorn.Server(uvicorn.Config(make_app(svc), host="127.0.0.1", port=port, log_level="error"))
    th = threading.Thread(target=srv.run, daemon=True)
    th.start()
    try:
        for _ in range(100):
            if srv.started:
                break
            time.sleep(0.05)
        get = lambda path: urllib.request.urlopen(f"http://127.0.0.1:{port}{path}", timeout=10).read().decode()  # noqa: E731
        assert json.loads(get("/health"))["candidates"] == 20
        got = json.loads(get("/pull?keys=0,3"))
        assert got["keys"] == [0, 3]
        assert all(abs(a - b) < 1e-6 for a, b in zip(got["values"][0], table4[0]))
        assert all(abs(a - b) < 1e-6 for a, b in zip(got["values"][1], table4[3]))
        nb = json.loads(get("/topk?key=0&k=3"))["neighbours"]
        assert nb[0][0] == 2 and nb[0][1] > 0.999                   # the parallel row is the nearest neighbour
        assert "adapm_pull_local_total" in get("/metrics")
        try:
            get("/pull?keys=9999")
            assert False, "out-of-range key was served"
        except urllib.error.HTTPError as e:
            assert e.code == 400
    finally:
        srv.should_exit = True
        th.join(10)
        svc.worker.finalize()
        svc.server.shutdown()


def test_schedules_to_intents():
    """parallel/schedules.py: the DSGD Latin square, the column-wise ranged-intent plan and the look-ahead helper."""
    import numpy as np

    from adapm_b200.parallel.schedules import LookaheadIntents, column_intent_plan, wor_block_schedule

    s = wor_block_schedule(5, epoch=3, seed=7)
    assert sorted(s[0].tolist()) == list(range(5)) and all(sorted(s[:, w].tolist()) == list(range(5)) for w in range(5))
    cols, dur, ptr = column_intent_plan(np.array([0, 0, 0, 0, 0, 2, 2, 5, 5, 5, 5, 5, 5, 9]), batch=4)
    assert cols.tolist() == [0, 2, 5, 9]
    assert dur.tolist() == [2, 1, 3, 1]                       # column 0 spans batches 0-1, column 5 batches 1-3 ...
    assert [cols[ptr[b]:ptr[b + 1]].tolist() for b in range(4)] == [[0], [2, 5], [], [9]]

    class FakeWorker:
        def __init__(self):
            self.clock, self.calls = 10, []

        def current_clock(self):
            return self.clock

        def intent(self, keys, start, end=0):
            self.calls.append((sorted(keys.tolist()), start, end))

    w = FakeWorker()
    data = [torch.tensor([[3, 3, 1], [1, 7, 7]]) + 10 * b for b in range(5)]
    look = LookaheadIntents(w, num_batches=5, read_ahead=2, keys_of=lambda b: data[b])
    look.prime()
    assert w.calls == [([1, 3, 7], 10, 11), ([11, 13, 17], 11, 12)]
    for s_ in range(5):
        look.signal(s_)
        w.clock += 1
    assert [c[0][0] for c in w.calls[2:]] == [21, 31, 41] and [c[1] for c in w.calls[2:]] == [12, 13, 14]
    assert look.keys_signalled == 15
