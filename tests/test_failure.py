"""Failure detection: a rank that dies must make its peers' blocking calls raise, not hang.
The shm fabric runs a failure-detector thread (process liveness of the peers) and every wait has a watchdog."""
import multiprocessing as mp
import os
import time
import uuid

import pytest


def _rank(rank, world, job, q, die, timeout_s):
    import torch

    import adapm_b200 as ad

    try:
        server = ad.Server(2, num_keys=16, num_threads=1, rank=rank, world=world, backend="cpu", fabric="shm", job=job,
                           options={"wait_timeout_s": timeout_s})
        kv = ad.Worker(0, server)
        kv.barrier()
        if rank == die:
            os._exit(17)                       # crash: no finalize, no shutdown
        t0 = time.time()
        try:
            kv.wait(kv.push(torch.tensor([1]), torch.ones(2)))
            kv.barrier()                       # the dead peer never arrives
            q.put((rank, "no error", time.time() - t0))
        except Exception as e:                 # noqa
            q.put((rank, str(e), time.time() - t0))
        q.close(); q.join_thread()             # flush the result before the hard exit
        os._exit(0)                            # the job is broken: do not attempt the collective shutdown
    except Exception as e:                     # noqa
        q.put((rank, "setup failed: " + str(e), 0.0))
        q.close(); q.join_thread()
        os._exit(1)


@pytest.mark.parametrize("detector", [True, False])
def test_dead_peer_breaks_barriers(detector, monkeypatch):
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    job = f"fail{os.getpid()}_{uuid.uuid4().hex[:6]}"
    # with the detector the survivors learn about the death within a fraction of a second; without a dead *process*
    # (here: simulated by a very short watchdog timeout) the watchdog of the wait itself fires
    timeout_s = 60 if detector else 2
    ps = [ctx.Process(target=_rank, args=(r, 3, job, q, 2 if detector else -1, timeout_s)) for r in range(3)]
    if not detector:
        ps = ps[:2]                            # rank 2 never starts: the others time out in setup / first barrier
    [p.start() for p in ps]
    got = [q.get(timeout=90) for _ in range(2)]
    [p.join(30) for p in ps]
    for rank, msg, dt in got:
        if detector:
            assert "failed peer" in msg or "broken" in msg, (rank, msg)
            assert dt < 10, f"rank {rank} needed {dt:.1f}s to notice the dead peer"
        else:
            assert "timed out" in msg or "watchdog" in msg or "broken" in msg, (rank, msg)
    if not detector:                           # the first one to give up is the watchdog; it breaks the barrier for the rest
        assert any("watchdog" in m or "timed out" in m for _, m, _ in got), got


def test_stale_shared_memory_segments_are_collected():
    """A crashed job leaves its POSIX-shm segments behind; rank 0 of a later job removes segments whose processes are
    all dead (and that are older than two minutes - simulated here by back-dating the control block)."""
    import glob
    import subprocess
    import sys

    ctx = mp.get_context("fork")
    q = ctx.Queue()
    job = f"stale{os.getpid()}_{uuid.uuid4().hex[:6]}"
    ps = [ctx.Process(target=_rank, args=(r, 2, job, q, 1, 60)) for r in range(2)]   # rank 1 crashes, rank 0 hard-exits
    [p.start() for p in ps]
    q.get(timeout=60)
    [p.join(30) for p in ps]
    left = glob.glob(f"/dev/shm/adapm_{job}_*")
    assert left, "the crashed job should have left its segments behind"
    old = time.time() - 600
    for f in left:
        os.utime(f, (old, old))                 # ctime cannot be set: the collector is told to use a zero age threshold
    env = dict(os.environ, ADAPM_SHM_GC_AGE_S="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "adapm_b200.launch", "-s", "2", "--backend", "cpu", "-m", "adapm_b200.apps.simple",
                        "--", "-k", "10", "-t", "1", "-i", "1", "-v", "1"], cwd=root, env=env, capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert not glob.glob(f"/dev/shm/adapm_{job}_*"), "stale segments were not collected"
