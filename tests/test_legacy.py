"""Legacy PS-Lite API: SimpleApp RPC and KVWorker/KVServer on the shared-memory mailboxes
(reference include/ps/simple_app.h, include/ps/kv_app.h; reference tests/test_simple_app.cc, test_kv_app.cc idea:
every worker pushes the same values `repeat` times and pulls the sum back)."""
import numpy as np
import pytest

from harness import run_cluster


def _simple_app_worker(kv, server, wid):
    from adapm_b200.legacy import SimpleApp, kAllNodes

    world = server.num_servers()
    app = SimpleApp(0, 0, server)
    got_req, got_res = [], []

    def on_request(d, a):
        got_req.append((d.head, bytes(d.body), d.sender))
        a.response(d, b"ack:" + bytes(d.body)[:16])

    def on_response(d, a):
        got_res.append((d.head, bytes(d.body), d.sender))

    app.set_request_handle(on_request)
    app.set_response_handle(on_response)
    kv.barrier()
    # one short message to everybody, one long (fragmented: 5000 B > one 976 B slot) to the next rank
    ts1 = app.request(7, f"hello from {wid}", kAllNodes)
    long_body = bytes((i * 7 + wid) % 251 for i in range(5000))
    ts2 = app.request(9, long_body, (wid + 1) % world)
    app.wait(ts1)
    app.wait(ts2)
    assert app.num_response(ts1) == world and app.num_response(ts2) == 1
    kv.barrier()
    # every rank saw one short request from every rank and the long one from its predecessor, intact
    short = sorted(s for h, b, s in got_req if h == 7)
    assert short == list(range(world))
    prev = (wid - 1) % world
    longs = [b for h, b, s in got_req if h == 9 and s == prev]
    assert len(longs) == 1 and longs[0] == bytes((i * 7 + prev) % 251 for i in range(5000))
    assert sorted(s for h, b, s in got_res if h == 7) == list(range(world))
    assert all(b.startswith(b"ack:") for _, b, _ in got_res)
    kv.barrier()
    kv.finalize()
    return len(got_req)


@pytest.mark.parametrize("mode", ["threads", "procs"])
def test_simple_app_rpc(mode):
    res = run_cluster(_simple_app_worker, world=3, workers=1, mode=mode, value_lengths=1, num_keys=8)
    assert all(r[0] == 4 for r in res.values())


def _kv_app_worker(kv, server, wid):
    from adapm_b200.legacy import KVServer, KVServerDefaultHandle, KVWorker

    world = server.num_servers()
    num_keys, vlen, repeat = 1000, 3, 4
    srv = KVServer(1, server)
    handle = KVServerDefaultHandle()
    srv.set_request_handle(handle)
    w = KVWorker(1, 1, server, num_keys=num_keys)
    kv.barrier()
    rng = np.random.default_rng(5)                # same keys/values on every rank
    keys = np.sort(rng.choice(num_keys, 200, replace=False)).astype(np.int64)
    vals = rng.random(keys.size * vlen).astype(np.float32)
    fired = []
    ts = [w.push(keys, vals, callback=lambda: fired.append(1)) for _ in range(repeat)]
    for t in ts:
        w.wait(t)
    assert len(fired) == repeat
    kv.barrier()                                  # all ranks' pushes are applied
    out = np.zeros_like(vals)
    lens = np.zeros(keys.size, np.int32)
    w.wait(w.pull(keys, out, lens))
    assert np.all(lens == vlen)
    np.testing.assert_allclose(out, vals * repeat * world, rtol=1e-5)
    # ZPull + AddCallback (kv_app.h:182-258): the callback runs when the request completes, or at once if it already has
    out2, late = np.zeros_like(vals), []
    t2 = w.zpull(keys, out2)
    w.add_callback(t2, lambda: late.append("cb"))
    w.wait(t2)
    assert late == ["cb"]
    np.testing.assert_allclose(out2, out)
    w.add_callback(t2, lambda: late.append("after"))
    assert late == ["cb", "after"]
    # static range partition: this server only ever stored keys of its own range
    lo, hi = w.ranges[server.my_rank()]
    assert all(lo <= k < hi for k in handle.store)
    # variable-length values through `lens`
    k2 = np.array([3, 500, 999], np.int64)
    l2 = np.array([1, 4, 2], np.int32)
    v2 = np.arange(7, dtype=np.float32) + 1
    kv.barrier()
    if wid == 0:
        w.wait(w.push(k2 + 0, v2, lens=l2, cmd=3))
    kv.barrier()
    kv.finalize()
    return float(out.sum())


@pytest.mark.parametrize("mode", ["threads", "procs"])
def test_kv_worker_server(mode):
    res = run_cluster(_kv_app_worker, world=2, workers=1, mode=mode, value_lengths=1, num_keys=8)
    vals = [r[0] for r in res.values()]
    assert vals[0] == vals[1] and vals[0] > 0


def test_default_slicer():
    from adapm_b200.legacy import KVPairs, default_slicer, server_key_ranges

    ranges = server_key_ranges(100, 4)
    assert ranges == [(0, 25), (25, 50), (50, 75), (75, 100)]
    kv = KVPairs(np.array([1, 2, 30, 80, 99], np.int64), np.arange(10, dtype=np.float32))
    s = default_slicer(kv, ranges)
    assert s[2] is None
    assert s[0].keys.tolist() == [1, 2] and s[0].vals.tolist() == [0, 1, 2, 3]
    assert s[1].keys.tolist() == [30] and s[1].vals.tolist() == [4, 5]
    assert s[3].keys.tolist() == [80, 99] and s[3].vals.tolist() == [6, 7, 8, 9]
    kv = KVPairs(np.array([1, 60], np.int64), np.arange(5, dtype=np.float32), np.array([2, 3], np.int32))
    s = default_slicer(kv, ranges)
    assert s[0].vals.tolist() == [0, 1] and s[2].vals.tolist() == [2, 3, 4] and s[2].lens.tolist() == [3]
    with pytest.raises(ValueError):
        default_slicer(KVPairs(np.array([5, 1], np.int64)), ranges)


def _big_reply_worker(kv, server, wid):
    """Two ranks answer each other with bodies far larger than a mailbox (64 slots x 976 B): the handlers run on the
    router threads, which must keep emptying their own mailbox while they wait for a slot in the peer's."""
    from adapm_b200.legacy import SimpleApp

    world = server.num_servers()
    app = SimpleApp(0, 0, server)
    big = bytes((i * 13 + wid) % 251 for i in range(200_000))
    got = []

    def on_request(d, a):
        a.response(d, big)

    def on_response(d, a):
        got.append((d.sender, len(d.body), bytes(d.body[:8])))

    app.set_request_handle(on_request)
    app.set_response_handle(on_response)
    kv.barrier()
    ts = [app.request(1, b"x" * 100_000, (wid + 1) % world) for _ in range(3)]
    for t in ts:
        app.wait(t)
    kv.barrier()
    peer = (wid + 1) % world
    assert len(got) == 3 and all(s == peer and n == 200_000 for s, n, _ in got)
    assert got[0][2] == bytes((i * 13 + peer) % 251 for i in range(8))
    kv.barrier()
    kv.finalize()
    return len(got)


def test_simple_app_large_mutual_replies():
    res = run_cluster(_big_reply_worker, world=2, workers=1, mode="threads", value_lengths=1, num_keys=8,
                      options={"wait_timeout_s": 30})
    assert all(r[0] == 3 for r in res.values())
