"""Property test of the protocol on the CUDA backend (not collected by default: run it explicitly on a GPU box)

    python -m pytest tests/manual_gpu_property.py -q            # ADAPM_HYP_EXAMPLES=200 for a longer hunt

Same random programs as tests/test_protocol_property.py, float32 rows holding small integers (exact), three logical
ranks on one GPU (inproc fabric) so that it also runs on a single-GPU box."""
import functools
import os

import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import test_protocol_property as pp
from harness import run_cluster

pytestmark = pytest.mark.gpu


def _run_f32(kv, server, wid, programs=None):
    # the CPU worker uses int64 rows; here everything is float32 (integers below 2^24 are exact)
    prog = programs[wid]
    rounds = max(len(p) for p in programs)
    mine = torch.zeros(pp.NUM_KEYS)
    errs = []
    for r in range(rounds):
        for o in (prog[r] if r < len(prog) else []):
            if o[0] == "intent":
                kv.intent(torch.tensor(o[1]), kv.current_clock() + o[2], kv.current_clock() + o[2] + o[3])
            elif o[0] == "push":
                k = torch.tensor(o[1])
                kv.wait(kv.push(k, torch.ones(len(o[1]) * pp.VPK)))
                mine[k] += 1
            elif o[0] == "pull":
                k = torch.tensor(o[1])
                v = torch.zeros(len(o[1]) * pp.VPK)
                kv.wait(kv.pull(k, v))
                if bool((v.view(-1, pp.VPK) < mine[k].view(-1, 1)).any()):
                    errs.append(f"w{wid} round {r}: read-your-writes violated: {v.tolist()} < {mine[k].tolist()}")
            elif o[0] == "clock":
                kv.advance_clock()
            elif o[0] == "sync":
                kv.wait_sync()
        kv.barrier()
    kv.waitall(); kv.wait_sync(); kv.barrier(); kv.wait_sync(); kv.barrier()
    out = torch.zeros(pp.NUM_KEYS * pp.VPK)
    kv.wait(kv.pull(torch.arange(pp.NUM_KEYS), out))
    kv.barrier()
    kv.finalize()
    return errs, out.view(-1, pp.VPK)[:, 0].tolist(), mine.tolist()


@settings(max_examples=int(os.environ.get("ADAPM_HYP_EXAMPLES", "15")), deadline=None, suppress_health_check=list(HealthCheck))
@given(programs=st.lists(pp.program, min_size=3, max_size=3),
       technique=st.sampled_from(["all", "replication_only", "relocation_only"]))
def test_random_programs_are_exact_cuda(programs, technique):
    res = run_cluster(functools.partial(_run_f32, programs=programs), world=3, workers=1, mode="threads",
                      value_lengths=pp.VPK, num_keys=pp.NUM_KEYS, dtype="float32", backend="cuda",
                      options={"sys.techniques": technique})
    total = [0.0] * pp.NUM_KEYS
    for r in res.values():
        errs, final, mine = r[0]
        assert not errs, errs
        total = [a + b for a, b in zip(total, mine)]
    for r in res.values():
        assert r[0][1] == total, (r[0][1], total)
        assert r["counters"]["protocol_errors"] == 0
