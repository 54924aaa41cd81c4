"""The CPU contract tests re-run on the CUDA backend: several logical ranks share one GPU
(inproc fabric), so the full relocation/replication protocol runs through the sm_100a kernels
with peer pointers that happen to live on the same device."""
import pytest

from harness import run_cluster
import test_contract_dynamic as dyn
import test_contract_locality as loc
import test_contract_many_keys as mk

pytestmark = pytest.mark.gpu


def _errs(res):
    return [e for r in res.values() for k, v in r.items() if k != "counters" for e in v]


def test_locality_api_cuda():
    res = run_cluster(loc._worker, world=3, workers=loc.NUM_LOCAL, mode="threads", value_lengths=1, num_keys=12,
                      dtype="float32", backend="cuda")
    assert not _errs(res), "\n".join(_errs(res))


@pytest.mark.parametrize("technique", ["all", "replication_only", "relocation_only"])
def test_dynamic_allocation_cuda(technique):
    world, workers = 3, 2
    res = run_cluster(dyn._dyn_worker, world=world, workers=workers, mode="threads", value_lengths=2, num_keys=20,
                      dtype="float32", backend="cuda", options={"sys.techniques": technique})
    total = world * workers * dyn.RUNS
    assert res[0][0] == [total, 2 * total], f"lost or duplicated updates: {res[0][0]}"
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())
    moved = sum(r["counters"]["relocations"] + r["counters"]["replica_setups"] for r in res.values())
    assert moved > 0


@pytest.mark.parametrize("technique", ["all", "replication_only", "relocation_only"])
def test_many_key_operations_cuda(technique):
    res = run_cluster(mk._worker, world=4, workers=2, mode="threads", value_lengths=mk.VPK, num_keys=mk.NUM_KEYS,
                      dtype="float32", backend="cuda", options={"sys.techniques": technique})
    errs = _errs(res)
    if technique == "relocation_only":
        errs = [e for e in errs if "were not local" not in e]
    assert not errs, "\n".join(errs)
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())


def test_set_operation_cuda():
    res = run_cluster(dyn._set_worker, world=4, workers=2, mode="threads", value_lengths=2, num_keys=20,
                      dtype="float32", backend="cuda")
    assert not _errs(res), "\n".join(_errs(res))


def test_set_under_relocation_cuda():
    """Set on keys that two other ranks keep relocating: the kernel reports keys in flight and the worker repeats them
    after a round (the device-side grace period would otherwise wait for a kernel that waits for the round)."""
    res = run_cluster(dyn._set_under_relocation_worker, world=3, workers=1, mode="threads", value_lengths=2,
                      num_keys=dyn.SET_KEYS + 8, dtype="float32", backend="cuda")
    assert not _errs(res), "\n".join(_errs(res))
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())


def test_host_sequenced_round_cuda(monkeypatch):
    """ADAPM_HOST_ROUND=1 keeps the host-sequenced round (control-plane barriers, event-based grace) working."""
    monkeypatch.setenv("ADAPM_HOST_ROUND", "1")
    res = run_cluster(mk._worker, world=3, workers=2, mode="threads", value_lengths=mk.VPK, num_keys=mk.NUM_KEYS,
                      dtype="float32", backend="cuda")
    assert not _errs(res), "\n".join(_errs(res))


def test_many_distinct_value_lengths_cuda():
    """More distinct value lengths than size classes: keys of different lengths share a slab class (per-key length table)."""
    import torch

    nk = 150
    lens = torch.tensor([(k * 7) % 97 + 1 for k in range(nk)], dtype=torch.int64)
    res = run_cluster(loc._many_lengths_worker, world=3, workers=2, mode="threads", value_lengths=lens, num_keys=nk,
                      dtype="float32", backend="cuda")
    assert not _errs(res), "\n".join(_errs(res))
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())


@pytest.mark.parametrize("dtype", ["int64", "float64"])
def test_exact_value_types_cuda(dtype):
    """The reference's contract tests use `long` values so that sums are exact (tests/test_many_key_operations.cc:11) and
    its applications `double`: both run on the CUDA backend through the generic Pull/Push/Set kernels and the sync
    round (element-wise protocol path; the fused application kernels are float32)."""
    res = run_cluster(mk._worker, world=3, workers=2, mode="threads", value_lengths=mk.VPK, num_keys=mk.NUM_KEYS,
                      dtype=dtype, backend="cuda")
    assert not _errs(res), "\n".join(_errs(res))
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())
    world, workers = 3, 2
    res = run_cluster(dyn._dyn_worker, world=world, workers=workers, mode="threads", value_lengths=2, num_keys=20,
                      dtype=dtype, backend="cuda")
    total = world * workers * dyn.RUNS
    assert res[0][0] == [total, 2 * total], f"lost or duplicated updates: {res[0][0]}"
