"""The device-resident sync round (cross-rank barrier kernels, device-side grace period, stop/sweep agreement on the
device) on a box with ONE GPU: two rank processes share the device, their contexts are time-sliced, so the barrier
kernels of the two ranks never run at the same time and every barrier costs a few time slices - slow, but it is the
production code path (one process per rank, shm control block, VMM heaps passed as file descriptors), which the
in-process multi-rank tests of test_gpu_contract.py cannot take (they use the host-sequenced round)."""
import pytest

from harness import run_cluster
import test_contract_dynamic as dyn
import test_contract_many_keys as mk

pytestmark = pytest.mark.gpu


def _errs(res):
    return [e for r in res.values() for k, v in r.items() if k != "counters" for e in v]


def test_device_round_two_processes_one_gpu(monkeypatch):
    monkeypatch.setenv("ADAPM_DEVICE_ROUND", "1")
    res = run_cluster(mk._worker, world=2, workers=2, mode="procs", value_lengths=mk.VPK, num_keys=mk.NUM_KEYS,
                      dtype="float32", backend="cuda", timeout=240)
    assert not _errs(res), "\n".join(_errs(res))
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())
    assert sum(r["counters"]["relocations"] + r["counters"]["replica_setups"] for r in res.values()) > 0


def test_device_round_exact_updates_two_processes(monkeypatch):
    """No update lost or duplicated while a hot key relocates / replicates (test_dynamic_allocation) on the device round."""
    monkeypatch.setenv("ADAPM_DEVICE_ROUND", "1")
    world, workers = 2, 2
    res = run_cluster(dyn._dyn_worker, world=world, workers=workers, mode="procs", value_lengths=2, num_keys=20,
                      dtype="float32", backend="cuda", timeout=240)
    total = world * workers * dyn.RUNS
    assert res[0][0] == [total, 2 * total], f"lost or duplicated updates: {res[0][0]}"
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())
