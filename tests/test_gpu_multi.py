"""Multi-GPU contract tests: one process per GPU (shm control block + CUDA-IPC peer-mapped heaps),
i.e. the production topology. Needs >= 2 visible GPUs (gpurun --gpus 2/4/8)."""
import pytest
import torch

from harness import run_cluster
import test_contract_dynamic as dyn
import test_contract_locality as loc
import test_contract_many_keys as mk

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                                 reason="needs at least 2 GPUs")]


def _errs(res):
    return [e for r in res.values() for k, v in r.items() if k != "counters" for e in v]


def _world(n):
    return min(n, torch.cuda.device_count())


def test_locality_api_multi_gpu():
    if torch.cuda.device_count() < 3:
        pytest.skip("the locality contract is written for exactly 3 nodes")
    res = run_cluster(loc._worker, world=3, workers=loc.NUM_LOCAL, mode="procs", value_lengths=1, num_keys=12,
                      dtype="float32", backend="cuda")
    assert not _errs(res), "\n".join(_errs(res))


def test_dynamic_allocation_multi_gpu():
    world, workers = _world(4), 2
    res = run_cluster(dyn._dyn_worker, world=world, workers=workers, mode="procs", value_lengths=2, num_keys=20,
                      dtype="float32", backend="cuda")
    total = world * workers * dyn.RUNS
    assert res[0][0] == [total, 2 * total], f"lost or duplicated updates: {res[0][0]}"
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())


def test_many_key_operations_multi_gpu():
    res = run_cluster(mk._worker, world=_world(4), workers=2, mode="procs", value_lengths=mk.VPK, num_keys=mk.NUM_KEYS,
                      dtype="float32", backend="cuda")
    assert not _errs(res), "\n".join(_errs(res))
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())
