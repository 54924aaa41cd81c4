"""Experiment (not collected by pytest: no test_ prefix in the file name): the device-resident round with TWO processes
that share ONE GPU (contexts are time-sliced, the barrier kernels of the two ranks never run at the same time).
    python -m pytest tests/experimental_one_gpu_procs.py -q
"""
import os

import pytest
import torch

from harness import run_cluster
import test_contract_many_keys as mk

pytestmark = pytest.mark.gpu


def test_device_round_two_processes_one_gpu(monkeypatch):
    monkeypatch.setenv("ADAPM_DEVICE_ROUND", "1")
    res = run_cluster(mk._worker, world=2, workers=2, mode="procs", value_lengths=mk.VPK, num_keys=mk.NUM_KEYS,
                      dtype="float32", backend="cuda", timeout=240)
    errs = [e for r in res.values() for k, v in r.items() if k != "counters" for e in v]
    assert not errs, "\n".join(errs)
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())
