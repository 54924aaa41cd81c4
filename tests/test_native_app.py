"""The Python-free C++ application on the core API: simple app + the reference's dynamic-allocation
contract at its original scale (100 000 asynchronous Push+Pull per worker, 3 nodes x 2 workers)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "adapm_b200", "adapm_simple")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        from adapm_b200 import _build

        _build.build()
    return EXE


def test_simple_app_native(exe):
    out = subprocess.run([exe, "-s", "2", "-t", "2", "-k", "10", "-i", "3", "-v", "2"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("iteration") == 2 * 2 * 3


@pytest.mark.parametrize("rep", range(3))
def test_dynamic_allocation_full_scale_native(exe, rep):
    out = subprocess.run([exe, "-s", "3", "-t", "2", "--stress", "100000"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "Dynamic Allocation: PASSED" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("technique", ["all", "replication_only", "relocation_only"])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_programs_native(exe, technique, seed):
    """C++ fuzz: 6 workers x 5 000 random Intent / Push / Pull / clock / WaitSync operations over 24 keys with asynchronous
    pushes: read-your-writes at every pull, exact sums on every rank at the end (scripts/sanitize.sh runs the same
    binary under ThreadSanitizer / AddressSanitizer)."""
    out = subprocess.run([exe, "--fuzz", "5000", "-s", "3", "-t", "2", "--seed", str(seed), "--techniques", technique],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "errors: PASSED" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
