"""Register / spill budget of the hot kernels, read from the `-Xptxas -v` output of the in-tree build (no GPU
needed). A change in a shared header (protocol.h is inlined into every fused kernel) that pushes a hot kernel over
its register budget silently costs 30 % of the step time; this test makes that loud."""
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOGS = os.path.join(ROOT, "build", "adapm_b200")


def _kernels():
    logs = glob.glob(os.path.join(LOGS, "*.cu.o.log"))
    if not logs:
        from adapm_b200 import _build

        _build.build()
        logs = glob.glob(os.path.join(LOGS, "*.cu.o.log"))
    out = {}
    pat = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:ptxas info\s*:\s*Function properties[^\n]*\n"
                     r"\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n)?"
                     r"ptxas info\s*:\s*Used (\d+) registers")
    for f in logs:
        for name, stack, st, ld, regs in pat.findall(open(f).read()):
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            dem = dem.replace("(anonymous namespace)::", "").replace("adapm::cudaops::", "").replace("adapm::", "")
            dem = re.sub(r"^void ", "", dem).split("(")[0]
            out[dem] = {"regs": int(regs), "spill": int(st or 0) + int(ld or 0)}
    return out


def test_hot_kernels_stay_within_their_register_budget():
    k = _kernels()
    if not k:
        pytest.skip("no ptxas logs (extension was not built in-tree)")
    # full-register variants of the fused steps: 2 blocks x 256 threads x 128 registers = the register file, no spills
    # (sgns_step_tma_kernel<VPL, MAXREG, BULK>: the bulk-reduction variant is the default)
    for name in ("sgns_step_tma_kernel<3, 128, true>", "sgns_step_tma_kernel<2, 128, true>",
                 "sgns_step_tma_kernel<1, 128, true>",
                 "sgns_step_kernel<3, 2>", "kge_step_kernel<2, 128>", "kge_step_kernel<2, 104>", "mf_step_kernel<1>"):
        assert name in k, sorted(k)
        assert k[name]["regs"] <= 128 and k[name]["spill"] == 0, (name, k[name])
    # lean multi-GPU variant: 104 registers leave 12 K registers per SM for one block of the round kernels
    lean = k["sgns_step_tma_kernel<3, 104, true>"]
    assert lean["regs"] <= 104 and lean["spill"] <= 256, lean
    # the round's kernels must fit NEXT to two lean training blocks: meta passes (128-thread blocks, no spills) and
    # the row passes (register variant: 128 threads, TMA-engine variant: 64 threads)
    room = 65536 - 2 * 256 * 104
    for name, v in k.items():
        if name.startswith(("phase_meta_kernel", "phase_scan_kernel", "phase_b_")):
            assert v["regs"] * 128 <= room and v["spill"] == 0, (name, v)
    assert k["phase_row_kernel<float>"]["regs"] * 128 <= room, k["phase_row_kernel<float>"]
    assert k["phase_row_tma_kernel"]["regs"] * 64 <= room, k["phase_row_tma_kernel"]
    # tensor-core kernels: no spills, enough room for 1 CTA/SM with large smem tiles
    for name, v in k.items():
        if name.startswith(("gemm_nt_tcgen05", "gather_gemm_kernel")):
            assert v["spill"] == 0, (name, v)
