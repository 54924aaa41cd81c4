#!/usr/bin/env python
"""Registers / spills of every kernel from the ptxas -v logs of the last build (build/adapm_b200/*.o.log)."""
import glob, re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob("build/adapm_b200/cuda_*.o.log")):
    txt = open(f).read()
    for e in re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]:
        name = e.split("'")[0]
        used = re.search(r"Used (\d+) registers", e)
        spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", e)
        stack = re.search(r"(\d+) bytes stack frame", e)
        dn = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dn = re.sub(r"adapm::|\(anonymous namespace\)::|cudaops::", "", dn)
        dn = dn.split("(")[0]
        if pat and not re.search(pat, dn):
            continue
        g = lambda m, i=1: m.group(i) if m else "?"
        print(f"{g(used):>4s} regs  stack {g(stack):>4s}  spill {g(spill)}/{g(spill, 2)}  {dn}")
