#!/usr/bin/env python
"""Source hygiene checks (the reference's tests/lint.py role): python files must compile, no tabs / trailing
whitespace / over-long lines in python, C++ and CUDA sources, every CUDA/C++ file starts with a comment header.

    python scripts/lint.py            # exit code 1 when something is found
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAX_LEN = {".py": 140, ".cc": 165, ".cu": 165, ".h": 165, ".cuh": 165}
SKIP_DIRS = {".git", "build", "gpurun_out", "baseline", "__pycache__", ".pytest_cache", "data", "profiles"}


def sources():
    for dp, dns, fns in os.walk(ROOT):
        dns[:] = [d for d in dns if d not in SKIP_DIRS]
        for fn in fns:
            ext = os.path.splitext(fn)[1]
            if ext in MAX_LEN:
                yield os.path.join(dp, fn), ext


def main() -> int:
    problems = []
    for path, ext in sources():
        rel = os.path.relpath(path, ROOT)
        if ext == ".py":
            try:
                compile(open(path, encoding="utf-8").read(), path, "exec")
            except SyntaxError as e:
                problems.append(f"{rel}:{e.lineno}: does not compile: {e.msg}")
        with open(path, encoding="utf-8") as f:
            lines = f.read().split("\n")
        if ext != ".py" and lines and not lines[0].startswith(("//", "/*", "#pragma", "#include")):
            problems.append(f"{rel}:1: missing comment header")
        for i, line in enumerate(lines, 1):
            if "\t" in line:
                problems.append(f"{rel}:{i}: tab character")
            if line != line.rstrip():
                problems.append(f"{rel}:{i}: trailing whitespace")
            if len(line) > MAX_LEN[ext]:
                problems.append(f"{rel}:{i}: line longer than {MAX_LEN[ext]} characters ({len(line)})")
    for p in problems:
        print(p)
    print(f"[lint] {len(problems)} problem(s)")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
