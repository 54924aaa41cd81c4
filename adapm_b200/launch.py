"""Job launcher (the reference's ``tracker/dmlc_local.py`` / ``dmlc_mpi.py`` equivalent).

    python -m adapm_b200.launch -s N [--backend cpu|cuda] <script.py | -m module> [args...]

Starts N ranks on this host (one per GPU for ``--backend cuda``) with ``RANK / LOCAL_RANK /
WORLD_SIZE / MASTER_ADDR / MASTER_PORT / ADAPM_JOB`` (plus the reference's ``DMLC_NUM_SERVER /
DMLC_ROLE / DMLC_RANK / DMLC_PS_ROOT_URI / DMLC_PS_ROOT_PORT`` for scripts that read them) and no
scheduler process: ranks rendezvous through a POSIX-shm control block. A rank that exits with
code 254 is restarted (``keepalive`` of tracker/dmlc_local.py:15-26). ``torchrun`` and ``mpirun``
work too - the package also reads ``OMPI_COMM_WORLD_*``, ``PMI_*`` and ``SLURM_*``.
The fabric is single-node (shm + CUDA IPC over NVLink/NVSwitch): one 8xB200 box is the target.
"""
from __future__ import annotations

import argparse
import os
import socket
import subprocess
import sys
import threading
import uuid


def _free_port() -> int:
    from .utils.net import get_available_port

    return get_available_port("127.0.0.1")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="adapm_b200.launch", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-s", "--num-servers", type=int, required=True, help="number of ranks (= PS nodes = GPUs)")
    ap.add_argument("-n", "--num-workers", type=int, default=0, help="accepted for parity with dmlc_local.py (ignored)")
    ap.add_argument("--backend", default=None, choices=["cpu", "cuda"])
    ap.add_argument("--log-dir", default=None, help="write one log file per rank")
    ap.add_argument("--env", action="append", default=[], help="extra KEY=VALUE for every rank")
    ap.add_argument("-m", dest="module", default=None, help="run a module instead of a script")
    ap.add_argument("command", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    if a.module is None and not a.command:
        ap.error("nothing to launch")
    cmd = [sys.executable] + (["-m", a.module] if a.module else []) + a.command
    port = _free_port()
    job = f"L{os.getpid()}_{uuid.uuid4().hex[:8]}"
    codes = [None] * a.num_servers

    def run(rank: int) -> None:
        env = dict(os.environ)
        env.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(a.num_servers),
                    "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "ADAPM_JOB": job,
                    "DMLC_NUM_SERVER": str(a.num_servers), "DMLC_NUM_WORKER": "0", "DMLC_ROLE": "server",
                    "DMLC_RANK": str(rank), "DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(port)})
        if a.backend:
            env["ADAPM_BACKEND"] = a.backend
        for kv in a.env:
            k, _, v = kv.partition("=")
            env[k] = v
        while True:
            out = open(os.path.join(a.log_dir, f"rank{rank}.log"), "ab") if a.log_dir else None
            rc = subprocess.call(cmd, env=env, stdout=out, stderr=subprocess.STDOUT if out else None)
            if out:
                out.close()
            if rc != 254:  # keepalive: restart on 254
                codes[rank] = rc
                return

    if a.log_dir:
        os.makedirs(a.log_dir, exist_ok=True)
    ths = [threading.Thread(target=run, args=(r,)) for r in range(a.num_servers)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    bad = [(r, c) for r, c in enumerate(codes) if c != 0]
    if bad:
        print(f"[launch] ranks failed: {bad}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
