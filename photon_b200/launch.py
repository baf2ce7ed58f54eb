"""Rank launcher for fault-tolerant jobs: ``python -m photon_b200.launch --nproc N [-m module | script.py] [args...]``.

``torchrun`` tears the whole job down as soon as one worker exits abnormally — the opposite of what a federation wants: the
reference's nodes are independent ``flower-client-app`` processes and the server keeps going with the ones that are left
(ref: photon/server_app.py:285,346; scripts/photon_llm_125M.sh:137-153). This launcher starts one process per GPU with the same
environment contract (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT) and simply waits: a rank that dies is reported,
the others are left alone (the control plane rules the dead rank out, photon_b200/server/control.py). The job's exit code is
rank 0's (the server).
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys


def main() -> None:
    ap = argparse.ArgumentParser(prog="python -m photon_b200.launch")
    ap.add_argument("--nproc", type=int, required=True)
    ap.add_argument("--master-addr", default="127.0.0.1")
    ap.add_argument("--master-port", type=int, default=29500)
    ap.add_argument("-m", dest="module", default=None)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    cmd = [sys.executable] + (["-m", a.module] if a.module else []) + [x for x in a.rest if x != "--"]
    if len(cmd) == 1:
        ap.error("nothing to launch")
    procs = []
    for r in range(a.nproc):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(a.nproc), LOCAL_WORLD_SIZE=str(a.nproc),
                   MASTER_ADDR=a.master_addr, MASTER_PORT=str(a.master_port))
        # a dead peer must not make NCCL's watchdog tear the survivors down: nothing is ever in flight towards it
        env.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        procs.append(subprocess.Popen(cmd, env=env))
    rc0 = procs[0].wait()          # the server decides when the job is over
    for r, p in enumerate(procs[1:], start=1):
        try:
            rc = p.wait(timeout=30.0)
        except subprocess.TimeoutExpired:     # e.g. a hung rank the server had already ruled out
            p.kill()
            rc = p.wait()
        if rc != 0:
            print(f"[launch] rank {r} exited with code {rc}", file=sys.stderr, flush=True)
    sys.exit(rc0)


if __name__ == "__main__":
    main()
