#!/usr/bin/env python
"""JSON job runner: reads the reference's job files (/root/reference/jobs/*/slab/*.json,
schema of /root/reference/launch.py:168-247 — `size`, `global_test_settings`, `tests`, keys prefixed with `$`
survive overrides) and turns every (test, size) pair into a `torchrun ... tests/cli.py` command instead of
`mpiexec ... slab|pencil`.  MPI-only keys (`additional-flags`, host/rank files) are ignored; the coordinator rank
the reference adds for testcase 1 is not needed.  Test infrastructure.

    python tests/launch_jobs.py jobs.json --gpus 8 [--dry-run] [--max-size 512]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LONG2SHORT = {"--input-dim-x": "-nx", "--input-dim-y": "-ny", "--input-dim-z": "-nz", "--testcase": "-t", "--iterations": "-i",
              "--warmup-rounds": "-w", "--double_prec": "-d", "--cuda_aware": "-c", "--comm-method": "-comm", "--send-method": "-snd",
              "--sequence": "-s", "--opt": "-o", "--benchmark_dir": "-b", "--partition1": "-p1", "--partition2": "-p2", "--fft-dim": "-f",
              "--comm-method2": "-comm2", "--send-method2": "-snd2", "--partition": "-p"}


def commands(job: dict, gpus: int, max_size: int | None = None):
    sizes = job.get("size", [0])
    glob = {k.lstrip("$"): v for k, v in job.get("global_test_settings", {}).items()}
    for test in job.get("tests", []):
        for size in sizes:
            t = dict(test)
            t.update(glob)
            t = {LONG2SHORT.get(k.lstrip("$"), k.lstrip("$")): v for k, v in t.items()}
            name = str(t.pop("name", "slab")).lower()
            if name not in ("slab", "pencil"):
                continue  # "Reference" = MPI bandwidth micro-benchmarks: out of scope
            if size:
                dims = size if isinstance(size, list) else [size] * 3
                t["-nx"], t["-ny"], t["-nz"] = dims
            if max_size and max(int(t["-nx"]), int(t["-ny"]), int(t["-nz"])) > max_size:
                continue
            if t.get("-s") == "Y_Then_ZX":
                continue  # experimental forward-only sequence, not provided
            ranks = int(t.pop("-p", gpus)) if name == "slab" else int(t.get("-p1", 1)) * int(t.get("-p2", 1))
            ranks = min(ranks, gpus) if name == "slab" else ranks
            if ranks > gpus:
                continue
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
                   "--master-port", "29544", os.path.join(HERE, "cli.py"), name]
            for k, v in t.items():
                if isinstance(v, bool):
                    if v:
                        cmd.append(k)
                else:
                    cmd += [k, str(v)]
            yield cmd


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("job")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--max-size", type=int, default=None)
    a = ap.parse_args(argv)
    job = json.load(open(a.job))
    rc = 0
    for cmd in commands(job, a.gpus, a.max_size):
        print(" ".join(cmd), flush=True)
        if not a.dry_run:
            r = subprocess.run(cmd)
            rc |= r.returncode
    return rc


if __name__ == "__main__":
    sys.exit(main())
