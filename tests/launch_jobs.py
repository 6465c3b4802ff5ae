#!/usr/bin/env python
"""JSON job runner: reads the reference's job files (/root/reference/jobs/*/slab/*.json,
schema of /root/reference/launch.py:168-247 — `size`, `global_test_settings`, `tests`, keys prefixed with `$`
survive overrides) and turns every (test, size) pair into a `torchrun ... tests/cli.py` command instead of
`mpiexec ... slab|pencil`.  MPI-only keys (`additional-flags`, host/rank files) are ignored; the coordinator rank
the reference adds for testcase 1 is not needed.  Test infrastructure.

    python tests/launch_jobs.py jobs.json --gpus 8 [--dry-run] [--max-size 512]
    torchrun --nproc-per-node 2 tests/launch_jobs.py tests/jobs/validation_slab.json --in-process --report gpurun_out/validation.json

`--in-process` (under torchrun): every (test, size) pair that uses all ranks of the launch runs inside THIS process
group through `cli.main(argv)` — one torch import and one NCCL bootstrap for the whole sweep instead of one per pair —
and rank 0 writes a result table.  `tests/jobs/*.json` are this repo's own job files in the reference's schema (the
sweeps of /root/reference/jobs/bwunicluster/{slab,pencil}/validation.json, which does not exist on a GPU box).
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
              "--comm-method1": "-comm1", "--send-method1": "-snd1", "--comm-method2": "-comm2", "--send-method2": "-snd2", "--partition": "-p"}


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


def run_in_process(jobs, max_size, report):
    """All commands of `jobs` whose rank count equals the launch's world size, through cli.main in this process."""
    import time
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    import cli
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cli.DEVICE_CACHE = {}  # the analytic testcase-4 fields depend on (shape, block) only
    rows, rc = [], 0
    todo = []
    for path in jobs:
        for cmd in commands(json.load(open(path)), world, max_size):
            if f"--nproc-per-node={world}" in cmd:
                argv = cmd[cmd.index(os.path.join(HERE, "cli.py")) + 1:]
                opt = dict(zip(argv[1:], argv[2:]))  # flag -> following token
                todo.append(((int(opt["-nx"]) * int(opt["-ny"]) * int(opt["-nz"]), argv[0], opt.get("-p1", ""), opt.get("-s", "")), len(todo), path, argv))
    for _, _, path, argv in sorted(todo):  # size-major, so consecutive cases share the cached input fields
        t0 = time.time()
        try:
            status = cli.main(argv)
            res = dict(cli.main.last)
        except (Exception, SystemExit) as ex:  # noqa: BLE001 - report and keep sweeping
            status, res = 1, {"status": 1, "error": repr(ex)}
        rc |= int(status != 0)
        rows.append({"job": os.path.basename(path), "argv": " ".join(argv), "seconds": round(time.time() - t0, 2), **res})
        if rank == 0:
            print("SWEEP", "ok  " if status == 0 else "FAIL", rows[-1]["argv"], res.get("rel"), res.get("error", ""), flush=True)
    if rank == 0:
        bad = [r for r in rows if r["status"] != 0]
        print(f"SWEEP summary: {len(rows) - len(bad)}/{len(rows)} cases within tolerance on {world} GPUs", flush=True)
        if report:
            json.dump({"world": world, "cases": rows, "failed": len(bad)}, open(report, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()
    return rc


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("job", nargs="+")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--max-size", type=int, default=None)
    ap.add_argument("--in-process", action="store_true", help="under torchrun: run every pair inside this process group")
    ap.add_argument("--report", default=None, help="--in-process: JSON result table written by rank 0")
    a = ap.parse_args(argv)
    if a.in_process:
        return run_in_process(a.job, a.max_size, a.report)
    rc = 0
    for path in a.job:
        rc |= run_commands(json.load(open(path)), a)
    return rc


def run_commands(job, a):
    rc = 0
    for cmd in commands(job, a.gpus, a.max_size):
        print(" ".join(cmd), flush=True)
        if not a.dry_run:
            r = subprocess.run(cmd)
            rc |= r.returncode
    return rc


if __name__ == "__main__":
    sys.exit(main())
