"""Several bench.py configurations in ONE torchrun (one torch import / process-group setup instead of one per
configuration — an 8-GPU box is charged 8x per second).  Usage:
    torchrun --nproc-per-node 8 tools/multi_bench.py <name>:<bench args> [<name>:<bench args> ...]
Each configuration prints bench.py's JSON line prefixed by its name on rank 0."""
import io
import os
import sys
from contextlib import redirect_stdout

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.makedirs("gpurun_out", exist_ok=True)
    for spec in sys.argv[1:]:
        name, _, rest = spec.partition(":")
        argv = ["--gpus", str(world)] + rest.split()
        envs = [a for a in argv if "=" in a and a.split("=")[0].isupper()]
        for e in envs:
            argv.remove(e)
            k, v = e.split("=", 1)
            os.environ[k] = v
        buf = io.StringIO()
        try:
            with redirect_stdout(buf):
                bench.main(argv)
            line = buf.getvalue().strip().splitlines()[-1] if buf.getvalue().strip() else ""
            if rank == 0:
                open(f"gpurun_out/mb{world}_{name}.json", "w").write(line + "\n")
                import json
                d = json.loads(line)
                r = d["roofline"]
                par = d["config"].get("parity") or {}
                print(name, "ms", round(d["ms_per_step"], 3), "inv_ms", round(d["config"]["ms_inverse"], 3), "GFLOP/s", round(d["value"]), "seq_ms", round(d["config"]["sequential_schedule_ms"], 3),
                      "parity", par.get("ok"), (par.get("small_grid") or {}).get("rel_l2_forward_max_over_ranks"),
                      "tuned", d["config"].get("tuned_schedule"),
                      "steps", {k: round(v, 3) for k, v in r["steps_ms"].items()}, "nvlink", [x.get("gbs_per_direction") for x in (r["nvlink"] if isinstance(r.get("nvlink"), list) else [r.get("nvlink", {})])], flush=True)
        except (Exception, SystemExit) as ex:  # keep going: the box is expensive
            if rank == 0:
                print(name, "FAILED", repr(ex), buf.getvalue()[-600:], flush=True)
        for e in envs:
            os.environ.pop(e.split("=", 1)[0], None)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
