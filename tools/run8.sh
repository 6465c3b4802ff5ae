#!/bin/bash
# One 8-GPU session: parity, then the headline configurations.  Usage: bash tools/run8.sh <ngpus>
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --master-port 29511 tests/mgpu_parity.py --quick > gpurun_out/mgpu$N.log 2>&1
grep -E "FAIL|mgpu_parity|rror" gpurun_out/mgpu$N.log | tail -8
run() { # name, args...
  name=$1; shift
  $TR --master-port 29512 bench.py --gpus $N "$@" 2> gpurun_out/b${N}_$name.err | grep "^{" > gpurun_out/b${N}_$name.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/b${N}_$name.json"))
    r = d["roofline"]
    print("$name", "ms", round(d["ms_per_step"], 3), "GFLOP/s", round(d["value"]), "seq_ms", round(d["config"]["sequential_schedule_ms"], 3),
          "steps", {k: round(v, 3) for k, v in r["steps_ms"].items()}, "nvlink", r.get("nvlink", {}).get("gbs_per_direction"))
except Exception as e:
    print("$name FAILED", e)
    print(open("gpurun_out/b${N}_$name.err").read()[-1500:])
PY
}
run streams128 --steps 10 --warmup 3 --no-e2e --no-cpu
DFFT_XCHG_CTAS=64 run streams64 --steps 10 --warmup 3 --no-e2e --no-cpu
DFFT_XCHG_CTAS=200 run streams200 --steps 10 --warmup 3 --no-e2e --no-cpu
run sync --steps 10 --warmup 3 --no-e2e --no-cpu --send Sync
run a2a --steps 10 --warmup 3 --no-e2e --no-cpu --send Sync --comm All2All
run r2c --steps 10 --warmup 3 --no-e2e --no-cpu --transform r2c
if [ "$N" = "8" ]; then
  run pencil_f32 --steps 5 --warmup 3 --no-e2e --no-cpu --send Sync --decomp pencil --prec f32 --shape 2048,2048,1024 --p1 2 --p2 4
  run pencil_f64 --steps 5 --warmup 3 --no-e2e --no-cpu --send Sync --decomp pencil --p1 2 --p2 4
fi
run e2e --steps 10 --warmup 3 --no-cpu
