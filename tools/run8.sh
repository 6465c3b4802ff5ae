#!/bin/bash
# One short multi-GPU session (an N-GPU box is charged N x wall time): parity, the headline configurations in a
# single torchrun, then full-size property checks.  Usage: bash tools/run8.sh <ngpus>
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 240 $TR --master-port 29511 tests/mgpu_parity.py --quick > gpurun_out/mgpu$N.log 2>&1
grep -E "FAIL|mgpu_parity|rror" gpurun_out/mgpu$N.log | tail -8
COMMON="--steps 10 --warmup 3 --no-e2e --no-cpu"
CFG=("sync:$COMMON --send Sync" "streams48:$COMMON DFFT_XCHG_CTAS=48" "streams72:$COMMON DFFT_XCHG_CTAS=72" "streams96:$COMMON DFFT_XCHG_CTAS=96" "r2c_sync:$COMMON --send Sync --transform r2c" "r2c_streams72:$COMMON --transform r2c DFFT_XCHG_CTAS=72")
if [ "$N" = "8" ]; then
  CFG+=("pencil_f32:--steps 5 --warmup 3 --no-e2e --no-cpu --send Sync --decomp pencil --prec f32 --shape 2048,2048,1024 --p1 2 --p2 4")
fi
timeout 420 $TR --master-port 29512 tools/multi_bench.py "${CFG[@]}" 2> gpurun_out/mb$N.err | tee gpurun_out/mb$N.log
tail -3 gpurun_out/mb$N.err
timeout 240 $TR --master-port 29513 tests/mgpu_parity.py --full 2>&1 | grep -E "full-size|FAIL|rror" | tail -8 | tee gpurun_out/mgpu${N}_full.log
