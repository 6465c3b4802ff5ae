#!/bin/bash
# One short multi-GPU session (an N-GPU box is charged N x wall time): parity, then the headline configurations
# in a single torchrun.  Usage: bash tools/run8.sh <ngpus>
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 240 $TR --master-port 29511 tests/mgpu_parity.py --quick > gpurun_out/mgpu$N.log 2>&1
grep -E "FAIL|mgpu_parity|rror" gpurun_out/mgpu$N.log | tail -8
COMMON="--steps 10 --warmup 3 --no-e2e --no-cpu"
CFG=("streams:$COMMON" "sync:$COMMON --send Sync" "streams64:$COMMON DFFT_XCHG_CTAS=64" "streams200:$COMMON DFFT_XCHG_CTAS=200" "a2a:$COMMON --send Sync --comm All2All" "r2c:$COMMON --transform r2c")
if [ "$N" = "8" ]; then
  CFG+=("pencil_f32:--steps 5 --warmup 3 --no-e2e --no-cpu --send Sync --decomp pencil --prec f32 --shape 2048,2048,1024 --p1 2 --p2 4")
fi
CFG+=("e2e:--steps 5 --warmup 3 --no-cpu")
timeout 420 $TR --master-port 29512 tools/multi_bench.py "${CFG[@]}" 2> gpurun_out/mb$N.err | tee gpurun_out/mb$N.log
tail -5 gpurun_out/mb$N.err
