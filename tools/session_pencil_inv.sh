#!/bin/bash
# Short N-GPU session for the overlapped pencil schedules (forward and inverse).  Parity of the Streams cases with the
# overlapped schedules forced (DFFT_PENCIL_OVERLAP=2: first exec untuned, then dfft_plan_tune, then again), then one
# bench configuration with the default mode: the tuning report lists the sequential and every overlapped candidate for
# both directions.  N=2 runs the 1x2 / 2x1 grids (one transposition local), N>=4 adds 2x(N/2).
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
DFFT_PENCIL_OVERLAP=2 timeout 300 $TR --master-port 29541 tests/mgpu_parity.py --only Pencil/Streams > gpurun_out/pencil_ovl_parity_n$N.log 2>&1; echo "parity rc=$?"
grep "^ok\|^FAIL\|tune\|mgpu_parity:\|rror" gpurun_out/pencil_ovl_parity_n$N.log | cut -c1-400 | head -30
B="--no-e2e --no-cpu --steps 10 --warmup 3 --decomp pencil --p1 2 --p2 $((N/2))"
timeout 600 $TR --master-port 29542 tools/multi_bench.py "pencil_f32_streams:$B --prec f32 --shape 1024,1024,512" > gpurun_out/pencil_ovl_mb_n$N.log 2>&1; echo "multi_bench rc=$?"
grep -v "^\[\|^\*\|^Setting\|NCCL version\|^$\|UserWarning\|e_in = " gpurun_out/pencil_ovl_mb_n$N.log | cut -c1-3000
