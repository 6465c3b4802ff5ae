#!/bin/bash
# final 8-GPU session: parity after the pruning, the BASELINE configs with the plan-time tuned schedule (wider grid),
# a narrow-exchange variant, both pencil grids, and the headline bench line with e2e.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export DFFT_PENCIL_OVERLAP=1
timeout 600 $TR --master-port 29511 tests/mgpu_parity.py --quick > gpurun_out/r02b_mgpu${N}_parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/r02b_mgpu${N}_parity.log
grep -c "^ok" gpurun_out/r02b_mgpu${N}_parity.log; grep "FAIL\|failed\|Error\|error" gpurun_out/r02b_mgpu${N}_parity.log | head -10
B="--no-e2e --no-cpu --steps 10 --warmup 3"
if [ "$N" = "8" ]; then PSHAPE="2048,2048,1024"; else PSHAPE="1024,1024,1024"; fi
P1=2; P2=$((N/2))
timeout 1500 $TR --master-port 29513 tools/multi_bench.py \
  "tuned:$B" \
  "tuned_narrow4:$B DFFT_BLOCKED=4 DFFT_XCHG_WIDE=0" \
  "tuned_bulk:$B DFFT_BULK_STORE=1" \
  "r2c_tuned:$B --transform r2c" \
  "pencil_f32_tuned:$B --decomp pencil --p1 $P1 --p2 $P2 --prec f32 --shape $PSHAPE" \
  "pencil_f32_T_tuned:$B --decomp pencil --p1 $P2 --p2 $P1 --prec f32 --shape $PSHAPE" \
  > gpurun_out/r02b_mb${N}.log 2>&1; echo "multi_bench rc=$?"
grep -v "^\[\|^\*\|^Setting\|NCCL version\|^$\|UserWarning\|e_in = " gpurun_out/r02b_mb${N}.log | cut -c1-2600
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02b_bench_n${N}.json 2> gpurun_out/r02b_bench_n${N}.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02b_bench_n${N}.json').read().strip().splitlines()[-1])
    print('bench', d['ms_per_step'], d['value'], d['e2e'], d['config']['parity']['ok'])
    for e in (d['roofline'].get('overlap_timeline') or []): print('   ', e['stream'], e['step'].ljust(16), e['begin_ms'], e['end_ms'])
except Exception as ex: print('bench parse failed', ex)
PY
