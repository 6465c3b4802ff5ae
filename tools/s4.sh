#!/bin/bash
# GPU session on 4 (or 8) GPUs: parity (small cases + full size vs cuFFT incl. the pencil grids), peer-copy baseline,
# the BASELINE configurations through bench.py (plan-time tuned schedule) and a few fixed variants for the record.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export DFFT_PENCIL_OVERLAP=1
timeout 600 $TR --master-port 29511 tests/mgpu_parity.py --quick > gpurun_out/r02_mgpu${N}_parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/r02_mgpu${N}_parity.log
grep -c "^ok" gpurun_out/r02_mgpu${N}_parity.log; grep "FAIL\|failed\|Error\|error" gpurun_out/r02_mgpu${N}_parity.log | head -20
timeout 900 $TR --master-port 29512 tests/mgpu_parity.py --cufft > gpurun_out/r02_mgpu${N}_cufft.log 2>&1; echo "cufft rc=$?" | tee -a gpurun_out/r02_mgpu${N}_cufft.log
grep "full-size\|failed\|Error" gpurun_out/r02_mgpu${N}_cufft.log | head -20
timeout 300 python tools/peer_bw.py 256 > gpurun_out/r02_peer_bw_${N}.log 2>&1; tail -1 gpurun_out/r02_peer_bw_${N}.log
B="--no-e2e --no-cpu --steps 10 --warmup 3"
if [ "$N" = "8" ]; then PSHAPE="2048,2048,1024"; else PSHAPE="1024,1024,1024"; fi
P1=2; P2=$((N/2))
timeout 1500 $TR --master-port 29513 tools/multi_bench.py \
  "tuned:$B" \
  "sync:$B --send Sync" \
  "ovl98:$B --no-tune" \
  "staged64:$B --no-tune DFFT_STAGED=1 DFFT_PUSH_CTAS=64" \
  "r2c_tuned:$B --transform r2c" \
  "r2c_sync:$B --transform r2c --send Sync" \
  "pencil_f32_sync:$B --decomp pencil --p1 $P1 --p2 $P2 --prec f32 --shape $PSHAPE --send Sync" \
  "pencil_f32_tuned:$B --decomp pencil --p1 $P1 --p2 $P2 --prec f32 --shape $PSHAPE" \
  "pencil_f32_T_sync:$B --decomp pencil --p1 $P2 --p2 $P1 --prec f32 --shape $PSHAPE --send Sync" \
  "pencil_f32_plain:$B --decomp pencil --p1 $P1 --p2 $P2 --prec f32 --shape $PSHAPE --send Sync DFFT_BLOCKED=0" \
  "a2a:$B --send Sync --comm All2All" \
  > gpurun_out/r02_mb${N}.log 2>&1; echo "multi_bench rc=$?"
grep -v "^\[\|^\*\|^Setting\|NCCL version\|^$" gpurun_out/r02_mb${N}.log | cut -c1-1200
# headline line with e2e (NUMA-local pinned buffers), as the driver runs it
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n${N}.json').read().strip().splitlines()[-1])
    print('bench', d['ms_per_step'], d['value'], d['config'].get('tuned_schedule'), d['e2e'], d['config']['parity']['ok'])
    for e in (d['roofline'].get('overlap_timeline') or []): print('   ', e['stream'], e['step'].ljust(16), e['begin_ms'], e['end_ms'])
except Exception as ex: print('bench parse failed', ex)
PY
