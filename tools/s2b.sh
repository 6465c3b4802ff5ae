#!/bin/bash
# GPU session 2b (two GPUs): per-SM NVLink push rates (st.global vs cp.async.bulk), then the slab at N=2 with the
# exchange pass spread over all SMs at a small footprint.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
timeout 300 tools/micro/nvlink_micro > gpurun_out/r02_nvlink_micro.csv 2> gpurun_out/r02_nvlink_micro.err; echo "micro rc=$?"
python - <<'PY'
import csv
rows=list(csv.DictReader(open('gpurun_out/r02_nvlink_micro.csv')))
best={}
for r in rows:
    k=(r['target'],r['mode'],int(r['grid']))
    v=float(r['GBps'])
    if k not in best or v>best[k][0]: best[k]=(v,r['chunk_bytes'],r['depth'])
for k in sorted(best): print(k, best[k])
PY
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="--no-e2e --no-cpu --steps 10 --warmup 3"
timeout 1500 $TR --master-port 29513 tools/multi_bench.py \
  "sync:$B --send Sync" \
  "streams96:$B" \
  "streams148:$B DFFT_XCHG_CTAS=148" \
  "streams148_bulk:$B DFFT_XCHG_CTAS=148 DFFT_BULK_STORE=1" \
  "streams222:$B DFFT_XCHG_CTAS=222" \
  "streams296:$B DFFT_XCHG_CTAS=296" \
  "streams148_g2:$B DFFT_XCHG_CTAS=148 DFFT_OVL_GROUPS=2" \
  "streams148_g8c8:$B DFFT_XCHG_CTAS=148 DFFT_OVL_GROUPS=8 DFFT_OVL_CHUNKS=8" \
  "streams148_c2:$B DFFT_XCHG_CTAS=148 DFFT_OVL_CHUNKS=2" \
  "streams148_notma:$B DFFT_XCHG_CTAS=148 DFFT_TMA=0" \
  "sync_bulk:$B --send Sync DFFT_BULK_STORE=1" \
  "r2c_sync:$B --transform r2c --send Sync" \
  "r2c_streams148:$B --transform r2c DFFT_XCHG_CTAS=148" \
  "r2c_streams296:$B --transform r2c DFFT_XCHG_CTAS=296" \
  > gpurun_out/r02_mb${N}b.log 2>&1; echo "multi_bench rc=$?"
grep -v "^\[\|^\*\|^Setting\|NCCL version\|^$" gpurun_out/r02_mb${N}b.log | cut -c1-330
for f in streams148 streams296 r2c_streams148; do python - "$f" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/mb2_{name}.json').read().strip().splitlines()[-1])
    print(name, d['ms_per_step'])
    for e in d['roofline']['overlap_timeline']: print('   ', e['stream'], e['step'].ljust(16), e['begin_ms'], e['end_ms'], round(e['end_ms']-e['begin_ms'],3))
except Exception as ex: print(name, 'no timeline', ex)
PY
done
