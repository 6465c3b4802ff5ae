#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="--no-e2e --no-cpu --steps 10 --warmup 3"
timeout 1500 $TR --master-port 29513 tools/multi_bench.py \
  "tuned:$B" \
  "ovl98:$B --no-tune DFFT_XCHG_CTAS=98" \
  "ovl148:$B --no-tune DFFT_XCHG_CTAS=148" \
  "ovl_full_swap:$B --no-tune DFFT_XCHG_CTAS=0 DFFT_OVL_PRIO_SWAP=1" \
  "ovl148_swap:$B --no-tune DFFT_XCHG_CTAS=148 DFFT_OVL_PRIO_SWAP=1" \
  "ovl_full_swap_c8:$B --no-tune DFFT_XCHG_CTAS=0 DFFT_OVL_PRIO_SWAP=1 DFFT_OVL_CHUNKS=8" \
  "ovl_full_swap_g1:$B --no-tune DFFT_XCHG_CTAS=0 DFFT_OVL_PRIO_SWAP=1 DFFT_OVL_GROUPS=1" \
  "staged32:$B --no-tune DFFT_STAGED=1" \
  "staged64:$B --no-tune DFFT_STAGED=1 DFFT_PUSH_CTAS=64" \
  "staged16:$B --no-tune DFFT_STAGED=1 DFFT_PUSH_CTAS=16" \
  "staged32_c8:$B --no-tune DFFT_STAGED=1 DFFT_OVL_CHUNKS=8" \
  "staged32_g2:$B --no-tune DFFT_STAGED=1 DFFT_OVL_GROUPS=2" \
  "r2c_staged32:$B --no-tune --transform r2c DFFT_STAGED=1" \
  "r2c_tuned:$B --transform r2c" \
  "r2c_ovl_full_swap:$B --no-tune --transform r2c DFFT_XCHG_CTAS=0 DFFT_OVL_PRIO_SWAP=1" \
  > gpurun_out/r02_mb${N}c.log 2>&1; echo "multi_bench rc=$?"
grep -v "^\[\|^\*\|^Setting\|NCCL version\|^$" gpurun_out/r02_mb${N}c.log | cut -c1-330
for f in tuned ovl98 ovl148 ovl_full_swap staged32 staged32_c8 r2c_staged32; do python - "$f" "$N" <<'PY'
import json,sys
name=sys.argv[1]; n=sys.argv[2]
try:
    d=json.loads(open(f'gpurun_out/mb{n}_{name}.json').read().strip().splitlines()[-1])
    print(name, d['ms_per_step'], d['config'].get('tuned_schedule'))
    for e in d['roofline']['overlap_timeline']: print('   ', e['stream'], e['step'].ljust(16), e['begin_ms'], e['end_ms'], round(e['end_ms']-e['begin_ms'],3))
except Exception as ex: print(name, 'no timeline', ex)
PY
done
