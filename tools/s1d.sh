#!/bin/bash
# one GPU: full parity suite at the current state, single-rank hand-over layout experiment, then the ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r02_s1d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_s1d_pytest.log; tail -4 gpurun_out/r02_s1d_pytest.log
for v in 1 0; do
  DFFT_N1_LAYOUT=$v timeout 300 python bench.py --no-e2e --no-cpu > gpurun_out/r02_s1d_bench_n1layout$v.json 2> gpurun_out/r02_s1d_bench_n1layout$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02_s1d_bench_n1layout{v}.json').read().strip().splitlines()[-1])
print('N1_LAYOUT', v, 'ms', round(d['ms_per_step'],4), 'inv', round(d['config']['ms_inverse'],4), [(q['step'],round(q['ms'],3),round(q['gbs'])) for q in d['roofline']['all_passes']], d['config']['parity']['ok'], 'cufft', d['cufft_1gpu_ms'])
PY
done
DFFT_N1_LAYOUT=0 DFFT_X_SWZ=0 timeout 300 python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N1_LAYOUT 0 X_SWZ 0 ms', round(d['ms_per_step'],4), [(q['step'],round(q['ms'],3),round(q['gbs'])) for q in d['roofline']['all_passes']])"
DFFT_N1_LAYOUT=0 timeout 300 python bench.py --no-e2e --no-cpu --transform r2c --shape 1024,1024,1024 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r2c 1024^3 N1_LAYOUT 0 ms', round(d['ms_per_step'],4), 'inv', round(d['config']['ms_inverse'],4), [(q['step'],round(q['ms'],3),round(q['gbs'])) for q in d['roofline']['all_passes']])"
DFFT_N1_LAYOUT=1 timeout 300 python bench.py --no-e2e --no-cpu --transform r2c --shape 1024,1024,1024 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r2c 1024^3 N1_LAYOUT 1 ms', round(d['ms_per_step'],4), 'inv', round(d['config']['ms_inverse'],4), [(q['step'],round(q['ms'],3),round(q['gbs'])) for q in d['roofline']['all_passes']])"
bash tools/s_ncu.sh
