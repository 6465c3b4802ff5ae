#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r02_s1e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_s1e_pytest.log; tail -6 gpurun_out/r02_s1e_pytest.log
timeout 300 python tools/axis_bench.py --prec f64 --sizes 256,512,1024,2048 --tag "_r02e" > gpurun_out/r02_s1e_axis_f64.log 2>&1; grep "r2c\|contig" gpurun_out/r02_s1e_axis_f64.log
timeout 300 python tools/axis_bench.py --prec f32 --sizes 512,1024,2048 --elems 29 --tag "_r02e" > gpurun_out/r02_s1e_axis_f32.log 2>&1; grep "r2c\|contig" gpurun_out/r02_s1e_axis_f32.log
timeout 300 python bench.py --no-e2e --no-cpu --transform r2c --shape 1024,1024,1024 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r2c 1024^3 ms', round(d['ms_per_step'],4), 'inv', round(d['config']['ms_inverse'],4), [(q['step'],round(q['ms'],3),round(q['gbs'])) for q in d['roofline']['all_passes']], 'cufft', d['cufft_1gpu_ms'], d['config']['parity']['ok'])"
timeout 300 python bench.py > gpurun_out/r02_s1e_bench.json 2> gpurun_out/r02_s1e_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_s1e_bench.json').read().strip().splitlines()[-1]); print('bench ms', round(d['ms_per_step'],4), 'inv', round(d['config']['ms_inverse'],4), [(q['step'],round(q['ms'],3),round(q['gbs'])) for q in d['roofline']['all_passes']], 'traffic', d['roofline']['traffic'], 'e2e', d['e2e']['ms_per_step'], 'cufft', d['cufft_1gpu_ms'])"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-600
