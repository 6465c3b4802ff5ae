#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_axis.py tests/test_gpu_plan.py -m gpu -q --timeout=300 -k "variants or shim or ordered" > gpurun_out/r02_s1b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_s1b_pytest.log; tail -15 gpurun_out/r02_s1b_pytest.log
run() { tag=$1; shift; env "$@" timeout 200 python tools/axis_bench.py --prec f64 --sizes 512,1024 --no-r2c --tag "_r02b_$tag" > gpurun_out/r02_s1b_axis_f64_$tag.log 2>&1; echo "== $tag $@"; grep -v "^copy" gpurun_out/r02_s1b_axis_f64_$tag.log; }
run base A=1
run tma DFFT_TMA=1
run tma_p2 DFFT_TMA=1 DFFT_TMA_L2PROMO=2
run tma_p3 DFFT_TMA=1 DFFT_TMA_L2PROMO=3
run cl2 DFFT_CLUSTER=2
run cl4 DFFT_CLUSTER=4
run cl8 DFFT_CLUSTER=8
runf() { tag=$1; shift; env "$@" timeout 200 python tools/axis_bench.py --prec f32 --sizes 1024,2048 --elems 29 --no-r2c --tag "_r02b_$tag" > gpurun_out/r02_s1b_axis_f32_$tag.log 2>&1; echo "== f32 $tag $@"; grep -v "^copy" gpurun_out/r02_s1b_axis_f32_$tag.log; }
runf tma DFFT_TMA=1
runf cl4 DFFT_CLUSTER=4
DFFT_TMA=1 timeout 300 python bench.py --no-e2e --no-cpu > gpurun_out/r02_s1b_bench_tma.json 2> gpurun_out/r02_s1b_bench_tma.err; echo "bench tma rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02_s1b_bench_tma.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['config']['ms_inverse'], [(q['step'],round(q['ms'],3),round(q['gbs'])) for q in d['roofline']['all_passes']], d['config']['parity']['ok'])"
