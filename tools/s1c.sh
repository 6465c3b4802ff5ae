#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out


timeout 600 python -m pytest tests/test_gpu_axis.py tests/test_gpu_plan.py -m gpu -q --timeout=300 -k "variants or forward_inverse or r2c_c2r" > gpurun_out/r02_s1c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_s1c_pytest.log; tail -6 gpurun_out/r02_s1c_pytest.log
run() { tag=$1; sizes=$2; shift; shift; env "$@" timeout 200 python tools/axis_bench.py --prec f64 --sizes $sizes --no-r2c --tag "_r02c_$tag" > gpurun_out/r02_s1c_axis_f64_$tag.log 2>&1; echo "== $tag $@"; grep -v "^copy" gpurun_out/r02_s1c_axis_f64_$tag.log; }
run tma0 1024,2048,4096 DFFT_TMA=0
run tma1 1024,2048,4096 DFFT_TMA=1
run ring 256,512,1024 DFFT_TMA=1 DFFT_TMA_RING=1
run tma_wide 1024 DFFT_TMA=1 DFFT_WIDE_TILES=1
run swz1 512,1024 DFFT_TILE_SWZ=1
run swz2 512,1024 DFFT_TILE_SWZ=2
run swz3 512,1024 DFFT_TILE_SWZ=3
run swz2_tma0 1024 DFFT_TILE_SWZ=2 DFFT_TMA=0
run ring_swz2 1024 DFFT_TMA=1 DFFT_TMA_RING=1 DFFT_TILE_SWZ=2
run tma256 256 DFFT_TMA=1
run base256 256 DFFT_TMA=0
runf() { tag=$1; shift; env "$@" timeout 200 python tools/axis_bench.py --prec f32 --sizes 512,1024,2048 --elems 29 --no-r2c --tag "_r02c_$tag" > gpurun_out/r02_s1c_axis_f32_$tag.log 2>&1; echo "== f32 $tag $@"; grep -v "^copy" gpurun_out/r02_s1c_axis_f32_$tag.log; }
runf ring DFFT_TMA=1 DFFT_TMA_RING=1
runf base DFFT_TMA=0
# ncu: where does the time go in the register-fed and the TMA-fed 1024-point strided pass
cat > /tmp/prof2.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import distributedfft_b200 as dfft
tot = 1 << 27; n = 1024
x = torch.randn(tot, dtype=torch.complex128, device="cuda"); y = torch.empty_like(x)
s = torch.cuda.current_stream()
for mode in ("0", "1", "ring"):
    os.environ["DFFT_TMA"] = "0" if mode == "0" else "1"
    os.environ["DFFT_TMA_RING"] = "1" if mode == "ring" else "0"
    for _ in range(2):
        dfft.fft1d_strided(dfft.F64, dfft.FORWARD, tot // (n * 1024), n, 1024, y, x, s)
        oy = 128; ch = 8; nz = tot // (n * oy)
        dfft.fft1d_general(dfft.F64, dfft.FORWARD, n, oy, nz // ch, ch, y, [nz, ch, oy * nz], x, [ch, n * oy * ch, oy * ch], s)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_c2c -o gpurun_out/r02_prof_tma python /tmp/prof2.py > gpurun_out/r02_prof_tma.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_prof_tma.log
