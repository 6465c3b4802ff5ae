#!/bin/bash
# One GPU: the committed ncu evidence for the bench command — launch list of `bench.py` itself and a --set full capture
# of its three pass kernels (and of the 1024-point / f32 / R2C kernels of the multi-GPU configs).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
grep -c fft_ gpurun_out/r02_bench_launches.csv
cat > /tmp/prof_bench.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import distributedfft_b200 as dfft
# the bench workload: 512^3 complex-double slab plan on one GPU, forward x2 + inverse x1
shape = (512, 512, 512)
plan = dfft.MPIcuFFT_Slab(dfft.Configurations(), dfft.Comm(), precision="double", transform="c2c")
plan.initFFT(dfft.GlobalSize(*shape), None, True)
x = torch.randn(shape, dtype=torch.complex128, device="cuda"); out = torch.empty_like(x); back = torch.empty_like(x)
for _ in range(2):
    plan.execC2C(out, x, dfft.FORWARD)
plan.execC2C(back, out, dfft.INVERSE)
plan.destroy(); del x, out, back
# R2C plan 1024^3 (config 5 on one GPU): z (R2C), y, x
shape = (1024, 1024, 1024)
plan = dfft.MPIcuFFT_Slab(dfft.Configurations(), dfft.Comm(), precision="double", transform="r2c")
plan.initFFT(dfft.GlobalSize(*shape), None, True)
x = torch.randn(shape, dtype=torch.float64, device="cuda"); out = torch.empty((1024, 1024, 513), dtype=torch.complex128, device="cuda")
plan.execR2C(out, x); plan.execR2C(out, x)
plan.destroy(); del x, out
# f32 pencil-like passes at 2048 points
tot = 1 << 28; n = 2048
x = torch.randn(tot, dtype=torch.complex64, device="cuda"); y = torch.empty_like(x); s = torch.cuda.current_stream()
for _ in range(2):
    dfft.fft1d_strided(dfft.F32, dfft.FORWARD, tot // (n * 1024), n, 1024, y, x, s)
    oy = 128; ch = 8; nz = tot // (n * oy)
    dfft.fft1d_general(dfft.F32, dfft.FORWARD, n, oy, nz // ch, ch, y, [nz, ch, oy * nz], x, [ch, n * oy * ch, oy * ch], s)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_ -o gpurun_out/r02_prof_final python /tmp/prof_bench.py > gpurun_out/r02_prof_final.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r02_prof_final.log
