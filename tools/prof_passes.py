"""Launch each pass kind a few times for one length so that ncu can capture them:
   ncu --set full --clock-control none --import-source on -k regex:fft_ -o gpurun_out/prof python tools/prof_passes.py --n 1024"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributedfft_b200 as dfft

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--elems", type=int, default=27)
ap.add_argument("--prec", default="f64")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
prec = dfft.F64 if a.prec == "f64" else dfft.F32
cdt = torch.complex128 if prec == dfft.F64 else torch.complex64
tot = 1 << a.elems
n = a.n
x = torch.randn(tot, dtype=cdt, device="cuda")
y = torch.empty_like(x)
s = torch.cuda.current_stream()
for _ in range(a.reps):
    dfft.fft1d_contig(prec, 0, dfft.FORWARD, n, tot // n, y, n, x, n, s)
    dfft.fft1d_strided(prec, dfft.FORWARD, tot // (n * 1024), n, 1024, y, x, s)
    dfft.fft1d_strided(prec, dfft.FORWARD, 1, n, tot // n, y, x, s)
    xr = x.view(torch.float64 if prec == dfft.F64 else torch.float32)
    nzo = n // 2 + 1
    lines = tot // n  # half the real data so the output fits in y
    dfft.fft1d_contig(prec, 1, dfft.FORWARD, n, lines, y, nzo, xr, n, s)
torch.cuda.synchronize()
print("done")
