"""Per-pass HBM throughput of the FFT kernels (read + write of the array per pass = algorithmic bytes),
CUDA-event timed.  Usage: python tools/axis_bench.py [--prec f64|f32] [--elems LOG2]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributedfft_b200 as dfft


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prec", default="f64")
    ap.add_argument("--elems", type=int, default=28)  # total complex elements = 2^elems (f64: 4 GiB)
    ap.add_argument("--sizes", default="128,256,512,1024,2048,4096")
    ap.add_argument("--tag", default="")
    ap.add_argument("--no-r2c", action="store_true")
    args = ap.parse_args()
    prec = dfft.F64 if args.prec == "f64" else dfft.F32
    cdt = torch.complex128 if prec == dfft.F64 else torch.complex64
    tot = 1 << args.elems
    es = 16 if prec == dfft.F64 else 8
    x = torch.randn(tot, dtype=cdt, device="cuda")
    y = torch.empty_like(x)
    s = torch.cuda.current_stream()
    res = []
    t = timeit(lambda: y.copy_(x))
    print(f"copy: {2*tot*es/t/1e6:.0f} GB/s")
    for n in [int(v) for v in args.sizes.split(",")]:
        lines = tot // n
        t = timeit(lambda: dfft.fft1d_contig(prec, 0, dfft.FORWARD, n, lines, y, n, x, n, s))
        gbs = 2 * tot * es / t / 1e6
        res.append(dict(kind="contig", n=n, ms=t, gbs=gbs))
        print(f"contig  n={n:5d} lines={lines:8d} {t:8.3f} ms {gbs:7.0f} GB/s", flush=True)
        # strided, y-like: [a][n][b] with b = 1024 (plane) and x-like: a = 1, b = tot/n
        for label, b in (("tiled-y", 1024), ("tiled-x", tot // n)):
            a = tot // (n * b)
            if a < 1:
                continue
            t = timeit(lambda: dfft.fft1d_strided(prec, dfft.FORWARD, a, n, b, y, x, s))
            gbs = 2 * tot * es / t / 1e6
            res.append(dict(kind=label, n=n, ms=t, gbs=gbs))
            print(f"{label} n={n:5d} a={a:6d} b={b:8d} {t:8.3f} ms {gbs:7.0f} GB/s", flush=True)
        # the slab's x pass on the blocked hand-over layout [nz/CH][n][oy][CH] -> out [n][oy][nz] (oy = 128)
        for ch in (4, 8):
            oy = 128
            nz = tot // (n * oy)
            if nz < 4 * ch:
                continue
            ins = [ch, n * oy * ch, oy * ch]
            outs = [nz, ch, oy * nz]
            t = timeit(lambda: dfft.fft1d_general(prec, dfft.FORWARD, n, oy, nz // ch, ch, y, outs, x, ins, s))
            gbs = 2 * tot * es / t / 1e6
            res.append(dict(kind=f"tiled-xb{ch}", n=n, ms=t, gbs=gbs))
            print(f"tiled-xb{ch} n={n:5d} oy={oy} nz={nz:6d} {t:8.3f} ms {gbs:7.0f} GB/s", flush=True)
        if not args.no_r2c:
            xr = x.view(torch.float64 if prec == dfft.F64 else torch.float32)
            nzo = n // 2 + 1
            lines_r = (2 * tot) // n
            out = torch.empty(lines_r * nzo, dtype=cdt, device="cuda") if lines_r * nzo <= 2 * tot else None
            if out is not None:
                t = timeit(lambda: dfft.fft1d_contig(prec, 1, dfft.FORWARD, n, lines_r, out, nzo, xr, n, s))
                gbs = (2 * tot * es / 2 + lines_r * nzo * es) / t / 1e6
                print(f"r2c     n={n:5d} lines={lines_r:8d} {t:8.3f} ms {gbs:7.0f} GB/s", flush=True)
                res.append(dict(kind="r2c", n=n, ms=t, gbs=gbs))
                del out
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/axis_bench_{args.prec}{args.tag}.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
