"""x pass with far rows: is it the TLB?  Compares (a) out-of-place [Nx][B] -> [Nx][B], (b) in place,
(c) blocked input layout [B/CH][Nx][CH] -> [Nx][B], and the y pass that produces the blocked layout."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributedfft_b200 as dfft


def timeit(fn, iters=8, warm=2):
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


prec, es = dfft.F64, 16
s = torch.cuda.current_stream()
for (NX, NYP, NZ) in ((1024, 128, 1024), (1024, 256, 1024), (512, 512, 512)):
    B = NYP * NZ
    tot = NX * B
    x = torch.randn(tot, dtype=torch.complex128, device="cuda")
    y = torch.empty_like(x)
    gb = 2 * tot * es / 1e6
    t = timeit(lambda: dfft.fft1d_strided(prec, dfft.FORWARD, 1, NX, B, y, x, s))
    print(f"x pass {NX} x (B={B}) out-of-place        : {t:7.3f} ms {gb/t:6.0f} GB/s", flush=True)
    t = timeit(lambda: dfft.fft1d_strided(prec, dfft.FORWARD, 1, NX, B, x, x, s))
    print(f"x pass {NX} x (B={B}) in place            : {t:7.3f} ms {gb/t:6.0f} GB/s", flush=True)
    for CH in (16, 64, 256):
        # in: [B/CH][NX][CH]: a1 = chunk (stride NX*CH), n stride CH;  out: [NX][B]: a1 stride CH, n stride B
        t = timeit(lambda: dfft.fft1d_general(prec, dfft.FORWARD, NX, 1, B // CH, CH, y, (0, CH, B), x, (0, NX * CH, CH), s))
        print(f"x pass {NX} blocked-in CH={CH:4d} -> [Nx][B]     : {t:7.3f} ms {gb/t:6.0f} GB/s", flush=True)
    # y pass on nxp planes: [nxp][NY][NZ] -> blocked layout of a [NX][NYP][NZ] slot (one destination, y range = NYP)
    NY = NYP
    nxp = tot // (NY * NZ)
    if NY >= 128:
        t = timeit(lambda: dfft.fft1d_strided(prec, dfft.FORWARD, nxp, NY, NZ, y, x, s))
        print(f"y pass {NY} ({nxp} planes) plain              : {t:7.3f} ms {gb/t:6.0f} GB/s", flush=True)
        for CH in (16, 64, 256):
            # in: a0 = x plane (NY*NZ), a1 = z chunk (CH), n = y (NZ);  out: (y*(NZ/CH) + zc) * NX*CH + x*CH + zi
            t = timeit(lambda: dfft.fft1d_general(prec, dfft.FORWARD, NY, nxp, NZ // CH, CH, y, (CH, nxp * CH, (NZ // CH) * nxp * CH), x, (NY * NZ, CH, NZ), s))
            print(f"y pass {NY} -> blocked CH={CH:4d}                 : {t:7.3f} ms {gb/t:6.0f} GB/s", flush=True)
    del x, y
