"""Does running the z pass and the y pass plane-group by plane-group keep the intermediate in the 126 MB L2?
Compares whole-array z+y passes with grouped launches (captured in a CUDA graph to hide launch cost)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributedfft_b200 as dfft

prec, cdt, es = dfft.F64, torch.complex128, 16
for (X, NY, NZ) in ((128, 1024, 1024), (512, 512, 512)):
    x = torch.randn(X * NY * NZ, dtype=cdt, device="cuda")
    y = torch.empty_like(x)
    s = torch.cuda.Stream()
    plane = NY * NZ

    def run(G, inplace_y):
        with torch.cuda.stream(s):
            for g0 in range(0, X, G):
                off = g0 * plane
                xin = x[off:off + G * plane]
                mid = y[off:off + G * plane]
                dfft.fft1d_contig(prec, 0, dfft.FORWARD, NZ, G * NY, mid, NZ, xin, NZ, s)
                dst = mid if inplace_y else xin
                dfft.fft1d_strided(prec, dfft.FORWARD, G, NY, NZ, dst, mid, s)

    for inplace in (True, False):
        for G in (X, 16, 8, 4, 2, 1):
            if G > X:
                continue
            run(G, inplace)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                run(G, inplace)
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            bytes_alg = 4 * X * plane * es
            print(f"{X}x{NY}x{NZ} y-{'inplace' if inplace else 'outofplace'} G={G:4d} planes ({G*plane*es/2**20:7.1f} MiB/group): {ms:7.3f} ms "
                  f"-> {bytes_alg/ms/1e6:7.0f} GB/s algorithmic (2 passes)", flush=True)
    del x, y
