"""One process, two GPUs: the exchanging y pass kernel (TILED C2C storing its output rows into ANOTHER GPU's
memory over NVLink) launched through the C ABI's single-axis entry point, with no rendezvous and no second
process — so it can run under `ncu` (kernel replay is safe) to read the NVLink byte counters of the very kernel
that does the transposition.  Input on cuda:0, output on cuda:1 in the blocked hand-over layout
[nz/CH][nx][ny][CH] (rows of one destination adjacent).  Prints the CUDA-event GB/s as well.
    python tools/nvlink_probe.py [n=1024] [ch=8] [planes=64] [prec=f64]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributedfft_b200 as dfft


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    planes = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    f64 = (sys.argv[4] if len(sys.argv) > 4 else "f64") == "f64"
    prec = dfft.F64 if f64 else dfft.F32
    cdt = torch.complex128 if f64 else torch.complex64
    es = 16 if f64 else 8
    nz = 1024
    assert torch.cuda.device_count() >= 2
    torch.cuda.set_device(0)
    x = torch.randn(planes * n * nz, dtype=cdt, device="cuda:0")
    y = torch.zeros(planes * n * nz, dtype=cdt, device="cuda:1")
    y[:16].copy_(x[:16])  # makes torch enable peer access 0 <-> 1
    torch.cuda.synchronize()
    s = torch.cuda.current_stream()
    # in [planes][n][nz]: a0 = plane, a1 = z chunk, n = y, b = CH;  out [nz/CH][planes][n][CH]
    ins = [n * nz, ch, nz]
    outs = [n * ch, planes * n * ch, ch]
    if os.environ.get("DFFT_XCHG_WIDE", "1") != "0":
        os.environ.setdefault("DFFT_WIDE_TILES", "1")

    def run():
        dfft.fft1d_general(prec, dfft.FORWARD, n, planes, nz // ch, ch, y, outs, x, ins, s)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    b = planes * n * nz * es
    print(f"nvlink_probe n={n} ch={ch} planes={planes} {'f64' if f64 else 'f32'}: {ms:.3f} ms, {b / ms / 1e6:.0f} GB/s stored into the peer ({b / 1e9:.3f} GB per launch)")
    # verify against a local run
    ref = torch.empty_like(x)
    dfft.fft1d_general(prec, dfft.FORWARD, n, planes, nz // ch, ch, ref, outs, x, ins, s)
    torch.cuda.synchronize()
    ok = torch.equal(ref.cpu(), y.cpu())
    print("peer result equals local result:", ok)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
