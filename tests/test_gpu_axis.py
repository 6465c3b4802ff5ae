"""Single-axis passes (what one cufftExec* of the reference's plans computes) vs the numpy oracle."""
import numpy as np
import pytest
import torch

import distributedfft_b200 as dfft
from oracle import dft_oracle as O
from common import CDT, NPC, NPR, RDT, TOL, dev, host

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", [dfft.F64, dfft.F32])
@pytest.mark.parametrize("log2n", list(range(1, 14)))
def test_contig_c2c(prec, log2n):
    n = 1 << log2n
    lines = 37 if n >= 64 else 301
    rng = np.random.default_rng(log2n)
    x = (rng.standard_normal((lines, n)) + 1j * rng.standard_normal((lines, n))).astype(NPC[prec])
    xin = dev(x)
    out = torch.empty_like(xin)
    for direction in (dfft.FORWARD, dfft.INVERSE):
        dfft.fft1d_contig(prec, 0, direction, n, lines, out, n, xin, n)
        torch.cuda.synchronize()
        ref = np.fft.fft(x.astype(np.complex128), axis=1) if direction == dfft.FORWARD else np.fft.ifft(x.astype(np.complex128), axis=1) * n
        assert O.rel_l2(host(out), ref) < TOL[prec]
    # in place
    buf = xin.clone()
    dfft.fft1d_contig(prec, 0, dfft.FORWARD, n, lines, buf, n, buf, n)
    torch.cuda.synchronize()
    assert O.rel_l2(host(buf), np.fft.fft(x.astype(np.complex128), axis=1)) < TOL[prec]


@pytest.mark.parametrize("prec", [dfft.F64, dfft.F32])
@pytest.mark.parametrize("log2n", list(range(2, 15)))
def test_contig_r2c_c2r(prec, log2n):
    n = 1 << log2n
    lines = 29
    rng = np.random.default_rng(100 + log2n)
    x = rng.standard_normal((lines, n)).astype(NPR[prec])
    nzo = n // 2 + 1
    xin = dev(x)
    spec = torch.empty((lines, nzo), dtype=CDT[prec], device="cuda")
    dfft.fft1d_contig(prec, 1, dfft.FORWARD, n, lines, spec, nzo, xin, n)
    torch.cuda.synchronize()
    ref = np.fft.rfft(x.astype(np.float64), axis=1)
    assert O.rel_l2(host(spec), ref) < TOL[prec]
    back = torch.empty_like(xin)
    dfft.fft1d_contig(prec, 2, dfft.INVERSE, n, lines, back, n, dev(ref.astype(NPC[prec])), nzo)
    torch.cuda.synchronize()
    assert O.rel_l2(host(back), x.astype(np.float64) * n) < TOL[prec]


@pytest.mark.parametrize("prec", [dfft.F64, dfft.F32])
@pytest.mark.parametrize("a,n,b", [(3, 8, 5), (2, 64, 33), (1, 128, 129), (2, 256, 16), (1, 512, 40), (2, 1024, 7), (1, 2048, 9),
                                   (1, 4096, 4), (1, 8192, 3), (5, 2, 17), (1, 16, 1000)])
def test_strided_c2c(prec, a, n, b):
    rng = np.random.default_rng(n + b)
    x = (rng.standard_normal((a, n, b)) + 1j * rng.standard_normal((a, n, b))).astype(NPC[prec])
    xin = dev(x)
    out = torch.empty_like(xin)
    for direction in (dfft.FORWARD, dfft.INVERSE):
        dfft.fft1d_strided(prec, direction, a, n, b, out, xin)
        torch.cuda.synchronize()
        ref = np.fft.fft(x.astype(np.complex128), axis=1) if direction == dfft.FORWARD else np.fft.ifft(x.astype(np.complex128), axis=1) * n
        assert O.rel_l2(host(out), ref) < TOL[prec]
    buf = xin.clone()
    dfft.fft1d_strided(prec, dfft.FORWARD, a, n, b, buf, buf)
    torch.cuda.synchronize()
    assert O.rel_l2(host(buf), np.fft.fft(x.astype(np.complex128), axis=1)) < TOL[prec]


@pytest.mark.parametrize("env", [{"DFFT_WIDE_TILES": "1"}, {"DFFT_PIPE": "1"}, {"DFFT_PIPE": "1", "DFFT_WIDE_TILES": "1"}])
def test_kernel_variants(env):
    """The alternative kernel variants (wide tiles, persistent register-prefetch) are selected by environment
    variables read once per process, so they are checked in a subprocess."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np, torch, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import distributedfft_b200 as dfft
from oracle import dft_oracle as O
worst = 0.0
for prec, cdt, tol in ((dfft.F64, np.complex128, 1e-10), (dfft.F32, np.complex64, 1e-5)):
    for (a, n, b) in ((2, 1024, 37), (1, 2048, 16), (3, 256, 40), (1, 512, 129), (2, 128, 64)):
        rng = np.random.default_rng(n)
        x = (rng.standard_normal((a, n, b)) + 1j * rng.standard_normal((a, n, b))).astype(cdt)
        xin = torch.from_numpy(x).cuda(); out = torch.empty_like(xin)
        for d in (dfft.FORWARD, dfft.INVERSE):
            dfft.fft1d_strided(prec, d, a, n, b, out, xin); torch.cuda.synchronize()
            ref = np.fft.fft(x.astype(np.complex128), axis=1) if d == dfft.FORWARD else np.fft.ifft(x.astype(np.complex128), axis=1) * n
            e = O.rel_l2(out.cpu().numpy(), ref); worst = max(worst, e / tol)
    for n in (64, 1024, 4096):
        lines = 77
        rng = np.random.default_rng(n)
        x = (rng.standard_normal((lines, n)) + 1j * rng.standard_normal((lines, n))).astype(cdt)
        xin = torch.from_numpy(x).cuda(); out = torch.empty_like(xin)
        dfft.fft1d_contig(prec, 0, dfft.FORWARD, n, lines, out, n, xin, n); torch.cuda.synchronize()
        e = O.rel_l2(out.cpu().numpy(), np.fft.fft(x.astype(np.complex128), axis=1)); worst = max(worst, e / tol)
    shape = (64, 1024, 32)
    plan = dfft.MPIcuFFT_Slab(dfft.Configurations(), dfft.Comm(), precision="double" if prec == dfft.F64 else "float", transform="c2c")
    plan.initFFT(dfft.GlobalSize(*shape), None, True)
    xc = O.complex_input(shape, dtype=cdt); outc = torch.empty(shape, dtype=torch.complex128 if prec == dfft.F64 else torch.complex64, device="cuda")
    plan.execC2C(outc, torch.from_numpy(xc).cuda(), dfft.FORWARD)
    worst = max(worst, O.rel_l2(outc.cpu().numpy(), O.fft_c2c(xc)) / tol)
print("WORST", worst)
assert worst < 1.0
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
