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


@pytest.mark.parametrize("env", [{"DFFT_WIDE_TILES": "1"}, {"DFFT_WIDE_TILES": "-1"}, {"DFFT_TMA": "1"}, {"DFFT_TMA": "1", "DFFT_WIDE_TILES": "1"},
                                 {"DFFT_TMA": "0"}, {"DFFT_TILE_SWZ": "2"}, {"DFFT_TMA": "1", "DFFT_TILE_SWZ": "1"}])
def test_kernel_variants(env, monkeypatch):
    """The alternative kernel variants (wide tiles, TMA-fed persistent kernel, tile-order blocking) are selected by
    environment variables that the launcher reads at every launch."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    worst = 0.0
    for prec in (dfft.F64, dfft.F32):
        tol = TOL[prec]
        for (a, n, b) in ((2, 1024, 37), (1, 2048, 16), (3, 256, 40), (1, 512, 129), (2, 128, 64)):
            rng = np.random.default_rng(n)
            x = (rng.standard_normal((a, n, b)) + 1j * rng.standard_normal((a, n, b))).astype(NPC[prec])
            xin = dev(x)
            out = torch.empty_like(xin)
            for d in (dfft.FORWARD, dfft.INVERSE):
                dfft.fft1d_strided(prec, d, a, n, b, out, xin)
                torch.cuda.synchronize()
                ref = np.fft.fft(x.astype(np.complex128), axis=1) if d == dfft.FORWARD else np.fft.ifft(x.astype(np.complex128), axis=1) * n
                worst = max(worst, O.rel_l2(host(out), ref) / tol)
        for n in (64, 1024, 4096):
            lines = 77
            rng = np.random.default_rng(n)
            x = (rng.standard_normal((lines, n)) + 1j * rng.standard_normal((lines, n))).astype(NPC[prec])
            xin = dev(x)
            out = torch.empty_like(xin)
            dfft.fft1d_contig(prec, 0, dfft.FORWARD, n, lines, out, n, xin, n)
            torch.cuda.synchronize()
            worst = max(worst, O.rel_l2(host(out), np.fft.fft(x.astype(np.complex128), axis=1)) / tol)
        shape = (64, 1024, 32)
        plan = dfft.MPIcuFFT_Slab(dfft.Configurations(), dfft.Comm(), precision="double" if prec == dfft.F64 else "float", transform="c2c")
        plan.initFFT(dfft.GlobalSize(*shape), None, True)
        xc = O.complex_input(shape, dtype=NPC[prec])
        outc = torch.empty(shape, dtype=CDT[prec], device="cuda")
        plan.execC2C(outc, dev(xc), dfft.FORWARD)
        worst = max(worst, O.rel_l2(host(outc), O.fft_c2c(xc)) / tol)
        plan.destroy()
    assert worst < 1.0
