"""Pins the CPU oracle (oracle/dft_oracle.py): against an O(N^2) long-double DFT (oracle/naive_dft.c),
against the closed form of the reference's testcase 4, and against the committed golden vectors."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import dft_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def naive():
    path = os.path.join(ROOT, "oracle", "_ref", "libnaive_dft.so")
    if not os.path.exists(path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", path, os.path.join(ROOT, "oracle", "naive_dft.c"), "-lm"], check=True)
    lib = C.CDLL(path)
    lib.naive_dft_lines.restype = None
    lib.naive_dft_lines.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
    return lib


def naive_fftn(lib, x, sign=-1, axes=(0, 1, 2)):
    """separable 3D DFT by the naive 1D routine, one axis at a time"""
    a = np.ascontiguousarray(x, dtype=np.complex128)
    for ax in axes:
        moved = np.ascontiguousarray(np.moveaxis(a, ax, -1))
        out = np.empty_like(moved)
        n = moved.shape[-1]
        lib.naive_dft_lines(moved.ctypes.data, out.ctypes.data, n, moved.size // n, 1, n, sign)
        a = np.moveaxis(out, -1, ax)
    return np.ascontiguousarray(a)


@pytest.mark.parametrize("shape", [(8, 16, 32), (4, 4, 64), (16, 2, 8)])
def test_c2c_matches_naive(naive, shape):
    x = O.complex_input(shape)
    assert O.rel_l2(O.fft_c2c(x), naive_fftn(naive, x)) < 1e-14
    assert O.rel_l2(O.fft_c2c(x, inverse=True), naive_fftn(naive, x, +1)) < 1e-14
    assert O.rel_l2(O.fft_c2c(x, d=1), naive_fftn(naive, x, axes=(2,))) < 1e-14
    assert O.rel_l2(O.fft_c2c(x, d=2), naive_fftn(naive, x, axes=(1, 2))) < 1e-14


@pytest.mark.parametrize("shape", [(8, 16, 32), (4, 4, 64)])
def test_r2c_c2r_match_naive(naive, shape):
    x = O.real_input(shape)
    nzo = shape[2] // 2 + 1
    full = naive_fftn(naive, x.astype(np.complex128))
    assert O.rel_l2(O.fft_r2c(x), full[:, :, :nzo]) < 1e-14
    for d in (1, 2):
        part = naive_fftn(naive, x.astype(np.complex128), axes={1: (2,), 2: (1, 2)}[d])
        assert O.rel_l2(O.fft_r2c(x, d), part[:, :, :nzo]) < 1e-14
    # unnormalised inverse: round trip = N * x (reference testcase 3, random_dist_default.cu:592)
    assert O.rel_l2(O.fft_c2r(O.fft_r2c(x), shape[2]), x * np.prod(shape)) < 1e-14
    assert O.rel_l2(O.fft_c2r(O.fft_r2c(x, 2), shape[2], 2), x * shape[1] * shape[2]) < 1e-14


def test_laplacian_closed_form():
    """reference testcase 4: inverse(coeff * forward(f)) == -3 sqrt(N) f (random_dist_default.cu:686-743)."""
    shape = (16, 32, 8)
    f = O.sine_input(shape)
    spec = O.fft_r2c(f) * O.laplacian_coefficients(*shape, (0, 0, 0), (16, 32, 5))
    back = O.fft_c2r(spec, 8)
    assert np.abs(back - O.laplacian_expected(shape)).max() < 1e-9 * np.sqrt(np.prod(shape)) * 3


def test_inputs_are_pure_functions_of_global_index():
    shape = (6, 10, 12)
    full = O.real_input(shape)
    sub = O.real_input(shape, (2, 3, 4), (3, 5, 6))
    assert np.array_equal(sub, full[2:5, 3:8, 4:10])
    assert full.min() >= 0 and full.max() < 255
    c = O.complex_input(shape)
    assert np.array_equal(c.real, full) and not np.array_equal(c.imag, full)


def test_golden_vectors():
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_small.npz"))
    for name in ("a", "b", "c"):
        shape = tuple(int(v) for v in g[f"{name}_shape"])
        xr = O.real_input(shape, seed=1234)
        assert np.array_equal(xr.ravel()[:16], g[f"{name}_real_head"])
        assert O.rel_l2(O.fft_r2c(xr), g[f"{name}_r2c"]) < 1e-15
        assert O.rel_l2(O.fft_c2c(O.complex_input(shape, seed=1234)), g[f"{name}_c2c"]) < 1e-15


def test_split_rule():
    # size[p] = n/P + (p < n%P): mpicufft_slab.cpp:112-128
    assert O.split(10, 4) == ([3, 3, 2, 2], [0, 3, 6, 8])
    assert O.split(513, 4) == ([129, 128, 128, 128], [0, 129, 257, 385])
    assert O.split(8, 8) == ([1] * 8, list(range(8)))
