"""CPU oracle for the distributed 3D FFT hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (distributedfft_b200, libdfft.so) never does.

What it restates
----------------
The reference (eggersn/DistributedFFT) contains no arithmetic of its own on this path: every transform
is a cuFFT call (closed source, CUDA toolkit; the repo pins no version beyond `module load
devel/cuda/11.0` in jobs/bwunicluster/**.sbatch), wrapped in MPI data movement.  The oracle therefore
restates (a) the published definition cuFFT implements and (b) the reference's data distribution:

 * math: unnormalised DFT, forward sign -1,  X[k] = sum_n x[n] exp(-2 pi i k n / N) per axis;
   R2C keeps kz in [0, Nz/2] (/root/reference/include/params.hpp:30); the inverse is unnormalised too
   (tests scale by Nx*Ny*Nz: /root/reference/tests/src/slab/random_dist_default.cu:592).
   numpy.fft (pocketfft) computes exactly this definition in float64.
 * split rule: size[p] = n/P + (p < n%P), starts = prefix sums
   (/root/reference/src/slab/default/mpicufft_slab.cpp:112-128,
    /root/reference/src/pencil/mpicufft_pencil.cpp:89-110, rank -> grid :84-85).
 * per-rank blocks: slab out [0:Nx, y0[p]:+Ny_p, 0:Nzo] (include/mpicufft_slab.hpp:124-125);
   z_then_yx out [0:Nx, 0:Ny, z0[p]:+Nz_p] (include/mpicufft_slab_z_then_yx.hpp:43-44);
   pencil out [0:Nx, y0[i]:+Ny_i, z0[j]:+Nz_j], after d=2 [x0[i]:+Nx_i, 0:Ny, z0[j]:+Nz_j]
   (include/mpicufft_pencil.hpp:119-122 — the reference's getOutStart uses start_x for z, a bug that is
   NOT replicated; tests/src/pencil/random_dist_3D.cu:386-394 assembles with the correct offsets).
 * the reference's three checks: t1 distributed == single-GPU 3D transform
   (random_dist_default.cu:226-459), t3 forward->inverse round trip (:528-623), t4 spectral Laplacian
   of sin*sin*sin against the closed form (:625-758).

Pinning
-------
The reference holds no golden vectors (inputs are cuRAND seeded with clock()); parity is pinned by
 1. oracle/naive_dft.c — an O(N^2) long-double DFT, compared with this module in tests/test_oracle.py;
 2. the analytic Laplacian of t4 (closed form, no FFT library involved);
 3. on the GPU box, single-GPU cuFFT (cufftPlan3d) through oracle/cufft_ref.cu — the very oracle the
    reference's own testcase 1 uses;
 4. tests/golden/*.npz — small seeded input/output vectors produced by this module
    (tests/golden/make_golden.py) so drift of numpy itself would be caught.
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------------------------------
# deterministic inputs (the reference draws uniform[0,1)*255 with an unseeded cuRAND:
# /root/reference/tests/src/slab/base.cu:40-53; we keep the range and make it reproducible per
# global index so every rank can fill its own block)
# ---------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform_block(global_shape, start, size, seed=1234, stream=0, dtype=np.float64) -> np.ndarray:
    """uniform[0,255) values of the sub-block [start, start+size) of a global array, a pure function of
    the global linear index."""
    nx, ny, nz = global_shape
    with np.errstate(over="ignore"):
        ix = (np.arange(size[0], dtype=np.uint64) + np.uint64(start[0]))[:, None, None]
        iy = (np.arange(size[1], dtype=np.uint64) + np.uint64(start[1]))[None, :, None]
        iz = (np.arange(size[2], dtype=np.uint64) + np.uint64(start[2]))[None, None, :]
        lin = (ix * np.uint64(ny) + iy) * np.uint64(nz) + iz
        h = _splitmix64(lin * np.uint64(2) + np.uint64(stream) + (np.uint64(seed) << np.uint64(40)))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (u * 255.0).astype(dtype)


def real_input(global_shape, start=(0, 0, 0), size=None, seed=1234, dtype=np.float64):
    size = global_shape if size is None else size
    return uniform_block(global_shape, start, size, seed, 0, dtype)


def complex_input(global_shape, start=(0, 0, 0), size=None, seed=1234, dtype=np.complex128):
    size = global_shape if size is None else size
    rdt = np.float64 if dtype == np.complex128 else np.float32
    re = uniform_block(global_shape, start, size, seed, 0, rdt)
    im = uniform_block(global_shape, start, size, seed, 1, rdt)
    return (re + 1j * im).astype(dtype)


def sine_input(global_shape, start=(0, 0, 0), size=None, dtype=np.float64):
    """f = sin(2 pi x/Nx) sin(2 pi y/Ny) sin(2 pi z/Nz) — input of the reference's testcase 4
    (/root/reference/tests/src/slab/random_dist_default.cu:689)."""
    size = global_shape if size is None else size
    nx, ny, nz = global_shape
    x = np.sin(2 * np.pi * (np.arange(size[0]) + start[0]) / nx)[:, None, None]
    y = np.sin(2 * np.pi * (np.arange(size[1]) + start[1]) / ny)[None, :, None]
    z = np.sin(2 * np.pi * (np.arange(size[2]) + start[2]) / nz)[None, None, :]
    return (x * y * z).astype(dtype)


# ---------------------------------------------------------------------------------------------------
# transforms (always evaluated in float64 / complex128, whatever the precision under test)
# ---------------------------------------------------------------------------------------------------
def fft_c2c(x: np.ndarray, d: int = 3, inverse: bool = False) -> np.ndarray:
    """Unnormalised complex transform over the last d of (x,y,z) in the reference's order z, y, x."""
    axes = {1: (2,), 2: (1, 2), 3: (0, 1, 2)}[d]
    x = np.asarray(x, dtype=np.complex128)
    if inverse:
        n = 1
        for a in axes:
            n *= x.shape[a]
        return np.fft.ifftn(x, axes=axes) * n
    return np.fft.fftn(x, axes=axes)


def fft_r2c(x: np.ndarray, d: int = 3) -> np.ndarray:
    """execR2C(out, in, d): real -> complex along z, then complex along y (d>=2) and x (d==3)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.fft.rfft(x, axis=2)
    if d >= 2:
        out = np.fft.fft(out, axis=1)
    if d >= 3:
        out = np.fft.fft(out, axis=0)
    return out


def fft_c2r(X: np.ndarray, nz: int, d: int = 3) -> np.ndarray:
    """execC2R(out, in, d): unnormalised inverse of fft_r2c (x, then y, then complex->real along z)."""
    X = np.asarray(X, dtype=np.complex128)
    if d >= 3:
        X = np.fft.ifft(X, axis=0) * X.shape[0]
    if d >= 2:
        X = np.fft.ifft(X, axis=1) * X.shape[1]
    return np.fft.irfft(X, n=nz, axis=2) * nz


# ---------------------------------------------------------------------------------------------------
# data distribution
# ---------------------------------------------------------------------------------------------------
def split(n: int, parts: int):
    sizes = [n // parts + (1 if p < n % parts else 0) for p in range(parts)]
    starts = [sum(sizes[:p]) for p in range(parts)]
    return sizes, starts


SLAB_ZY_THEN_X, SLAB_Z_THEN_YX, PENCIL = 0, 1, 2
R2C, C2C = 0, 1


def layout(decomp, transform, nx, ny, nz, p1, p2, rank, which):
    """(size, start) of `rank`'s block. which: 0 input, 1 after z, 2 after z and y, 3 output."""
    nzc = nz if transform == C2C else nz // 2 + 1
    if decomp == SLAB_ZY_THEN_X:
        P = p1
        sx, x0 = split(nx, P)
        oy, oy0 = split(ny, P)
        if which <= 1:
            return [sx[rank], ny, nz if which == 0 else nzc], [x0[rank], 0, 0]
        if which == 2:
            return [sx[rank], ny, nzc], [x0[rank], 0, 0]
        return [nx, oy[rank], nzc], [0, oy0[rank], 0]
    if decomp == SLAB_Z_THEN_YX:
        P = p1
        sx, x0 = split(nx, P)
        sz, z0 = split(nzc, P)
        if which <= 1:
            return [sx[rank], ny, nz if which == 0 else nzc], [x0[rank], 0, 0]
        return [nx, ny, sz[rank]], [0, 0, z0[rank]]
    i, j = rank // p2, rank % p2
    sx, x0 = split(nx, p1)
    sy, y0 = split(ny, p2)
    sz, z0 = split(nzc, p2)
    oy, oy0 = split(ny, p1)
    if which <= 1:
        return [sx[i], sy[j], nz if which == 0 else nzc], [x0[i], y0[j], 0]
    if which == 2:
        return [sx[i], ny, sz[j]], [x0[i], 0, z0[j]]
    return [nx, oy[i], sz[j]], [0, oy0[i], z0[j]]


def block(a: np.ndarray, start, size) -> np.ndarray:
    return a[start[0]:start[0] + size[0], start[1]:start[1] + size[1], start[2]:start[2] + size[2]]


def assemble(global_shape, blocks, dtype=np.complex128) -> np.ndarray:
    """Inverse of block(): blocks = [(start, size, array), ...] -> global array (the coordinator's
    scatter-assemble of /root/reference/tests/src/slab/random_dist_default.cu:352-360)."""
    out = np.zeros(global_shape, dtype=dtype)
    for start, size, arr in blocks:
        block(out, start, size)[...] = np.asarray(arr).reshape(size)
    return out


# ---------------------------------------------------------------------------------------------------
# the reference's checks, with tolerances
# ---------------------------------------------------------------------------------------------------
def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, dtype=np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64)
    b = np.asarray(b, dtype=a.dtype)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


TOL = {"f64": 1e-10, "f32": 1e-5}  # BASELINE.json north_star / SURVEY.md §8(c)


def laplacian_coefficients(nx, ny, nz, start, size) -> np.ndarray:
    """-(k1^2+k2^2+k3^2)/sqrt(Nx Ny Nz) on a block of the R2C spectrum, wave numbers folded like
    derivativeCoefficients (/root/reference/tests/src/slab/random_dist_default.cu:71-119)."""
    def k(n, s, c):
        idx = np.arange(c) + s
        return np.where(idx <= n // 2, idx, idx - n).astype(np.float64)
    kx = k(nx, start[0], size[0])[:, None, None]
    ky = k(ny, start[1], size[1])[None, :, None]
    kz = (np.arange(size[2]) + start[2]).astype(np.float64)[None, None, :]
    return -(kx * kx + ky * ky + kz * kz) / np.sqrt(float(nx) * ny * nz)


def laplacian_expected(global_shape, start=(0, 0, 0), size=None) -> np.ndarray:
    """Closed form of testcase 4: inverse(coeff * forward(f)) = -3 sqrt(N) f
    (/root/reference/tests/src/slab/random_dist_default.cu:697)."""
    nx, ny, nz = global_shape
    return -3.0 * np.sqrt(float(nx) * ny * nz) * sine_input(global_shape, start, size)
