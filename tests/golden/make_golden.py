"""Generates tests/golden/golden_small.npz with oracle/dft_oracle.py (numpy pocketfft, float64).
The reference has no golden vectors; these pin the oracle itself and give the GPU tests
fixed expected outputs.  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dft_oracle as O  # noqa: E402

out = {}
for name, shape in (("a", (8, 16, 32)), ("b", (32, 4, 8)), ("c", (16, 16, 16))):
    xr = O.real_input(shape, seed=1234)
    xc = O.complex_input(shape, seed=1234)
    out[f"{name}_shape"] = np.array(shape)
    out[f"{name}_r2c"] = O.fft_r2c(xr)
    out[f"{name}_c2c"] = O.fft_c2c(xc)
    out[f"{name}_r2c_d1"] = O.fft_r2c(xr, 1)
    out[f"{name}_r2c_d2"] = O.fft_r2c(xr, 2)
    out[f"{name}_real_head"] = xr.ravel()[:16]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
