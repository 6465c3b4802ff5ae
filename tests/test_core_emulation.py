"""Runs the CUDA FFT core's index/twiddle arithmetic on the CPU (tests/emu_fft_core.cpp includes the very
header the kernels are built from) against a naive long-double DFT — every size 2..4096, both
directions, three per-thread radices."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def test_core_matches_naive_dft():
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "emu")
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "emu_fft_core.cpp")], check=True)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
        assert "fails=0" in r.stdout
