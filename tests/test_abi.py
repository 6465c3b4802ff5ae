"""The C-ABI library loads without a GPU, exports every symbol include/dfft.h declares, and its host-only
entry points (partition arithmetic, layouts, argument checking) agree with the oracle's restatement."""
import ctypes as C
import os
import re

import pytest

import distributedfft_b200 as dfft
from distributedfft_b200 import _lib
from oracle import dft_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "dfft.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfft_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 40
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/dfft.h but not exported by libdfft.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert _lib.lib().dfft_version() == 100


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdfft.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.lib()


@pytest.mark.parametrize("n,parts", [(10, 4), (513, 4), (1024, 8), (7, 7), (129, 2)])
def test_partition_matches_oracle(n, parts):
    assert dfft.partition_sizes(n, parts) == O.split(n, parts)


CASES = [
    (dfft.SLAB_ZY_THEN_X, 64, 32, 16, 4, 1), (dfft.SLAB_ZY_THEN_X, 30, 20, 18, 8, 1), (dfft.SLAB_Z_THEN_YX, 64, 32, 16, 4, 1),
    (dfft.SLAB_Z_THEN_YX, 17, 9, 30, 3, 1), (dfft.PENCIL, 64, 64, 64, 2, 4), (dfft.PENCIL, 64, 64, 64, 4, 2),
    (dfft.PENCIL, 10, 12, 14, 3, 2), (dfft.PENCIL, 1024, 1024, 1024, 2, 4), (dfft.PENCIL, 8, 8, 8, 1, 1),
]


@pytest.mark.parametrize("decomp,nx,ny,nz,p1,p2", CASES)
@pytest.mark.parametrize("transform", [dfft.R2C, dfft.C2C])
def test_layouts_match_oracle_and_tile_the_domain(decomp, nx, ny, nz, p1, p2, transform):
    P = p1 * p2
    nzc = nz if transform == dfft.C2C else nz // 2 + 1
    for which in range(4):
        cover = 0
        for r in range(P):
            got = dfft.layout(decomp, transform, nx, ny, nz, p1, p2, r, which)
            want = O.layout(decomp, transform, nx, ny, nz, p1, p2, r, which)
            assert (list(got[0]), list(got[1])) == (list(want[0]), list(want[1])), (which, r)
            cover += got[0][0] * got[0][1] * got[0][2]
        assert cover == nx * ny * (nz if which == 0 else nzc)


def test_argument_errors_have_codes_and_messages():
    l = _lib.lib()
    size = (C.c_size_t * 3)()
    start = (C.c_size_t * 3)()
    assert l.dfft_layout(dfft.PENCIL, dfft.R2C, 8, 8, 8, 2, 2, 7, 3, size, start) == -1
    assert b"rank" in l.dfft_last_error_string()
    assert l.dfft_layout(9, dfft.R2C, 8, 8, 8, 2, 2, 0, 3, size, start) == -1
    assert l.dfft_partition(8, 0, size, start) == -1
    with pytest.raises(_lib.DfftError):
        _lib.check(l.dfft_exec_r2c(None, None, None))


def test_job_runner_maps_reference_job_files():
    """tests/launch_jobs.py turns the reference's job JSON (launch.py schema) into torchrun + cli.py commands."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("launch_jobs", os.path.join(ROOT, "tests", "launch_jobs.py"))
    lj = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lj)
    job = {"size": [128, [128, 128, 256]], "global_test_settings": {"$-t": 4, "--warmup-rounds": 1, "--iterations": 0, "--double_prec": True},
           "tests": [{"name": "Slab", "-comm": "Peer2Peer", "-snd": "Streams", "--cuda_aware": False, "-p": 4},
                     {"name": "Pencil", "-comm": "All2All", "-snd": "Sync", "-p1": 2, "-p2": 2},
                     {"name": "Reference", "-t": 2, "-p1": 2, "-p2": 2}]}
    cmds = list(lj.commands(job, 4))
    assert len(cmds) == 4
    first = " ".join(cmds[0])
    assert "--nproc-per-node=4" in first and "cli.py slab" in first and "-comm Peer2Peer" in first and "-snd Streams" in first
    assert "-t 4" in first and "-d" in cmds[0] and "-c" not in cmds[0] and "-nx 128" in first and "-w 1" in first
    assert "-nz 256" in " ".join(cmds[1])
    assert "cli.py pencil" in " ".join(cmds[2]) and "-p1 2" in " ".join(cmds[2])


def test_repo_job_files_and_pencil_flag_aliases():
    """tests/jobs/*.json (this repo's validation sweeps in the reference's schema) expand to argv sets that tests/cli.py
    parses, including the pencil executable's -comm1/-snd1 spelling (tests/src/pencil/main.cpp:173-178)."""
    import glob
    import importlib.util
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    spec = importlib.util.spec_from_file_location("launch_jobs", os.path.join(ROOT, "tests", "launch_jobs.py"))
    lj = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lj)
    import cli
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "jobs", "*.json")))
    assert len(files) >= 4
    seen = 0
    for f in files:
        for cmd in lj.commands(json.load(open(f)), 8, 256):
            a = cli.parse(cmd[cmd.index(os.path.join(ROOT, "tests", "cli.py")) + 1:])
            assert a.testcase == 4 and a.double_prec and a.nx in (128, 256)
            seen += 1
    assert seen >= 2 * (9 + 40)
    a = cli.parse("pencil -nx 8 -ny 8 -nz 8 -comm1 All2All -snd1 Streams -comm2 Peer2Peer -snd2 MPI_Type -p1 2 -p2 2".split())
    assert (a.comm, a.snd, a.comm2, a.snd2) == ("All2All", "Streams", "Peer2Peer", "MPI_Type")


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference (the CPU arm: oracle port, pocketfft on the host cores) prints the contract's JSON
    line and needs no GPU."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--shape", "32,32,32"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GFLOP/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    # other ranks of a torchrun launch exit without work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=120, env=env)
    assert r2.returncode == 0 and r2.stdout.strip() == ""


def test_cpp_shim_header_compiles():
    """include/dfft.hpp (the reference's class names over the C ABI) is valid C++17 on its own."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cpp")
        open(src, "w").write('#include "dfft.hpp"\nint main() { GlobalSize g(8, 8, 8); Pencil_Partition p(2, 4); '
                             'MPIcuFFT_Slab_Opt1<double>* a = nullptr; MPIcuFFT_Pencil_Opt1<float>* b = nullptr; (void)a; (void)b; return int(g.Nz_out + p.P2) == 9 ? 0 : 1; }\n')
        subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), src], check=True)


def test_cpp_shim_covers_the_reference_api_surface():
    """tests/cpp/shim_api_surface.cpp: every public member of the reference's classes / parameter structs, in the ways its
    test drivers use them, compiles against include/dfft.hpp; the parameter structs behave like params.hpp."""
    import subprocess
    import tempfile
    src = os.path.join(ROOT, "tests", "cpp", "shim_api_surface.cpp")
    inc = "-I" + os.path.join(ROOT, "include")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", inc, src], check=True)
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "params_api")
        subprocess.run(["g++", "-std=c++17", "-DPARAMS_ONLY", inc, src, "-o", exe], check=True)
        assert subprocess.run([exe]).returncode == 0


def _dry_plan(P, rank, decomp, transform, shape, p1=0, p2=1):
    comm = C.c_void_p()
    _lib.check(_lib.lib().dfft_comm_create_dry(rank, P, C.byref(comm)))
    plan = C.c_void_p()
    rc = _lib.lib().dfft_plan_create(comm, None, decomp, dfft.F64, transform, shape[0], shape[1], shape[2], p1 or P, p2, 1, C.byref(plan))
    msg = (_lib.lib().dfft_last_error_string() or b"").decode()
    if rc == 0:
        _lib.lib().dfft_plan_destroy(plan)
    _lib.lib().dfft_comm_destroy(comm)
    return rc, msg


def test_plan_argument_validation_without_gpu():
    """initFFT-time checks through geometry-only plans: unsupported lengths, empty shares, bad grids."""
    assert _dry_plan(1, 0, dfft.SLAB_ZY_THEN_X, dfft.R2C, (64, 64, 64))[0] == 0
    rc, msg = _dry_plan(1, 0, dfft.SLAB_ZY_THEN_X, dfft.R2C, (96, 64, 64))
    assert rc == -5 and "powers of two" in msg                      # DFFT_ERR_UNSUPPORTED
    rc, msg = _dry_plan(1, 0, dfft.SLAB_ZY_THEN_X, dfft.R2C, (64, 64, 2))
    assert rc == -5                                                 # R2C needs Nz >= 4
    assert _dry_plan(1, 0, dfft.SLAB_ZY_THEN_X, dfft.C2C, (64, 64, 2))[0] == 0
    rc, msg = _dry_plan(8, 3, dfft.SLAB_ZY_THEN_X, dfft.R2C, (4, 64, 64))
    assert rc == -1 and "without data" in msg                       # more ranks than x planes
    rc, msg = _dry_plan(8, 0, dfft.PENCIL, dfft.R2C, (64, 64, 64), p1=3, p2=2)
    assert rc == -1 and "P1*P2" in msg
    assert _dry_plan(8, 7, dfft.PENCIL, dfft.R2C, (64, 64, 64), p1=2, p2=4)[0] == 0
    rc, msg = _dry_plan(16, 0, dfft.SLAB_ZY_THEN_X, dfft.R2C, (16384, 64, 64))
    assert rc == -5                                                 # lengths above 8192
    # geometry-only plans refuse to execute
    comm = C.c_void_p()
    _lib.check(_lib.lib().dfft_comm_create_dry(0, 1, C.byref(comm)))
    plan = C.c_void_p()
    _lib.check(_lib.lib().dfft_plan_create(comm, None, dfft.SLAB_ZY_THEN_X, dfft.F64, dfft.R2C, 8, 8, 8, 1, 1, 1, C.byref(plan)))
    assert _lib.lib().dfft_exec_r2c(plan, C.c_void_p(16), C.c_void_p(16)) == -4
    _lib.lib().dfft_plan_destroy(plan)
    _lib.lib().dfft_comm_destroy(comm)
