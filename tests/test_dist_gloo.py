"""world_size-2 (and 4) CPU runs over gloo of the *distribution logic*: every rank transforms its block
with numpy, the transpositions move exactly the per-peer sub-blocks that the C library's layouts
(dfft_layout) imply, and every rank's result must equal its block of the global transform.  This covers
the host-side arithmetic of the N>1 path (who sends which rows to whom, in which order they land)
without a GPU; the CUDA kernels reuse the same layouts on the device."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange(group_ranks, send_blocks):
    """all-to-all-v of numpy arrays inside a sub-group, by pairwise isend/irecv (gloo)."""
    me = dist.get_rank()
    out = [None] * len(group_ranks)
    reqs = []
    shapes = [None] * len(group_ranks)
    # exchange shapes first (deterministic from layouts in the real library; here sent for simplicity)
    for q, r in enumerate(group_ranks):
        if r == me:
            out[q] = send_blocks[q]
    for q, r in enumerate(group_ranks):
        if r == me:
            continue
        t = torch.from_numpy(np.ascontiguousarray(send_blocks[q]).view(np.float64).copy())
        reqs.append(dist.isend(t, r))
    return out, reqs


def _worker(rank, world, port, decomp, transform, shape, p1, p2, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import distributedfft_b200 as dfft
        from oracle import dft_oracle as O

        nx, ny, nz = shape
        lay = lambda r, w: dfft.layout(decomp, transform, nx, ny, nz, p1, p2, r, w)
        c2c = transform == dfft.C2C
        size0, start0 = lay(rank, 0)
        x = (O.complex_input if c2c else O.real_input)(shape, start0, size0)
        a = np.fft.fft(x, axis=2) if c2c else np.fft.rfft(x, axis=2)

        def transpose(a, w_from, w_to, group):
            """re-block from layout w_from to layout w_to inside `group` (global ranks)."""
            s_from, o_from = lay(rank, w_from)
            s_to, o_to = lay(rank, w_to)
            recv = np.zeros(s_to, dtype=np.complex128)
            sends, reqs, bufs = [], [], []
            for r in group:
                sr, orr = lay(r, w_to)  # what r holds afterwards
                lo = [max(o_from[k], orr[k]) for k in range(3)]
                hi = [min(o_from[k] + s_from[k], orr[k] + sr[k]) for k in range(3)]
                blk = a[lo[0] - o_from[0]:hi[0] - o_from[0], lo[1] - o_from[1]:hi[1] - o_from[1], lo[2] - o_from[2]:hi[2] - o_from[2]]
                if r == rank:
                    recv[lo[0] - o_to[0]:hi[0] - o_to[0], lo[1] - o_to[1]:hi[1] - o_to[1], lo[2] - o_to[2]:hi[2] - o_to[2]] = blk
                else:
                    t = torch.from_numpy(np.ascontiguousarray(blk).view(np.float64).copy())
                    reqs.append(dist.isend(t, r))
                    sends.append(t)
            for r in group:
                if r == rank:
                    continue
                sr, orr = lay(r, w_from)  # what r held before
                lo = [max(o_to[k], orr[k]) for k in range(3)]
                hi = [min(o_to[k] + s_to[k], orr[k] + sr[k]) for k in range(3)]
                shp = [hi[k] - lo[k] for k in range(3)]
                t = torch.empty(int(np.prod(shp)) * 2, dtype=torch.float64)
                dist.recv(t, r)
                recv[lo[0] - o_to[0]:hi[0] - o_to[0], lo[1] - o_to[1]:hi[1] - o_to[1], lo[2] - o_to[2]:hi[2] - o_to[2]] = \
                    t.numpy().view(np.complex128).reshape(shp)
            for rq in reqs:
                rq.wait()
            return recv

        if decomp == dfft.PENCIL:
            i, j = rank // p2, rank % p2
            g1 = [i * p2 + k for k in range(p2)]
            g2 = [k * p2 + j for k in range(p1)]
        elif decomp == dfft.SLAB_ZY_THEN_X:
            g1, g2 = [rank], list(range(world))
        else:
            g1, g2 = list(range(world)), [rank]
        a = transpose(a, 1, 2, g1)
        a = np.fft.fft(a, axis=1)
        a = transpose(a, 2, 3, g2)
        a = np.fft.fft(a, axis=0)
        size3, start3 = lay(rank, 3)
        full = O.fft_c2c(O.complex_input(shape)) if c2c else O.fft_r2c(O.real_input(shape))
        err = O.rel_l2(a, O.block(full, start3, size3))
        q.put((rank, err))
    finally:
        dist.destroy_process_group()


CASES = [
    ("slab", 0, 0, (8, 12, 10), 2, 1),
    ("slab-c2c", 0, 1, (8, 8, 8), 2, 1),
    ("z_then_yx", 1, 0, (6, 4, 16), 2, 1),
    ("pencil-1x2", 2, 0, (8, 8, 12), 1, 2),
    ("pencil-2x1", 2, 0, (8, 8, 12), 2, 1),
    ("pencil-2x2", 2, 0, (10, 12, 14), 2, 2),
]


@pytest.mark.parametrize("name,decomp,transform,shape,p1,p2", CASES)
def test_distribution_logic_over_gloo(name, decomp, transform, shape, p1, p2):
    world = p1 * p2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, decomp, transform, shape, p1, p2, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-13, (name, rank, err)
