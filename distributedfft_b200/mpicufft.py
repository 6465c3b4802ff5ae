"""Host-side mirror of the reference's plan classes over the C ABI (include/dfft.h).

Class and method names follow /root/reference/include/mpicufft.hpp:55-105,
mpicufft_slab.hpp:88-125, mpicufft_slab_z_then_yx.hpp and mpicufft_pencil.hpp:71-122, so caller code
written against the reference reads the same; MPI_Comm becomes `Comm`.
Buffers are CUDA device memory: torch tensors (their .data_ptr() is used) or raw integer addresses.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

from . import _lib
from ._lib import check, lib
from .params import CommunicationMethod, Configurations, GlobalSize, Partition, SendMethod

SLAB_ZY_THEN_X, SLAB_Z_THEN_YX, PENCIL = 0, 1, 2
F32, F64 = 0, 1
R2C, C2C = 0, 1
FORWARD, INVERSE = -1, 1


def _ptr(buf) -> int:
    if buf is None:
        return 0
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "data_ptr"):
        if hasattr(buf, "is_contiguous") and not buf.is_contiguous():
            raise ValueError("buffers must be contiguous ([x][y][z], z fastest)")
        if hasattr(buf, "is_cuda") and not buf.is_cuda:
            raise ValueError("buffers must live in CUDA device memory (use HostExecutor for host buffers)")
        return int(buf.data_ptr())
    raise TypeError(f"unsupported buffer type {type(buf)}")


def _stream_ptr(stream) -> int:
    if stream is None:
        return 0
    if isinstance(stream, int):
        return stream
    return int(stream.cuda_stream)


class Comm:
    """Replaces MPI_Comm: one rank per process/GPU, NCCL underneath (src/mpicufft.cpp:42-50)."""

    def __init__(self, rank: int = 0, nranks: int = 1, unique_id: Optional[bytes] = None, device: int = 0):
        h = C.c_void_p()
        idbuf = None
        if nranks > 1:
            if unique_id is None or len(unique_id) != _lib.UNIQUE_ID_BYTES:
                raise ValueError("unique_id of 128 bytes required for nranks > 1")
            idbuf = C.create_string_buffer(unique_id, _lib.UNIQUE_ID_BYTES)
        check(lib().dfft_comm_create(rank, nranks, idbuf, device, C.byref(h)))
        self._h = h
        self.rank, self.size, self.device = rank, nranks, device

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_lib.UNIQUE_ID_BYTES)
        check(lib().dfft_get_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: Optional[int] = None) -> "Comm":
        """Build the communicator of the current torch.distributed world (rank 0 creates the NCCL id,
        torch broadcasts it — the role MPI_Bcast would play in the reference's launcher)."""
        import torch
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size() == 1:
            return cls(0, 1, None, torch.cuda.current_device() if device is None else device)
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(rank, world, box[0], torch.cuda.current_device() if device is None else device)

    def destroy(self):
        if self._h:
            lib().dfft_comm_destroy(self._h)
            self._h = None


class MPIcuFFT:
    """Abstract plan (include/mpicufft.hpp:55-105). `precision`: 'double' | 'float' replaces the C++
    template argument; `transform`: 'r2c' (the reference's) or 'c2c'."""

    _decomp = SLAB_ZY_THEN_X

    def __init__(self, config: Configurations, comm: Optional[Comm] = None, max_world_size: int = -1,
                 precision: str = "double", transform: str = "r2c"):
        self.config = config
        self.comm = comm if comm is not None else Comm()
        if precision in ("double", "f64", "float64"):
            self.precision = F64
        elif precision in ("float", "f32", "float32"):
            self.precision = F32
        else:
            raise ValueError(f"precision must be 'double' or 'float', got {precision!r}")
        if transform in ("r2c", "R2C"):
            self.transform = R2C
        elif transform in ("c2c", "C2C"):
            self.transform = C2C
        else:
            raise ValueError(f"transform must be 'r2c' or 'c2c', got {transform!r}")
        self._h = None
        self.initialized = False

    # -- plan ------------------------------------------------------------------------------------
    def initFFT(self, global_size: GlobalSize, partition: Optional[Partition] = None, allocate: bool = True):
        if global_size is None:
            raise RuntimeError("GlobalSize or Partition not initialized!")
        cfg = _lib.dfft_config(int(self.config.cuda_aware), int(self.config.warmup_rounds), int(self.config.comm_method),
                               int(self.config.send_method),
                               self.config.benchmark_dir.encode() if self.config.benchmark_dir else None,
                               int(self.config.comm_method2), int(self.config.send_method2))
        p1 = partition.P1 if partition is not None else self.comm.size
        p2 = partition.P2 if partition is not None else 1
        if self._decomp == PENCIL and partition is None:
            raise RuntimeError("GlobalSize or Partition not initialized!")
        h = C.c_void_p()
        check(lib().dfft_plan_create(self.comm._h, C.byref(cfg), self._decomp, self.precision, self.transform,
                                     global_size.Nx, global_size.Ny, global_size.Nz, p1, p2, 1 if allocate else 0, C.byref(h)))
        self._h = h
        self.global_size = global_size
        self.initialized = True

    def setWorkArea(self, device=None, host=None):
        check(lib().dfft_set_work_area(self._h, _ptr(device), _ptr(host)))

    def destroy(self):
        if self._h:
            lib().dfft_plan_destroy(self._h)
            self._h = None

    # -- exec ------------------------------------------------------------------------------------
    def _need(self):
        if not self.initialized:
            raise RuntimeError("plan not initialised (call initFFT)")

    def execR2C(self, out, in_, d: int = 3, stream=None):
        self._need()
        if stream is not None:
            if d != 3:
                raise ValueError("async partial transforms are not exposed")
            return check(lib().dfft_exec_r2c_async(self._h, _ptr(out), _ptr(in_), _stream_ptr(stream)))
        if d == 3:
            return check(lib().dfft_exec_r2c(self._h, _ptr(out), _ptr(in_)))
        return check(lib().dfft_exec_r2c_partial(self._h, _ptr(out), _ptr(in_), d))

    def execC2R(self, out, in_, d: int = 3, stream=None):
        self._need()
        if stream is not None:
            if d != 3:
                raise ValueError("async partial transforms are not exposed")
            return check(lib().dfft_exec_c2r_async(self._h, _ptr(out), _ptr(in_), _stream_ptr(stream)))
        if d == 3:
            return check(lib().dfft_exec_c2r(self._h, _ptr(out), _ptr(in_)))
        return check(lib().dfft_exec_c2r_partial(self._h, _ptr(out), _ptr(in_), d))

    def execC2C(self, out, in_, direction: int = FORWARD, d: int = 3, stream=None):
        """Complex transform (not in the reference; BASELINE configs 1-4 need it)."""
        self._need()
        if stream is not None:
            if d != 3:
                raise ValueError("async partial transforms are not exposed")
            return check(lib().dfft_exec_c2c_async(self._h, _ptr(out), _ptr(in_), direction, _stream_ptr(stream)))
        if d == 3:
            return check(lib().dfft_exec_c2c(self._h, _ptr(out), _ptr(in_), direction))
        return check(lib().dfft_exec_c2c_partial(self._h, _ptr(out), _ptr(in_), direction, d))

    def wait(self):
        return check(lib().dfft_plan_wait(self._h))

    def tune(self, out, in_, direction: int = FORWARD, reps: int = 3) -> str:
        """Plan-time measurement: try the sequential and the overlapped schedules on these buffers and keep the fastest
        (collective).  Returns the report string."""
        self._need()
        check(lib().dfft_plan_tune(self._h, _ptr(out), _ptr(in_), 1 if direction > 0 else 0, reps))
        return (lib().dfft_plan_tune_report(self._h) or b"").decode()

    # -- getters ---------------------------------------------------------------------------------
    def _triple(self, fn, *a):
        v = (C.c_size_t * 3)()
        check(fn(self._h, *a, v))
        return [int(x) for x in v]

    def getInSize(self): return self._triple(lib().dfft_get_in_size)
    def getInStart(self): return self._triple(lib().dfft_get_in_start)
    def getOutSize(self): return self._triple(lib().dfft_get_out_size)
    def getOutStart(self): return self._triple(lib().dfft_get_out_start)
    def getPartialSize(self, d): return self._triple(lib().dfft_get_partial_size, d)
    def getPartialStart(self, d): return self._triple(lib().dfft_get_partial_start, d)
    def getDomainSize(self): return int(lib().dfft_get_domain_size(self._h))
    def getWorkSizeDevice(self): return int(lib().dfft_get_work_size_device(self._h))
    def getWorkSizeHost(self): return int(lib().dfft_get_work_size_host(self._h))
    def getWorkAreaDevice(self): return lib().dfft_get_work_area_device(self._h)
    def getWorkAreaHost(self): return None
    def getRank(self): return int(lib().dfft_get_rank(self._h))
    def getWorldSize(self): return int(lib().dfft_get_world_size(self._h))

    # -- timer -----------------------------------------------------------------------------------
    def enableTimer(self, on: bool = True):
        check(lib().dfft_timer_enable(self._h, 1 if on else 0))

    def phaseTimes(self):
        n = lib().dfft_get_phase_count(self._h)
        ms = (C.c_double * max(n, 1))()
        n = check(lib().dfft_get_phase_times(self._h, ms, n))
        return [(lib().dfft_get_phase_name(self._h, i).decode(), float(ms[i])) for i in range(n)]

    def stepTimes(self):
        """[(label, ms)] per launched step of the last timed exec."""
        n = lib().dfft_get_step_count(self._h)
        ms = (C.c_double * max(n, 1))()
        n = check(lib().dfft_get_step_times(self._h, ms, n))
        return [(lib().dfft_get_step_label(self._h, i).decode(), float(ms[i])) for i in range(n)]

    def timeline(self):
        """[(label, stream, begin_ms, end_ms)] of the last timed exec — every step on its own plan stream."""
        cap = 256
        b = (C.c_double * cap)(); e = (C.c_double * cap)(); st = (C.c_int * cap)()
        n = check(lib().dfft_get_timeline(self._h, b, e, st, cap))
        return [(lib().dfft_get_timeline_label(self._h, i).decode(), int(st[i]), float(b[i]), float(e[i])) for i in range(min(n, cap))]

    def lastBreakdown(self):
        f, x, t = C.c_double(), C.c_double(), C.c_double()
        check(lib().dfft_get_last_breakdown(self._h, C.byref(f), C.byref(x), C.byref(t)))
        return {"fft_ms": f.value, "exchange_ms": x.value, "total_ms": t.value}

    def lastLaunchCount(self):
        return int(lib().dfft_get_last_launch_count(self._h))


class MPIcuFFT_Slab(MPIcuFFT):
    """2D (y,z) FFT -> transpose (x split -> y split) -> 1D x FFT. include/mpicufft_slab.hpp."""
    _decomp = SLAB_ZY_THEN_X


class MPIcuFFT_Slab_Z_Then_YX(MPIcuFFT):
    """1D z FFT -> transpose (x split -> z split) -> 2D (y,x) FFT. include/mpicufft_slab_z_then_yx.hpp."""
    _decomp = SLAB_Z_THEN_YX


class MPIcuFFT_Pencil(MPIcuFFT):
    """z FFT -> transpose -> y FFT -> transpose -> x FFT on a P1 x P2 grid. include/mpicufft_pencil.hpp."""
    _decomp = PENCIL

    def initFFT(self, global_size, partition=None, allocate=True):
        if partition is None or global_size is None:
            raise RuntimeError("GlobalSize or Partition not initialized!")
        if partition.P1 * partition.P2 != self.comm.size:
            raise RuntimeError("Invalid Input Partition!")
        self.partition = partition
        super().initFFT(global_size, partition, allocate)

    def getPartitionDimensions(self):
        """(input_dim, transposed_dim, output_dim), each a dict with size_x/y/z and start_x/y/z lists
        (Partition_Dimensions, params.hpp:58-81; mpicufft_pencil.hpp:112-116)."""
        from .params import partition_sizes
        g, p = self.global_size, self.partition
        nzc = g.Nz if self.transform == C2C else g.Nz // 2 + 1

        def dims(nx_parts, ny_parts, nz_parts, nz):
            d = {}
            for ax, n, parts in (("x", g.Nx, nx_parts), ("y", g.Ny, ny_parts), ("z", nz, nz_parts)):
                d[f"size_{ax}"], d[f"start_{ax}"] = partition_sizes(n, parts)
            return d
        return dims(p.P1, p.P2, 1, g.Nz), dims(p.P1, 1, p.P2, nzc), dims(1, p.P1, p.P2, nzc)


# "Realigned" (Opt1) classes of the reference — include/mpicufft_slab_opt1.hpp, mpicufft_slab_z_then_yx_opt1.hpp,
# mpicufft_pencil_opt1.hpp: same results and output layouts as the default classes; the transposing store they add is
# always fused into the FFT passes here, so they are aliases.
MPIcuFFT_Slab_Opt1 = MPIcuFFT_Slab
MPIcuFFT_Slab_Z_Then_YX_Opt1 = MPIcuFFT_Slab_Z_Then_YX
MPIcuFFT_Pencil_Opt1 = MPIcuFFT_Pencil


def layout(decomp: int, transform: int, nx: int, ny: int, nz: int, p1: int, p2: int, rank: int, which: int):
    """Geometry of any rank without a device (dfft_layout): which = 0 in, 1 after z, 2 after z,y, 3 out."""
    size = (C.c_size_t * 3)()
    start = (C.c_size_t * 3)()
    check(lib().dfft_layout(decomp, transform, nx, ny, nz, p1, p2, rank, which, size, start))
    return [int(x) for x in size], [int(x) for x in start]


def fft1d_contig(precision: int, kind: int, direction: int, n: int, lines: int, out, out_pitch: int, in_, in_pitch: int, stream=None):
    check(lib().dfft_fft1d_contig(precision, kind, direction, n, lines, _ptr(out), out_pitch, _ptr(in_), in_pitch, _stream_ptr(stream)))


def fft1d_strided(precision: int, direction: int, a: int, n: int, b: int, out, in_, stream=None):
    check(lib().dfft_fft1d_strided(precision, direction, a, n, b, _ptr(out), _ptr(in_), _stream_ptr(stream)))


def fft1d_general(precision: int, direction: int, n: int, a0: int, a1: int, b: int, out, out_strides, in_, in_strides, stream=None):
    """Batched C2C along n over a general view: element (i0,i1,n,ib) at i0*s[0] + i1*s[1] + n*s[2] + ib."""
    os_ = (C.c_longlong * 3)(*out_strides)
    is_ = (C.c_longlong * 3)(*in_strides)
    check(lib().dfft_fft1d_general(precision, direction, n, a0, a1, b, _ptr(out), os_, _ptr(in_), is_, _stream_ptr(stream)))
