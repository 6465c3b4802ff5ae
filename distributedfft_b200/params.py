"""Plan inputs — Python mirror of /root/reference/include/params.hpp:24-93 (same names, same fields)."""
from __future__ import annotations

import enum
from dataclasses import dataclass


class GlobalSize:
    def __init__(self, Nx: int, Ny: int, Nz: int):
        self.Nx, self.Ny, self.Nz = int(Nx), int(Ny), int(Nz)
        self.Nz_out = self.Nz // 2 + 1


class Partition:
    def __init__(self, P1: int = 1, P2: int = 1):
        self.P1, self.P2 = int(P1), int(P2)


class Slab_Partition(Partition):
    def __init__(self, P1: int):
        super().__init__(P1, 1)


class Pencil_Partition(Partition):
    def __init__(self, P1: int, P2: int):
        super().__init__(P1, P2)


class CommunicationMethod(enum.IntEnum):
    Peer2Peer = 0
    All2All = 1


class SendMethod(enum.IntEnum):
    Sync = 0
    Streams = 1
    MPI_Type = 2


@dataclass
class Configurations:
    cuda_aware: bool = True
    warmup_rounds: int = 0
    comm_method: CommunicationMethod = CommunicationMethod.Peer2Peer
    send_method: SendMethod = SendMethod.Sync
    benchmark_dir: str = ""
    comm_method2: CommunicationMethod = CommunicationMethod.Peer2Peer
    send_method2: SendMethod = SendMethod.Sync


def partition_sizes(n: int, parts: int):
    """size[p] = n/parts + (p < n%parts), start = prefix sums — the rule of
    /root/reference/src/slab/default/mpicufft_slab.cpp:112-128, evaluated by the C library."""
    import ctypes as C

    from ._lib import check, lib

    sizes = (C.c_size_t * parts)()
    starts = (C.c_size_t * parts)()
    check(lib().dfft_partition(n, parts, sizes, starts))
    return list(sizes), list(starts)
