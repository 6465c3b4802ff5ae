"""Host-buffer front end: transforms whose input and output live in (pinned) host memory.

The reference's exec calls take device pointers only; its test drivers stage data by hand
(/root/reference/tests/src/slab/random_dist_default.cu:204-209).  `HostExecutor` is the convenience a caller with
host-resident data needs: every `submit` enqueues  H2D copy -> exec -> D2H copy  on three streams with two sets of
device buffers, so the device->host copy of transform i overlaps the host->device copy of transform i+1 (PCIe is
full duplex) and the FFT itself hides under the copies.  torch is used for memory and streams only.
"""
from __future__ import annotations

import contextlib
import os

import torch

from .mpicufft import C2C, FORWARD, MPIcuFFT


def gpu_local_cpus(device: int):
    """CPUs of the NUMA node the GPU's PCIe root port hangs off (sysfs local_cpulist of the GPU's PCI function),
    or None when the platform does not say."""
    try:
        pr = torch.cuda.get_device_properties(device)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:
        return None


@contextlib.contextmanager
def numa_local(device: int):
    """Run the enclosed allocation on the GPU's NUMA node: pinned pages are placed by first touch of the allocating
    thread, so staging buffers end up next to the PCIe root port of their GPU (each GPU's H2D / D2H traffic then
    stays off the inter-socket link — with 8 GPUs streaming at once that link is the bottleneck otherwise)."""
    old = None
    cpus = gpu_local_cpus(device)
    try:
        if cpus:
            old = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus)
        yield cpus
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)


def pinned_empty(n: int, dtype: torch.dtype, device: int = None) -> torch.Tensor:
    """Pinned host tensor whose pages live on the NUMA node of `device` (default: the current CUDA device)."""
    dev = torch.cuda.current_device() if device is None else device
    with numa_local(dev):
        t = torch.empty(n, dtype=dtype, pin_memory=True)
        t.view(torch.uint8)[:: 4096].fill_(0)  # first touch, in case the allocator handed out untouched pages
    return t


class HostExecutor:
    def __init__(self, plan: MPIcuFFT, direction: int = FORWARD):
        self.plan = plan
        self.direction = direction
        f64 = plan.precision == 1
        self.cdt = torch.complex128 if f64 else torch.complex64
        self.rdt = torch.float64 if f64 else torch.float32
        es = 16 if f64 else 8
        isz = plan.getInSize()
        osz = plan.getOutSize()
        self.c2c = plan.transform == C2C
        self.n_in = isz[0] * isz[1] * isz[2]
        self.n_out = osz[0] * osz[1] * osz[2]
        dom = plan.getDomainSize() // es
        forward = direction == FORWARD
        in_dtype = self.cdt if (self.c2c or not forward) else self.rdt
        out_dtype = self.cdt if (self.c2c or forward) else self.rdt
        n_src = self.n_in if forward else dom
        n_dst = dom if forward else self.n_in
        self.n_src_valid = self.n_in if forward else self.n_out
        self.n_dst_valid = self.n_out if forward else self.n_in
        self.d_in = [torch.empty(n_src, dtype=in_dtype, device="cuda") for _ in range(2)]
        self.d_out = [torch.empty(n_dst, dtype=out_dtype, device="cuda") for _ in range(2)]
        self.s_in, self.s_fft, self.s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_fft = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.count = 0

    def submit(self, host_out: torch.Tensor, host_in: torch.Tensor) -> None:
        """Enqueue one transform host_in -> host_out (flat pinned tensors of the plan's block sizes)."""
        k = self.count & 1
        self.count += 1
        # the device buffers of set k are free once the transform that used them two submits ago has been copied out
        self.s_in.wait_event(self.ev_fft[k]) if self.count > 2 else None
        with torch.cuda.stream(self.s_in):
            self.d_in[k][: self.n_src_valid].copy_(host_in.reshape(-1)[: self.n_src_valid], non_blocking=True)
            self.ev_in[k].record(self.s_in)
        self.s_fft.wait_event(self.ev_in[k])
        if self.count > 2:
            self.s_fft.wait_event(self.ev_out[k])
        with torch.cuda.stream(self.s_fft):
            if self.c2c:
                self.plan.execC2C(self.d_out[k], self.d_in[k], self.direction, stream=self.s_fft)
            elif self.direction == FORWARD:
                self.plan.execR2C(self.d_out[k], self.d_in[k], stream=self.s_fft)
            else:
                self.plan.execC2R(self.d_out[k], self.d_in[k], stream=self.s_fft)
            self.ev_fft[k].record(self.s_fft)
        self.s_out.wait_event(self.ev_fft[k])
        with torch.cuda.stream(self.s_out):
            host_out.reshape(-1)[: self.n_dst_valid].copy_(self.d_out[k][: self.n_dst_valid], non_blocking=True)
            self.ev_out[k].record(self.s_out)

    def wait(self) -> None:
        self.s_in.synchronize()
        self.s_fft.synchronize()
        self.s_out.synchronize()
        self.plan.wait()
