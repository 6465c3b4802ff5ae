"""Helpers shared by the GPU parity tests: the CUDA path is always reached through the C ABI
(distributedfft_b200 -> libdfft.so); oracle/ is only the checker."""
import numpy as np
import torch

import distributedfft_b200 as dfft
from oracle import dft_oracle as O

CDT = {dfft.F64: torch.complex128, dfft.F32: torch.complex64}
RDT = {dfft.F64: torch.float64, dfft.F32: torch.float32}
NPC = {dfft.F64: np.complex128, dfft.F32: np.complex64}
NPR = {dfft.F64: np.float64, dfft.F32: np.float32}
TOL = {dfft.F64: O.TOL["f64"], dfft.F32: O.TOL["f32"]}
PNAME = {dfft.F64: "double", dfft.F32: "float"}


def dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def make_plan(cls, prec, transform, shape, partition=None, comm=None, comm_method=dfft.CommunicationMethod.Peer2Peer):
    cfg = dfft.Configurations(comm_method=comm_method, comm_method2=comm_method)
    plan = cls(cfg, comm if comm is not None else dfft.Comm(), precision=PNAME[prec], transform="c2c" if transform == dfft.C2C else "r2c")
    plan.initFFT(dfft.GlobalSize(*shape), partition, True)
    return plan
