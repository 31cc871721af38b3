"""Single-GPU Lanczos driver (ed_lanczos).  The reference has no eigensolver of its own (it passes `mul!` to Arpack,
docs/src/examples/spinhalf.md:26).  The row-sharded multi-GPU loop lives in the library too (ed_lanczos_sharded); its thin
caller is edcuda.distributed.ShardedOperator.lanczos."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from ._lib import ED_C128, ED_F64, ED_SIDE_LEFT, check, lib


@dataclass
class LanczosResult:
    alpha: np.ndarray
    beta: np.ndarray
    ritz: np.ndarray
    steps: int


def tridiag_eigvals(alpha, beta) -> np.ndarray:
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    b = np.ascontiguousarray(beta, dtype=np.float64)
    out = np.empty(len(a), dtype=np.float64)
    if len(a):
        check(lib.ed_tridiag_eigvals(a.ctypes.data, b.ctypes.data if len(b) else None, len(a), out.ctypes.data))
    return out


def lanczos(opr, n_steps: int, v0: Optional[np.ndarray] = None, seed: int = 0, dtype=None, n_ritz: int = 4) -> LanczosResult:
    """Single-GPU Lanczos: `n_steps` steps from `v0` (numpy, host) or from a Philox-seeded normal vector."""
    if dtype is None:
        dtype = np.complex128 if (opr.is_complex or (v0 is not None and np.iscomplexobj(v0))) else np.float64
    code = ED_C128 if np.dtype(dtype) == np.complex128 else ED_F64
    ptr = None
    if v0 is not None:
        v0 = np.ascontiguousarray(v0, dtype=dtype)
        if v0.size != opr.dimension:
            from ._lib import DimensionMismatch
            raise DimensionMismatch("start vector length differs from the dimension")
        ptr = v0.ctypes.data
    alpha = np.zeros(n_steps)
    beta = np.zeros(n_steps)
    n_ritz = min(n_ritz, n_steps)
    ritz = np.zeros(n_ritz)
    done = C.c_int32()
    check(lib.ed_lanczos(opr._handle, n_steps, ptr, code, seed, alpha.ctypes.data, beta.ctypes.data, ritz.ctypes.data,
                         n_ritz, C.byref(done)))
    k = done.value
    return LanczosResult(alpha[:k], beta[:k], ritz[: min(n_ritz, k)], k)


class LanczosState:
    """ed_lanczos_state: the Lanczos loop as a resumable object (run k steps, save, load, continue bit for bit)."""

    def __init__(self, opr, v0: Optional[np.ndarray] = None, seed: int = 0, dtype=None, _handle=None):
        self.opr = opr
        if _handle is not None:
            self._handle = _handle
            return
        if dtype is None:
            dtype = np.complex128 if (opr.is_complex or (v0 is not None and np.iscomplexobj(v0))) else np.float64
        code = ED_C128 if np.dtype(dtype) == np.complex128 else ED_F64
        ptr = None
        if v0 is not None:
            v0 = np.ascontiguousarray(v0, dtype=dtype)
            ptr = v0.ctypes.data
        h = C.c_void_p()
        check(lib.ed_lanczos_state_create(opr._handle, code, ptr, seed, C.byref(h)))
        self._handle = h
        self.steps = 0

    def step(self, n_steps: int):
        check(lib.ed_lanczos_state_step(self._handle, n_steps))
        self.steps = getattr(self, "steps", 0) + n_steps
        return self

    def result(self, n_ritz: int = 4) -> LanczosResult:
        cap = max(getattr(self, "steps", 0), 1)
        alpha, beta = np.zeros(cap), np.zeros(cap)
        n_ritz = min(n_ritz, cap)
        ritz = np.zeros(n_ritz)
        done = C.c_int32()
        check(lib.ed_lanczos_state_result(self._handle, cap, alpha.ctypes.data, beta.ctypes.data, ritz.ctypes.data, n_ritz, C.byref(done)))
        k = done.value
        return LanczosResult(alpha[:k], beta[:k], ritz[: min(n_ritz, k)], k)

    def save(self, path: str):
        check(lib.ed_lanczos_state_save(self._handle, str(path).encode()))

    @classmethod
    def load(cls, opr, path: str, steps_taken: int):
        h = C.c_void_p()
        check(lib.ed_lanczos_state_load(opr._handle, str(path).encode(), C.byref(h)))
        st = cls(opr, _handle=h)
        st.steps = steps_taken
        return st

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and lib is not None:
            try:
                lib.ed_lanczos_state_destroy(h)
            except Exception:
                pass
