"""Thin callers of the library's multi-GPU context (include/edcuda.h: ed_ctx_*, ed_sharded_*, ed_dvec_*,
ed_apply_sharded, ed_lanczos_sharded).  Everything on the data path -- row partition, halo copies over NVLink, NCCL
all-gather / all-reduce, the Lanczos loop -- runs inside libedcuda.so; this module only creates the handles.

The reference parallelises inside apply! (Threads.@threads over split rows,
Representation/abstract_operator_representation.jl:260-267, 358-378); a caller never partitions anything.  Same here:

    ctx = Context.single_process([0, 1, 2, 3])            # one process drives 4 GPUs (ncclCommInitAll)
    ctx = Context.from_env()                               # one process per GPU under torchrun (ncclCommInitRank)
    sh  = ShardedOperator(ctx, lambda: ed.represent(ed.represent(sector), h))   # built once per local GPU
    x, y = sh.vector(), sh.vector();  x.randn(seed);  sh.apply(y, x);  res = sh.lanczos(100)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Optional, Sequence

import numpy as np

from ._lib import ED_C128, ED_F64, check, lib
from .lanczos import LanczosResult


_STORES = []      # rendezvous stores are kept alive for the life of the process
_UID_COUNTER = [0]


def _bootstrap_uid(rank: int, world: int) -> bytes:
    """Rank 0 asks the library for an NCCL unique id; the launcher's rendezvous store hands the 128 bytes to the other
    processes: under torchrun the agent's TCP store on MASTER_ADDR:MASTER_PORT (every worker is a client), otherwise a
    store hosted by rank 0 on that port.  Bootstrap only: no vector data and no collective of the solver ever goes
    through torch.distributed."""
    from datetime import timedelta
    from torch.distributed import TCPStore
    host = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("MASTER_PORT", "29500"))
    agent = os.environ.get("TORCHELASTIC_USE_AGENT_STORE", "") == "True"
    store = TCPStore(host, port, world, is_master=(rank == 0 and not agent), timeout=timedelta(seconds=600), wait_for_workers=False)
    _STORES.append(store)
    key = "edcuda/uid/%s/%s/%d" % (os.environ.get("TORCHELASTIC_RUN_ID", "-"), os.environ.get("TORCHELASTIC_RESTART_COUNT", "0"), _UID_COUNTER[0])
    _UID_COUNTER[0] += 1
    if rank == 0:
        uid = (C.c_uint8 * 128)()
        check(lib.ed_ctx_unique_id(uid))
        store.set(key, bytes(uid))
        return bytes(uid)
    return store.get(key)


class Context:
    """ed_ctx: a communicator of `world` ranks, one per GPU of the node."""

    def __init__(self, handle):
        self._handle = handle
        w, nl, fr, ver = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.ed_ctx_info(handle, C.byref(w), C.byref(nl), C.byref(fr), C.byref(ver)))
        self.world, self.n_local, self.first_rank, self.nccl_version = w.value, nl.value, fr.value, ver.value

    @classmethod
    def single_process(cls, devices: Sequence[int]):
        """One process, len(devices) GPUs.  All ids equal = loopback test context (ranks share one GPU, no NCCL)."""
        ids = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        check(lib.ed_ctx_create(len(devices), ids, C.byref(h)))
        return cls(h)

    @classmethod
    def from_env(cls):
        """One process per GPU as launched by torchrun (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT)."""
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        check(lib.ed_set_device(local_rank))
        h = C.c_void_p()
        if world == 1:
            check(lib.ed_ctx_create_rank(1, 0, local_rank, None, C.byref(h)))
        else:
            uid = _bootstrap_uid(rank, world)
            check(lib.ed_ctx_create_rank(world, rank, local_rank, (C.c_uint8 * 128).from_buffer_copy(uid), C.byref(h)))
        return cls(h)

    @property
    def rank(self) -> int:
        return self.first_rank

    def device(self, i: int = 0) -> int:
        d = C.c_int32()
        check(lib.ed_ctx_device_stream(self._handle, i, C.byref(d), None))
        return d.value

    def stream(self, i: int = 0) -> int:
        s = C.c_void_p()
        check(lib.ed_ctx_device_stream(self._handle, i, None, C.byref(s)))
        return s.value or 0

    def sync(self):
        check(lib.ed_ctx_sync(self._handle))

    def barrier(self):
        check(lib.ed_ctx_barrier(self._handle))

    def timer_record(self, slot: int):
        check(lib.ed_ctx_timer_record(self._handle, slot))

    def timer_elapsed(self, a: int, b: int) -> float:
        ms = C.c_double()
        check(lib.ed_ctx_timer_elapsed(self._handle, a, b, C.byref(ms)))
        return ms.value

    def allreduce(self, values, op: str = "sum"):
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64))
        check(lib.ed_ctx_allreduce_host(self._handle, v.ctypes.data, v.size, 0 if op == "sum" else 1))
        return v

    def close(self):
        if self._handle is not None:
            check(lib.ed_ctx_destroy(self._handle))
            self._handle = None


class DistributedVector:
    """ed_dvec: every rank holds its rows (layout: ShardedOperator.ranges)."""

    def __init__(self, sh: "ShardedOperator"):
        self.sh = sh
        h = C.c_void_p()
        check(lib.ed_dvec_create(sh._handle, C.byref(h)))
        self._handle = h

    def local(self, i: int = 0):
        """(device pointer, rows) of local rank i"""
        p, n = C.c_void_p(), C.c_int64()
        check(lib.ed_dvec_local(self._handle, i, C.byref(p), C.byref(n)))
        return p.value or 0, n.value

    def tensor(self, i: int = 0):
        """torch view of local rank i's rows (plumbing for host<->device copies in tests and the bench)"""
        import torch
        ptr, n = self.local(i)
        dt = self.sh.np_dtype

        class _View:
            __cuda_array_interface__ = {"shape": (n,), "typestr": np.dtype(dt).str, "data": (ptr, False), "version": 2, "strides": None}

        if n == 0:
            return torch.empty(0, dtype=torch.complex128 if dt == np.complex128 else torch.float64, device=torch.device("cuda", self.sh.ctx.device(i)))
        return torch.as_tensor(_View(), device=torch.device("cuda", self.sh.ctx.device(i)))

    def randn(self, seed: int, scale: float = 1.0):
        check(lib.ed_dvec_randn(self._handle, seed, scale))
        return self

    def upload(self, host_full: np.ndarray):
        a = np.ascontiguousarray(host_full, dtype=self.sh.np_dtype)
        if a.size != self.sh.dimension:
            from ._lib import DimensionMismatch
            raise DimensionMismatch("vector length differs from the dimension")
        check(lib.ed_dvec_upload(self._handle, a.ctypes.data))
        return self

    def download(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """rows owned by this process's ranks are filled in (all rows when one process drives every GPU)"""
        if out is None:
            out = np.zeros(self.sh.dimension, dtype=self.sh.np_dtype)
        check(lib.ed_dvec_download(self._handle, out.ctypes.data))
        return out

    def close(self):
        if self._handle is not None:
            check(lib.ed_dvec_destroy(self._handle))
            self._handle = None


class ShardedOperator:
    """ed_sharded: an operator representation whose rows are distributed over the ranks of a Context.
    `make_opr` is called once per local GPU (with that device current) and returns the representation built there."""

    def __init__(self, ctx: Context, make_opr: Callable[[], object], dtype=None, exchange: str = "auto", n_chunks: int = 0):
        self.ctx = ctx
        self.oprs: List[object] = []
        for i in range(ctx.n_local):
            check(lib.ed_set_device(ctx.device(i)))
            self.oprs.append(make_opr())
        check(lib.ed_set_device(ctx.device(0)))
        opr = self.oprs[0]
        self.np_dtype = np.dtype(dtype or (np.complex128 if opr.is_complex else np.float64))
        self.code = ED_C128 if self.np_dtype == np.complex128 else ED_F64
        self.dimension = opr.dimension
        handles = (C.c_void_p * ctx.n_local)(*[o._handle for o in self.oprs])
        h = C.c_void_p()
        check(lib.ed_sharded_create(ctx._handle, handles, self.code, {"auto": 0, "allgather": 1, "halo": 2, "pull": 2, "push": 3, "cepush": 4, "nccl": 5}[exchange], n_chunks, C.byref(h)))
        self._handle = h

    def info(self, i: int = 0) -> dict:
        nl, nh, nr, npl, nc, he = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.ed_sharded_info(self._handle, i, C.byref(nl), C.byref(nh), C.byref(nr), C.byref(npl), C.byref(nc), C.byref(he)))
        return {"n_local": nl.value, "n_halo": nh.value, "n_ranges": nr.value, "n_pulls": npl.value, "n_chunks": nc.value,
                "exchange": ("allgather", "halo", "halo", "halo", "halo")[he.value],
                "halo_transport": (None, "copy-engine pulls", "owner pushes", "owner copy-engine pushes", "nccl send/recv")[he.value]}

    def ranges(self, i: int = 0):
        n = self.info(i)["n_ranges"]
        lo, hi = (C.c_int64 * max(n, 1))(), (C.c_int64 * max(n, 1))()
        check(lib.ed_sharded_ranges(self._handle, i, lo, hi))
        return [(lo[k], hi[k]) for k in range(n)]

    def vector(self) -> DistributedVector:
        return DistributedVector(self)

    def apply(self, y: DistributedVector, x: DistributedVector, fence: bool = True, dot: bool = False):
        """y = H x (mul!).  Asynchronous unless dot=True, which returns <x, Hx> (complex)."""
        if dot:
            d = (C.c_double * 2)()
            check(lib.ed_apply_sharded(self._handle, y._handle, x._handle, 0 if fence else 1, d))
            return complex(d[0], d[1])
        check(lib.ed_apply_sharded(self._handle, y._handle, x._handle, 0 if fence else 1, None))
        return None

    def profile(self, y: DistributedVector, x: DistributedVector) -> dict:
        """one matvec taken apart (halo exchange): ms of pack / fence / peer copies alone / kernels alone, max over ranks"""
        ms = (C.c_double * 4)()
        check(lib.ed_sharded_profile(self._handle, y._handle, x._handle, ms))
        return {"pack_ms": ms[0], "fence_ms": ms[1], "pull_ms": ms[2], "kernel_ms": ms[3]}

    def lanczos(self, n_steps: int, seed: int = 0, v0: Optional[DistributedVector] = None, n_ritz: int = 4):
        alpha, beta = np.zeros(n_steps), np.zeros(n_steps)
        n_ritz = min(n_ritz, n_steps)
        ritz = np.zeros(n_ritz)
        done, ms = C.c_int32(), C.c_double()
        check(lib.ed_lanczos_sharded(self._handle, n_steps, seed, v0._handle if v0 is not None else None, alpha.ctypes.data,
                                     beta.ctypes.data, ritz.ctypes.data, n_ritz, C.byref(done), C.byref(ms)))
        k = done.value
        res = LanczosResult(alpha[:k], beta[:k], ritz[: min(n_ritz, k)], k)
        res.ms_per_step = ms.value
        return res

    def close(self):
        if self._handle is not None:
            check(lib.ed_sharded_destroy(self._handle))
            self._handle = None
