"""Lanczos drivers on top of the engine's device kernels.

The reference has no eigensolver of its own (it passes `mul!` to Arpack, docs/src/examples/spinhalf.md:26);
north_star asks for a device-resident Lanczos whose matvec is the engine's apply and whose dot products are
NCCL all-reduces when the basis is row-sharded.

  lanczos(opr, n_steps, ...)            one GPU, the whole loop inside libedcuda (ed_lanczos)
  ShardedLanczos(opr_shard, ...)        one process per GPU: rows [lo, hi) per rank; every step
                                        all-gathers the Krylov vector (NCCL over NVLink through
                                        torch.distributed) and all-reduces the two scalars.
"""
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


def split_rows(dim: int, world: int):
    """Contiguous, count-balanced row ranges (the reference's splitrange, src/util.jl:102-121)."""
    base, rem = divmod(dim, world)
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    return [(int(offs[r]), int(offs[r + 1])) for r in range(world)]


class RowSharding:
    """Contiguous row shards of a length-`dim` vector over `world` ranks and the all-gather that rebuilds the full
    vector (device agnostic: NCCL on GPUs, gloo in the CPU tests).  Ragged shards are padded to the largest one
    for the collective and compacted afterwards."""

    def __init__(self, dim: int, rank: int, world: int, t_dtype, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.dim, self.rank, self.world, self.group = dim, rank, world, group
        self.ranges = split_rows(dim, world)
        self.lo, self.hi = self.ranges[rank]
        self.max_rows = max(h - l for l, h in self.ranges)
        self.uniform = all(h - l == self.max_rows for l, h in self.ranges)
        self.x_full = torch.zeros(dim if world > 1 else 0, dtype=t_dtype, device=device)
        if not self.uniform:
            self.x_pad = torch.zeros(self.max_rows * world, dtype=t_dtype, device=device)
            self.send = torch.zeros(self.max_rows, dtype=t_dtype, device=device)

    def gather(self, x_local):
        """all-gather the rank-local rows into the full-length vector."""
        dist = self.dist
        if self.world == 1:
            return x_local                    # one rank owns every row: the local vector is the full vector, no copy
        elif self.uniform:
            dist.all_gather_into_tensor(self.x_full, x_local, group=self.group)
        else:
            self.send[: self.hi - self.lo].copy_(x_local)
            dist.all_gather_into_tensor(self.x_pad, self.send, group=self.group)
            for r, (l, h) in enumerate(self.ranges):
                self.x_full[l:h].copy_(self.x_pad[r * self.max_rows: r * self.max_rows + (h - l)])
        return self.x_full


class ShardedMatvec:
    """Row-sharded y = H x over torch.distributed: x shards are all-gathered (NCCL over NVLink on GPUs),
    then every rank applies its rows.  `opr` is this rank's representation (set_rows is called here)."""

    def __init__(self, opr, rank: int, world: int, dtype=None, group=None):
        import torch
        self.torch = torch
        self.opr, self.rank, self.world, self.group = opr, rank, world, group
        self.dim = opr.dimension
        self.np_dtype = np.dtype(dtype or (np.complex128 if opr.is_complex else np.float64))
        self.t_dtype = torch.complex128 if self.np_dtype == np.complex128 else torch.float64
        self.code = ED_C128 if self.np_dtype == np.complex128 else ED_F64
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.sharding = RowSharding(self.dim, rank, world, self.t_dtype, self.dev, group)
        self.dist = self.sharding.dist
        self.ranges = self.sharding.ranges
        self.lo, self.hi = self.sharding.lo, self.sharding.hi
        self.local_ranges = [(self.lo, self.hi, 0)]      # (global lo, global hi, offset in the local vector)
        self.n_local = self.hi - self.lo
        opr.set_rows(self.lo, self.hi)

    @property
    def x_full(self):
        return self.sharding.x_full

    def gather(self, x_local):
        return self.sharding.gather(x_local)

    def apply_local(self, y_local, x_full, dot_out=None):
        torch = self.torch
        check(lib.ed_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream), 1))
        try:
            check(lib.ed_apply_async(self.opr._handle, y_local.data_ptr(), x_full.data_ptr(), self.code, ED_SIDE_LEFT, 0,
                                     dot_out.data_ptr() if dot_out is not None else None))
        finally:
            lib.ed_set_stream(None, 0)
        return y_local

    def matvec(self, y_local, x_local, dot_out=None):
        return self.apply_local(y_local, self.gather(x_local), dot_out)


class DeviceBuffer:
    """Device memory allocated by the library (cudaMalloc) so that it can be exported to the other per-GPU
    processes of the node through CUDA IPC; viewed as a torch tensor through __cuda_array_interface__."""

    def __init__(self, n: int, np_dtype):
        self.n, self.np_dtype = int(n), np.dtype(np_dtype)
        p = C.c_void_p()
        check(lib.ed_device_malloc(max(self.n, 1) * self.np_dtype.itemsize, C.byref(p)))
        self.ptr = p.value
        self.__cuda_array_interface__ = {"shape": (self.n,), "typestr": self.np_dtype.str, "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def tensor(self):
        import torch
        return torch.as_tensor(self, device=torch.device("cuda", torch.cuda.current_device()))

    def ipc_handle(self) -> bytes:
        h = (C.c_uint8 * 64)()
        check(lib.ed_ipc_get_handle(C.c_void_p(self.ptr), h))
        return bytes(h)

    def __del__(self):
        if getattr(self, "ptr", None) and lib is not None:
            try:
                lib.ed_device_free(C.c_void_p(self.ptr))
            except Exception:
                pass


class P2PShardedMatvec:
    """Row-sharded y = H x WITHOUT an all-gather: every rank keeps its rows of x in a buffer that the other ranks map
    through CUDA IPC, and the matvec kernel pulls the few neighbour tiles it needs straight over NVLink (peer loads),
    overlapped with its local work.  Shards are tile aligned and wrap aware (ed_oprep_suggest_row_ranges): for a ring
    every rank owns the same range of high bits in BOTH halves of the basis (top site empty / occupied), i.e. two row
    ranges stored back to back in its local vectors, so that the periodic bond never leaves the rank.  `n_buffers`
    shared buffers are kept so that a Lanczos loop can ping-pong between them.  Only the U(1) fast-path kernel consumes
    segmented inputs."""

    def __init__(self, opr, rank: int, world: int, dtype=None, group=None, n_buffers: int = 2, exchange: str = "p2p"):
        """exchange = "p2p": the kernel loads peer tiles itself over NVLink; "dma": split exchange -- copy engines pull the
        needed peer rows into a local mirror vector while a first kernel pass does everything that is rank-local, a
        second pass adds the contributions of the mirrored rows (ed_oprep_set_exchange / ed_oprep_remote_rows)."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.opr, self.rank, self.world, self.group = opr, rank, world, group
        self.dim = opr.dimension
        self.np_dtype = np.dtype(dtype or (np.complex128 if opr.is_complex else np.float64))
        self.t_dtype = torch.complex128 if self.np_dtype == np.complex128 else torch.float64
        self.code = ED_C128 if self.np_dtype == np.complex128 else ED_F64
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.dma = exchange == "dma"
        self.rank_ranges = []          # per rank: its 1 or 2 (lo, hi) row ranges
        for r in range(world):
            lo, hi, n = (C.c_int64 * 2)(), (C.c_int64 * 2)(), C.c_int32()
            check(lib.ed_oprep_suggest_row_ranges(opr._handle, self.code, world, r, lo, hi, C.byref(n)))
            self.rank_ranges.append([(lo[i], hi[i]) for i in range(n.value)])
        self.local_ranges, off = [], 0   # (global lo, global hi, offset in the local vector)
        for lo, hi in self.rank_ranges[rank]:
            self.local_ranges.append((lo, hi, off))
            off += hi - lo
        self.n_local = off
        self.lo, self.hi = self.rank_ranges[rank][0]
        self.ranges = [rr[0] for rr in self.rank_ranges]
        n_local = self.n_local
        self.bufs = [DeviceBuffer(n_local, self.np_dtype) for _ in range(n_buffers)]
        self.views = [b.tensor() for b in self.bufs]
        for v in self.views:
            v.zero_()
        torch.cuda.synchronize()
        mine = [b.ipc_handle() for b in self.bufs]
        if world > 1:
            allh = [None] * world
            dist.all_gather_object(allh, mine, group=group)
        else:
            allh = [mine]
        self._opened = []
        base_ptr = []      # per buffer: base pointer of every rank's local vector as seen from this process
        for b in range(n_buffers):
            ptrs = []
            for r in range(world):
                if r == rank:
                    ptrs.append(self.bufs[b].ptr)
                else:
                    p = C.c_void_p()
                    check(lib.ed_ipc_open_handle((C.c_uint8 * 64).from_buffer_copy(allh[r][b]), C.byref(p)))
                    self._opened.append(p.value)
                    ptrs.append(p.value)
            base_ptr.append(ptrs)
        # segments in ascending global row order: (lo, rank, byte offset inside that rank's vector)
        segs = []
        for r, rr in enumerate(self.rank_ranges):
            o = 0
            for lo, hi in rr:
                if hi > lo:
                    segs.append((lo, r, o * self.np_dtype.itemsize))
                o += hi - lo
        segs.sort()
        self.n_seg = len(segs)
        self.seg_lo = (C.c_int64 * (self.n_seg + 1))(*([s_[0] for s_ in segs] + [self.dim]))
        self.seg_ptr = [[base_ptr[b][r] + o for (_, r, o) in segs] for b in range(n_buffers)]
        self._token = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self._dot_tmp = torch.zeros(4, 2, dtype=torch.float64, device=self.dev)
        self.local_seg_mask = sum(1 << i for i, (_, r, _) in enumerate(segs) if r == rank)
        if self.dma:
            self._setup_dma(segs, base_ptr, n_buffers)

    def _setup_dma(self, segs, base_ptr, n_buffers):
        """Mirror vector + the list of peer copies that fill the rows the local tiles read from other ranks."""
        torch = self.torch
        nr = len(self.local_ranges)
        lo = (C.c_int64 * nr)(*[r[0] for r in self.local_ranges])
        hi = (C.c_int64 * nr)(*[r[1] for r in self.local_ranges])
        n = C.c_int32()
        check(lib.ed_oprep_remote_rows(self.opr._handle, self.code, nr, lo, hi, 0, None, None, C.byref(n)))
        out_lo, out_hi = (C.c_int64 * max(n.value, 1))(), (C.c_int64 * max(n.value, 1))()
        check(lib.ed_oprep_remote_rows(self.opr._handle, self.code, nr, lo, hi, n.value, out_lo, out_hi, C.byref(n)))
        self.mirror = torch.empty(self.dim, dtype=self.t_dtype, device=self.dev)
        item = self.np_dtype.itemsize
        seg_hi = [s_[0] for s_ in segs[1:]] + [self.dim]
        rows = [sum(h - l for l, h in rr) for rr in self.rank_ranges]

        class _Peer:          # torch view of a peer rank's whole local vector (IPC-mapped pointer)
            def __init__(self, ptr, n_, typestr):
                self.__cuda_array_interface__ = {"shape": (n_,), "typestr": typestr, "data": (ptr, False), "version": 2, "strides": None}

        self._peer_views = [[torch.as_tensor(_Peer(base_ptr[b][r], max(rows[r], 1), self.np_dtype.str), device=self.dev)
                             for r in range(self.world)] for b in range(n_buffers)]
        self.copies = []       # (mirror start, rank, start inside that rank's vector, length)
        for a, b_ in zip(list(out_lo)[: n.value], list(out_hi)[: n.value]):
            for (s_lo, r, off_bytes), s_hi in zip(segs, seg_hi):
                c0, c1 = max(a, s_lo), min(b_, s_hi)
                if c1 > c0:
                    assert r != self.rank
                    self.copies.append((c0, r, off_bytes // item + (c0 - s_lo), c1 - c0))
        self.remote_rows = sum(c[3] for c in self.copies)
        self._copy_streams = [torch.cuda.Stream(device=self.dev) for _ in range(4)]

    def x_buffer(self, which: int = 0):
        """This rank's rows of shared input buffer `which` (a torch tensor; write x here; layout = local_ranges)."""
        return self.views[which]

    def fence(self):
        """Stream-ordered barrier across ranks: call after writing a shared buffer, before peers read it."""
        if self.world > 1:
            self.dist.all_reduce(self._token, group=self.group)

    def _apply_ranges(self, y_local, dots):
        """one ed_apply_async per local row range; dots: list of (2,) tensors or None"""
        for i, (lo, hi, off) in enumerate(self.local_ranges):
            dot = None if dots is None else dots[i]
            if hi <= lo:
                if dot is not None:
                    dot.zero_()
                continue
            self.opr.set_rows(lo, hi)
            check(lib.ed_apply_async(self.opr._handle, y_local[off:].data_ptr(), None, self.code, ED_SIDE_LEFT, 0,
                                     dot.data_ptr() if dot is not None else None))

    def matvec(self, y_local, which: int = 0, dot_out=None):
        torch = self.torch
        ptrs = (C.c_void_p * self.n_seg)(*self.seg_ptr[which])
        nr = len(self.local_ranges)
        main = torch.cuda.current_stream()
        check(lib.ed_oprep_set_x_segments(self.opr._handle, self.n_seg, self.seg_lo, ptrs))
        check(lib.ed_set_stream(C.c_void_p(main.cuda_stream), 1))
        n_dots = nr
        try:
            if not self.dma:
                self._apply_ranges(y_local, None if dot_out is None else [self._dot_tmp[i] for i in range(nr)])
            else:
                ready = torch.cuda.Event()
                ready.record(main)                      # x is written and fenced at this point of the stream
                check(lib.ed_oprep_set_exchange(self.opr._handle, 1, None, self.local_seg_mask))
                self._apply_ranges(y_local, None if dot_out is None else [self._dot_tmp[i] for i in range(nr)])
                used = set()
                for i, (m0, r, s0, ln) in enumerate(self.copies):   # copy engines, concurrent with the local pass
                    st = self._copy_streams[i % len(self._copy_streams)]
                    if i % len(self._copy_streams) not in used:
                        st.wait_event(ready)
                        used.add(i % len(self._copy_streams))
                    with torch.cuda.stream(st):
                        self.mirror[m0:m0 + ln].copy_(self._peer_views[which][r][s0:s0 + ln], non_blocking=True)
                for j in used:
                    ev = torch.cuda.Event()
                    ev.record(self._copy_streams[j])
                    main.wait_event(ev)
                check(lib.ed_oprep_set_exchange(self.opr._handle, 2, C.c_void_p(self.mirror.data_ptr()), self.local_seg_mask))
                self._apply_ranges(y_local, None if dot_out is None else [self._dot_tmp[nr + i] for i in range(nr)])
                n_dots = 2 * nr
        finally:
            lib.ed_oprep_set_exchange(self.opr._handle, 0, None, 0)
            lib.ed_set_stream(None, 0)
            lib.ed_oprep_set_x_segments(self.opr._handle, 0, None, None)
        if dot_out is not None:
            torch.sum(self._dot_tmp[:n_dots], dim=0, out=dot_out)
        return y_local

    def close(self):
        """Collective: every rank stops using the shared buffers, then unmaps its peers' memory.  Call it before the
        object goes away -- an owner that frees a buffer its peers still map, and later exports a new buffer that
        reuses the allocation, makes the peers' next ed_ipc_open_handle fail with "resource already mapped"."""
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)
        self._unmap()
        if self.world > 1:
            self.dist.barrier(group=self.group)      # nobody frees its buffers before every peer has unmapped them

    def _unmap(self):
        for p in self._opened:
            lib.ed_ipc_close_handle(C.c_void_p(p))
        self._opened = []

    def __del__(self):
        # last resort when close() was not called: at least drop this process's mappings of the peers' buffers
        try:
            self._unmap()
        except Exception:
            pass


class ShardedLanczos:
    """Three-term Lanczos with unnormalised, device-resident, row-sharded Krylov vectors (see csrc/lanczos.cu).
    Per step: one all-gather (x), two scalar all-reduces (<u,Hu> and |u_next|^2); no host synchronisation."""

    def __init__(self, opr, rank: int = 0, world: int = 1, dtype=None, group=None, exchange: str = "allgather"):
        """exchange = "allgather" (NCCL all-gather of x per step) or "p2p" (peer loads of the far tiles, no gather)."""
        self.p2p = exchange in ("p2p", "dma")
        self.mv = P2PShardedMatvec(opr, rank, world, dtype, group, exchange=exchange) if self.p2p else ShardedMatvec(opr, rank, world, dtype, group)
        torch = self.mv.torch
        n = self.mv.n_local
        self.n_local = n
        mk = lambda: torch.zeros(max(n, 1), dtype=self.mv.t_dtype, device=self.mv.dev)[:n]
        if self.p2p:
            self.u_cur, self.u_prev, self.w = self.mv.x_buffer(0), self.mv.x_buffer(1), mk()
            self._cur = 0
        else:
            self.u_cur, self.u_prev, self.w = mk(), mk(), mk()

    def _ed_stream(self):
        torch = self.mv.torch
        check(lib.ed_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream), 1))

    def close(self):
        """Collective (see P2PShardedMatvec.close); a no-op for the all-gather exchange."""
        if self.p2p:
            self.mv.close()

    def run(self, n_steps: int, seed: int = 0, v0_local=None, n_ritz: int = 4) -> LanczosResult:
        mv, torch, dist = self.mv, self.mv.torch, self.mv.dist
        dots = torch.zeros(n_steps, 2, dtype=torch.float64, device=mv.dev)
        norms = torch.zeros(n_steps + 1, 2, dtype=torch.float64, device=mv.dev)
        self.u_prev.zero_()
        self._ed_stream()
        try:
            if v0_local is not None:
                self.u_cur.copy_(v0_local)
            else:
                for lo, hi, off in mv.local_ranges:      # keyed by the global row index: shard-count independent
                    if hi > lo:
                        check(lib.ed_vector_randn_async(self.u_cur[off:].data_ptr(), hi - lo, mv.code, seed, lo))
            check(lib.ed_vector_norm2_async(self.u_cur.data_ptr(), self.n_local, mv.code, norms[0].data_ptr()))
        finally:
            lib.ed_set_stream(None, 0)
        if mv.world > 1:
            dist.all_reduce(norms[0], group=mv.group)
        for j in range(n_steps):
            if self.p2p:
                # peers may read u_cur only after its owner finished writing it: the all-reduce of norms[j] above
                # (stream ordered, issued after the update kernel) is that fence
                mv.matvec(self.w, self._cur, dots[j])
                self._cur ^= 1
            else:
                mv.matvec(self.w, self.u_cur, dots[j])
            if mv.world > 1:
                dist.all_reduce(dots[j], group=mv.group)
            self._ed_stream()
            try:
                check(lib.ed_lanczos_update_async(self.u_prev.data_ptr(), self.w.data_ptr(), self.u_cur.data_ptr(), self.n_local,
                                                  mv.code, dots[j].data_ptr(), norms[j].data_ptr(),
                                                  norms[j - 1].data_ptr() if j > 0 else None, norms[j + 1].data_ptr()))
            finally:
                lib.ed_set_stream(None, 0)
            if mv.world > 1:
                dist.all_reduce(norms[j + 1], group=mv.group)
            self.u_cur, self.u_prev = self.u_prev, self.u_cur
        torch.cuda.synchronize()
        hd, hn = dots.cpu().numpy(), norms.cpu().numpy()
        alpha, beta = [], []
        for j in range(n_steps):
            if not (hn[j, 0] > 0.0) or not np.isfinite(hn[j, 0]):
                break
            alpha.append(hd[j, 0] / hn[j, 0])
            beta.append(float(np.sqrt(hn[j + 1, 0])))
            if not (beta[-1] > 1e-13 * abs(alpha[-1]) + 1e-300):
                break
        alpha, beta = np.array(alpha), np.array(beta)
        ritz = tridiag_eigvals(alpha, beta[:-1] if len(beta) else beta)[:n_ritz] if len(alpha) else np.array([])
        return LanczosResult(alpha, beta, ritz, len(alpha))
