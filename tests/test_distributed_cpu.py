"""world_size-2 gloo test (CPU) of the row-sharding plumbing used by the multi-GPU path: contiguous balanced
shards, ragged all-gather, and the sharded Lanczos recurrence with all-reduced scalars.  The compute kernels are
GPU-only (no CPU fallback), so the per-rank matvec here is the oracle's C twin restricted to the rank's rows --
exactly the row-owner contract ed_oprep_set_rows gives the GPU kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ed_oracle as O
import ed_oracle_c as OC
from helpers import oracle_spin_chain


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "exactdiagonalization.jl_b200"))
    from edcuda.lanczos import RowSharding, split_rows
    hs, h = oracle_spin_chain(n)
    terms = O.term_arrays(h)
    basis = OC.basis_fixed_popcount(n, n // 2)
    dim = len(basis)
    sh = RowSharding(dim, rank, world, torch.float64, torch.device("cpu"))
    lo, hi = sh.lo, sh.hi
    rng = np.random.default_rng(7)
    v0 = rng.standard_normal(dim)
    u_cur = torch.from_numpy(v0[lo:hi].copy())
    u_prev = torch.zeros(hi - lo, dtype=torch.float64)
    n2 = [None] * (steps + 1)
    t = torch.tensor([float(u_cur @ u_cur)])
    dist.all_reduce(t)
    n2[0] = float(t)
    alphas, betas = [], []
    for j in range(steps):
        x_full = sh.gather(u_cur).numpy()
        w = np.zeros(hi - lo)
        OC.apply(basis, terms, x_full, w, lo, hi)          # this rank's rows only
        d = torch.tensor([float(u_cur.numpy() @ w)])
        dist.all_reduce(d)
        nc = np.sqrt(n2[j])
        alpha = float(d) / n2[j]
        c3 = nc / np.sqrt(n2[j - 1]) if j > 0 else 0.0
        u_next = (torch.from_numpy(w) - alpha * u_cur) / nc - c3 * u_prev
        t = torch.tensor([float(u_next @ u_next)])
        dist.all_reduce(t)
        n2[j + 1] = float(t)
        alphas.append(alpha); betas.append(np.sqrt(n2[j + 1]))
        u_prev, u_cur = u_cur, u_next
    if rank == 0:
        q.put((sh.ranges, alphas, betas, v0))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11])
def test_sharded_lanczos_gloo_world2(n):
    from edcuda.lanczos import split_rows
    assert split_rows(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert split_rows(7, 2) == [(0, 4), (4, 7)]
    steps = 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    ranges, alphas, betas, v0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hs, h = oracle_spin_chain(n)
    hsr = O.represent(O.HilbertSpaceSector(hs, n - 2 * (n // 2)))
    dim = hsr.dimension
    assert ranges[0][0] == 0 and ranges[-1][1] == dim and ranges[0][1] == ranges[1][0]
    assert abs((ranges[0][1] - ranges[0][0]) - (ranges[1][1] - ranges[1][0])) <= 1      # n=11: ragged shards
    opr = O.OperatorRepresentation(hsr, h)
    a_ref, b_ref = O.lanczos(lambda v: O.apply_vectorized(np.zeros(dim), opr, v), v0, steps)
    assert np.allclose(alphas, a_ref, atol=1e-9) and np.allclose(betas, b_ref, atol=1e-9)
