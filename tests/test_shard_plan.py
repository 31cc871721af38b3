"""CPU checks of the multi-GPU row partition and halo exchange that ed_sharded_create builds for the tiled U(1) kernel
(ed_u1_shard_layout, csrc/apply_u1.cu), through the host-only C-ABI entry ed_shard_plan_describe -- no GPU needed:

  * every rank's plan replayed on numpy arrays: the ranges partition the basis, every tile a rank reads resolves -- through
    its directory -- to exactly the right global rows of x, in its own vector or in the halo buffer filled by its pulls,
    and no launch chunk reads halo data that an earlier-or-equal chunk's pulls did not bring;
  * a world-2 `gloo` run (one process per rank): pulls executed as real inter-process transfers, each rank's rows computed
    by the oracle's C twin from ONLY its local + halo data (everything else poisoned with NaN), Lanczos scalars
    all-reduced; the result must equal the unsharded oracle.  The GPU kernels cannot run here (no CPU fallback): the C
    twin stands in for the per-rank kernel, exactly the row-owner contract the kernel implements.
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import ed_oracle as O
import ed_oracle_c as OC
from helpers import oracle_spin_chain

HALO = 1 << 62


def _terms(n, model):
    if model == "xxz":
        bonds = [(i, (i + 1) % n) for i in range(n)]
        return O.term_arrays(oracle_spin_chain(n, bonds, jz=0.37)[1])
    if model == "open":
        return O.term_arrays(oracle_spin_chain(n, [(i, i + 1) for i in range(n - 1)], jz=-1.1, jxy=0.8)[1])
    if model == "j1j2":
        _, a = oracle_spin_chain(n, [(i, (i + 1) % n) for i in range(n)])
        _, b = oracle_spin_chain(n, [(i, (i + 2) % n) for i in range(n)], jz=0.5, jxy=0.5)
        return O.term_arrays(O.simplify(a + b))
    raise ValueError(model)


def describe(ed, terms, n, n_dn, world, rank, n_chunks=4, policy=0, dtype=0):
    from edcuda._lib import lib, check
    op = ed.Operator([(int(m), int(r), int(c), float(a)) for m, r, c, a in zip(*terms)])
    counts = (C.c_int64 * 12)()
    args = (op.handle(), n, n_dn, dtype, world, rank, n_chunks, policy)
    check(lib.ed_shard_plan_describe(*args, counts, None, None, None, None, None, None))
    n_local, n_halo, n_ranges, n_tiles, n_pulls, n_reads, chunks, dim, n_packs, n_send, n_pushes, _ = list(counts)
    pushes = np.zeros((max(n_pushes, 1), 5), dtype=np.int64)
    ranges = np.zeros((max(n_ranges, 1), 2), dtype=np.int64)
    tiles = np.zeros((max(n_tiles, 1), 4), dtype=np.int64)
    pulls = np.zeros((max(n_pulls, 1), 5), dtype=np.int64)
    packs = np.zeros((max(n_packs, 1), 3), dtype=np.int64)
    reads = np.zeros((max(n_reads, 1), 4), dtype=np.int64)
    check(lib.ed_shard_plan_describe(*args, counts, ranges.ctypes.data, tiles.ctypes.data, pulls.ctypes.data, packs.ctypes.data,
                                     reads.ctypes.data, pushes.ctypes.data))
    return dict(n_local=n_local, n_halo=n_halo, n_send=n_send, dim=dim, chunks=chunks, ranges=ranges[:n_ranges], tiles=tiles[:n_tiles],
                pulls=pulls[:n_pulls], packs=packs[:n_packs], reads=reads[:n_reads], pushes=pushes[:n_pushes])


def send_buffer(plan, u_local):
    """what the owner-side pack kernel gathers for the peers"""
    send = np.full(plan["n_send"], np.nan)
    for src, dst, ln in plan["packs"]:
        send[dst:dst + ln] = u_local[src:src + ln]
    assert not np.isnan(send).any()
    return send


def local_vector(plan, x):
    return np.concatenate([x[lo:hi] for lo, hi in plan["ranges"]]) if len(plan["ranges"]) else x[:0]


@pytest.mark.parametrize("n,n_dn,model,world,chunks,policy", [
    (20, 10, "xxz", 2, 4, 0), (20, 10, "xxz", 4, 3, 0), (22, 11, "xxz", 8, 8, 0), (20, 7, "j1j2", 3, 2, 0),
    (18, 9, "open", 2, 1, 0), (20, 10, "xxz", 4, 4, 1), (24, 12, "xxz", 8, 8, 0), (16, 8, "xxz", 16, 2, 0),
    (20, 10, "xxz", 2, 2, 2), (22, 11, "xxz", 4, 4, 3), (20, 7, "j1j2", 8, 3, 3)])
def test_plans_replayed_on_numpy(ed, n, n_dn, model, world, chunks, policy):
    terms = _terms(n, model)
    plans = [describe(ed, terms, n, n_dn, world, r, chunks, policy) for r in range(world)]
    dim = plans[0]["dim"]
    assert dim == len(OC.basis_fixed_popcount(n, n_dn))
    x = np.arange(dim, dtype=np.float64) * 1.25 + 3.0
    # the ranges partition the basis
    cover = np.zeros(dim, dtype=np.int32)
    for p in plans:
        assert p["n_local"] == sum(hi - lo for lo, hi in p["ranges"])
        for lo, hi in p["ranges"]:
            cover[lo:hi] += 1
    assert np.all(cover == 1)
    rows = [p["n_local"] for p in plans]
    assert max(rows) <= 1.35 * dim / world + 7000          # balanced up to tile granularity
    local = [local_vector(p, x) for p in plans]
    send = [send_buffer(p, local[r]) for r, p in enumerate(plans)]
    assert sum(p["n_send"] for p in plans) == sum(p["n_halo"] for p in plans)
    for r, p in enumerate(plans):
        # own tiles: local offsets follow the range layout
        for gbase, size, off, chunk in p["tiles"]:
            assert np.array_equal(local[r][off:off + size], x[gbase:gbase + size])
        assert np.all(np.diff(p["tiles"][:, 3]) >= 0) and (len(p["tiles"]) == 0 or p["tiles"][-1, 3] < p["chunks"])
        assert len(p["pulls"]) <= (world - 1) * p["chunks"]         # one contiguous piece per peer and chunk
        halo = np.full(p["n_halo"], np.nan)
        ready = np.full(p["n_halo"], 1 << 30, dtype=np.int64)      # chunk whose pulls bring each halo element
        assert np.all(np.diff(p["pulls"][:, 1]) >= 0)
        for peer, chunk, src, dst, ln in p["pulls"]:
            assert peer != r and 0 <= peer < world
            halo[dst:dst + ln] = send[peer][src:src + ln]
            ready[dst:dst + ln] = chunk
        assert not np.isnan(halo).any()                             # the pulls fill the halo completely, nothing twice
        assert sum(pl[4] for pl in p["pulls"]) == p["n_halo"]
        # push exchange: the peers write the same tiles straight into this rank's halo, earliest chunk first
        halo2 = np.full(p["n_halo"], np.nan)
        ready2 = np.full(p["n_halo"], 1 << 30, dtype=np.int64)
        for s_rank, q in enumerate(plans):
            assert np.all(np.diff(q["pushes"][:, 1]) >= 0)
            for recv, chunk, src, dst, ln in q["pushes"]:
                if recv == r:
                    assert s_rank != r and np.isnan(halo2[dst:dst + ln]).all()
                    halo2[dst:dst + ln] = local[s_rank][src:src + ln]
                    ready2[dst:dst + ln] = chunk
        assert np.array_equal(halo2, halo) and np.array_equal(ready2, ready)
        for ti, gbase, size, where in p["reads"]:
            assert where >= 0
            if where & HALO:
                off = where & (HALO - 1)
                assert np.array_equal(halo[off:off + size], x[gbase:gbase + size])
                assert ready[off:off + size].max() <= p["tiles"][ti, 3]    # brought by this chunk's pulls or earlier
            else:
                assert np.array_equal(local[r][where:where + size], x[gbase:gbase + size])
    if model == "xxz" and policy == 0 and n >= 20:
        # the planner's choice never moves more than the plain or the wrap-aware contiguous split
        for pol in (1, 2):
            other = [describe(ed, terms, n, n_dn, world, r, chunks, pol) for r in range(world)]
            assert max(p["n_halo"] for p in plans) <= max(p["n_halo"] for p in other)
    if world == 2 and model == "xxz" and policy == 2:
        assert all(len(p["ranges"]) == 2 for p in plans)        # wrap-aware shards: two ranges per rank


def test_describe_rejects_unsupported(ed):
    from edcuda._lib import lib
    hs, pauli = ed.spin_half_system(8)
    op = ed.simplify(pauli(0, "x") * pauli(1, "x") * pauli(2, "x"))      # three-site term: not a bond Hamiltonian
    counts = (C.c_int64 * 12)()
    assert lib.ed_shard_plan_describe(op.handle(), 8, 4, 0, 2, 0, 2, 0, counts, None, None, None, None, None, None) == ed._lib.ED_ERR_UNSUPPORTED


# ------------------------------------------------------------------ world-2 gloo run
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, steps, q):
    import sys
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, os.path.join(root, "exactdiagonalization.jl_b200"))
        import edcuda as ed
        terms = _terms(n, "xxz")
        basis = OC.basis_fixed_popcount(n, n // 2)
        dim = len(basis)
        plan = describe(ed, terms, n, n // 2, world, rank, 3, 0)
        rows = [None] * world
        dist.all_gather_object(rows, int(plan["n_send"]))
        cap = max(rows)

        def exchange(u_local):
            """the pulls as real transfers: every rank publishes its packed send buffer, peers slice what their plan lists"""
            sb = send_buffer(plan, u_local)
            pad = torch.zeros(cap, dtype=torch.float64)
            pad[: len(sb)] = torch.from_numpy(sb)
            allv = [torch.zeros(cap, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(allv, pad)
            halo = np.full(plan["n_halo"], np.nan)
            for peer, chunk, src, dst, ln in plan["pulls"]:
                halo[dst:dst + ln] = allv[peer][src:src + ln].numpy()
            return halo

        def matvec(u_local):
            halo = exchange(u_local)
            xv = np.full(dim, np.nan)                       # everything the rank does not hold is poison
            off = 0
            for lo, hi in plan["ranges"]:
                xv[lo:hi] = u_local[off:off + hi - lo]
                off += hi - lo
            for ti, gbase, size, where in plan["reads"]:
                if where & HALO:
                    o = where & (HALO - 1)
                    xv[gbase:gbase + size] = halo[o:o + size]
            out = []
            for lo, hi in plan["ranges"]:
                w = np.zeros(hi - lo)
                OC.apply(basis, terms, xv, w, lo, hi)
                out.append(w)
            w = np.concatenate(out)
            assert not np.isnan(w).any(), "the kernel would read a row that is neither local nor in the halo"
            return w

        rng = np.random.default_rng(7)
        v0 = rng.standard_normal(dim)
        u_cur = local_vector(plan, v0)
        u_prev = np.zeros_like(u_cur)

        def allsum(v):
            t = torch.tensor([float(v)], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t)

        n2 = [allsum(u_cur @ u_cur)]
        alphas, betas = [], []
        for j in range(steps):
            w = matvec(u_cur)
            d = allsum(u_cur @ w)
            nc = np.sqrt(n2[j])
            alpha = d / n2[j]
            c3 = nc / np.sqrt(n2[j - 1]) if j > 0 else 0.0
            u_next = (w - alpha * u_cur) / nc - c3 * u_prev
            n2.append(allsum(u_next @ u_next))
            alphas.append(alpha); betas.append(np.sqrt(n2[j + 1]))
            u_prev, u_cur = u_cur, u_next
        # one more matvec gathered by global rows for a direct comparison
        y = matvec(local_vector(plan, v0))
        parts = [None] * world
        dist.all_gather_object(parts, (plan["ranges"].tolist(), y))
        if rank == 0:
            full = np.zeros(dim)
            for rg, yy in parts:
                off = 0
                for lo, hi in rg:
                    full[lo:hi] = yy[off:off + hi - lo]
                    off += hi - lo
            q.put(("ok", alphas, betas, full, v0))
    except Exception:
        import traceback
        q.put(("error", traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


def test_world2_gloo_halo_exchange_and_lanczos():
    import torch.multiprocessing as mp
    n, steps, world = 16, 30, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert res[0] == "ok", res[1]
    _, alphas, betas, full, v0 = res
    terms = _terms(n, "xxz")
    basis = OC.basis_fixed_popcount(n, n // 2)
    exp = np.zeros(len(basis))
    OC.apply(basis, terms, v0, exp)
    assert np.max(np.abs(full - exp)) <= 1e-13 * np.max(np.abs(exp))

    def mv(v):
        out = np.zeros_like(v)
        OC.apply(basis, terms, v, out)
        return out

    a_ref, b_ref = O.lanczos(mv, v0, steps)
    assert np.allclose(alphas, a_ref, atol=1e-9) and np.allclose(betas, b_ref, atol=1e-9)


def test_refinement_trades_row_balance_for_halo(ed, monkeypatch):
    """the local refinement of the tile cut (csrc/apply_u1.cu: u1_global_layout) lowers the largest halo where the transfer
    dominates, keeps every shard within a bounded row imbalance, covers the basis exactly once and is the same on every rank"""
    n, n_dn, world, chunks = 26, 13, 8, 8
    terms = _terms(n, "xxz")
    monkeypatch.setenv("EDCUDA_SHARD_REFINE", "0")
    plain = [describe(ed, terms, n, n_dn, world, r, chunks, 0) for r in range(world)]
    monkeypatch.delenv("EDCUDA_SHARD_REFINE")
    fine = [describe(ed, terms, n, n_dn, world, r, chunks, 0) for r in range(world)]
    dim = fine[0]["dim"]
    assert sum(p["n_local"] for p in fine) == dim == sum(p["n_local"] for p in plain)
    cover = np.zeros(dim, dtype=np.int32)
    for p in fine:
        for lo, hi in p["ranges"]:
            cover[lo:hi] += 1
    assert cover.min() == 1 and cover.max() == 1
    halo_plain, halo_fine = max(p["n_halo"] for p in plain), max(p["n_halo"] for p in fine)
    rows_fine = max(p["n_local"] for p in fine)
    assert halo_fine <= halo_plain
    assert max(rows_fine, 1.5 * halo_fine) <= max(max(p["n_local"] for p in plain), 1.5 * halo_plain)   # the objective it minimises
    assert rows_fine <= 1.25 * dim / world
    again = describe(ed, terms, n, n_dn, world, 3, chunks, 0)                # a second evaluation (another rank's view): identical
    assert np.array_equal(again["ranges"], fine[3]["ranges"]) and again["n_halo"] == fine[3]["n_halo"]
