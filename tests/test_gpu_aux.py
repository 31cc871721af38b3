"""vector symmetry_reduce / symmetry_unreduce (K8) against the oracle on random vectors, the
isinvariant gate of represent(rhsr, op), basis / reduced-basis checkpoints and the resumable Lanczos state."""
import os

import numpy as np
import pytest

import ed_oracle as O
import ed_oracle_np as ON
from helpers import oracle_spin_chain, rel_err, to_oracle_symops

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("n,qn,ks", [(10, 0, (0, 3, 5)), (9, 1, (0, 2)), (12, 2, (1, 6))])
def test_vector_reduce_unreduce_vs_oracle(gpu_ed, n, qn, ks):
    """symmetry_reduce(rhsr, large) / symmetry_reduce!(out, ...) / symmetry_unreduce (Symmetry/symmetry_reduce.jl:38-56,
    208-225): small[idx[p]] += conj(amp[p]) * large[p] and large[p] = amp[p] * small[idx[p]] on random vectors with many
    orbits per sector -- real and complex large vectors, accumulate on and off, complex characters."""
    ed = gpu_ed
    hs, h = ed.models.heisenberg_chain(n)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, qn))
    basis = hsr.basis_list
    rng = np.random.default_rng(n + qn)
    for k in ks:
        symops = ed.lattices.chain_translation_irrep(n, k)
        if k == ks[0] and n % 2 == 0:       # non-abelian-looking list: translations x inversion (still 1-D characters)
            symops = ed.lattices.symmorphic_product(ed.lattices.chain_translation_irrep(n, 0), ed.lattices.chain_inversion_irrep(n, -1))
        perms = [ed.symmetry._flatten(op, n)[0] for op, _ in symops]
        chis = [c for _, c in symops]
        red, idx, amp = ON.symmetry_reduce(basis, perms, chis)
        rhsr = ed.symmetry_reduce(hsr, symops)
        d = rhsr.dimension
        assert d == len(red) and d > 3
        for cplx in (False, True):
            large = rng.standard_normal(len(basis)) + (1j * rng.standard_normal(len(basis)) if cplx else 0.0)
            exp = ON.vector_reduce(idx, amp, d, large)
            got = ed.symmetry_reduce(rhsr, large)
            assert got.dtype == np.complex128 and rel_err(got, exp) < TOL
            out = rng.standard_normal(d) + 1j * rng.standard_normal(d)      # symmetry_reduce!(out, ...) adds
            base = out.copy()
            ed.symmetry_reduce_b(out, rhsr, large)
            assert rel_err(out, base + exp) < TOL
            small = rng.standard_normal(d) + (1j * rng.standard_normal(d) if cplx else 0.0)
            un = ed.symmetry_unreduce(rhsr, small)
            assert rel_err(un, ON.vector_unreduce(idx, amp, small)) < TOL
            # unreduce is an isometry onto the sector and reduce its adjoint
            assert abs(np.linalg.norm(un) - np.linalg.norm(small)) < 1e-12 * np.linalg.norm(small)
            assert rel_err(ed.symmetry_reduce(rhsr, un), small.astype(complex)) < TOL
        with pytest.raises(ed.DimensionMismatch):
            ed.symmetry_reduce(rhsr, np.zeros(len(basis) + 1))
        with pytest.raises(ed.DimensionMismatch):
            ed.symmetry_unreduce(rhsr, np.zeros(d + 1))


def test_represent_reduced_rejects_non_invariant_operator(gpu_ed, monkeypatch):
    ed = gpu_ed
    n = 8
    hs, pauli = ed.spin_half_system(n)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    rhsr = ed.symmetry_reduce(hsr, ed.lattices.chain_translation_irrep(n, 2))
    ring = ed.models.heisenberg_bonds(hs, ed.lattices.chain_bonds(n))
    ed.represent(rhsr, ring)                                              # invariant: fine
    bad = ed.models.heisenberg_bonds(hs, ed.lattices.chain_bonds(n, periodic=False))
    with pytest.raises(ValueError, match="not invariant under symmetry element 1"):
        ed.represent(rhsr, bad)
    monkeypatch.setenv("EDCUDA_SKIP_INVARIANCE_CHECK", "1")               # the reference's behaviour: no check
    ed.represent(rhsr, bad)


def test_basis_checkpoints(gpu_ed, tmp_path, golden):
    ed = gpu_ed
    hs, h = ed.models.heisenberg_chain(14)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 2))
    p = str(tmp_path / "sector.edb")
    hsr.save(p)
    assert os.path.getsize(p) < 200                                       # no words: the generation parameters
    back = ed.HilbertSpaceRepresentation.load(hs, p)
    assert back.kind == hsr.kind == ed.ED_BASIS_COMBINADIC and np.array_equal(back.basis_list, hsr.basis_list)
    hs2, _ = ed.spin_half_system(13)
    with pytest.raises(ValueError):                                       # another Hilbert space
        ed.HilbertSpaceRepresentation.load(hs2, p)
    words = np.array(sorted(set(range(64)) - {5, 9, 33}), dtype=np.uint64)
    hs6, _ = ed.spin_half_system(6)
    lst = ed.represent(hs6, words)
    p2 = str(tmp_path / "list.edb")
    lst.save(p2)
    back2 = ed.HilbertSpaceRepresentation.load(hs6, p2)
    assert back2.kind == ed.ED_BASIS_LIST and np.array_equal(back2.basis_list, words)
    # tJ sector (DP-rank tables) round trip
    g = golden["tj"]
    sites = [ed.Site([ed.State(str(i), tuple(q)) for i, q in enumerate(states)]) for states in g["site_states"]]
    tj = ed.HilbertSpace(sites)
    sec = ed.represent(ed.HilbertSpaceSector(tj, [(3, 1), (2, 0)]))
    p3 = str(tmp_path / "tj.edb")
    sec.save(p3)
    assert np.array_equal(ed.HilbertSpaceRepresentation.load(tj, p3).basis_list, sec.basis_list)
    # reduced basis: same operator representation from the checkpoint, bound to the symmetry it was made with
    symops = ed.lattices.chain_translation_irrep(14, 3)
    rhsr = ed.symmetry_reduce(hsr, symops)
    pr = str(tmp_path / "reduced.edr")
    rhsr.save(pr)
    rback = ed.ReducedHilbertSpaceRepresentation.load(hsr, symops, pr)
    assert np.array_equal(rback.basis_list, rhsr.basis_list) and np.array_equal(rback.orbit_sizes(), rhsr.orbit_sizes())
    x = np.random.default_rng(0).standard_normal(rhsr.dimension) + 0j
    assert np.array_equal(ed.represent(rback, h) * x, ed.represent(rhsr, h) * x)
    assert np.array_equal(rback.basis_mapping_index, rhsr.basis_mapping_index)
    with pytest.raises(ValueError):                                       # other characters
        ed.ReducedHilbertSpaceRepresentation.load(hsr, ed.lattices.chain_translation_irrep(14, 4), pr)
    with pytest.raises(ValueError):                                       # not a reduced-basis file
        ed.ReducedHilbertSpaceRepresentation.load(hsr, symops, p)


@pytest.mark.parametrize("cplx", [False, True])
def test_lanczos_resume_bit_for_bit(gpu_ed, tmp_path, cplx):
    """run 25 steps, save, load, run 35 more == 60 uninterrupted steps: alpha and beta identical to the last bit."""
    ed = gpu_ed
    from edcuda.lanczos import LanczosState, lanczos
    hs, h = ed.models.j1j2_chain(16, 0.5)
    opr = ed.represent(ed.represent(ed.HilbertSpaceSector(hs, 0)), h)
    dtype = np.complex128 if cplx else np.float64
    full = LanczosState(opr, seed=5, dtype=dtype).step(60).result()
    assert full.steps == 60
    one_call = lanczos(opr, 60, seed=5, dtype=dtype)
    assert np.array_equal(full.alpha, one_call.alpha) and np.array_equal(full.beta, one_call.beta)
    st = LanczosState(opr, seed=5, dtype=dtype).step(25)
    p = str(tmp_path / "lanczos.edl")
    st.save(p)
    del st
    opr2 = ed.represent(ed.represent(ed.HilbertSpaceSector(hs, 0)), h)     # a fresh representation, as after a restart
    res = LanczosState.load(opr2, p, 25).step(35).result()
    assert res.steps == 60
    assert np.array_equal(res.alpha, full.alpha) and np.array_equal(res.beta, full.beta)
    assert abs(res.ritz[0] + 24.0) < 1e-8                                    # Majumdar-Ghosh: -1.5 L
    hs2, h2 = ed.models.heisenberg_chain(12)
    with pytest.raises(ed.DimensionMismatch):
        LanczosState.load(ed.represent(ed.represent(ed.HilbertSpaceSector(hs2, 0)), h2), p, 25)


def test_lanczos_stops_at_invariant_subspace(gpu_ed):
    """a start vector inside a small invariant subspace -- the loop must report the steps that are valid
    instead of continuing on noise and returning ghost Ritz values."""
    ed = gpu_ed
    from edcuda.lanczos import lanczos
    n = 10
    hs, h = ed.models.heisenberg_chain(n)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    opr = ed.represent(hsr, h)
    # momentum-0, fully symmetric start vector lives in the k=0 sector (dimension << 252)
    rhsr = ed.symmetry_reduce(hsr, ed.lattices.chain_translation_irrep(n, 0))
    v0 = ed.symmetry_unreduce(rhsr, np.ones(rhsr.dimension)).real.copy()
    res = lanczos(opr, 120, v0=v0)
    assert res.steps <= rhsr.dimension + 1 < 120
    full = np.linalg.eigvalsh(opr.matrix())
    sector = np.linalg.eigvalsh(ed.represent(rhsr, h).matrix())
    assert abs(res.ritz[0] - sector[0]) < 1e-9                               # lowest level of the k=0 sector
    assert all(np.min(np.abs(full - r)) < 1e-7 for r in res.ritz)           # no ghosts


def test_reduced_lookup_hash_and_buckets_agree(gpu_ed, monkeypatch):
    """the reduced index is looked up through the hash table (default) or the bucketed binary search (EDCUDA_RBASIS_HASH=0):
    same matvec bit for bit, same basis mapping, same reduce / unreduce of a vector; the staged and the row-per-thread kernels"""
    ed = gpu_ed
    n = 16
    hs, h = ed.models.heisenberg_chain(n)
    symops = ed.lattices.chain_translation_irrep(n, 3)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("EDCUDA_RBASIS_HASH", mode)
        hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
        rhsr = ed.symmetry_reduce(hsr, symops)
        x = np.random.default_rng(1).standard_normal(rhsr.dimension) * (1 + 0.5j)
        res = []
        for min_rows in ("1", "1000000"):                                # staged kernel / row-per-thread kernel
            monkeypatch.setenv("EDCUDA_K6_MIN_ROWS", min_rows)
            res.append(ed.represent(rhsr, h) * x)
        big = np.random.default_rng(2).standard_normal(hsr.dimension) + 0j
        out[mode] = (res, rhsr.basis_mapping_index.copy(), ed.symmetry_reduce(rhsr, big), ed.symmetry_unreduce(rhsr, x))
    for a, b in zip(out["1"][0] + [out["1"][1], out["1"][3]], out["0"][0] + [out["0"][1], out["0"][3]]):
        assert np.array_equal(a, b)
    assert rel_err(out["1"][2], out["0"][2]) < 1e-13                     # the reduce kernel sums orbit members with atomics
    assert rel_err(out["1"][0][0], out["1"][0][1]) < 1e-12
