"""The oracle (oracle/ed_oracle.py) pinned against the reference's own golden vectors
(tests/golden/reference_tests.json, transcribed from /root/reference/test/*.jl).  CPU only."""
import cmath
import math

import numpy as np
import pytest

import ed_oracle as O
from helpers import oracle_heisenberg_xyz, oracle_spin_chain, chain_translation_irrep, oracle_from_terms


def spin_space(n):
    return O.HilbertSpace([O.Site([O.State("Up", 1), O.State("Dn", -1)])] * n)


def test_sector_basis_spin_half(golden):
    g = golden["sector_basis_spin_half_4_sz0"]
    hs = spin_space(g["n_sites"])
    assert O.hs_get_basis_list(O.HilbertSpaceSector(hs, g["qn"])) == g["basis"]
    hsr1 = O.represent(hs, g["basis"])
    hsr2 = O.represent(O.HilbertSpaceSector(hs, 0))
    assert list(hsr1.basis_list) == list(hsr2.basis_list)


def test_width_errors(golden):
    g = golden["width_errors"]
    hs = spin_space(g["n_sites"])
    with pytest.raises(ValueError):
        O.represent(hs, br_bits=g["br_bits"])
    with pytest.raises(ValueError):
        O.represent(O.HilbertSpaceSector(hs, 0), br_bits=g["br_bits"])
    with pytest.raises(ValueError):
        O.HilbertSpaceRepresentation(hs, [0b0101], br_bits=8)


def test_full_space(golden):
    g = golden["full_space_4"]
    hsr = O.represent(spin_space(g["n_sites"]), br_bits=32)
    assert hsr.dimension == g["dimension"] and hsr.hilbert_space.bitwidth == g["bitwidth"]


def tj_space(g):
    sites = [O.Site([O.State(str(i), tuple(q)) for i, q in enumerate(states)]) for states in g["site_states"]]
    return O.HilbertSpace(sites)


def test_tj(golden):
    g = golden["tj"]
    hs = tj_space(g)
    assert hs.quantum_number_sectors() == [tuple(q) for q in g["sectors"]]
    assert list(O.represent(hs).basis_list) == g["all"]
    assert list(O.represent(O.HilbertSpaceSector(hs, [tuple(q) for q in g["sectors"]])).basis_list) == g["all"]
    assert list(O.represent(O.HilbertSpaceSector(hs, [])).basis_list) == []
    for key, basis in g["per_sector"].items():
        qn = tuple(int(x) for x in key.split(","))
        hsr = O.represent(O.HilbertSpaceSector(hs, qn))
        assert list(hsr.basis_list) == basis
        assert all(hsr.lookup(b) == i + 1 for i, b in enumerate(basis))
    m = g["multi"]
    assert list(O.represent(O.HilbertSpaceSector(hs, [tuple(q) for q in m["sectors"]])).basis_list) == m["basis"]
    u = g["unsorted_input"]
    assert list(O.represent(hs, u["input"]).basis_list) == u["basis"]


def test_frozen_sorted_array_validation():
    # test/test_frozensortedarray.jl:5-12 : unsorted / duplicate keys are ArgumentErrors
    hs = spin_space(4)
    with pytest.raises(ValueError):
        O.HilbertSpaceRepresentation(hs, [3, 1])
    with pytest.raises(ValueError):
        O.HilbertSpaceRepresentation(hs, [1, 1, 2])


def test_term_walk_order(golden):
    g = golden["pure_iterators"]
    pop = O.PureOperator(*g["term"])
    for b, exp in g["row"].items():
        assert O.get_row_iterator(pop, int(b)) == [tuple(e) for e in exp]
    for b, exp in g["col"].items():
        assert O.get_column_iterator(pop, int(b)) == [tuple(e) for e in exp]
    for br, bc, v in g["element"]:
        assert O.get_element(pop, br, bc) == v
    g = golden["sum_iterators"]
    sop = oracle_from_terms(g["terms"])
    for b, exp in g["row"].items():
        assert O.get_row_iterator(sop, int(b)) == [tuple(e) for e in exp]
    for b, exp in g["col"].items():
        assert O.get_column_iterator(sop, int(b)) == [tuple(e) for e in exp]
    for br, bc, v in g["element"]:
        assert abs(O.get_element(sop, br, bc) - v) < 1e-6


def test_oprep_misses(golden):
    g = golden["oprep_misses"]
    hs = spin_space(g["n_sites"])
    hs2 = spin_space(2)
    op = O.pure_operator(hs2, 1, 0, 1) * 2.0 + O.pure_operator(hs2, 1, 1, 0) * 3.0
    assert [(t.bitmask, t.bitrow, t.bitcol, t.amplitude) for t in op.terms] == [tuple(t) for t in g["terms"]]
    opr = O.OperatorRepresentation(O.represent(hs, g["basis"]), op)
    for i, exp in g["row_iter"].items():
        assert opr.get_row_iterator(int(i)) == [tuple(e) for e in exp]
    for i, exp in g["col_iter"].items():
        assert opr.get_column_iterator(int(i)) == [tuple(e) for e in exp]


def test_sparse_dense_sigma_plus(golden):
    hs = spin_space(4)
    hsr = O.represent(hs)
    opr = O.OperatorRepresentation(hsr, O.pauli_matrix(hs, 1, "+"))
    sp, s0 = np.array([[0, 1], [0, 0]]), np.eye(2)
    H0 = np.kron(np.kron(np.kron(s0, s0), sp), s0)
    assert np.array_equal(O.dense_matrix(opr), H0)
    colptr, rowval, nzval = O.sparse_serial(opr)
    import scipy.sparse as sps
    assert np.array_equal(sps.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(16, 16)).toarray(), H0)
    for i in range(1, 17):
        for j in range(1, 17):
            assert opr.get_element(i, j) == H0[i - 1, j - 1]
    with pytest.raises(IndexError):
        opr.get_row_iterator(0)
    with pytest.raises(IndexError):
        opr.get_element(17, 1)


def test_apply_semantics():
    # test/test_operator_representation.jl:198-270: apply! accumulates (both sides), mul! overwrites, DimensionMismatch
    hs = spin_space(4)
    hsr = O.represent(hs)
    opr = O.OperatorRepresentation(hsr, O.pauli_matrix(hs, 1, "+"))
    H0 = O.dense_matrix(opr)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(16) + 1j * rng.standard_normal(16)
    out = np.zeros(16, dtype=complex)
    O.apply_serial(out, opr, x, "left")
    assert np.allclose(out, H0 @ x)
    O.apply_serial(out, opr, x, "left")
    assert np.allclose(out, 2 * (H0 @ x))
    out[:] = 0
    O.apply_serial(out, opr, x, "right")
    assert np.allclose(out, x @ H0)
    out2 = np.arange(16) * 10.0 + 0j
    O.mul(out2, opr, x)
    assert np.allclose(out2, H0 @ x)
    with pytest.raises(ValueError):
        O.apply_serial(out, opr, np.zeros(17, dtype=complex))
    with pytest.raises(ValueError):
        O.apply_serial(np.zeros(17, dtype=complex), opr, x)
    out3 = np.zeros(16, dtype=complex)
    O.apply_vectorized(out3, opr, x, "left")
    assert np.array_equal(out3, H0 @ x)


def test_symmetry_apply(golden):
    g = golden["symmetry_apply"]
    hs = spin_space(g["n_sites"])
    trans = O.SitePermutation([j - 1 for j in g["transop_1based"]])
    inv = O.SitePermutation([j - 1 for j in g["invop_1based"]])
    for w, im in g["words"]:
        assert O.symmetry_apply(hs, trans, w) == (im, 1)
    pop1 = O.PureOperator(*g["pop1"])
    assert O.symmetry_apply_operator(hs, trans, pop1) == O.PureOperator(*g["pop1_trans"])
    assert O.symmetry_apply_operator(hs, inv, pop1) == O.PureOperator(*g["pop1_inv"])
    _, j1 = oracle_heisenberg_xyz(4)
    assert O.isinvariant(hs, trans, j1) and O.isinvariant(hs, inv, j1)
    assert not O.isinvariant(hs, trans, pop1)


def test_bitflip(golden):
    g = golden["bitflip"]
    hs = spin_space(g["n_sites"])
    p = O.SitePermutation([j - 1 for j in g["perm_1based"]])
    b0, b1 = O.GlobalBitFlip(False), O.GlobalBitFlip(True)
    assert O.symmetry_apply(hs, b0, g["word"])[0] == g["b0"]
    assert O.symmetry_apply(hs, b1, g["word"])[0] == g["b1"]
    assert O.symmetry_apply(hs, p * b0, g["word"])[0] == g["pb0"]
    assert O.symmetry_apply(hs, p * b1, g["word"])[0] == g["pb1"]


def test_symmetry_reduce_chain4(golden):
    hs = spin_space(4)
    hsr = O.represent(O.HilbertSpaceSector(hs, 0))
    for irrep, basis in golden["reduce_translation_chain4"]["irrep_1based"].items():
        rhsr = O.symmetry_reduce(hsr, chain_translation_irrep(4, int(irrep) - 1))
        assert list(rhsr.basis_list) == basis
    inv = O.SitePermutation([0, 3, 2, 1])
    ident = O.SitePermutation([0, 1, 2, 3])
    for irrep, basis in golden["reduce_inversion_chain4"]["irrep_1based"].items():
        chi = 1.0 if irrep == "1" else -1.0
        rhsr = O.symmetry_reduce(hsr, [(ident, 1.0 + 0j), (inv, chi + 0j)])
        assert list(rhsr.basis_list) == basis
    g = golden["reduce_symmorphic_chain4"]
    t1 = chain_translation_irrep(4, 0)
    for key, chi in (("t1_p1", 1.0), ("t1_p2", -1.0)):
        ops = [(p * t, pp * pt) for (t, pt) in t1 for (p, pp) in [(ident, 1.0 + 0j), (inv, chi + 0j)]]
        assert list(O.symmetry_reduce(hsr, ops).basis_list) == g[key]
    # dimension sums agree across T, P, T x| P decompositions (test_symmetry_reduce.jl:161)
    dims_t = sum(O.symmetry_reduce(hsr, chain_translation_irrep(4, k)).dimension for k in range(4))
    dims_p = sum(O.symmetry_reduce(hsr, [(ident, 1.0 + 0j), (inv, c + 0j)]).dimension for c in (1.0, -1.0))
    assert dims_t == dims_p == hsr.dimension


def test_phase_convention_chain7(golden):
    g = golden["convention_chain7"]
    hs = spin_space(7)
    for key, qn in (("qn_plus5", 5), ("qn_minus5", -5)):
        gg = g[key]
        hsr = O.represent(O.HilbertSpaceSector(hs, qn))
        assert list(hsr.basis_list) == gg["basis"]
        rhsr = O.symmetry_reduce(hsr, chain_translation_irrep(7, gg["irrep_1based"] - 1))
        assert list(rhsr.basis_list) == gg["rbasis"]
        psi = O.symmetry_unreduce_vector(rhsr, np.array([1.0]))
        expect = np.array([cmath.exp(gg["phase_sign"] * 2j * math.pi * i / 7) / math.sqrt(7) for i in range(7)])
        assert np.allclose(psi, expect, atol=O.RTOL_DEFAULT)
        sv = np.array([0.3 - 0.4j])
        assert np.allclose(O.symmetry_reduce_vector(rhsr, O.symmetry_unreduce_vector(rhsr, sv)), sv)


def test_reduced_operator_chain4(golden):
    g = golden["reduced_operator_chain4"]
    hs, j1 = oracle_heisenberg_xyz(4)
    hsr = O.represent(O.HilbertSpaceSector(hs, 0))
    rhsr = O.symmetry_reduce(hsr, chain_translation_irrep(4, g["irrep_1based"] - 1))
    assert list(rhsr.basis_list) == g["rbasis"]
    j1_mat = O.dense_matrix(O.OperatorRepresentation(hsr, j1))
    psis = [np.array(p) / np.linalg.norm(p) for p in g["psis"]]
    H = np.array([[psis[i] @ (j1_mat @ psis[j]) for j in range(2)] for i in range(2)])
    ropr = O.ReducedOperatorRepresentation(rhsr, j1)
    assert np.allclose(O.dense_matrix(ropr), H, atol=O.RTOL_DEFAULT)
    assert np.allclose(H, [[0, 4 * math.sqrt(2)], [4 * math.sqrt(2), -4]])
    for i in (1, 2):
        row = np.zeros(2, dtype=complex)
        for j, a in ropr.get_row_iterator(i):
            if j > 0:
                row[j - 1] += a
        assert np.allclose(row, H[i - 1, :])
        col = np.zeros(2, dtype=complex)
        for j, a in ropr.get_column_iterator(i):
            if j > 0:
                col[j - 1] += a
        assert np.allclose(col, H[:, i - 1])
        for j in (1, 2):
            assert abs(ropr.get_element(i, j) - H[i - 1, j - 1]) < 1e-12


def test_spectrum_union_chain(golden):
    # test/test_reduced_representation.jl:236-252 (n=4) and the complex-phase case n=7 (:256-301)
    for n, qn in ((4, 0), (7, 1)):
        hs, j1 = oracle_heisenberg_xyz(n)
        hsr = O.represent(O.HilbertSpaceSector(hs, qn))
        full = np.linalg.eigvalsh(O.dense_matrix(O.OperatorRepresentation(hsr, j1)))
        parts = []
        for k in range(n):
            rhsr = O.symmetry_reduce(hsr, chain_translation_irrep(n, k))
            if rhsr.dimension:
                m = O.dense_matrix(O.ReducedOperatorRepresentation(rhsr, j1))
                assert np.allclose(m, m.conj().T, atol=1e-10)
                parts.extend(np.linalg.eigvalsh(m))
        assert np.allclose(sorted(parts), full, atol=1e-8)


def test_heisenberg_terms(golden):
    # test/test_operator_simplify.jl:68-84
    _, h_xyz = oracle_heisenberg_xyz(4)
    _, h_pm = oracle_spin_chain(4)
    assert len(h_xyz.terms) == 6 * 4
    assert {t.key() for t in h_xyz.terms} == {t.key() for t in h_pm.terms}
    assert h_xyz == h_pm


def test_known_answers_l16(golden):
    k = golden["known_answers"]
    hs, h = oracle_spin_chain(16)
    assert len(h.terms) == k["L16_terms"]
    hsr = O.represent(O.HilbertSpaceSector(hs, 0))
    assert hsr.dimension == k["L16_dim"]
    # raw row-iterator hits and E0 through the vectorised apply (same arithmetic as apply_serial)
    opr = O.OperatorRepresentation(hsr, h)
    hits = sum(int(((hsr.basis_list & np.uint64(t.bitmask)) == np.uint64(t.bitrow)).sum()) for t in h.terms)
    assert hits == k["L16_raw_hits"]
    import scipy.sparse.linalg as spl
    lin = spl.LinearOperator((hsr.dimension,) * 2, dtype=np.float64,
                             matvec=lambda v: O.apply_vectorized(np.zeros(hsr.dimension), opr, np.asarray(v, dtype=np.float64).ravel()))
    e0 = spl.eigsh(lin, k=1, which="SA", tol=1e-12)[0][0]
    assert abs(e0 - k["L16_E0"]) < 1e-9
