"""GPU parity tests: libedcuda (through the C ABI via the edcuda host package) against the oracle
and the reference's golden vectors.  Bit-exact for basis words, indices and CSC structure; amplitudes and
matvec outputs within 1e-12 relative (north_star), Lanczos eigenvalues within 1e-10."""
import cmath
import math

import numpy as np
import pytest

import ed_oracle as O
from helpers import (oracle_heisenberg_xyz, oracle_spin_chain, oracle_from_terms, to_oracle_symops, rel_err)

pytestmark = pytest.mark.gpu

TOL = 1e-12


def spin_space_o(n):
    return O.HilbertSpace([O.Site([O.State("Up", 1), O.State("Dn", -1)])] * n)


def terms_of(op_o):
    return [(t.bitmask, t.bitrow, t.bitcol, t.amplitude) for t in op_o.terms]


# ------------------------------------------------------------------ K1: bases
def test_sector_basis_goldens(gpu_ed, golden):
    ed = gpu_ed
    g = golden["sector_basis_spin_half_4_sz0"]
    hs, _ = ed.spin_half_system(4)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    assert hsr.kind == ed.ED_BASIS_COMBINADIC
    assert list(hsr.basis_list) == g["basis"]
    assert hsr == ed.represent(hs, np.array(g["basis"], dtype=np.uint64))
    full = ed.represent(hs, np.uint32)
    assert full.dimension == 16 and full.bitwidth == 4 and list(full.basis_list) == list(range(16))
    assert list(hsr.lookup(g["basis"])) == [1, 2, 3, 4, 5, 6]
    assert list(hsr.lookup([0, 1, 7, 15, 16, 2 ** 40])) == [-1] * 6
    assert hsr.basis_lookup[0b0101] == 2 and hsr.basis_lookup.get(0b0111, -1) == -1
    with pytest.raises(KeyError):
        hsr.basis_lookup[0b0111]


def test_tj_goldens(gpu_ed, golden):
    ed = gpu_ed
    g = golden["tj"]
    sites = [ed.Site([ed.State(str(i), tuple(q)) for i, q in enumerate(states)]) for states in g["site_states"]]
    hs = ed.HilbertSpace(sites)
    allr = ed.represent(hs)
    assert allr.kind == ed.ED_BASIS_DPRANK
    assert list(allr.basis_list) == g["all"]
    assert list(allr.lookup(g["all"])) == list(range(1, 13))
    assert list(allr.lookup([3, 7, 11, 15])) == [-1] * 4   # invalid code of the 3-state site
    assert list(ed.represent(ed.HilbertSpaceSector(hs)).basis_list) == g["all"]
    assert list(ed.represent(ed.HilbertSpaceSector(hs, [])).basis_list) == []
    for key, basis in g["per_sector"].items():
        qn = tuple(int(x) for x in key.split(","))
        hsr = ed.represent(ed.HilbertSpaceSector(hs, qn))
        assert list(hsr.basis_list) == basis
        assert list(hsr.lookup(basis)) == list(range(1, len(basis) + 1))
        assert hsr == ed.represent(hs, np.array(basis, dtype=np.uint64))
    m = g["multi"]
    assert list(ed.represent(ed.HilbertSpaceSector(hs, [tuple(q) for q in m["sectors"]])).basis_list) == m["basis"]
    u = g["unsorted_input"]
    assert list(ed.represent(hs, np.array(u["input"], dtype=np.uint64)).basis_list) == u["basis"]
    with pytest.raises(ValueError):
        ed.represent(hs, np.array([1, 2, 2], dtype=np.uint64))     # duplicates: ArgumentError


@pytest.mark.parametrize("n", [1, 2, 5, 10, 13])
def test_spin_half_all_sectors_vs_oracle(gpu_ed, n):
    ed = gpu_ed
    hs, _ = ed.spin_half_system(n)
    hs_o = spin_space_o(n)
    total = 0
    for qn in range(-n, n + 1, 2):
        got = ed.represent(ed.HilbertSpaceSector(hs, qn)).basis_list
        exp = np.array(O.hs_get_basis_list(O.HilbertSpaceSector(hs_o, qn)), dtype=np.uint64)
        assert np.array_equal(got, exp)
        total += len(got)
    assert total == 2 ** n
    multi = ed.represent(ed.HilbertSpaceSector(hs, [n, n - 2, -n]))
    exp = np.array(O.hs_get_basis_list(O.HilbertSpaceSector(hs_o, [n, n - 2, -n])), dtype=np.uint64)
    assert np.array_equal(multi.basis_list, exp)
    assert np.array_equal(multi.lookup(exp), np.arange(1, len(exp) + 1))


def test_mixed_sites_two_qn_vs_oracle(gpu_ed):
    ed = gpu_ed
    # spin-1 sites (3 states) + spin-1/2 sites with (charge, 2Sz) quantum numbers
    s1 = [("p", (0, 2)), ("0", (0, 0)), ("m", (0, -2))]
    sh = [("e", (0, 0)), ("u", (1, 1)), ("d", (1, -1)), ("ud", (2, 0))]
    layout = [s1, sh, s1, sh, sh]
    hs = ed.HilbertSpace([ed.Site([ed.State(a, q) for a, q in st]) for st in layout])
    hs_o = O.HilbertSpace([O.Site([O.State(a, q) for a, q in st]) for st in layout])
    assert hs.quantum_number_sectors() == hs_o.quantum_number_sectors()
    for qn in hs_o.quantum_number_sectors()[::3]:
        got = ed.represent(ed.HilbertSpaceSector(hs, qn))
        exp = np.array(O.hs_get_basis_list(O.HilbertSpaceSector(hs_o, qn)), dtype=np.uint64)
        assert np.array_equal(got.basis_list, exp)
        assert np.array_equal(got.lookup(exp), np.arange(1, len(exp) + 1))
    allg = ed.represent(hs)
    exp = np.array(O.hs_get_basis_list(hs_o), dtype=np.uint64)
    assert np.array_equal(allg.basis_list, exp)


def test_known_dimensions(gpu_ed, golden):
    ed = gpu_ed
    k = golden["known_answers"]
    for n, key in ((16, "L16_dim"), (28, "L28_dim"), (32, "L32_dim")):
        hs, _ = ed.spin_half_system(n)
        hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
        assert hsr.dimension == k[key]
    hs, _ = ed.spin_half_system(28)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    # size-independent properties at full size: ascending, popcount, rank(unrank(i)) = i on a window
    lo = 30_000_000
    w = hsr.download(lo, 200_000)
    assert np.all(w[1:] > w[:-1])
    assert all(bin(int(x)).count("1") == 14 for x in w[::997])
    assert np.array_equal(hsr.lookup(w), np.arange(lo + 1, lo + len(w) + 1))
    assert int(hsr.download(hsr.dimension - 1, 1)[0]) == ((1 << 14) - 1) << 14


# ------------------------------------------------------------------ K2: apply / mul
def _check_apply(ed, hsr, hsr_o, op, op_o, cplx_vec, seed=0, generic=None):
    opr = ed.represent(hsr, op)
    if generic is not None:
        opr.set_kernel(1 if generic else 0)
    opr_o = O.OperatorRepresentation(hsr_o, op_o)
    n = hsr.dimension
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n)
    if cplx_vec:
        x = x + 1j * rng.standard_normal(n)
    dt = x.dtype
    for side, so in ((0, "left"), (1, "right")):
        exp = O.apply_vectorized(np.zeros(n, dtype=dt), opr_o, x, so)
        out = np.zeros(n, dtype=dt)
        if side == 0:
            ed.apply_b(out, opr, x)
        else:
            ed.apply_b(out, x, opr)
        assert rel_err(out, exp) < TOL
        # apply! adds
        if side == 0:
            ed.apply_b(out, opr, x)
        else:
            ed.apply_b(out, x, opr)
        assert rel_err(out, 2 * exp) < TOL
    exp = O.apply_vectorized(np.zeros(n, dtype=dt), opr_o, x, "left")
    out = (np.arange(n) * 10.0).astype(dt)
    ed.mul_b(out, opr, x)          # mul! overwrites
    assert rel_err(out, exp) < TOL
    assert rel_err(opr * x, exp) < TOL
    return opr, opr_o


@pytest.mark.parametrize("n,generic", [(8, True), (12, True), (12, False), (16, False), (16, True)])
def test_apply_heisenberg_chain(gpu_ed, n, generic):
    ed = gpu_ed
    hs, h = ed.models.heisenberg_chain(n)
    hs_o, h_o = oracle_spin_chain(n)
    assert h.terms == terms_of(h_o)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    hsr_o = O.represent(O.HilbertSpaceSector(hs_o, 0))
    _check_apply(ed, hsr, hsr_o, h, h_o, False, generic=generic)
    _check_apply(ed, hsr, hsr_o, h, h_o, True, generic=generic)


@pytest.mark.parametrize("qn", [2, -6])
def test_apply_other_sectors_xxz_j1j2(gpu_ed, qn):
    ed = gpu_ed
    n = 14
    hs, _ = ed.spin_half_system(n)
    bonds = ed.lattices.chain_bonds(n, 1) + ed.lattices.chain_bonds(n, 2)
    h = ed.simplify(ed.models.xxz_bonds(hs, ed.lattices.chain_bonds(n, 1), 1.0, 0.5)
                    + ed.models.heisenberg_bonds(hs, ed.lattices.chain_bonds(n, 2), 0.5))
    hs_o, a = oracle_spin_chain(n, [(i, (i + 1) % n) for i in range(n)], jz=0.5)
    _, b = oracle_spin_chain(n, [(i, (i + 2) % n) for i in range(n)], jz=0.5, jxy=0.5)
    h_o = O.simplify(a + b)
    assert h.terms == terms_of(h_o)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, qn))
    hsr_o = O.represent(O.HilbertSpaceSector(hs_o, qn))
    _check_apply(ed, hsr, hsr_o, h, h_o, False)


def test_apply_complex_operator_full_space_and_list_basis(gpu_ed, golden):
    ed = gpu_ed
    n = 6
    hs, pauli = ed.spin_half_system(n)
    hs_o, pauli_o = O.spin_half_system(n)
    # non-Hermitian complex operator that does not conserve Sz: exercises misses and ED_BASIS_FULL
    op = pauli(0, "y") * pauli(3, "x") + (0.3 - 0.7j) * pauli(2, "+") + pauli(1, "z") * pauli(5, "-") * 1.5
    op_o = pauli_o(0, "y") * pauli_o(3, "x") + (0.3 - 0.7j) * pauli_o(2, "+") + (pauli_o(1, "z") * pauli_o(5, "-")) * 1.5
    assert op.terms == terms_of(op_o)
    full, full_o = ed.represent(hs), O.represent(hs_o)
    assert full.kind == ed.ED_BASIS_FULL
    _check_apply(ed, full, full_o, op, op_o, True)
    with pytest.raises(ValueError):   # complex operator with real vectors
        ed.apply_b(np.zeros(64), ed.represent(full, op), np.zeros(64))
    # arbitrary user basis (binary-search lookup), many misses
    rng = np.random.default_rng(3)
    words = np.sort(rng.choice(64, size=37, replace=False)).astype(np.uint64)
    lst, lst_o = ed.represent(hs, words), O.represent(hs_o, words)
    assert lst.kind == ed.ED_BASIS_LIST
    _check_apply(ed, lst, lst_o, op, op_o, True)
    sec, sec_o = ed.represent(ed.HilbertSpaceSector(hs, 0)), O.represent(O.HilbertSpaceSector(hs_o, 0))
    _check_apply(ed, sec, sec_o, op, op_o, True)
    # golden: misses are -1 in the iterators
    g = golden["oprep_misses"]
    hs4, _ = ed.spin_half_system(4)
    opr = ed.represent(ed.represent(hs4, np.array(g["basis"], dtype=np.uint64)), ed.Operator([tuple(t) for t in g["terms"]]))
    for i, exp in g["row_iter"].items():
        assert opr.get_row_iterator(int(i)) == [tuple(e) for e in exp]
    for i, exp in g["col_iter"].items():
        assert opr.get_column_iterator(int(i)) == [tuple(e) for e in exp]


def test_apply_generic_sector_dprank(gpu_ed):
    ed = gpu_ed
    # t-J like chain: 3-state sites, hopping + Sz Sz, two quantum numbers
    states = [("e", (0, 0)), ("u", (1, 1)), ("d", (1, -1))]
    n = 6
    hs = ed.HilbertSpace([ed.Site([ed.State(a, q) for a, q in states])] * n)
    hs_o = O.HilbertSpace([O.Site([O.State(a, q) for a, q in states])] * n)
    op, op_o = ed.Operator(), None
    for i in range(n):
        j = (i + 1) % n
        for s in (1, 2):
            t = ed.pure_operator(hs, i, s, 0, -1.0) * ed.pure_operator(hs, j, 0, s, 1.0)
            t_o = O.pure_operator(hs_o, i, s, 0, -1.0) * O.pure_operator(hs_o, j, 0, s, 1.0)
            op = op + t + t.adjoint()
            tt = t_o + O.PureOperator(t_o.bitmask, t_o.bitcol, t_o.bitrow, t_o.amplitude)
            op_o = tt if op_o is None else op_o + tt
        z = 0.25 * ed.pure_operator(hs, i, 1, 1, 1.0) * ed.pure_operator(hs, j, 2, 2, 1.0)
        z_o = 0.25 * (O.pure_operator(hs_o, i, 1, 1, 1.0) * O.pure_operator(hs_o, j, 2, 2, 1.0))
        op, op_o = op + z, op_o + z_o
    op, op_o = ed.simplify(op), O.simplify(op_o)
    assert op.terms == terms_of(op_o)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, (4, 0)))
    hsr_o = O.represent(O.HilbertSpaceSector(hs_o, (4, 0)))
    assert hsr.kind == ed.ED_BASIS_DPRANK and np.array_equal(hsr.basis_list, hsr_o.basis_list)
    _check_apply(ed, hsr, hsr_o, op, op_o, False)


def test_apply_errors_and_sharding(gpu_ed):
    ed = gpu_ed
    n = 12
    hs, h = ed.models.heisenberg_chain(n)
    hs_o, h_o = oracle_spin_chain(n)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    opr = ed.represent(hsr, h)
    d = hsr.dimension
    x = np.random.default_rng(5).standard_normal(d)
    for bad_out, bad_x in ((d + 1, d), (d, d + 1)):
        with pytest.raises(ed.DimensionMismatch):
            ed.apply_b(np.zeros(bad_out), opr, np.zeros(bad_x))
        with pytest.raises(ed.DimensionMismatch):
            ed.apply_b(np.zeros(bad_out), np.zeros(bad_x), opr)
    exp = O.apply_vectorized(np.zeros(d), O.OperatorRepresentation(O.represent(O.HilbertSpaceSector(hs_o, 0)), h_o), x)
    # row shards (the multi-GPU partition) reproduce the unsharded result piecewise
    for generic in (1, 0):
        pieces = []
        for lo, hi in ((0, 100), (100, 611), (611, d)):
            shard = ed.represent(hsr, h).set_rows(lo, hi).set_kernel(generic)
            out = np.zeros(hi - lo)
            ed.mul_b(out, shard, x)
            pieces.append(out)
        assert rel_err(np.concatenate(pieces), exp) < TOL
    # device-resident vectors (torch) go through the same entry point without staging
    import torch
    xt = torch.from_numpy(x).cuda()
    yt = torch.zeros(d, dtype=torch.float64, device="cuda")
    ed.mul_b(yt, opr, xt)
    torch.cuda.synchronize()
    assert rel_err(yt.cpu().numpy(), exp) < TOL


def test_get_element_row_column_dense(gpu_ed):
    ed = gpu_ed
    hs, pauli = ed.spin_half_system(4)
    hsr = ed.represent(hs)
    opr = ed.represent(hsr, pauli(1, "+"))
    sp, s0 = np.array([[0, 1], [0, 0]]), np.eye(2)
    H0 = np.kron(np.kron(np.kron(s0, s0), sp), s0)        # test_operator_representation.jl:137-147
    assert np.array_equal(opr.matrix(), H0)
    assert np.array_equal(opr.sparse().toarray(), H0)
    for i in range(1, 17):
        row, col = opr.get_row(i), opr.get_column(i)
        assert row == {j + 1: H0[i - 1, j] for j in range(16) if H0[i - 1, j] != 0}
        assert col == {j + 1: H0[j, i - 1] for j in range(16) if H0[j, i - 1] != 0}
        for j in range(1, 17):
            assert opr.get_element(i, j) == H0[i - 1, j - 1]
            assert opr[i, j] == H0[i - 1, j - 1]
    for bad in ((0, 1), (17, 1), (1, 0), (1, 17)):
        with pytest.raises(IndexError):
            opr.get_element(*bad)
    for bad in (0, 17):
        with pytest.raises(IndexError):
            opr.get_row_iterator(bad)
        with pytest.raises(IndexError):
            opr.get_column_iterator(bad)


# ------------------------------------------------------------------ K3/K4: sparse
@pytest.mark.parametrize("n,qn", [(8, 0), (10, 2), (12, 0)])
def test_sparse_csc_bit_exact_vs_oracle(gpu_ed, n, qn):
    ed = gpu_ed
    hs, h = ed.models.heisenberg_chain(n)
    hs_o, h_o = oracle_spin_chain(n)
    opr = ed.represent(ed.represent(ed.HilbertSpaceSector(hs, qn)), h)
    opr_o = O.OperatorRepresentation(O.represent(O.HilbertSpaceSector(hs_o, qn)), h_o)
    colptr, rowval, nzval = opr.sparse_csc()
    colptr_o, rowval_o, nzval_o = O.sparse_serial(opr_o)
    assert colptr.dtype == np.int64 and rowval.dtype == np.int64
    assert np.array_equal(colptr, colptr_o)
    assert np.array_equal(rowval, rowval_o)
    assert rel_err(nzval, nzval_o) < TOL
    # tol = 0 keeps exact-zero diagonals out only if |v| < 0 (never): structure grows
    colptr0, _, _ = opr.sparse_csc(tol=0.0)
    colptr0_o, _, _ = O.sparse_serial(opr_o, tol=0.0)
    assert np.array_equal(colptr0, colptr0_o) and colptr0[-1] >= colptr[-1]


def test_sparse_complex_and_list(gpu_ed):
    ed = gpu_ed
    n = 5
    hs, pauli = ed.spin_half_system(n)
    hs_o, pauli_o = O.spin_half_system(n)
    op = pauli(0, "y") * pauli(3, "x") + (0.3 - 0.7j) * pauli(2, "+") + 1e-9 * pauli(4, "x") + pauli(1, "z")
    op_o = pauli_o(0, "y") * pauli_o(3, "x") + (0.3 - 0.7j) * pauli_o(2, "+") + 1e-9 * pauli_o(4, "x") + pauli_o(1, "z")
    words = np.array(sorted(set(range(32)) - {3, 9, 27}), dtype=np.uint64)
    opr = ed.represent(ed.represent(hs, words), op)
    opr_o = O.OperatorRepresentation(O.represent(hs_o, words), op_o)
    for tol in (None, 1e-12):
        got = opr.sparse_csc(tol)
        exp = O.sparse_serial(opr_o) if tol is None else O.sparse_serial(opr_o, tol)
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
        assert rel_err(got[2], exp[2]) < TOL
    assert rel_err(opr.matrix(), O.dense_matrix(opr_o)) < TOL


def test_sparse_known_answers(gpu_ed, golden):
    ed = gpu_ed
    k = golden["known_answers"]
    hs, h = ed.models.heisenberg_chain(16)
    assert len(h.terms) == k["L16_terms"]
    opr = ed.represent(ed.represent(ed.HilbertSpaceSector(hs, 0)), h)
    colptr, rowval, nzval = opr.sparse_csc()
    assert colptr[-1] - 1 == k["L16_nnz_sparse"] == len(rowval)
    # rows ascending inside each column, symmetric matrix, matvec through CSC equals the matrix-free apply
    m = opr.sparse()
    assert all(np.all(np.diff(rowval[colptr[j] - 1: colptr[j + 1] - 1]) > 0) for j in range(0, 12870, 131))
    assert abs(m - m.T).max() < 1e-14
    x = np.random.default_rng(0).standard_normal(12870)
    assert rel_err(m @ x, opr * x) < TOL
    hs2, h2 = ed.models.heisenberg_square(4, 4)
    assert len(h2.terms) == k["sq4x4_terms"]
    opr2 = ed.represent(ed.represent(ed.HilbertSpaceSector(hs2, 0)), h2)
    assert opr2.sparse_csc()[0][-1] - 1 == k["sq4x4_nnz_sparse"]


# ------------------------------------------------------------------ symmetry_apply
def test_symmetry_apply_goldens(gpu_ed, golden):
    ed = gpu_ed
    g = golden["symmetry_apply"]
    hs, _ = ed.spin_half_system(4)
    trans = ed.SitePermutation([j - 1 for j in g["transop_1based"]])
    inv = ed.SitePermutation([j - 1 for j in g["invop_1based"]])
    for w, im in g["words"]:
        assert ed.symmetry_apply(hs, trans, w) == (im, 1)
        assert ed.symmetry_apply(ed.HilbertSpaceSector(hs, 0), trans, w) == (im, 1)
    pop1 = ed.Operator([tuple(g["pop1"])])
    assert ed.symmetry_apply_operator(hs, trans, pop1).terms == [tuple(g["pop1_trans"])]
    assert ed.symmetry_apply_operator(hs, inv, pop1).terms == [tuple(g["pop1_inv"])]
    _, j1 = ed.models.heisenberg_chain(4)
    assert ed.isinvariant(hs, trans, j1) and ed.isinvariant(hs, inv, j1) and not ed.isinvariant(hs, trans, pop1)
    b = golden["bitflip"]
    p = ed.SitePermutation([j - 1 for j in b["perm_1based"]])
    b0, b1 = ed.GlobalBitFlip(False), ed.GlobalBitFlip(True)
    assert ed.symmetry_apply(hs, b0, b["word"])[0] == b["b0"]
    assert ed.symmetry_apply(hs, b1, b["word"])[0] == b["b1"]
    assert ed.symmetry_apply(hs, p * b0, b["word"])[0] == b["pb0"]
    assert ed.symmetry_apply(hs, p * b1, b["word"])[0] == b["pb1"]
    assert ed.symmetry_apply(hs, b1 * p, b["word"])[0] == b["pb1"]
    # multi-bit sites: whole fields move
    states = [("a", 0), ("b", 1), ("c", 2)]
    hs3 = ed.HilbertSpace([ed.Site([ed.State(s, q) for s, q in states])] * 3)
    hs3_o = O.HilbertSpace([O.Site([O.State(s, q) for s, q in states])] * 3)
    perm = [2, 0, 1]
    words = np.array([hs3.compress(c) for c in ((0, 1, 2), (2, 2, 0), (1, 0, 0))], dtype=np.uint64)
    got, _ = ed.symmetry_apply(hs3, ed.SitePermutation(perm), words)
    assert list(got) == [O.symmetry_apply(hs3_o, O.SitePermutation(perm), int(w))[0] for w in words]


# ------------------------------------------------------------------ K5: symmetry_reduce
def _check_rhsr(ed, hsr, hsr_o, symops, tol=None):
    rhsr = ed.symmetry_reduce(hsr, symops) if tol is None else ed.symmetry_reduce(hsr, symops, tol)
    rhsr_o = O.symmetry_reduce(hsr_o, to_oracle_symops(symops)) if tol is None else O.symmetry_reduce(hsr_o, to_oracle_symops(symops), tol)
    assert np.array_equal(rhsr.basis_list, rhsr_o.basis_list)
    assert np.array_equal(rhsr.basis_mapping_index, rhsr_o.basis_mapping_index)
    assert np.max(np.abs(rhsr.basis_mapping_amplitude - rhsr_o.basis_mapping_amplitude), initial=0.0) < 1e-14
    return rhsr, rhsr_o


def test_symmetry_reduce_goldens(gpu_ed, golden):
    ed = gpu_ed
    L = ed.lattices
    hs, _ = ed.spin_half_system(4)
    hs_o = spin_space_o(4)
    hsr, hsr_o = ed.represent(ed.HilbertSpaceSector(hs, 0)), O.represent(O.HilbertSpaceSector(hs_o, 0))
    for irrep, basis in golden["reduce_translation_chain4"]["irrep_1based"].items():
        rhsr, _ = _check_rhsr(ed, hsr, hsr_o, L.chain_translation_irrep(4, int(irrep) - 1))
        assert list(rhsr.basis_list) == basis and rhsr.parent is hsr
        sv = np.random.default_rng(1).standard_normal(rhsr.dimension) + 0.5j
        lv = ed.symmetry_unreduce(rhsr, sv)
        assert np.allclose(ed.symmetry_reduce(rhsr, lv), sv, atol=1e-14)
        out = np.zeros(rhsr.dimension, dtype=complex)
        assert ed.symmetry_reduce_b(out, rhsr, lv) is out and np.allclose(out, sv, atol=1e-14)
        with pytest.raises(ed.DimensionMismatch):
            ed.symmetry_unreduce(rhsr, np.zeros(rhsr.dimension + 1, dtype=complex))
        with pytest.raises(ed.DimensionMismatch):
            ed.symmetry_reduce(rhsr, np.zeros(hsr.dimension + 1, dtype=complex))
    for irrep, basis in golden["reduce_inversion_chain4"]["irrep_1based"].items():
        rhsr, _ = _check_rhsr(ed, hsr, hsr_o, L.chain_inversion_irrep(4, 1 if irrep == "1" else -1))
        assert list(rhsr.basis_list) == basis
    g = golden["reduce_symmorphic_chain4"]
    for key, par in (("t1_p1", 1), ("t1_p2", -1)):
        ops = L.symmorphic_product(L.chain_translation_irrep(4, 0), L.chain_inversion_irrep(4, par))
        rhsr, _ = _check_rhsr(ed, hsr, hsr_o, ops)
        assert list(rhsr.basis_list) == g[key]


def test_phase_convention_chain7(gpu_ed, golden):
    ed = gpu_ed
    g = golden["convention_chain7"]
    hs, _ = ed.spin_half_system(7)
    for key, qn in (("qn_plus5", 5), ("qn_minus5", -5)):
        gg = g[key]
        hsr = ed.represent(ed.HilbertSpaceSector(hs, qn))
        assert list(hsr.basis_list) == gg["basis"]
        rhsr = ed.symmetry_reduce(hsr, ed.lattices.chain_translation_irrep(7, gg["irrep_1based"] - 1))
        assert list(rhsr.basis_list) == gg["rbasis"]
        psi = ed.symmetry_unreduce(rhsr, np.array([1.0]))
        expect = np.array([cmath.exp(gg["phase_sign"] * 2j * math.pi * i / 7) / math.sqrt(7) for i in range(7)])
        assert np.allclose(psi, expect, atol=1e-14)


@pytest.mark.parametrize("n,qn", [(8, 0), (9, 1), (10, 0)])
def test_symmetry_reduce_chain_all_momenta(gpu_ed, n, qn):
    ed = gpu_ed
    L = ed.lattices
    hs, _ = ed.spin_half_system(n)
    hs_o = spin_space_o(n)
    hsr, hsr_o = ed.represent(ed.HilbertSpaceSector(hs, qn)), O.represent(O.HilbertSpaceSector(hs_o, qn))
    total = 0
    for k in range(n):
        rhsr, _ = _check_rhsr(ed, hsr, hsr_o, L.chain_translation_irrep(n, k))
        total += rhsr.dimension
        if rhsr.dimension:
            assert np.array_equal(rhsr.orbit_sizes(), [len({O.symmetry_apply(hs_o, O.SitePermutation([(i + x) % n for i in range(n)]), int(b))[0] for x in range(n)}) for b in rhsr.basis_list])
    assert total == hsr.dimension
    if qn == 0:   # spin flip x translation x inversion (examples/spinhalf_triangular.jl:89-96 style product irreps)
        for flip_par in (1, -1):
            ops = [(f * p, cf * cp) for (p, cp) in L.symmorphic_product(L.chain_translation_irrep(n, 0), L.chain_inversion_irrep(n, 1))
                   for (f, cf) in ((ed.GlobalBitFlip(False), 1.0), (ed.GlobalBitFlip(True), float(flip_par)))]
            _check_rhsr(ed, hsr, hsr_o, ops)
    # a user-supplied (list) parent and the full space work the same way
    lst, lst_o = ed.represent(hs, hsr.basis_list), O.represent(hs_o, hsr_o.basis_list)
    _check_rhsr(ed, lst, lst_o, L.chain_translation_irrep(n, 1))


def test_symmetry_reduce_errors(gpu_ed):
    ed = gpu_ed
    hs, _ = ed.spin_half_system(4)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    T = ed.lattices.chain_translation_irrep(4, 1)
    with pytest.raises(ValueError):     # identity must come first
        ed.symmetry_reduce(hsr, [T[1], T[0], T[2], T[3]])
    with pytest.raises(ed.UnsupportedError):   # not closed under composition
        ed.symmetry_reduce(hsr, [T[0], T[1]])
    with pytest.raises(KeyError):       # images leave the parent basis: KeyError in the reference (:81)
        ed.symmetry_reduce(ed.represent(hs, np.array([0b0011, 0b0101], dtype=np.uint64)), T)
    # tol changes which stabiliser characters count as 1 (symmetry_reduce_generic.jl:62)
    hs_o = spin_space_o(4)
    hsr_o = O.represent(O.HilbertSpaceSector(hs_o, 0))
    sloppy = [(op, chi * cmath.exp(1e-3j)) if i == 2 else (op, chi) for i, (op, chi) in enumerate(ed.lattices.chain_translation_irrep(4, 0))]
    for tol in (1e-8, 1e-2):
        a = ed.symmetry_reduce(hsr, sloppy, tol)
        b = O.symmetry_reduce(hsr_o, to_oracle_symops(sloppy), tol)
        assert np.array_equal(a.basis_list, b.basis_list)


def test_square_4x4_sector_dimensions(gpu_ed, golden):
    ed = gpu_ed
    k = golden["known_answers"]
    hs, h = ed.models.heisenberg_square(4, 4)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    dims = []
    for k2 in range(4):
        for k1 in range(4):
            dims.append(ed.symmetry_reduce(hsr, ed.lattices.torus_translation_irrep(4, 4, k1, k2)).dimension)
    assert dims == k["sq4x4_sector_dims"] and sum(dims) == 12870


# ------------------------------------------------------------------ K6: reduced operator representation
def test_reduced_operator_chain4_golden(gpu_ed, golden):
    ed = gpu_ed
    g = golden["reduced_operator_chain4"]
    hs, pauli = ed.spin_half_system(4)
    j1 = sum(pauli(i, mu) * pauli((i + 1) % 4, mu) for mu in "xyz" for i in range(4))   # unsimplified, complex
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    rhsr = ed.symmetry_reduce(hsr, ed.lattices.chain_translation_irrep(4, g["irrep_1based"] - 1))
    assert list(rhsr.basis_list) == g["rbasis"] and rhsr.dimension == 2
    j1_mat = ed.represent(hsr, j1).matrix()
    psis = [np.array(p) / np.linalg.norm(p) for p in g["psis"]]
    H = np.array([[psis[i] @ (j1_mat @ psis[j]) for j in range(2)] for i in range(2)])
    ropr = ed.represent(rhsr, j1)
    assert ropr.is_complex
    assert np.allclose(ropr.matrix(), H, atol=1e-13)
    assert np.allclose(H, [[0, 4 * math.sqrt(2)], [4 * math.sqrt(2), -4]])
    for i in (1, 2):
        row = np.zeros(2, dtype=complex)
        err = 0
        for j, a in ropr.get_row_iterator(i):
            if j > 0:
                row[j - 1] += a
            else:
                err += a
        assert abs(err) < 1e-12 and np.allclose(row, H[i - 1, :], atol=1e-13)
        col = np.zeros(2, dtype=complex)
        for j, a in ropr.get_column_iterator(i):
            if j > 0:
                col[j - 1] += a
        assert np.allclose(col, H[:, i - 1], atol=1e-13)
        for j in (1, 2):
            assert abs(ropr.get_element(i, j) - H[i - 1, j - 1]) < 1e-13
    for bad in ((0, 1), (3, 1), (1, 0), (1, 3)):
        with pytest.raises(IndexError):
            ropr.get_element(*bad)


@pytest.fixture(params=["simple", "staged"])
def k6_path(request, monkeypatch):
    """Run the reduced-representation tests through both K6 implementations (row-per-thread / word-parallel staged)."""
    monkeypatch.setenv("EDCUDA_K6_MIN_ROWS", "1" if request.param == "staged" else "1000000000000")
    return request.param


@pytest.mark.parametrize("n,qn", [(7, 1), (8, 0), (10, 0)])
def test_reduced_apply_vs_oracle(gpu_ed, n, qn, k6_path):
    ed = gpu_ed
    L = ed.lattices
    hs, h = ed.models.j1j2_chain(n, 0.5)
    hs_o, a = oracle_spin_chain(n)
    _, b = oracle_spin_chain(n, [(i, (i + 2) % n) for i in range(n)], jz=0.5, jxy=0.5)
    h_o = O.simplify(a + b)
    assert h.terms == terms_of(h_o)
    hsr, hsr_o = ed.represent(ed.HilbertSpaceSector(hs, qn)), O.represent(O.HilbertSpaceSector(hs_o, qn))
    full = np.linalg.eigvalsh(O.dense_matrix(O.OperatorRepresentation(hsr_o, h_o))) if hsr.dimension <= 300 else None
    spectrum = []
    for k in range(n):
        symops = L.chain_translation_irrep(n, k)
        rhsr = ed.symmetry_reduce(hsr, symops)
        if rhsr.dimension == 0:
            continue
        rhsr_o = O.symmetry_reduce(hsr_o, to_oracle_symops(symops))
        ropr, ropr_o = ed.represent(rhsr, h), O.ReducedOperatorRepresentation(rhsr_o, h_o)
        d = rhsr.dimension
        rng = np.random.default_rng(k)
        x = rng.standard_normal(d) + 1j * rng.standard_normal(d)
        for side in ("left", "right"):
            exp = O.apply_serial(np.zeros(d, dtype=complex), ropr_o, x, side)
            out = np.zeros(d, dtype=complex)
            if side == "left":
                ed.apply_b(out, ropr, x)
            else:
                ed.apply_b(out, x, ropr)
            assert rel_err(out, exp) < TOL
        out = np.full(d, 7.0 + 0j)
        ed.mul_b(out, ropr, x)
        assert rel_err(out, O.apply_serial(np.zeros(d, dtype=complex), ropr_o, x, "left")) < TOL
        with pytest.raises(ValueError):
            ed.apply_b(np.zeros(d), ropr, np.zeros(d))           # reduced representations are ComplexF64
        with pytest.raises(ed.DimensionMismatch):
            ed.apply_b(np.zeros(d + 1, dtype=complex), ropr, x)
        # iterators, sparse and dense agree with the oracle's
        i = 1 + (k % d)
        for got, exp in ((ropr.get_row_iterator(i), ropr_o.get_row_iterator(i)), (ropr.get_column_iterator(i), ropr_o.get_column_iterator(i))):
            assert [g_[0] for g_ in got] == [e[0] for e in exp]
            assert np.allclose([g_[1] for g_ in got], [e[1] for e in exp], atol=1e-13)
        cp, rv, nz = ropr.sparse_csc()
        cp_o, rv_o, nz_o = O.sparse_serial(ropr_o)
        assert np.array_equal(cp, cp_o) and np.array_equal(rv, rv_o) and rel_err(nz, nz_o) < TOL
        m = ropr.matrix()
        assert rel_err(m, O.dense_matrix(ropr_o)) < TOL
        assert np.allclose(m, m.conj().T, atol=1e-12)
        spectrum.extend(np.linalg.eigvalsh(m))
    if full is not None:
        assert np.allclose(sorted(spectrum), full, atol=1e-9)     # spectra union (test_reduced_representation.jl:236-252)


def test_reduced_square_4x4_momentum_sectors(gpu_ed, golden, k6_path):
    """Config 2: 4x4 square Heisenberg, all 16 momentum sectors: CSC structure bit-exact vs the oracle on two
    sectors (the oracle's Python loops are slow), ground state over all sectors = known answer."""
    ed = gpu_ed
    import scipy.sparse.linalg as spl
    L = ed.lattices
    hs, h = ed.models.heisenberg_square(4, 4)
    hs_o, h_o = oracle_spin_chain(16, L.square_bonds(4, 4))
    assert h.terms == terms_of(h_o)
    hsr, hsr_o = ed.represent(ed.HilbertSpaceSector(hs, 0)), O.represent(O.HilbertSpaceSector(hs_o, 0))
    e_min = []
    for k2 in range(4):
        for k1 in range(4):
            symops = L.torus_translation_irrep(4, 4, k1, k2)
            rhsr = ed.symmetry_reduce(hsr, symops)
            ropr = ed.represent(rhsr, h)
            m = ropr.sparse()
            assert abs(m - m.getH()).max() < 1e-12
            x = np.random.default_rng(k1 + 4 * k2).standard_normal(rhsr.dimension) + 0j
            assert rel_err(ropr * x, m @ x) < TOL
            e_min.append(spl.eigsh(m, k=1, which="SA", tol=1e-12)[0][0])
            if (k1, k2) in ((0, 0), (1, 2)):
                rhsr_o = O.symmetry_reduce(hsr_o, to_oracle_symops(symops))
                assert np.array_equal(rhsr.basis_list, rhsr_o.basis_list)
                cp, rv, nz = ropr.sparse_csc()
                cp_o, rv_o, nz_o = O.sparse_serial(O.ReducedOperatorRepresentation(rhsr_o, h_o))
                assert np.array_equal(cp, cp_o) and np.array_equal(rv, rv_o) and rel_err(nz, nz_o) < TOL
    assert abs(min(e_min) - golden["known_answers"]["sq4x4_E0"]) < 1e-9
    assert abs(e_min[0] - golden["known_answers"]["sq4x4_E0"]) < 1e-9       # ground state sits at k = 0


def test_triangular_space_group_small(gpu_ed, k6_path):
    """3x3... the 6x6 configuration's machinery (T x| C6v, k=0 A1, 432-element groups) on a 4x4 torus where the
    oracle can follow: representatives, mapping and reduced matvec agree; Burnside count checks the dimension."""
    ed = gpu_ed
    L = ed.lattices
    n = 4
    hs, h = ed.models.heisenberg_triangular(n)
    hs_o, h_o = oracle_spin_chain(n * n, L.triangular_bonds(n, n), jz=0.25, jxy=0.25)
    assert h.terms == terms_of(h_o)
    hsr, hsr_o = ed.represent(ed.HilbertSpaceSector(hs, 0)), O.represent(O.HilbertSpaceSector(hs_o, 0))
    symops = L.triangular_space_group_irrep(n, "A1")
    assert len(symops) == 12 * n * n
    rhsr = ed.symmetry_reduce(hsr, symops)
    # Burnside: number of orbits = average number of fixed points
    perms = [ed.symmetry._flatten(op, n * n)[0] for op, _ in symops]
    fixed = 0
    words = hsr.basis_list
    for p in perms:
        img = np.zeros_like(words)
        for i, j in enumerate(p):
            img |= ((words >> np.uint64(i)) & np.uint64(1)) << np.uint64(j)
        fixed += int((img == words).sum())
    assert rhsr.dimension == fixed // len(perms)
    rhsr_o = O.symmetry_reduce(hsr_o, to_oracle_symops(symops))
    assert np.array_equal(rhsr.basis_list, rhsr_o.basis_list)
    assert np.array_equal(rhsr.basis_mapping_index, rhsr_o.basis_mapping_index)
    assert np.max(np.abs(rhsr.basis_mapping_amplitude - rhsr_o.basis_mapping_amplitude)) < 1e-14
    ropr, ropr_o = ed.represent(rhsr, h), O.ReducedOperatorRepresentation(rhsr_o, h_o)
    d = rhsr.dimension
    x = np.random.default_rng(4).standard_normal(d) + 0j
    exp = O.apply_serial(np.zeros(d, dtype=complex), ropr_o, x, "left")
    assert rel_err(ropr * x, exp) < TOL


# ------------------------------------------------------------------ K7: Lanczos
def test_lanczos_vs_oracle_and_known_answers(gpu_ed, golden):
    ed = gpu_ed
    from edcuda.lanczos import lanczos
    k = golden["known_answers"]
    hs, h = ed.models.heisenberg_chain(16)
    hs_o, h_o = oracle_spin_chain(16)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    opr = ed.represent(hsr, h)
    opr_o = O.OperatorRepresentation(O.represent(O.HilbertSpaceSector(hs_o, 0)), h_o)
    d = hsr.dimension
    v0 = np.random.default_rng(20260717 + 1).standard_normal(d)
    res = lanczos(opr, 120, v0=v0)
    a_o, b_o = O.lanczos(lambda v: O.apply_vectorized(np.zeros(d), opr_o, v), v0, 40)
    assert np.allclose(res.alpha[:40], a_o, rtol=0, atol=1e-9) and np.allclose(res.beta[:40], b_o, rtol=0, atol=1e-9)
    assert abs(res.ritz[0] - k["L16_E0"]) < 1e-10
    assert abs(O.tridiag_eigvals(res.alpha, res.beta)[0] - res.ritz[0]) < 1e-10
    # seeded start vector generated on device; complex vectors; reduced representation
    res2 = lanczos(opr, 150, seed=7, dtype=np.complex128)
    assert abs(res2.ritz[0] - k["L16_E0"]) < 1e-10
    rhsr = ed.symmetry_reduce(hsr, ed.lattices.chain_translation_irrep(16, 0))
    res3 = lanczos(ed.represent(rhsr, h), 100, seed=11)
    assert abs(res3.ritz[0] - k["L16_E0"]) < 1e-10
    # Majumdar-Ghosh point: E0 = -1.5 L exactly
    hs12, mg = ed.models.j1j2_chain(12, 0.5)
    res4 = lanczos(ed.represent(ed.represent(ed.HilbertSpaceSector(hs12, 0)), mg), 200, seed=3)
    assert abs(res4.ritz[0] + 18.0) < 1e-10


# ------------------------------------------------------------------ K2 fast path (apply_u1.cu)
def _c_oracle_apply(n, n_dn, op, x, side=0):
    import ed_oracle_c as OC
    basis = OC.basis_fixed_popcount(n, n_dn)
    out = np.zeros_like(x)
    OC.apply(basis, op.arrays(), x, out, side=side)
    return basis, out


@pytest.mark.parametrize("n,n_dn,model", [(20, 10, "xxz"), (20, 7, "j1j2_field"), (16, 8, "square"), (16, 5, "triangular"),
                                          (24, 12, "xxz"), (13, 6, "open_chain_disorderfree"), (24, 12, "long_range")])
def test_fast_path_vs_c_oracle(gpu_ed, n, n_dn, model):
    """The tiled U(1) kernel against the oracle's C twin (reference algorithm) at sizes the Python oracle cannot
    reach, for every bond geometry the lowering distinguishes (several distance classes, wrap bonds, fields)."""
    ed = gpu_ed
    L = ed.lattices
    hs, pauli = ed.spin_half_system(n)
    if model == "xxz":
        h = ed.models.xxz_bonds(hs, L.chain_bonds(n), 1.0, 0.37)
    elif model == "j1j2_field":
        h = ed.simplify(ed.models.heisenberg_bonds(hs, L.chain_bonds(n, 1)) + ed.models.heisenberg_bonds(hs, L.chain_bonds(n, 2), 0.5)
                        + sum(0.3 * pauli(i, "z") for i in range(0, n, 2)) + 1.25 * ed.Operator([(0, 0, 0, 1.0)]))
    elif model == "square":
        h = ed.models.heisenberg_bonds(hs, L.square_bonds(4, 4))
    elif model == "triangular":
        h = ed.models.heisenberg_bonds(hs, L.triangular_bonds(4, 4), 0.25)
    elif model == "long_range":
        # one amplitude, 63 bonds inside the low 14 bits: the ELL class is split at 60 bonds
        h = ed.simplify(sum((ed.models.heisenberg_bonds(hs, L.chain_bonds(n, dist)) for dist in range(2, 7)),
                            ed.models.heisenberg_bonds(hs, L.chain_bonds(n, 1))))
    else:
        h = ed.models.xxz_bonds(hs, L.chain_bonds(n, 1, periodic=False), 0.8, -1.1)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, n - 2 * n_dn))
    assert hsr.kind == ed.ED_BASIS_COMBINADIC
    d = hsr.dimension
    rng = np.random.default_rng(n + n_dn)
    x = rng.standard_normal(d)
    basis, exp = _c_oracle_apply(n, n_dn, h, x)
    assert np.array_equal(hsr.download(0, d), basis)
    fast, gen = ed.represent(hsr, h), ed.represent(hsr, h).set_kernel(1)
    y_fast, y_gen = np.zeros(d), np.zeros(d)
    ed.mul_b(y_fast, fast, x)
    ed.mul_b(y_gen, gen, x)
    assert rel_err(y_gen, exp) < TOL
    assert rel_err(y_fast, exp) < TOL
    # apply! accumulates, x*H == H*x for the symmetric real operator, complex vectors
    ed.apply_b(y_fast, x, fast)
    assert rel_err(y_fast, 2 * exp) < TOL
    xc = x + 1j * rng.standard_normal(d)
    _, expc = _c_oracle_apply(n, n_dn, h, xc)
    yc = np.zeros(d, dtype=complex)
    ed.mul_b(yc, fast, xc)
    assert rel_err(yc, expc) < TOL
    # row shards through the fast path
    cuts = [0, d // 7, d // 2 + 3, d]
    parts = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        out = np.zeros(hi - lo)
        ed.mul_b(out, ed.represent(hsr, h).set_rows(lo, hi), x)
        parts.append(out)
    assert rel_err(np.concatenate(parts), exp) < TOL


def test_fast_path_falls_back_for_unsupported_operators(gpu_ed):
    ed = gpu_ed
    n = 10
    hs, pauli = ed.spin_half_system(n)
    hs_o, pauli_o = O.spin_half_system(n)
    # three-site term and asymmetric hopping: not expressible as symmetric bond exchange -> generic kernel, same answer
    op = ed.simplify(pauli(0, "z") * pauli(1, "+") * pauli(2, "-") + 0.5 * pauli(3, "+") * pauli(4, "-")
                     + ed.models.heisenberg_bonds(hs, ed.lattices.chain_bonds(n)))
    op_o = O.simplify(pauli_o(0, "z") * pauli_o(1, "+") * pauli_o(2, "-") + 0.5 * (pauli_o(3, "+") * pauli_o(4, "-"))
                      + oracle_spin_chain(n)[1])
    assert op.terms == terms_of(op_o)
    hsr, hsr_o = ed.represent(ed.HilbertSpaceSector(hs, 0)), O.represent(O.HilbertSpaceSector(hs_o, 0))
    _check_apply(ed, hsr, hsr_o, op, op_o, False)


def test_full_size_properties_l28(gpu_ed):
    """Config 3 (J1-J2 chain L=28, Sz=0, D = 40,116,600) at full size through size-independent properties:
    fast kernel == generic kernel, symmetry <x,Hy> = <Hx,y>, linearity, Majumdar-Ghosh ground energy -1.5 L."""
    ed = gpu_ed
    import torch
    from edcuda.lanczos import lanczos
    hs, h = ed.models.j1j2_chain(28, 0.5)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    d = hsr.dimension
    assert d == 40116600
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(d, dtype=torch.float64, device="cuda", generator=g)
    z = torch.randn(d, dtype=torch.float64, device="cuda", generator=g)
    fast, gen = ed.represent(hsr, h), ed.represent(hsr, h).set_kernel(1)
    hx_f, hx_g, hz = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    ed.mul_b(hx_f, fast, x)
    ed.mul_b(hx_g, gen, x)
    ed.mul_b(hz, fast, z)
    torch.cuda.synchronize()
    assert float((hx_f - hx_g).abs().max() / hx_g.abs().max()) < TOL
    assert abs(float(torch.dot(z, hx_f) - torch.dot(hz, x))) < 1e-10 * float(hx_f.norm() * z.norm())
    comb = torch.empty_like(x)
    ed.mul_b(comb, fast, 2.0 * x - 0.5 * z)
    assert float((comb - (2.0 * hx_f - 0.5 * hz)).abs().max() / hx_f.abs().max()) < TOL
    res = lanczos(fast, 160, seed=5)
    assert abs(res.ritz[0] + 42.0) < 1e-9


# ------------------------------------------------------------------ config 4 at full size
def test_triangular_6x6_full_size(gpu_ed, golden):
    """Config 4: 6x6 triangular Heisenberg, T x| C6v, k=0 A1, Sz=0.  The parent sector has 9,075,135,300 states and is
    never materialised; the reduced dimension must equal the Burnside count 21,029,820.  The reduced
    matvec is checked through size-independent properties: Hermiticity <z,Hx> = conj(<x,Hz>), linearity, and a
    sampled row against the row iterator."""
    ed = gpu_ed
    import torch
    hs, h = ed.models.heisenberg_triangular(6)
    assert len(h.terms) == 648
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    assert hsr.dimension == 9075135300 and hsr.kind == ed.ED_BASIS_COMBINADIC
    symops = ed.lattices.triangular_space_group_irrep(6, "A1")
    rhsr = ed.symmetry_reduce(hsr, symops)
    d = rhsr.dimension
    assert d == golden["known_answers"]["tri6x6_k0A1_dim"]
    words = rhsr.basis_list
    assert np.all(words[1:] > words[:-1]) and int(words[0]) == (1 << 18) - 1
    sizes = rhsr.orbit_sizes()
    assert sizes.max() == 432 and np.all(432 % sizes == 0)
    assert int(sizes.astype(np.int64).sum()) == hsr.dimension          # the orbits of the representatives tile the parent
    ropr = ed.represent(rhsr, h)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(d, dtype=torch.complex128, device="cuda", generator=g)
    z = torch.randn(d, dtype=torch.complex128, device="cuda", generator=g)
    hx, hz, hc = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    ed.mul_b(hx, ropr, x)
    ed.mul_b(hz, ropr, z)
    ed.mul_b(hc, ropr, (0.5 - 2j) * x + z)
    torch.cuda.synchronize()
    assert abs(complex(torch.vdot(z, hx)) - complex(torch.vdot(hz, x))) < 1e-9 * float(hx.norm() * z.norm())
    assert float((hc - ((0.5 - 2j) * hx + hz)).abs().max() / hx.abs().max()) < TOL
    i = 12345678
    xc = x.cpu().numpy()
    row = sum(a * xc[j - 1] for j, a in ropr.get_row_iterator(i) if j > 0)
    assert abs(row - complex(hx[i - 1])) < 1e-11 * abs(row)
    # cached CSR: assembled once on device, afterwards a bandwidth-bound SpMV that agrees with the matrix-free apply
    nnz = ropr.cache_matrix()
    assert 20 * d < nnz < 60 * d
    hx2 = torch.empty_like(x)
    ed.mul_b(hx2, ropr, x)
    torch.cuda.synchronize()
    assert float((hx2 - hx).abs().max() / hx.abs().max()) < TOL


# ------------------------------------------------------------------ cached CSR (SpMV) path
@pytest.mark.parametrize("block_cols", [None, 37])
def test_cached_matrix_apply_matches_matrix_free(gpu_ed, monkeypatch, block_cols):
    """block_cols = 37 forces the column-blocked form of the cached matrix (one SpMV pass per L2-sized window of x; at
    full size it only switches on when x outgrows the L2) on these small matrices: 25 and 3 blocks."""
    ed = gpu_ed
    from edcuda.lanczos import lanczos
    if block_cols:
        monkeypatch.setenv("EDCUDA_CSR_BLOCK_COLS", str(block_cols))
    n = 12
    hs, h = ed.models.j1j2_chain(n, 0.5)
    hs_o, a = oracle_spin_chain(n)
    _, b = oracle_spin_chain(n, [(i, (i + 2) % n) for i in range(n)], jz=0.5, jxy=0.5)
    h_o = O.simplify(a + b)
    hsr, hsr_o = ed.represent(ed.HilbertSpaceSector(hs, 0)), O.represent(O.HilbertSpaceSector(hs_o, 0))
    d = hsr.dimension
    rng = np.random.default_rng(0)
    x = rng.standard_normal(d)
    exp = O.apply_vectorized(np.zeros(d), O.OperatorRepresentation(hsr_o, h_o), x)
    opr = ed.represent(hsr, h)
    nnz = opr.cache_matrix()
    assert nnz == opr.sparse_csc(tol=0.0)[0][-1] - 1
    y = np.zeros(d)
    ed.mul_b(y, opr, x)
    assert rel_err(y, exp) < TOL
    ed.apply_b(y, opr, x)                      # accumulate through the cached path
    assert rel_err(y, 2 * exp) < TOL
    xc = x + 1j * rng.standard_normal(d)
    assert rel_err(opr * xc, O.apply_vectorized(np.zeros(d, dtype=complex), O.OperatorRepresentation(hsr_o, h_o), xc)) < TOL
    assert abs(lanczos(opr, 150, seed=1).ritz[0] + 18.0) < 1e-10      # Majumdar-Ghosh through the SpMV
    # row shard + cache; right side keeps using the matrix-free kernel until its own cache is built
    shard = ed.represent(hsr, h).set_rows(100, 700)
    shard.cache_matrix()
    ys = np.zeros(600)
    ed.mul_b(ys, shard, x)
    assert rel_err(ys, exp[100:700]) < TOL
    # reduced, complex characters
    symops = ed.lattices.chain_translation_irrep(n, 5)
    rhsr = ed.symmetry_reduce(hsr, symops)
    rhsr_o = O.symmetry_reduce(hsr_o, to_oracle_symops(symops))
    ropr, ropr_o = ed.represent(rhsr, h), O.ReducedOperatorRepresentation(rhsr_o, h_o)
    dr = rhsr.dimension
    xr = rng.standard_normal(dr) + 1j * rng.standard_normal(dr)
    free = ropr * xr
    ropr.cache_matrix()
    ropr.cache_matrix(1)
    assert rel_err(ropr * xr, O.apply_serial(np.zeros(dr, dtype=complex), ropr_o, xr, "left")) < TOL
    assert rel_err(xr * ropr, O.apply_serial(np.zeros(dr, dtype=complex), ropr_o, xr, "right")) < TOL
    assert rel_err(ropr * xr, free) < TOL
    ropr.drop_cache()
    assert rel_err(ropr * xr, free) < TOL
