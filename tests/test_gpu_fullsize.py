"""Parity at the FULL sizes of BASELINE.json's five configurations: the CUDA path (through the C ABI) against the oracle
on the SAME inputs.  Configs 1-2 are compared in full with the vectorised oracle twin (oracle/ed_oracle_np.py); configs
3-5 with the oracle's C twin (oracle/ed_oracle_c.c: the reference's apply_parallel! -- term walk in order + binary search
per hit, abstract_operator_representation.jl:358-378 -- and the reduced row iterator of
reduced_operator_representation.jl:57-85 evaluated on the fly): config 3 over all 40,116,600 rows, configs 4 and 5 on
sampled row blocks that include the first, a middle and the last kernel tile and the rows either side of the half-basis
wrap (where the byte offsets of the L=32 vectors pass 2^32).  Tolerance: bit-exact structure, 1e-12 relative values.
"""
import math

import numpy as np
import pytest

import ed_oracle as O
import ed_oracle_c as OC
import ed_oracle_np as ON
from helpers import oracle_spin_chain, rel_err, to_oracle_symops

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _oracle_terms(n, bonds, jz=1.0, jxy=1.0):
    """term arrays from the ORACLE's operator algebra (never the product's), simplify order."""
    _, h = oracle_spin_chain(n, bonds, jz=jz, jxy=jxy)
    return O.term_arrays(h)


def _same_terms(h, terms):
    m, r, c, a = h.arrays()
    return (np.array_equal(m, terms[0]) and np.array_equal(r, terms[1]) and np.array_equal(c, terms[2])
            and np.array_equal(np.asarray(a, dtype=terms[3].dtype), terms[3]))


def _blocks(d, fixed, n_random, size, seed):
    """sorted, disjoint [lo, hi) row blocks: the given ones plus n_random blocks of `size` rows."""
    rng = np.random.default_rng(seed)
    out = [(max(0, lo), min(d, hi)) for lo, hi in fixed]
    for lo in rng.integers(0, d - size, n_random):
        out.append((int(lo), int(lo) + size))
    out.sort()
    merged = []
    for lo, hi in out:
        if merged and lo < merged[-1][1]:
            merged[-1] = (merged[-1][0], max(hi, merged[-1][1]))
        else:
            merged.append((lo, hi))
    return merged


def _randn_device(ed, d, seed, cplx=False):
    """the library's Philox vector keyed by the global row (the bench's synthetic input) as a torch tensor"""
    import ctypes as C
    import torch
    from edcuda._lib import lib, check
    x = torch.empty(d, dtype=torch.complex128 if cplx else torch.float64, device="cuda")
    check(lib.ed_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream), 1))
    try:
        check(lib.ed_vector_randn_async(x.data_ptr(), d, ed.ED_C128 if cplx else ed.ED_F64, seed, 0))
    finally:
        lib.ed_set_stream(None, 0)
    torch.cuda.synchronize()
    return x


# ------------------------------------------------------------------ config 1
def test_config1_l16_csc_bit_exact_and_matvec(gpu_ed, golden):
    """Heisenberg chain L=16, Sz=0 (12,870 states, 96 terms): basis, sparse() CSC structure bit-exact over all 12,870
    columns (117,794 entries), values and matvec to 1e-12."""
    ed = gpu_ed
    n = 16
    hs, h = ed.models.heisenberg_chain(n)
    terms = _oracle_terms(n, None)
    assert _same_terms(h, terms)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    basis = OC.basis_fixed_popcount(n, n // 2)
    assert np.array_equal(hsr.basis_list, basis) and len(basis) == golden["known_answers"]["L16_dim"]
    opr = ed.represent(hsr, h)
    colptr, rowval, nzval = opr.sparse_csc()
    cp_o, rv_o, nz_o = ON.sparse_plain(basis, terms)
    assert colptr[-1] - 1 == golden["known_answers"]["L16_nnz_sparse"]
    assert np.array_equal(colptr, cp_o) and np.array_equal(rowval, rv_o)
    assert rel_err(nzval, nz_o) < TOL
    cp0, rv0, nz0 = opr.sparse_csc(tol=0.0)           # nothing chopped: exact-zero diagonals stay
    cp0_o, rv0_o, nz0_o = ON.sparse_plain(basis, terms, 0.0)
    assert np.array_equal(cp0, cp0_o) and np.array_equal(rv0, rv0_o) and rel_err(nz0, nz0_o) < TOL
    rng = np.random.default_rng(16)
    for x in (rng.standard_normal(len(basis)), rng.standard_normal(len(basis)) + 1j * rng.standard_normal(len(basis))):
        for side in (0, 1):
            exp = np.zeros_like(x)
            OC.apply(basis, terms, x, exp, side=side)
            got = (opr * x) if side == 0 else (x * opr)
            assert rel_err(got, exp) < TOL


# ------------------------------------------------------------------ config 2
def test_config2_square_4x4_all_16_sectors_bit_exact(gpu_ed, golden):
    """4x4 square Heisenberg, translation-reduced: for EVERY momentum (k1, k2) the reduced basis, basis_mapping_index and
    the CSC structure of sparse() are bit-exact against the oracle; amplitudes, CSC values and the reduced matvec 1e-12."""
    ed = gpu_ed
    L = ed.lattices
    hs, h = ed.models.heisenberg_square(4, 4)
    terms = _oracle_terms(16, L.square_bonds(4, 4))
    assert _same_terms(h, terms)
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    basis = OC.basis_fixed_popcount(16, 8)
    assert np.array_equal(hsr.basis_list, basis)
    dims = []
    rng = np.random.default_rng(44)
    for k2 in range(4):
        for k1 in range(4):
            symops = L.torus_translation_irrep(4, 4, k1, k2)
            perms = [op.map for op, _ in symops]
            chis = [c for _, c in symops]
            red_o, idx_o, amp_o = ON.symmetry_reduce(basis, perms, chis)
            rhsr = ed.symmetry_reduce(hsr, symops)
            assert np.array_equal(rhsr.basis_list, red_o), (k1, k2)
            assert np.array_equal(rhsr.basis_mapping_index, idx_o), (k1, k2)
            assert np.max(np.abs(rhsr.basis_mapping_amplitude - amp_o)) < 1e-14
            dims.append(rhsr.dimension)
            ropr = ed.represent(rhsr, h)
            cp, rv, nz = ropr.sparse_csc()
            cp_o, rv_o, nz_o = ON.sparse_reduced(basis, red_o, idx_o, amp_o, terms)
            assert np.array_equal(cp, cp_o) and np.array_equal(rv, rv_o), (k1, k2)
            assert rel_err(nz, nz_o) < TOL
            d = rhsr.dimension
            x = rng.standard_normal(d) + 1j * rng.standard_normal(d)
            assert rel_err(ropr * x, ON.apply_reduced(basis, red_o, idx_o, amp_o, terms, x)) < TOL
    assert dims == golden["known_answers"]["sq4x4_sector_dims"] and sum(dims) == 12870


# ------------------------------------------------------------------ config 3
def test_config3_j1j2_l28_full_vs_c_oracle(gpu_ed):
    """J1-J2 chain L=28, Sz=0: all 40,116,600 rows of y = H x (tiled kernel) against the C twin on the same x; the
    generic term-walk kernel, ComplexF64 vectors and x*H on sampled row blocks."""
    ed = gpu_ed
    import torch
    n = 28
    hs, h = ed.models.j1j2_chain(n, 0.5)
    bonds1 = [(i, (i + 1) % n) for i in range(n)]
    bonds2 = [(i, (i + 2) % n) for i in range(n)]
    _, a = oracle_spin_chain(n, bonds1)
    _, b = oracle_spin_chain(n, bonds2, jz=0.5, jxy=0.5)
    terms = O.term_arrays(O.simplify(a + b))
    assert _same_terms(h, terms) and len(terms[0]) == 336
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    d = hsr.dimension
    assert d == 40116600
    basis = OC.basis_fixed_popcount(n, n // 2)
    x = _randn_device(ed, d, 20260717 + 3)
    y = torch.empty_like(x)
    fast = ed.represent(hsr, h)
    ed.mul_b(y, fast, x)
    torch.cuda.synchronize()
    xh, yh = x.cpu().numpy(), y.cpu().numpy()
    exp = np.zeros(d)
    OC.apply(basis, terms, xh, exp)
    scale = np.max(np.abs(exp))
    assert np.max(np.abs(yh - exp)) / scale < TOL
    half = math.comb(n - 1, n // 2)
    blocks = _blocks(d, [(0, 20000), (half - 20000, half + 20000), (d - 20000, d)], 6, 20000, 28)
    xc = x.to(torch.complex128) + 1j * _randn_device(ed, d, 99)
    xch = xc.cpu().numpy()
    for lo, hi in blocks:
        assert np.array_equal(hsr.download(lo, hi - lo), basis[lo:hi])
        out = np.zeros(hi - lo)
        ed.mul_b(out, ed.represent(hsr, h).set_kernel(1).set_rows(lo, hi), xh)        # generic kernel, host vectors
        assert np.max(np.abs(out - exp[lo:hi])) / scale < TOL
        expc = np.zeros(hi - lo, dtype=complex)
        OC.apply(basis, terms, xch, expc, lo, hi, 1)                                  # column walk: x * H
        outc = torch.zeros(hi - lo, dtype=torch.complex128, device="cuda")
        ed.apply_b(outc, xc, ed.represent(hsr, h).set_rows(lo, hi))
        torch.cuda.synchronize()
        assert np.max(np.abs(outc.cpu().numpy() - expc)) / scale < TOL


# ------------------------------------------------------------------ config 5
def test_config5_xxz_l32_sampled_rows_vs_c_oracle(gpu_ed):
    """XXZ chain L=32, Sz=0 (601,080,390 states; the headline configuration): y = H x by the tiled kernel over the whole
    basis on the bench's Philox input, compared with the C twin on > 1e6 sampled rows: first / middle / last tiles, the
    rows either side of the half-basis wrap, the rows where the byte offset of a vector element passes 2^32, and random
    blocks; the generic kernel on a subset.  Basis words of the sampled rows are compared bit-exact as well."""
    ed = gpu_ed
    import torch
    n = 32
    hs, h = ed.models.xxz_chain(n, 1.0)
    terms = _oracle_terms(n, None)
    assert _same_terms(h, terms) and len(terms[0]) == 192
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    d = hsr.dimension
    assert d == 601080390
    x = _randn_device(ed, d, 20260717 + 5)
    y = torch.empty_like(x)
    opr = ed.represent(hsr, h)
    ed.mul_b(y, opr, x)
    torch.cuda.synchronize()
    checksum = float(torch.dot(x, y))
    xh = x.cpu().numpy()
    del x
    basis = OC.basis_fixed_popcount(n, n // 2)
    half = math.comb(n - 1, n // 2)          # first row whose top site is occupied
    four_gb = (1 << 32) // 8                 # element whose byte offset is 2^32
    fixed = [(0, 30000), (d - 30000, d), (half - 50000, half + 50000), (four_gb - 30000, four_gb + 30000),
             (d // 2 - 30000, d // 2 + 30000), (d // 3, d // 3 + 30000)]
    blocks = _blocks(d, fixed, 32, 30000, 32)
    total = 0
    scale = float(y.abs().max())
    worst = 0.0
    for i, (lo, hi) in enumerate(blocks):
        exp = np.zeros(hi - lo)
        OC.apply(basis, terms, xh, exp, lo, hi)
        got = y[lo:hi].cpu().numpy()
        worst = max(worst, float(np.max(np.abs(got - exp))) / scale)
        assert np.array_equal(hsr.download(lo, 1000), basis[lo:lo + 1000])
        if i % 8 == 0:                        # generic term-walk kernel on the same rows (device binary ranks, no tiles)
            out = torch.zeros(hi - lo, dtype=torch.float64, device="cuda")
            xd = torch.from_numpy(xh).cuda()
            ed.mul_b(out, ed.represent(hsr, h).set_kernel(1).set_rows(lo, hi), xd)
            torch.cuda.synchronize()
            worst = max(worst, float(np.max(np.abs(out.cpu().numpy() - exp))) / scale)
            del xd
        total += hi - lo
    assert total >= 1_000_000
    assert worst < TOL, worst
    # the bench's checksum <x, Hx> of this input (profiles/r01_bench_line_n1.json: -1.0313662088158513 after 1/D scaling)
    assert abs(checksum / d + 1.0313662088158513) < 1e-9


# ------------------------------------------------------------------ config 4
@pytest.mark.parametrize("irrep", ["A1", "B2"])
def test_config4_tri6x6_sampled_rows_vs_c_oracle(gpu_ed, golden, irrep):
    """6x6 triangular Heisenberg, T x| C6v, k=0, Sz=0 (parent 9,075,135,300 states, |G| = 432): > 1e4 sampled rows of the
    reduced matvec -- matrix-free (K6) and through the cached CSR -- against the C twin's on-the-fly evaluation of the
    reference's reduced row iterator (orbit scan over all 432 elements per hit) on the same x.  A1 (all characters 1) is
    BASELINE's sector; B2 exercises the signs of the characters at full size."""
    ed = gpu_ed
    import torch
    L = ed.lattices
    hs, h = ed.models.heisenberg_triangular(6)
    terms = _oracle_terms(36, L.triangular_bonds(6, 6), jz=0.25, jxy=0.25)
    assert _same_terms(h, terms) and len(terms[0]) == 648
    hsr = ed.represent(ed.HilbertSpaceSector(hs, 0))
    symops = L.triangular_space_group_irrep(6, irrep)
    rhsr = ed.symmetry_reduce(hsr, symops)
    d = rhsr.dimension
    if irrep == "A1":
        assert d == golden["known_answers"]["tri6x6_k0A1_dim"]
    rbasis = rhsr.basis_list
    orbit = rhsr.orbit_sizes()
    perms = np.array([op.map for op, _ in symops], dtype=np.int32)
    chis = [c for _, c in symops]
    x = _randn_device(ed, d, 20260717 + 4, cplx=True)
    y = torch.empty_like(x)
    ropr = ed.represent(rhsr, h)
    ed.mul_b(y, ropr, x)
    torch.cuda.synchronize()
    xh = x.cpu().numpy()
    scale = float(y.abs().max())
    blocks = _blocks(d, [(0, 1000), (d - 1000, d), (d // 2, d // 2 + 1000)], 9, 1000, 36)
    exps = []
    total = 0
    for lo, hi in blocks:
        exp = np.zeros(hi - lo, dtype=complex)
        OC.apply_reduced_onthefly(rbasis, orbit, perms, chis, terms, xh, exp, lo, hi)
        exps.append(exp)
        assert np.max(np.abs(y[lo:hi].cpu().numpy() - exp)) / scale < TOL
        total += hi - lo
    assert total >= 10_000
    if irrep == "A1":
        nnz = ropr.cache_matrix()
        assert 20 * d < nnz < 60 * d
        y2 = torch.empty_like(x)
        ed.mul_b(y2, ropr, x)
        torch.cuda.synchronize()
        for (lo, hi), exp in zip(blocks, exps):
            assert np.max(np.abs(y2[lo:hi].cpu().numpy() - exp)) / scale < TOL
