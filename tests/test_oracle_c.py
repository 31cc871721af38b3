"""The oracle's C/OpenMP twin (oracle/ed_oracle_c.c -- the timed CPU baseline) against the Python oracle,
which is itself pinned to the reference's golden vectors.  CPU only."""
import math

import numpy as np
import pytest

import ed_oracle as O
import ed_oracle_c as OC
from helpers import oracle_spin_chain, chain_translation_irrep


def test_basis_matches_python_oracle(golden):
    g = golden["sector_basis_spin_half_4_sz0"]
    assert list(OC.basis_fixed_popcount(4, 2)) == g["basis"]
    for n in (1, 5, 9, 12):
        hs = O.HilbertSpace([O.Site([O.State("Up", 1), O.State("Dn", -1)])] * n)
        for n_dn in range(n + 1):
            exp = O.hs_get_basis_list(O.HilbertSpaceSector(hs, n - 2 * n_dn))
            assert list(OC.basis_fixed_popcount(n, n_dn)) == exp
            assert list(OC.sector_basis_dp([2] * n, [[1, -1]] * n, n - 2 * n_dn)) == exp
    # mixed widths (one-number version of the tJ golden): 3-state site then two 2-state sites
    hs = O.HilbertSpace([O.Site([O.State("e", 0), O.State("u", 1), O.State("d", -1)]),
                         O.Site([O.State("u", 1), O.State("d", -1)]), O.Site([O.State("u", 1), O.State("d", -1)])])
    for q in (-3, -1, 0, 2):
        exp = O.hs_get_basis_list(O.HilbertSpaceSector(hs, q))
        assert list(OC.sector_basis_dp([3, 2, 2], [[0, 1, -1], [1, -1], [1, -1]], q)) == exp


@pytest.mark.parametrize("cplx", [False, True])
def test_apply_matches_python_oracle(cplx):
    n = 12
    hs, h = oracle_spin_chain(n, jz=0.7)
    hsr = O.represent(O.HilbertSpaceSector(hs, 0))
    opr = O.OperatorRepresentation(hsr, h)
    basis = OC.basis_fixed_popcount(n, n // 2)
    assert np.array_equal(basis, hsr.basis_list)
    rng = np.random.default_rng(2)
    x = rng.standard_normal(len(basis))
    if cplx:
        x = x + 1j * rng.standard_normal(len(basis))
    for side, so in ((0, "left"), (1, "right")):
        exp = O.apply_serial(np.zeros_like(x), opr, x, so)
        out = np.zeros_like(x)
        OC.apply(basis, O.term_arrays(h), x, out, side=side)
        assert np.array_equal(out, exp)          # same association as the reference loop: bit-identical
        part = np.zeros(300, dtype=x.dtype)
        OC.apply(basis, O.term_arrays(h), x, part, 200, 500, side)
        assert np.array_equal(part, exp[200:500])
    assert OC.count_hits(basis, O.term_arrays(h)) == sum(len(opr.get_row_iterator(i)) for i in range(1, len(basis) + 1))


def test_reduced_onthefly_matches_python_oracle():
    n = 10
    hs, h = oracle_spin_chain(n)
    hsr = O.represent(O.HilbertSpaceSector(hs, 0))
    for k in (0, 3):
        symops = chain_translation_irrep(n, k)
        rhsr = O.symmetry_reduce(hsr, symops)
        ropr = O.ReducedOperatorRepresentation(rhsr, h)
        d = rhsr.dimension
        x = np.random.default_rng(k).standard_normal(d) + 0.5j
        exp = O.apply_serial(np.zeros(d, dtype=complex), ropr, x, "left")
        _, sizes = O.reduced_representatives_on_the_fly(hs, [int(b) for b in hsr.basis_list], symops)
        perms = np.array([op.map for op, _ in symops], dtype=np.int32)
        out = np.zeros(d, dtype=complex)
        OC.apply_reduced_onthefly(rhsr.basis_list, np.array(sizes, dtype=np.int32), perms, [c for _, c in symops],
                                  O.term_arrays(h), x, out)
        assert np.max(np.abs(out - exp)) / np.max(np.abs(exp)) < 1e-13
