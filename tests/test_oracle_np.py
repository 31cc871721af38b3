"""The vectorised oracle twin (oracle/ed_oracle_np.py) against the loop oracle (oracle/ed_oracle.py), which is pinned to
the reference's golden vectors by tests/test_oracle_golden.py.  CPU only."""
import numpy as np
import pytest

import ed_oracle as O
import ed_oracle_np as ON
from helpers import oracle_spin_chain, chain_translation_irrep


def _perms_chis(symops):
    return [op.map for op, _ in symops], [c for _, c in symops]


@pytest.mark.parametrize("n,k", [(4, 0), (4, 1), (7, 1), (8, 3), (10, 0), (10, 5)])
def test_symmetry_reduce_matches_loop_oracle(n, k, golden):
    hs, h = oracle_spin_chain(n)
    hsr = O.represent(O.HilbertSpaceSector(hs, n % 2))
    symops = chain_translation_irrep(n, k)
    exp = O.symmetry_reduce(hsr, symops)
    red, idx, amp = ON.symmetry_reduce(hsr.basis_list, *_perms_chis(symops))
    assert np.array_equal(red, exp.basis_list)
    assert np.array_equal(idx, exp.basis_mapping_index)
    assert np.max(np.abs(amp - exp.basis_mapping_amplitude)) < 1e-15
    if n == 4:       # the reference's own golden: test/test_symmetry_reduce.jl:34-48
        assert [int(b) for b in red] == golden["reduce_translation_chain4"]["irrep_1based"][str(k + 1)]


def test_symmetry_reduce_with_reflection_and_stabilisers():
    n = 8
    hs, _ = oracle_spin_chain(n)
    hsr = O.represent(O.HilbertSpaceSector(hs, 0))
    inv = O.SitePermutation([(-i) % n for i in range(n)])
    for parity in (1, -1):
        symops = [(p * t, cp * ct) for (t, ct) in chain_translation_irrep(n, 0)
                  for (p, cp) in ((O.SitePermutation(range(n)), 1.0 + 0j), (inv, complex(parity)))]
        exp = O.symmetry_reduce(hsr, symops)
        red, idx, amp = ON.symmetry_reduce(hsr.basis_list, *_perms_chis(symops))
        assert np.array_equal(red, exp.basis_list) and np.array_equal(idx, exp.basis_mapping_index)
        assert np.max(np.abs(amp - exp.basis_mapping_amplitude)) < 1e-15


@pytest.mark.parametrize("n,qn", [(8, 0), (10, 2)])
def test_sparse_plain_matches_loop_oracle(n, qn):
    hs, h = oracle_spin_chain(n, jz=0.7)
    hsr = O.represent(O.HilbertSpaceSector(hs, qn))
    for tol in (ON.RTOL_DEFAULT, 0.0):
        exp = O.sparse_serial(O.OperatorRepresentation(hsr, h), tol)
        got = ON.sparse_plain(hsr.basis_list, O.term_arrays(h), tol)
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
        assert np.array_equal(got[2], exp[2])        # same addition order: bit-identical values


@pytest.mark.parametrize("n,k", [(8, 0), (8, 3), (10, 5), (7, 2)])
def test_reduced_sparse_apply_and_vectors_match_loop_oracle(n, k):
    hs, h = oracle_spin_chain(n)
    hsr = O.represent(O.HilbertSpaceSector(hs, n % 2))
    symops = chain_translation_irrep(n, k)
    rhsr = O.symmetry_reduce(hsr, symops)
    ropr = O.ReducedOperatorRepresentation(rhsr, h)
    red, idx, amp = ON.symmetry_reduce(hsr.basis_list, *_perms_chis(symops))
    terms = O.term_arrays(h)
    exp = O.sparse_serial(ropr)
    got = ON.sparse_reduced(hsr.basis_list, red, idx, amp, terms)
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
    assert np.max(np.abs(got[2] - exp[2])) < 1e-14
    d = len(red)
    rng = np.random.default_rng(n + k)
    x = rng.standard_normal(d) + 1j * rng.standard_normal(d)
    y = O.apply_serial(np.zeros(d, dtype=complex), ropr, x, "left")
    assert np.max(np.abs(ON.apply_reduced(hsr.basis_list, red, idx, amp, terms, x) - y)) < 1e-13 * np.max(np.abs(y))
    large = rng.standard_normal(hsr.dimension) + 1j * rng.standard_normal(hsr.dimension)
    assert np.max(np.abs(ON.vector_reduce(idx, amp, d, large) - O.symmetry_reduce_vector(rhsr, large))) < 1e-14
    assert np.max(np.abs(ON.vector_unreduce(idx, amp, x) - O.symmetry_unreduce_vector(rhsr, x))) < 1e-14
