"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol the header
declares, refuses to compute without a GPU (no CPU fallback), and the host operator/lattice helpers agree
with the oracle."""
import os
import re

import numpy as np
import pytest

import ed_oracle as O
from helpers import oracle_heisenberg_xyz, oracle_spin_chain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "edcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ed_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ed):
    import ctypes
    from edcuda import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 45
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/edcuda.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback(ed):
    if ed.device_count() > 0:
        pytest.skip("a GPU is present")
    hs, _ = ed.spin_half_system(4)
    with pytest.raises(ed.CudaError):
        ed.represent(ed.HilbertSpaceSector(hs, 0))
    with pytest.raises(ed.CudaError):
        ed.represent(hs)


def test_argument_errors_need_no_gpu(ed):
    hs, _ = ed.spin_half_system(9)
    with pytest.raises(ValueError):        # BR too small: ArgumentError (hilbert_space_representation.jl:63-69)
        ed.represent(hs, np.uint8)
    with pytest.raises(ValueError):
        ed.represent(ed.HilbertSpaceSector(hs, 1), np.uint8)
    with pytest.raises(ed.UnsupportedError):
        ed.represent(hs, "UInt128")
    with pytest.raises(ValueError):        # bits outside the mask (pure_operator.jl:32-36)
        ed.Operator([(0b01, 0b10, 0b00, 1.0)])
    import ctypes as C
    from edcuda._lib import lib, ED_ERR_ARGUMENT
    m = np.array([1], dtype=np.uint64); r = np.array([2], dtype=np.uint64); c = np.array([0], dtype=np.uint64)
    a = np.array([1.0]); h = C.c_void_p()
    assert lib.ed_operator_create(1, m.ctypes.data, r.ctypes.data, c.ctypes.data, a.ctypes.data, 0, C.byref(h)) == ED_ERR_ARGUMENT
    with pytest.raises(ValueError):        # |chi| != 1 (symmetry_reduce_generic.jl:27-29)
        ed.symmetry.SymmetryHandle(2, [(ed.SitePermutation([0, 1]), 1.0), (ed.SitePermutation([1, 0]), 0.5)])


def test_host_operator_algebra_matches_oracle(ed):
    for n in (4, 6):
        hs_o, h_o = oracle_heisenberg_xyz(n)
        hs, pauli = ed.spin_half_system(n)
        h = ed.simplify(sum(pauli(i, mu) * pauli((i + 1) % n, mu) for mu in "xyz" for i in range(n)))
        assert [(t.bitmask, t.bitrow, t.bitcol, t.amplitude) for t in h_o.terms] == h.terms
        assert not h.is_complex
        _, h2 = ed.models.heisenberg_chain(n)
        assert h2.terms == h.terms
    hs, pauli = ed.spin_half_system(3)
    hs_o, pauli_o = O.spin_half_system(3)
    a = pauli(0, "y") * pauli(1, "x") + 0.5 * pauli(2, "z") * pauli(0, "+")
    a_o = pauli_o(0, "y") * pauli_o(1, "x") + 0.5 * (pauli_o(2, "z") * pauli_o(0, "+"))
    assert [(t.bitmask, t.bitrow, t.bitcol, t.amplitude) for t in a_o.terms] == a.terms
    s, s_o = ed.simplify(a * a), O.simplify(a_o * a_o)
    assert [(t.bitmask, t.bitrow, t.bitcol, t.amplitude) for t in s_o.terms] == s.terms


def test_hilbert_space_descriptors(ed, golden):
    g = golden["tj"]
    sites = [ed.Site([ed.State(str(i), tuple(q)) for i, q in enumerate(states)]) for states in g["site_states"]]
    hs = ed.HilbertSpace(sites)
    assert hs.bitwidths == [2, 1, 1] and hs.bitoffsets == [0, 2, 3, 4]
    assert hs.quantum_number_sectors() == [tuple(q) for q in g["sectors"]]
    assert hs.get_bitmask(0) == 0b0011 and hs.get_bitmask(2) == 0b1000 and hs.get_bitmask() == 0b1111
    assert hs.compress([2, 1, 0]) == 0b0110 and hs.extract(0b0110) == (2, 1, 0)
    assert hs.get_quantum_number(0b0110) == (3, -1)
    assert ed.HilbertSpaceSector(hs, [(9, 9), (3, 1)]).allowed_quantum_numbers == {(3, 1)}


def test_lattice_groups_are_groups(ed):
    L = ed.lattices
    for ops in (L.chain_translation_irrep(6, 2), L.torus_translation_irrep(4, 4, 1, 3), L.triangular_space_group_irrep(6, "B2")):
        perms = [tuple(op.map) if isinstance(op, ed.SitePermutation) else tuple(ed.symmetry._flatten(op, len(ops[0][0].map))[0]) for op, _ in ops]
        assert perms[0] == tuple(range(len(perms[0])))
        assert len(set(perms)) == len(perms)
        index = {p: i for i, p in enumerate(perms)}
        chi = [c for _, c in ops]
        for a in range(0, len(perms), 7):
            for b in range(0, len(perms), 5):
                ab = tuple(perms[a][perms[b][i]] for i in range(len(perms[0])))
                assert ab in index
                assert abs(chi[index[ab]] - chi[a] * chi[b]) < 1e-12   # one-dimensional irrep: homomorphism
    assert len(L.triangular_space_group_irrep(6)) == 432
    bonds = L.triangular_bonds(6, 6)
    assert len(bonds) == 108
    bset = {frozenset(b) for b in bonds}
    for op, _ in L.triangular_space_group_irrep(6)[::17]:
        p = ed.symmetry._flatten(op, 36)[0]
        assert {frozenset((p[i], p[j])) for i, j in bonds} == bset


def test_term_walk_goldens_on_host_operators(ed, golden):
    # test/test_operator.jl:302-307, 560-577 through the product's host-side Operator
    g = golden["pure_iterators"]
    pop = ed.Operator([tuple(g["term"])])
    for b, exp in g["row"].items():
        assert ed.get_row_iterator(pop, int(b)) == [tuple(e) for e in exp]
    for b, exp in g["col"].items():
        assert ed.get_column_iterator(pop, int(b)) == [tuple(e) for e in exp]
    for br, bc, v in g["element"]:
        assert ed.get_element(pop, br, bc) == v
    g = golden["sum_iterators"]
    sop = ed.Operator([tuple(t) for t in g["terms"]])
    for b, exp in g["row"].items():
        assert ed.get_row_iterator(sop, int(b)) == [tuple(e) for e in exp]
    for b, exp in g["col"].items():
        assert ed.get_column_iterator(sop, int(b)) == [tuple(e) for e in exp]
    for br, bc, v in g["element"]:
        assert abs(ed.get_element(sop, br, bc) - v) < 1e-12


def test_isinvariant_gate_needs_no_gpu(ed):
    """ed_operator_isinvariant (Symmetry/symmetry_apply.jl:110-135 for every element of a symmetry) against the oracle's
    isinvariant, element by element; host only."""
    n = 8
    hs, pauli = ed.spin_half_system(n)
    hs_o, pauli_o = O.spin_half_system(n)
    L = ed.lattices
    ring = ed.models.heisenberg_bonds(hs, L.chain_bonds(n))
    _, ring_o = oracle_spin_chain(n)
    symops = L.chain_translation_irrep(n, 3)
    assert ed.isinvariant_all(hs, symops, ring) == (True, -1)
    assert all(O.isinvariant(hs_o, O.SitePermutation(op.map), ring_o) for op, _ in symops)
    # an open chain is not translation invariant: the first violating element is T^1
    open_chain = ed.models.heisenberg_bonds(hs, L.chain_bonds(n, periodic=False))
    _, open_o = oracle_spin_chain(n, [(i, i + 1) for i in range(n - 1)])
    ok, bad = ed.isinvariant_all(hs, symops, open_chain)
    assert not ok and bad == 1
    assert not O.isinvariant(hs_o, O.SitePermutation(symops[1][0].map), open_o)
    # a staggered field is invariant under even translations only
    stag = ed.simplify(ring + sum((-1.0) ** i * pauli(i, "z") for i in range(n)))
    ok, bad = ed.isinvariant_all(hs, symops, stag)
    assert not ok and bad == 1
    assert ed.isinvariant_all(hs, symops[::2], stag) == (True, -1)
    # the same operator written with different term lists is still invariant (matrix elements, not term lists)
    a = ed.simplify(ring + 0.5 * pauli(0, "z") * pauli(4, "z") + 0.5 * pauli(4, "z") * pauli(0, "z"))
    inv4 = [(L.chain_translation(n, 4 * k), 1.0) for k in range(2)]
    assert ed.isinvariant_all(hs, inv4, a) == (True, -1)
    # global bit flip: sz sz invariant, a uniform field is not
    flips = [(ed.SitePermutation(range(n)), 1.0), (ed.GlobalBitFlip(True), 1.0)]
    assert ed.isinvariant_all(hs, flips, ring) == (True, -1)
    assert ed.isinvariant_all(hs, flips, ed.simplify(ring + sum(pauli(i, "z") for i in range(n))))[0] is False
