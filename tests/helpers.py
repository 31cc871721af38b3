"""Shared builders for tests: the same model written once with the oracle's algebra and once with the product's."""
import cmath
import math

import numpy as np

import ed_oracle as O


def oracle_spin_chain(n, bonds=None, jz=1.0, jxy=1.0):
    hs, pauli = O.spin_half_system(n)
    if bonds is None:
        bonds = [(i, (i + 1) % n) for i in range(n)]
    h = None
    for (i, j) in bonds:
        t = (2.0 * jxy) * (pauli(i, "+") * pauli(j, "-")) + (2.0 * jxy) * (pauli(i, "-") * pauli(j, "+")) \
            + float(jz) * (pauli(i, "z") * pauli(j, "z"))
        h = t if h is None else h + t
    return hs, O.simplify(h)


def oracle_heisenberg_xyz(n, bonds=None):
    """sum sigma^mu sigma^mu written with x,y,z like the reference examples (complex intermediate)."""
    hs, pauli = O.spin_half_system(n)
    if bonds is None:
        bonds = [(i, (i + 1) % n) for i in range(n)]
    h = None
    for mu in ("x", "y", "z"):
        for (i, j) in bonds:
            t = pauli(i, mu) * pauli(j, mu)
            h = t if h is None else h + t
    return hs, O.simplify(h)


def oracle_from_terms(terms):
    ts = [O.PureOperator(m, r, c, a) for (m, r, c, a) in terms]
    return ts[0] if len(ts) == 1 else O.SumOperator(ts)


def to_oracle_symops(symops):
    """product-side (SitePermutation|GlobalBitFlip|DirectProduct, chi) list -> oracle objects."""
    import edcuda
    out = []

    def conv(op):
        if isinstance(op, edcuda.SitePermutation):
            return O.SitePermutation(op.map)
        if isinstance(op, edcuda.GlobalBitFlip):
            return O.GlobalBitFlip(op.value)
        if isinstance(op, edcuda.DirectProductOperation):
            return O.DirectProductOperation([conv(o) for o in op.operations])
        raise TypeError(op)

    for op, chi in symops:
        out.append((conv(op), complex(chi)))
    return out


def chain_translation_irrep(n, k):
    return [(O.SitePermutation([(i + x) % n for i in range(n)]), cmath.exp(-2j * math.pi * k * x / n)) for x in range(n)]


def rel_err(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / den)
