"""Plain-array lattice helper: bonds, site permutations and 1-D irrep characters for the lattices the
benchmark configurations name (chain, square, triangular).

The reference takes these objects from LatticeTools.jl (un-vendored; call sites
src/Symmetry/symmetry_reduce_translation.jl:26-27, symmetry_reduce_point.jl:24,
symmetry_reduce_symmorphic.jl:28-46).  The engine consumes plain (permutation, character) lists, so this
module fixes its own documented conventions:
  * sites are numbered x + n1*y (0-based), matching sub2ind in examples/spinhalfsquare_large.jl:35;
  * a translation by (a, b) sends site (x, y) to (x+a, y+b) and carries the character
    exp(-2 pi i (k1 a/n1 + k2 b/n2))  -- with amplitude conj(chi)/sqrt(N) this is the reference's
    psi_k = sum_x e^{+ikx}|x> convention pinned by test/test_symmetry_reduce.jl:186-193;
  * triangular lattice: a1=(1,0), a2=(-1/2, sqrt(3)/2); nearest-neighbour bonds (1,0), (1,1), (0,1)
    (examples/spinhalf_triangular.jl:17-21); C6: (n1,n2)->(n1-n2, n1); mirror: (n1,n2)->(n1-n2,-n2).
Every function returns `symops_and_amplitudes`: a list of (SitePermutation, character), identity first.
"""
from __future__ import annotations

import cmath
import math
from typing import List, Sequence, Tuple

from .symmetry import SitePermutation


# ------------------------------------------------------------------ chain
def chain_bonds(n: int, distance: int = 1, periodic: bool = True) -> List[Tuple[int, int]]:
    if periodic:
        return [(i, (i + distance) % n) for i in range(n)]
    return [(i, i + distance) for i in range(n - distance)]


def chain_translation(n: int, x: int) -> SitePermutation:
    return SitePermutation([(i + x) % n for i in range(n)])


def chain_translation_irrep(n: int, k: int):
    """Irrep k (0-based; the reference's irrep index is k+1) of the translation group of an n-chain."""
    return [(chain_translation(n, x), cmath.exp(-2j * math.pi * k * x / n)) for x in range(n)]


def chain_inversion(n: int) -> SitePermutation:
    """i -> -i mod n (the reference's SitePermutation([1,4,3,2]) at n=4, test/test_symmetry_apply.jl:21)."""
    return SitePermutation([(-i) % n for i in range(n)])


def chain_inversion_irrep(n: int, parity: int):
    """parity = +1 (irrep 1) or -1 (irrep 2) of {identity, inversion}."""
    return [(SitePermutation(range(n)), 1.0 + 0j), (chain_inversion(n), complex(parity))]


def symmorphic_product(t_irrep, p_irrep):
    """(p*t, phi_p*phi_t) with t outer, p inner (symmetry_reduce_symmorphic.jl:40-46)."""
    return [(p * t, phi_p * phi_t) for (t, phi_t) in t_irrep for (p, phi_p) in p_irrep]


# ------------------------------------------------------------------ square
def square_site(n1: int, n2: int, x: int, y: int) -> int:
    return (x % n1) + (y % n2) * n1


def square_bonds(n1: int, n2: int) -> List[Tuple[int, int]]:
    """Nearest-neighbour bonds in the order of examples/spinhalfsquare_large.jl:37-42."""
    out = []
    for i1 in range(n1):
        for i2 in range(n2):
            out.append((square_site(n1, n2, i1, i2), square_site(n1, n2, i1 + 1, i2)))
            out.append((square_site(n1, n2, i1, i2), square_site(n1, n2, i1, i2 + 1)))
    return out


def torus_translation(n1: int, n2: int, a: int, b: int) -> SitePermutation:
    return SitePermutation([square_site(n1, n2, (i % n1) + a, (i // n1) + b) for i in range(n1 * n2)])


def torus_translation_irrep(n1: int, n2: int, k1: int, k2: int):
    """Translation irrep (k1, k2) of an n1 x n2 torus; elements ordered with `a` fastest."""
    return [(torus_translation(n1, n2, a, b), cmath.exp(-2j * math.pi * (k1 * a / n1 + k2 * b / n2)))
            for b in range(n2) for a in range(n1)]


# ------------------------------------------------------------------ triangular
def triangular_bonds(n1: int, n2: int) -> List[Tuple[int, int]]:
    out = []
    for y in range(n2):
        for x in range(n1):
            for (dx, dy) in ((1, 0), (1, 1), (0, 1)):
                out.append((square_site(n1, n2, x, y), square_site(n1, n2, x + dx, y + dy)))
    return out


def _c6(x: int, y: int) -> Tuple[int, int]:
    return x - y, x


def _mirror(x: int, y: int) -> Tuple[int, int]:
    return x - y, -y


def triangular_point_ops(n: int):
    """The 12 elements C6^r * M^m of C6v on an n x n triangular torus -> [(perm, r, m)], identity first."""
    ops = []
    for m in range(2):
        for r in range(6):
            mapping = []
            for i in range(n * n):
                x, y = i % n, i // n
                if m:
                    x, y = _mirror(x, y)
                for _ in range(r):
                    x, y = _c6(x, y)
                mapping.append(square_site(n, n, x, y))
            ops.append((SitePermutation(mapping), r, m))
    return ops


_C6V_1D = {"A1": lambda r, m: 1.0, "A2": lambda r, m: (-1.0) ** m, "B1": lambda r, m: (-1.0) ** r,
           "B2": lambda r, m: (-1.0) ** (r + m)}


def triangular_space_group_irrep(n: int, irrep: str = "A1"):
    """k = 0 sector of T x| C6v on the n x n triangular torus with a one-dimensional C6v irrep:
    |G| = 12 n^2 elements (432 at n = 6), identity first, translation outer / point op inner."""
    chi = _C6V_1D[irrep]
    t_irrep = torus_translation_irrep(n, n, 0, 0)
    p_irrep = [(p, complex(chi(r, m))) for (p, r, m) in triangular_point_ops(n)]
    return symmorphic_product(t_irrep, p_irrep)
