"""Model Hamiltonians of the benchmark configurations, written the way the reference's examples write them
(Pauli normalisation, H = sum_bonds sum_mu sigma^mu sigma^mu; examples/spinhalfchain.jl:20,
examples/spinhalfsquare_large.jl:37-42, examples/spinhalf_triangular.jl:75-77) and passed through simplify
(src/Operator/operator_simplify.jl:32-81) so the term order is the reference's canonical one: 6 terms per bond.
"""
from __future__ import annotations

from typing import Iterable, Tuple

from .operators import Operator, pauli_matrix, simplify, spin_half_system


def xxz_bonds(hs, bonds: Iterable[Tuple[int, int]], jxy: float = 1.0, jz: float = 1.0) -> Operator:
    """sum_<ij> [ 2*jxy (s+_i s-_j + s-_i s+_j) + jz sz_i sz_j ]  (= jxy (sx sx + sy sy) + jz sz sz in Pauli matrices)."""
    acc = Operator()
    for (i, j) in bonds:
        acc = acc + (2.0 * jxy) * (pauli_matrix(hs, i, "+") * pauli_matrix(hs, j, "-"))
        acc = acc + (2.0 * jxy) * (pauli_matrix(hs, i, "-") * pauli_matrix(hs, j, "+"))
        acc = acc + float(jz) * (pauli_matrix(hs, i, "z") * pauli_matrix(hs, j, "z"))
    return simplify(acc)


def heisenberg_bonds(hs, bonds, j: float = 1.0) -> Operator:
    return xxz_bonds(hs, bonds, j, j)


def heisenberg_chain(n: int, j: float = 1.0):
    """(hs, H) of the periodic spin-1/2 Heisenberg chain, examples/spinhalfchain.jl:20."""
    from .lattices import chain_bonds
    hs, _ = spin_half_system(n)
    return hs, heisenberg_bonds(hs, chain_bonds(n), j)


def xxz_chain(n: int, delta: float = 1.0):
    from .lattices import chain_bonds
    hs, _ = spin_half_system(n)
    return hs, xxz_bonds(hs, chain_bonds(n), 1.0, delta)


def j1j2_chain(n: int, j2: float = 0.5):
    """J1-J2 chain (test/test_reduced_representation.jl:27-28)."""
    from .lattices import chain_bonds
    hs, _ = spin_half_system(n)
    h = heisenberg_bonds(hs, chain_bonds(n, 1), 1.0) + heisenberg_bonds(hs, chain_bonds(n, 2), j2)
    return hs, simplify(h)


def heisenberg_square(n1: int, n2: int):
    from .lattices import square_bonds
    hs, _ = spin_half_system(n1 * n2)
    return hs, heisenberg_bonds(hs, square_bonds(n1, n2))


def heisenberg_triangular(n: int, scale: float = 0.25):
    """0.25 * sum_NN sigma.sigma on the n x n triangular torus (examples/spinhalf_triangular.jl:75-77)."""
    from .lattices import triangular_bonds
    hs, _ = spin_half_system(n * n)
    return hs, heisenberg_bonds(hs, triangular_bonds(n, n), scale)
