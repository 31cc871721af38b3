"""Host-side operator containers: what the Julia shim would flatten before calling ed_operator_create.

Reference: src/Operator/pure_operator.jl:25-52,78-90,136-154,184-201 (PureOperator, isless, product,
pure_operator), sum_operator.jl:13-23,62-100 (SumOperator algebra), operator_simplify.jl:12-81
(simplify), Toolkit/spin_half.jl:9-42 (spin_half_system, pauli_matrix).
Only the term layout (bitmask, bitrow, bitcol, amplitude) matters to the engine; the small algebra
here exists so Hamiltonians can be written the way the reference's examples write them.
A term is stored as a plain tuple (mask, row, col, amp); an operator is an ordered list of terms.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from ._lib import check, lib

Term = Tuple[int, int, int, complex]
RTOL = float(np.sqrt(np.finfo(np.float64).eps))
_MASK64 = (1 << 64) - 1


class Operator:
    """Ordered sum of pure terms.  len(terms)==0 plays NullOperator, ==1 PureOperator, else SumOperator."""

    __slots__ = ("terms", "_handle")

    def __init__(self, terms: Sequence[Term] = ()):
        out = []
        for (m, r, c, a) in terms:
            m, r, c = int(m), int(r), int(c)
            if (~m & r) & _MASK64:
                raise ValueError("every bit of bitrow not in bitmask should be set to zero")
            if (~m & c) & _MASK64:
                raise ValueError("every bit of bitcol not in bitmask should be set to zero")
            out.append((m, r, c, a))
        self.terms: List[Term] = out
        self._handle = None

    # ---- algebra (term order follows sum_operator.jl) ---------------------------------
    def __add__(self, other):
        if isinstance(other, (int, float, complex)):
            if other == 0:
                return self
            other = Operator([(0, 0, 0, other)])
        return Operator(self.terms + other.terms)

    def __radd__(self, other):
        if isinstance(other, (int, float, complex)) and other == 0:
            return self  # lets sum(...) start from 0
        return Operator([(0, 0, 0, other)]) + self

    def __neg__(self):
        return Operator([(m, r, c, -a) for (m, r, c, a) in self.terms])

    def __sub__(self, other):
        return self + (-other)

    def __mul__(self, other):
        if isinstance(other, Operator):
            out = []
            for (m2, r2, c2, a2) in other.terms:      # vec([tl*tr for tl in lhs, tr in rhs]): lhs index fastest
                for (m1, r1, c1, a1) in self.terms:
                    inter = m1 & m2
                    if (c1 & inter) != (r2 & inter):  # pure_operator.jl:145-146
                        continue
                    out.append((m1 | m2, r1 | (r2 & ~m1 & m2), c2 | (c1 & m1 & ~m2), a1 * a2))
            return Operator(out)
        return Operator([(m, r, c, a * other) for (m, r, c, a) in self.terms])

    def __rmul__(self, other):
        return Operator([(m, r, c, other * a) for (m, r, c, a) in self.terms])

    def __truediv__(self, other):
        return Operator([(m, r, c, a / other) for (m, r, c, a) in self.terms])

    def __pow__(self, n: int):
        out = Operator([(0, 0, 0, 1)])
        for _ in range(n):
            out = out * self
        return out

    def adjoint(self):
        return Operator([(m, c, r, np.conj(a)) for (m, r, c, a) in self.terms])

    def transpose(self):
        return Operator([(m, c, r, a) for (m, r, c, a) in self.terms])

    def __eq__(self, other):
        return isinstance(other, Operator) and self.terms == other.terms

    def __len__(self):
        return len(self.terms)

    def __repr__(self):
        return f"Operator({len(self.terms)} terms)"

    @property
    def is_complex(self) -> bool:
        return any(isinstance(a, (complex, np.complexfloating)) for (_, _, _, a) in self.terms)

    # ---- flattening --------------------------------------------------------------------
    def arrays(self):
        m = np.array([t[0] for t in self.terms], dtype=np.uint64)
        r = np.array([t[1] for t in self.terms], dtype=np.uint64)
        c = np.array([t[2] for t in self.terms], dtype=np.uint64)
        a = np.array([t[3] for t in self.terms], dtype=np.complex128 if self.is_complex else np.float64)
        return m, r, c, a

    def handle(self):
        if self._handle is None:
            m, r, c, a = self.arrays()
            h = C.c_void_p()
            cplx = 1 if a.dtype == np.complex128 else 0
            check(lib.ed_operator_create(len(self.terms), m.ctypes.data, r.ctypes.data, c.ctypes.data,
                                         a.ctypes.data, cplx, C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and lib is not None:
            try:
                lib.ed_operator_destroy(h)
            except Exception:
                pass


NullOperator = Operator


def simplify(op: Operator, tol: float = RTOL) -> Operator:
    """operator_simplify.jl:32-81: drop |a| <= tol, sort by (mask,row,col,re,im), merge equal
    (mask,row,col), drop merged ~0, demote to real when every imaginary part vanishes."""
    cplx = op.is_complex
    kept = []
    for (m, r, c, a) in op.terms:
        if abs(a) <= tol:
            continue
        if cplx:
            a = complex(a)
            if abs(a.imag) <= tol:
                a = complex(a.real, 0.0)
        kept.append((m, r, c, a))
    kept.sort(key=lambda t: (t[0], t[1], t[2], complex(t[3]).real, complex(t[3]).imag))
    out = []
    for (m, r, c, a) in kept:
        if out and out[-1][:3] == (m, r, c):
            out[-1] = (m, r, c, out[-1][3] + a)
        else:
            out.append((m, r, c, a))
    out = [t for t in out if not abs(t[3]) <= tol]
    if cplx and out and max(abs(complex(t[3]).imag) for t in out) <= tol:
        out = [(m, r, c, complex(a).real) for (m, r, c, a) in out]
    return Operator(out)


def pure_operator(hs, isite: int, istate_row: int, istate_col: int, amplitude=1) -> Operator:
    """pure_operator.jl:184-201 (0-based site and local-state indices)."""
    site = hs.sites[isite]
    if not (0 <= istate_row < site.dimension and 0 <= istate_col < site.dimension):
        raise IndexError("local state index out of range")
    off = hs.bitoffsets[isite]
    return Operator([(hs.get_bitmask(isite), istate_row << off, istate_col << off, amplitude)])


def pauli_matrix(hs, isite: int, j: str) -> Operator:
    """Toolkit/spin_half.jl:28-42: Up = local state 0, Dn = local state 1."""
    if j == "x":
        return pure_operator(hs, isite, 0, 1, 1) + pure_operator(hs, isite, 1, 0, 1)
    if j == "y":
        return pure_operator(hs, isite, 0, 1, -1j) + pure_operator(hs, isite, 1, 0, 1j)
    if j == "z":
        return pure_operator(hs, isite, 0, 0, 1) + pure_operator(hs, isite, 1, 1, -1)
    if j == "+":
        return pure_operator(hs, isite, 0, 1, 1)
    if j == "-":
        return pure_operator(hs, isite, 1, 0, 1)
    raise ValueError(f"pauli matrix of type {j} not supported")


def spin_half_system(n_sites: int):
    """Toolkit/spin_half.jl:9-19 -> (hilbert_space, pauli)."""
    from .hilbert import HilbertSpace, Site, State
    if n_sites > 64:
        raise ValueError(f"spin half system of {n_sites} sites cannot be expressed using UInt64")
    site = Site([State("Up", 1), State("Dn", -1)])
    hs = HilbertSpace([site for _ in range(n_sites)])
    return hs, (lambda isite, j: pauli_matrix(hs, isite, j))


# ---- term walk on bare bit strings (src/Operator/operator_iterator.jl:34-83); host-side, used for inspection ----
def get_row_iterator(op: Operator, brow: int):
    """[(bcol, amplitude)] of row `brow`, in term order: match (b & m) == r -> ((b & ~m) | c, a)."""
    return [((brow & ~m) | c, a) for (m, r, c, a) in op.terms if (brow & m) == r]


def get_column_iterator(op: Operator, bcol: int):
    """[(brow, amplitude)] of column `bcol`, in term order: match (b & m) == c -> ((b & ~m) | r, a)."""
    return [((bcol & ~m) | r, a) for (m, r, c, a) in op.terms if (bcol & m) == c]


def get_element(op: Operator, br: int, bc: int):
    """operator_iterator.jl:71-83."""
    out = 0
    for (m, r, c, a) in op.terms:
        if (br & m) == r and ((br & ~m) | c) == bc:
            out = out + a
    return out
