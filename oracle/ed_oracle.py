"""CPU ORACLE -- test infrastructure only, never the product path.

A plain Python/numpy restatement of the Hamiltonian-application path of
ExactDiagonalization.jl v0.14.3 (reference tree: /root/reference).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import
this module; the product (`libedcuda.so` + the `edcuda` host package) never does.

Parity status: PINNED against the reference's own golden vectors (the literal
expectations in /root/reference/test/*.jl, transcribed in tests/golden/*.json with
file:line) -- the reference itself is Julia and cannot be executed in this image.
The LatticeTools.jl boundary (2-D lattice site numbering / irrep tables) is
"parity unpinned": the reference's tests hold no values for it.

Every function cites the reference file:line it follows.  Loops are kept in the
reference's order on purpose (term order, column order, group-element order) so
that floating-point association is the reference's.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

RTOL_DEFAULT = float(np.sqrt(np.finfo(np.float64).eps))  # Base.rtoldefault(Float64)
U64 = (1 << 64) - 1


# --------------------------------------------------------------------------- util
def make_bitmask(msb: int, lsb: int = 0) -> int:
    """src/util.jl:49-60."""
    return ((1 << msb) - 1) ^ ((1 << lsb) - 1)


def merge_vec(x: Sequence[int], y: Sequence[int]) -> List[int]:
    """src/util.jl:62-86 (equal heads are both kept)."""
    z: List[int] = []
    nx = ny = 0
    while nx < len(x) and ny < len(y):
        if x[nx] < y[ny]:
            z.append(x[nx]); nx += 1
        elif y[ny] < x[nx]:
            z.append(y[ny]); ny += 1
        else:
            z.append(x[nx]); z.append(y[ny]); nx += 1; ny += 1
    if nx < len(x):
        z.extend(x[nx:])
    elif ny < len(y):
        z.extend(y[ny:])
    return z


def choptol(d: Dict, tol: float) -> Dict:
    """src/util.jl:88-94: delete entries with abs(v) < tol (strict)."""
    for k in [k for k, v in d.items() if abs(v) < tol]:
        del d[k]
    return d


# --------------------------------------------------------------------------- Hilbert space
@dataclass(frozen=True)
class State:
    """src/HilbertSpace/site.jl:29-41."""
    name: str
    quantum_number: Tuple[int, ...]

    def __init__(self, name, qn):
        object.__setattr__(self, "name", name)
        object.__setattr__(self, "quantum_number", tuple(qn) if isinstance(qn, (tuple, list)) else (int(qn),))


class Site:
    """src/HilbertSpace/site.jl:69-72; bitwidth :93 = ceil(log2(n_states))."""

    def __init__(self, states: Sequence[State]):
        self.states = list(states)

    @property
    def bitwidth(self) -> int:
        return int(math.ceil(math.log2(len(self.states))))

    def __eq__(self, other):
        return self.states == other.states


class HilbertSpace:
    """src/HilbertSpace/hilbert_space.jl:25-41: site i occupies bits
    [bitoffsets[i], bitoffsets[i+1]); site 0 here (= Julia site 1) is the LSB."""

    def __init__(self, sites: Sequence[Site]):
        self.sites = list(sites)
        self.bitwidths = [s.bitwidth for s in self.sites]
        self.bitoffsets = [0]
        for w in self.bitwidths:
            self.bitoffsets.append(self.bitoffsets[-1] + w)

    @property
    def parent(self):
        return self

    @property
    def bitwidth(self) -> int:
        return self.bitoffsets[-1]

    @property
    def n_qn(self) -> int:
        return len(self.sites[0].states[0].quantum_number)

    def get_bitmask(self, isite: int | None = None) -> int:
        """hilbert_space.jl:107-113 (isite 0-based here)."""
        if isite is None:
            return make_bitmask(self.bitwidth)
        return make_bitmask(self.bitoffsets[isite + 1], self.bitoffsets[isite])

    def quantum_number_sectors(self) -> List[Tuple[int, ...]]:
        """hilbert_space.jl:119-129."""
        qns = {tuple([0] * self.n_qn)}
        for site in self.sites:
            nxt = set()
            for st in site.states:
                for q in qns:
                    nxt.add(tuple(a + b for a, b in zip(q, st.quantum_number)))
            qns = nxt
        return sorted(qns)

    def compress(self, indexarray: Sequence[int]) -> int:
        """hilbert_space.jl:199-214 (indexarray holds 0-based local state indices)."""
        b = 0
        for isite, idx in enumerate(indexarray):
            assert 0 <= idx < len(self.sites[isite].states)
            b |= idx << self.bitoffsets[isite]
        return b

    def get_state_index(self, b: int, isite: int) -> int:
        """hilbert_space.jl:249-251 (0-based result)."""
        return (b >> self.bitoffsets[isite]) & make_bitmask(self.bitwidths[isite])

    def get_quantum_number(self, b: int) -> Tuple[int, ...]:
        """hilbert_space.jl:135-145."""
        q = [0] * self.n_qn
        for isite, site in enumerate(self.sites):
            st = site.states[self.get_state_index(b, isite)]
            q = [a + c for a, c in zip(q, st.quantum_number)]
        return tuple(q)


class HilbertSpaceSector:
    """src/HilbertSpace/hilbert_space_sector.jl:13-64: allowed = requested ∩ reachable."""

    def __init__(self, parent: HilbertSpace, allowed=None):
        self.parent = parent
        sectors = set(parent.quantum_number_sectors())
        if allowed is None:
            req = sectors
        else:
            if isinstance(allowed, (int, np.integer)):
                allowed = [(int(allowed),)]
            elif isinstance(allowed, tuple) and all(isinstance(a, (int, np.integer)) for a in allowed):
                allowed = [tuple(allowed)]
            req = set()
            for a in allowed:
                req.add((int(a),) if isinstance(a, (int, np.integer)) else tuple(a))
        self.allowed_quantum_numbers = sectors & req

    @property
    def bitwidth(self):
        return self.parent.bitwidth


def basespace(hs):
    return hs.parent if isinstance(hs, HilbertSpaceSector) else hs


def hs_get_basis_list(hs, br_bits: int = 64) -> List[int]:
    """src/Representation/hilbert_space_representation.jl:109-119 (full space) and
    :127-206 (sector: per-site DP over partial quantum numbers, then merge)."""
    base = basespace(hs)
    if br_bits <= base.bitwidth:  # :110 / :129 (strict: MSB kept free)
        raise ValueError(f"type UInt{br_bits} not enough to represent the hilbert space (need {base.bitwidth} bits)")
    if isinstance(hs, HilbertSpace):
        out = []
        dims = [len(s.states) for s in hs.sites]
        idx = [0] * len(dims)
        total = 1
        for d in dims:
            total *= d
        for _ in range(total):
            out.append(hs.compress(idx))
            for k in range(len(dims)):
                idx[k] += 1
                if idx[k] < dims[k]:
                    break
                idx[k] = 0
        out.sort()
        return out
    hss = hs
    if not (hss.allowed_quantum_numbers & set(base.quantum_number_sectors())):
        return []
    qns = [[st.quantum_number for st in site.states] for site in base.sites]
    n_sites = len(base.sites)
    zero = tuple([0] * base.n_qn)
    add = lambda a, b: tuple(x + y for x, y in zip(a, b))
    sub = lambda a, b: tuple(x - y for x, y in zip(a, b))
    qn_possible = [None] * (n_sites + 1)
    qn_possible[0] = [zero]
    for i in range(n_sites):
        qn_possible[i + 1] = sorted({add(qa, qi) for qi in qns[i] for qa in qn_possible[i]})
    qn_requested = [None] * (n_sites + 1)
    qn_requested[n_sites] = sorted(hss.allowed_quantum_numbers)
    for i in range(n_sites - 1, -1, -1):
        qn_requested[i] = sorted({sub(qa, qi) for qi in qns[i] for qa in qn_requested[i + 1]})
    qn_schedule = [[q for q in x if q in set(y)] for x, y in zip(qn_requested, qn_possible)]
    sector = {zero: [0]}
    for i in range(n_sites):
        new = {}
        for q in qn_schedule[i + 1]:
            lst = []
            for i_state, q_curr in enumerate(qns[i]):
                q_prev = sub(q, q_curr)
                if q_prev in sector:
                    lst.extend(s | (i_state << base.bitoffsets[i]) for s in sector[q_prev])
            new[q] = lst
        sector = new
    basis: List[int] = []
    for q in list(sector.keys()):
        basis = merge_vec(basis, sector.pop(q))
    assert all(basis[i] < basis[i + 1] for i in range(len(basis) - 1))
    return basis


class HilbertSpaceRepresentation:
    """hilbert_space_representation.jl:16-73 with a FrozenSortedArrayIndex lookup
    (src/frozensortedarray.jl:11-48)."""

    def __init__(self, hs, basis_list: Sequence[int], br_bits: int = 64):
        base = basespace(hs)
        if br_bits <= base.bitwidth:
            raise ValueError("binary type too small")
        self.hilbert_space = base
        self.basis_list = np.asarray(list(basis_list), dtype=np.uint64)
        # FrozenSortedArrayIndex ctor: sorted + unique or ArgumentError (frozensortedarray.jl:14-22)
        if len(self.basis_list) > 1:
            if not np.all(self.basis_list[:-1] <= self.basis_list[1:]):
                raise ValueError("vals must be sorted")
            if np.any(self.basis_list[:-1] == self.basis_list[1:]):
                raise ValueError("vals contains duplicates")

    @property
    def dimension(self) -> int:
        return len(self.basis_list)

    def lookup(self, key: int, default: int = -1) -> int:
        """get(basis_lookup, key, default) -- frozensortedarray.jl:29-48. 1-based index."""
        idx = int(np.searchsorted(self.basis_list, np.uint64(key), side="left"))
        if idx < len(self.basis_list) and int(self.basis_list[idx]) == key:
            return idx + 1
        return default

    def lookup_array(self, keys: np.ndarray) -> np.ndarray:
        keys = np.asarray(keys, dtype=np.uint64)
        idx = np.searchsorted(self.basis_list, keys, side="left")
        ok = idx < len(self.basis_list)
        hit = np.zeros(len(keys), dtype=bool)
        hit[ok] = self.basis_list[idx[ok]] == keys[ok]
        return np.where(hit, idx + 1, -1).astype(np.int64)


def represent(hs, basis_list=None, br_bits: int = 64) -> HilbertSpaceRepresentation:
    """hilbert_space_representation.jl:215-256 (sorts a user list if unsorted, :251-253)."""
    if basis_list is None:
        return HilbertSpaceRepresentation(hs, hs_get_basis_list(hs, br_bits), br_bits)
    bl = list(int(b) for b in basis_list)
    if any(bl[i] > bl[i + 1] for i in range(len(bl) - 1)):
        bl = sorted(bl)
    return HilbertSpaceRepresentation(hs, bl, br_bits)


# --------------------------------------------------------------------------- operators
class NullOperator:
    """src/Operator/null_operator.jl:9."""
    terms: list = []

    def __eq__(self, other):
        return isinstance(other, NullOperator)


class PureOperator:
    """src/Operator/pure_operator.jl:25-52."""
    __slots__ = ("bitmask", "bitrow", "bitcol", "amplitude")

    def __init__(self, bitmask: int, bitrow: int, bitcol: int, amplitude):
        if (~bitmask) & bitrow & U64:
            raise ValueError("every bit of bitrow not in bitmask should be set to zero")
        if (~bitmask) & bitcol & U64:
            raise ValueError("every bit of bitcol not in bitmask should be set to zero")
        self.bitmask, self.bitrow, self.bitcol, self.amplitude = int(bitmask), int(bitrow), int(bitcol), amplitude

    @property
    def terms(self):
        return [self]

    def key(self):
        a = complex(self.amplitude)
        return (self.bitmask, self.bitrow, self.bitcol, a.real, a.imag)  # isless order, pure_operator.jl:78-90

    def __eq__(self, o):
        return isinstance(o, PureOperator) and self.key() == o.key()

    def __hash__(self):
        return hash(self.key())

    def __repr__(self):
        return f"PureOperator({self.bitmask:#b},{self.bitrow:#b},{self.bitcol:#b},{self.amplitude})"

    def __neg__(self):
        return PureOperator(self.bitmask, self.bitrow, self.bitcol, -self.amplitude)

    def __mul__(self, rhs):
        if isinstance(rhs, NullOperator):
            return rhs
        if isinstance(rhs, SumOperator):
            return SumOperator([t for t in (self * r for r in rhs.terms) if not isinstance(t, NullOperator)])
        if isinstance(rhs, PureOperator):
            return _pure_mul(self, rhs)
        return PureOperator(self.bitmask, self.bitrow, self.bitcol, self.amplitude * rhs)

    def __rmul__(self, lhs):
        return PureOperator(self.bitmask, self.bitrow, self.bitcol, lhs * self.amplitude)

    def __add__(self, rhs):
        return _op_add(self, rhs)

    def __radd__(self, lhs):
        return _op_add(lhs, self)

    def __sub__(self, rhs):
        return _op_add(self, -rhs)


def _pure_mul(lhs: PureOperator, rhs: PureOperator):
    """pure_operator.jl:136-154."""
    onlylhs = lhs.bitmask & ~rhs.bitmask
    onlyrhs = ~lhs.bitmask & rhs.bitmask
    inter = lhs.bitmask & rhs.bitmask
    if (lhs.bitcol & inter) != (rhs.bitrow & inter):
        return NullOperator()
    return PureOperator(lhs.bitmask | rhs.bitmask,
                        lhs.bitrow | (rhs.bitrow & onlyrhs),
                        rhs.bitcol | (lhs.bitcol & onlylhs),
                        lhs.amplitude * rhs.amplitude)


class SumOperator:
    """src/Operator/sum_operator.jl:13-23."""

    def __init__(self, terms: Sequence[PureOperator]):
        self.terms = list(terms)

    def __neg__(self):
        return SumOperator([-t for t in self.terms])

    def __mul__(self, rhs):
        if isinstance(rhs, NullOperator):
            return rhs
        if isinstance(rhs, PureOperator):
            return SumOperator([t for t in (l * rhs for l in self.terms) if not isinstance(t, NullOperator)])
        if isinstance(rhs, SumOperator):
            # sum_operator.jl:74-78: [tl*tr for tl in lhs, tr in rhs] then vec() = column-major: tl fastest
            out = []
            for tr in rhs.terms:
                for tl in self.terms:
                    p = tl * tr
                    if not isinstance(p, NullOperator):
                        out.append(p)
            return SumOperator(out)
        return SumOperator([t * rhs for t in self.terms])

    def __rmul__(self, lhs):
        return SumOperator([lhs * t for t in self.terms])

    def __add__(self, rhs):
        return _op_add(self, rhs)

    def __radd__(self, lhs):
        return _op_add(lhs, self)

    def __sub__(self, rhs):
        return _op_add(self, -rhs)

    def __eq__(self, o):
        return isinstance(o, SumOperator) and self.terms == o.terms


def _op_add(lhs, rhs):
    """sum_operator.jl:83-100 (terms concatenated in order; Null is the additive identity)."""
    if isinstance(lhs, (int, float, complex)) and lhs == 0:
        return rhs
    if isinstance(lhs, NullOperator):
        return rhs
    if isinstance(rhs, NullOperator):
        return lhs
    return SumOperator(list(lhs.terms) + list(rhs.terms))


def _isapprox0(x, tol):
    return abs(x) <= tol


def simplify(op, tol: float = RTOL_DEFAULT):
    """src/Operator/operator_simplify.jl:12-81: drop ~0, sort by isless, merge equal
    (mask,row,col), demote complex->real when all imaginary parts vanish."""
    if isinstance(op, NullOperator):
        return op
    if isinstance(op, PureOperator):
        a = op.amplitude
        if _isapprox0(a, tol):
            return NullOperator()
        if isinstance(a, complex) and _isapprox0(a.imag, tol):
            return PureOperator(op.bitmask, op.bitrow, op.bitcol, a.real)
        return op
    is_complex = any(isinstance(t.amplitude, complex) for t in op.terms)
    terms = []
    for t in op.terms:
        a = t.amplitude
        if _isapprox0(a, tol):
            continue  # simplify.(so.terms) drops ~0 terms (:33)
        if is_complex:
            a = complex(a.real, 0.0) if (isinstance(a, complex) and _isapprox0(a.imag, tol)) else complex(a)
        terms.append(PureOperator(t.bitmask, t.bitrow, t.bitcol, a))
    if not terms:
        return NullOperator()
    terms.sort(key=lambda t: t.key())
    new_terms = []
    bm, br, bc, am = terms[0].bitmask, terms[0].bitrow, terms[0].bitcol, terms[0].amplitude
    for t in terms[1:]:
        if (bm, br, bc) == (t.bitmask, t.bitrow, t.bitcol):
            am = am + t.amplitude
        else:
            if not _isapprox0(am, tol):
                new_terms.append(PureOperator(bm, br, bc, am))
            bm, br, bc, am = t.bitmask, t.bitrow, t.bitcol, t.amplitude
    if not _isapprox0(am, tol):
        new_terms.append(PureOperator(bm, br, bc, am))
    if not new_terms:
        return NullOperator()
    if is_complex and max(abs(complex(t.amplitude).imag) for t in new_terms) <= tol:
        new_terms = [PureOperator(t.bitmask, t.bitrow, t.bitcol, complex(t.amplitude).real) for t in new_terms]
    if len(new_terms) == 1:
        return new_terms[0]
    return SumOperator(new_terms)


def pure_operator(hs: HilbertSpace, isite: int, istate_row: int, istate_col: int, amplitude=1):
    """pure_operator.jl:184-201 (isite and local state indices 0-based here)."""
    assert 0 <= istate_row < len(hs.sites[isite].states) and 0 <= istate_col < len(hs.sites[isite].states)
    bm = hs.get_bitmask(isite)
    return PureOperator(bm, istate_row << hs.bitoffsets[isite], istate_col << hs.bitoffsets[isite], amplitude)


def pauli_matrix(hs: HilbertSpace, isite: int, j: str):
    """src/Toolkit/spin_half.jl:28-42 (Up = local state 0, Dn = local state 1; isite 0-based)."""
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
    raise ValueError(j)


def spin_half_system(n_sites: int):
    """src/Toolkit/spin_half.jl:9-19."""
    site = Site([State("Up", 1), State("Dn", -1)])
    hs = HilbertSpace([site] * n_sites)
    return hs, (lambda isite, j: pauli_matrix(hs, isite, j))


# --------------------------------------------------------------------------- term walk
def get_row_iterator(op, brow: int):
    """src/Operator/operator_iterator.jl:34-38, 48-54: in term order,
    match (b & m) == r  ->  ((b & ~m) | c, a)."""
    return [((brow & ~t.bitmask) | t.bitcol, t.amplitude) for t in op.terms if (brow & t.bitmask) == t.bitrow]


def get_column_iterator(op, bcol: int):
    """operator_iterator.jl:41-45, 57-63: match (b & m) == c -> ((b & ~m) | r, a)."""
    return [((bcol & ~t.bitmask) | t.bitrow, t.amplitude) for t in op.terms if (bcol & t.bitmask) == t.bitcol]


def get_element(op, br: int, bc: int):
    """operator_iterator.jl:66-83."""
    out = 0
    for t in op.terms:
        if (br & t.bitmask) == t.bitrow and ((br & ~t.bitmask) | t.bitcol) == bc:
            out = out + t.amplitude
    return out


def term_arrays(op):
    """Flatten an operator into the SoA layout handed to the engine."""
    terms = op.terms
    m = np.array([t.bitmask for t in terms], dtype=np.uint64)
    r = np.array([t.bitrow for t in terms], dtype=np.uint64)
    c = np.array([t.bitcol for t in terms], dtype=np.uint64)
    is_complex = any(isinstance(t.amplitude, complex) for t in terms)
    a = np.array([t.amplitude for t in terms], dtype=np.complex128 if is_complex else np.float64)
    return m, r, c, a


# --------------------------------------------------------------------------- operator representation
class OperatorRepresentation:
    """src/Representation/operator_representation.jl:13-26."""

    def __init__(self, hsr: HilbertSpaceRepresentation, op):
        self.hsr = hsr
        self.operator = op

    @property
    def space(self):
        return self.hsr

    @property
    def dimension(self):
        return self.hsr.dimension

    @property
    def is_complex(self):
        return any(isinstance(t.amplitude, complex) for t in self.operator.terms)

    def get_row_iterator(self, irow: int):
        """operator_representation.jl:66-79 (irow 1-based; misses give -1)."""
        if not (1 <= irow <= self.dimension):
            raise IndexError(irow)
        brow = int(self.hsr.basis_list[irow - 1])
        return [(self.hsr.lookup(bcol, -1), a) for bcol, a in get_row_iterator(self.operator, brow)]

    def get_column_iterator(self, icol: int):
        """operator_representation.jl:90-103."""
        if not (1 <= icol <= self.dimension):
            raise IndexError(icol)
        bcol = int(self.hsr.basis_list[icol - 1])
        return [(self.hsr.lookup(brow, -1), a) for brow, a in get_column_iterator(self.operator, bcol)]

    def get_element(self, irow: int, icol: int):
        """operator_representation.jl:109-119."""
        dim = self.dimension
        if irow <= 0 or irow > dim or icol <= 0 or icol > dim:
            raise IndexError((irow, icol))
        return get_element(self.operator, int(self.hsr.basis_list[irow - 1]), int(self.hsr.basis_list[icol - 1]))


def apply_serial(out: np.ndarray, opr, state: np.ndarray, side: str = "left"):
    """abstract_operator_representation.jl:296-316 (side='left': out += opr*state)
    and :327-347 (side='right': out += state*opr).  Adds, does not overwrite."""
    n = opr.dimension
    if len(out) != n:
        raise ValueError(f"DimensionMismatch: out has length {len(out)} != dimension {n}")
    if len(state) != n:
        raise ValueError(f"DimensionMismatch: state has length {len(state)} != dimension {n}")
    it = opr.get_row_iterator if side == "left" else opr.get_column_iterator
    for i in range(1, n + 1):
        for j, a in it(i):
            if 1 <= j <= n:
                out[i - 1] += a * state[j - 1]
    return out


def mul(out, opr, state):
    """abstract_operator_representation.jl:110-118: zero-fill then apply!."""
    out[:] = 0
    return apply_serial(out, opr, state, "left")


def apply_vectorized(out: np.ndarray, opr: OperatorRepresentation, state: np.ndarray, side: str = "left",
                     rows: slice | None = None):
    """Same arithmetic as apply_serial for a plain OperatorRepresentation, vectorised over
    rows with the term loop outermost-in-order, so each out[i] still accumulates its hits in
    `terms` order (operator_iterator.jl:52).  Used where the Python loop is too slow."""
    n = opr.dimension
    if len(out) != n or len(state) != n:
        raise ValueError("DimensionMismatch")
    basis = opr.hsr.basis_list
    sel = slice(0, n) if rows is None else rows
    b = basis[sel]
    o = out[sel]
    for t in opr.operator.terms:
        m, r, c = np.uint64(t.bitmask), np.uint64(t.bitrow), np.uint64(t.bitcol)
        if side == "right":
            r, c = c, r
        match = (b & m) == r
        if not match.any():
            continue
        tgt = (b[match] & ~m) | c
        idx = opr.hsr.lookup_array(tgt)
        ok = idx > 0
        rows_hit = np.nonzero(match)[0][ok]
        o[rows_hit] += t.amplitude * state[idx[ok] - 1]
    return out


def sparse_serial(opr, tol: float = RTOL_DEFAULT):
    """abstract_operator_representation.jl:145-169: per column, merge duplicates, chop
    |v| < tol, sort rows.  Returns (colptr, rowval, nzval) 1-based Int64 CSC like
    SparseMatrixCSC."""
    n = opr.dimension
    m = n
    colptr = np.zeros(n + 1, dtype=np.int64)
    rowval: List[int] = []
    nzval: List = []
    colptr[0] = 1
    for icol in range(1, n + 1):
        colvec: Dict[int, complex] = {}
        for irow, a in opr.get_column_iterator(icol):
            if 1 <= irow <= m:
                colvec[irow] = colvec.get(irow, 0) + a
        choptol(colvec, tol)
        colptr[icol] = colptr[icol - 1] + len(colvec)
        for irow in sorted(colvec):
            rowval.append(irow)
            nzval.append(colvec[irow])
    dtype = np.complex128 if getattr(opr, "is_complex", False) else np.float64
    return colptr, np.asarray(rowval, dtype=np.int64), np.asarray(nzval, dtype=dtype)


def dense_matrix(opr):
    """abstract_operator_representation.jl:121-132 (no chop)."""
    n = opr.dimension
    dtype = np.complex128 if getattr(opr, "is_complex", False) else np.float64
    out = np.zeros((n, n), dtype=dtype)
    for icol in range(1, n + 1):
        for irow, a in opr.get_column_iterator(icol):
            if 1 <= irow <= n:
                out[irow - 1, icol - 1] += a
    return out


def get_row(opr, irow):
    """abstract_operator_representation.jl:207-218 -> dict {icol: value} (chopped)."""
    items: Dict[int, complex] = {}
    for icol, v in opr.get_row_iterator(irow):
        if 1 <= icol <= opr.dimension:
            items[icol] = items.get(icol, 0) + v
    return choptol(items, RTOL_DEFAULT)


def get_column(opr, icol):
    """abstract_operator_representation.jl:221-232."""
    items: Dict[int, complex] = {}
    for irow, v in opr.get_column_iterator(icol):
        if 1 <= irow <= opr.dimension:
            items[irow] = items.get(irow, 0) + v
    return choptol(items, RTOL_DEFAULT)


# --------------------------------------------------------------------------- symmetry
class SitePermutation:
    """LatticeTools.SitePermutation stand-in: `map[i] = j` sends site i to site j (0-based)."""

    def __init__(self, mapping: Sequence[int]):
        self.map = list(int(x) for x in mapping)
        assert sorted(self.map) == list(range(len(self.map)))

    def __mul__(self, rhs):
        if isinstance(rhs, SitePermutation):  # (A*B)(psi) = A(B(psi))
            return SitePermutation([self.map[rhs.map[i]] for i in range(len(self.map))])
        return DirectProductOperation([self, rhs])


class GlobalBitFlip:
    """src/Symmetry/bitflipsymmetry.jl:8-12."""

    def __init__(self, value: bool = False):
        self.value = bool(value)

    def __mul__(self, rhs):
        if isinstance(rhs, GlobalBitFlip):
            return GlobalBitFlip(self.value ^ rhs.value)
        return DirectProductOperation([self, rhs])


class DirectProductOperation:
    """LatticeTools.DirectProductOperation stand-in (operations applied right-to-left)."""

    def __init__(self, operations):
        self.operations = list(operations)


def symmetry_apply(hs, symop, b: int) -> Tuple[int, int]:
    """src/Symmetry/symmetry_apply.jl:82-92 (permutation: field of site i moves to site
    map[i], sign 1), :65-77 (product: right-to-left), bitflipsymmetry.jl:23-35."""
    hs = basespace(hs)
    if isinstance(symop, SitePermutation):
        out = 0
        for i, j in enumerate(symop.map):
            out |= ((b >> hs.bitoffsets[i]) & make_bitmask(hs.bitwidths[i])) << hs.bitoffsets[j]
        return out, 1
    if isinstance(symop, GlobalBitFlip):
        if symop.value:
            return hs.get_bitmask() & ~b, 1
        return b, 1
    if isinstance(symop, DirectProductOperation):
        sign = 1
        for op in reversed(symop.operations):
            b, v = symmetry_apply(hs, op, b)
            sign *= v
        return b, sign
    raise TypeError(symop)


def symmetry_apply_operator(hs, symop, op):
    """symmetry_apply.jl:56-63, 96-106."""
    if isinstance(op, NullOperator):
        return op
    f = lambda x: symmetry_apply(hs, symop, x)[0]
    terms = [PureOperator(f(t.bitmask), f(t.bitrow), f(t.bitcol), t.amplitude) for t in op.terms]
    return terms[0] if isinstance(op, PureOperator) else SumOperator(terms)


def isinvariant(hs, symop, op) -> bool:
    """symmetry_apply.jl:110-116."""
    return isinstance(simplify(op - symmetry_apply_operator(hs, symop, op)), NullOperator)


class ReducedHilbertSpaceRepresentation:
    """src/Symmetry/reduced_hilbert_space_representation.jl:13-22."""

    def __init__(self, parent, basis_list, basis_mapping_index, basis_mapping_amplitude):
        self.parent = parent
        self.basis_list = np.asarray(basis_list, dtype=np.uint64)
        self.basis_mapping_index = np.asarray(basis_mapping_index, dtype=np.int64)  # 1-based / -1
        self.basis_mapping_amplitude = np.asarray(basis_mapping_amplitude, dtype=np.complex128)

    @property
    def dimension(self):
        return len(self.basis_list)


def symmetry_reduce(hsr: HilbertSpaceRepresentation, symops_and_amplitudes, tol: float = RTOL_DEFAULT):
    """src/Symmetry/symmetry_reduce_generic.jl:22-106 (serial).  Element 0 must be the
    identity (the loop starts at the second element, :56)."""
    if not all(abs(abs(y) - 1.0) <= RTOL_DEFAULT for _, y in symops_and_amplitudes):  # :27-29 isapprox(abs(y), 1)
        raise ValueError("all amplitudes need to have norm 1")
    hs = hsr.hilbert_space
    n_basis = hsr.dimension
    rep = np.full(n_basis, -1, dtype=np.int64)
    amp = np.zeros(n_basis, dtype=np.complex128)
    G = len(symops_and_amplitudes)
    reduced: List[int] = []
    visited = np.zeros(n_basis, dtype=bool)
    for ivec_p in range(n_basis):
        if visited[ivec_p]:
            continue
        bvec = int(hsr.basis_list[ivec_p])
        states = [bvec] * G
        phases = [1.0 + 0j] * G
        compatible = True
        for i in range(1, G):
            symop, ampl = symops_and_amplitudes[i]
            bp, sgn = symmetry_apply(hs, symop, bvec)
            if bp < bvec:
                compatible = False
                break
            if bp == bvec and not (abs(ampl * sgn - 1.0) <= tol):
                compatible = False
                break
            states[i] = bp
            phases[i] = np.conj(ampl * sgn)
        if not compatible:
            continue
        reduced.append(bvec)
        amps: Dict[int, complex] = {}
        for i in range(G):
            amps[states[i]] = phases[i]  # later duplicates overwrite (:74-78)
        inv_norm = 1.0 / math.sqrt(float(len(amps)))
        for bp, a in amps.items():
            ip = hsr.lookup(bp, 0)
            if ip <= 0:
                raise KeyError(bp)  # basis_lookup[bvec_prime] throws (:81)
            visited[ip - 1] = True
            rep[ip - 1] = ivec_p + 1
            amp[ip - 1] = a * inv_norm
    index = np.full(n_basis, -1, dtype=np.int64)
    for ivec_r, bvec in enumerate(reduced):
        index[hsr.lookup(bvec) - 1] = ivec_r + 1
    for ipp in range(n_basis):
        ip = rep[ipp]
        if ip <= 0 or ipp + 1 == ip:
            continue
        index[ipp] = index[ip - 1]
    return ReducedHilbertSpaceRepresentation(hsr, reduced, index, amp)


def symmetry_reduce_vector(rhsr, large_vector):
    """src/Symmetry/symmetry_reduce.jl:38-56."""
    if len(large_vector) != rhsr.parent.dimension:
        raise ValueError("DimensionMismatch")
    small = np.zeros(rhsr.dimension, dtype=np.complex128)
    for ip, ir in enumerate(rhsr.basis_mapping_index):
        if ir > 0:
            small[ir - 1] += np.conj(rhsr.basis_mapping_amplitude[ip]) * large_vector[ip]
    return small


def symmetry_unreduce_vector(rhsr, small_vector):
    """src/Symmetry/symmetry_reduce.jl:208-225."""
    if len(small_vector) != rhsr.dimension:
        raise ValueError("DimensionMismatch")
    large = np.zeros(rhsr.parent.dimension, dtype=np.complex128)
    for ip, ir in enumerate(rhsr.basis_mapping_index):
        if ir > 0:
            large[ip] += rhsr.basis_mapping_amplitude[ip] * small_vector[ir - 1]
    return large


class ReducedOperatorRepresentation:
    """src/Symmetry/reduced_operator_representation.jl:16-31 (scalar type always complex, :26)."""
    is_complex = True

    def __init__(self, rhsr: ReducedHilbertSpaceRepresentation, op):
        self.rhsr = rhsr
        self.operator = op

    @property
    def space(self):
        return self.rhsr

    @property
    def dimension(self):
        return self.rhsr.dimension

    def get_row_iterator(self, irow_r: int):
        """reduced_operator_representation.jl:57-85."""
        rhsr, hsr = self.rhsr, self.rhsr.parent
        if not (1 <= irow_r <= rhsr.dimension):
            raise IndexError(irow_r)
        brow = int(rhsr.basis_list[irow_r - 1])
        irow_p = hsr.lookup(brow, 0)
        inv_ampl_row = 1.0 / rhsr.basis_mapping_amplitude[irow_p - 1]
        out = []
        for bcol, ampl in get_row_iterator(self.operator, brow):
            ampl = complex(ampl)
            icol_p = hsr.lookup(bcol, -1)
            if icol_p <= 0:
                out.append((-1, ampl)); continue
            icol_r = int(rhsr.basis_mapping_index[icol_p - 1])
            if icol_r <= 0:
                out.append((-1, ampl)); continue
            out.append((icol_r, ampl * rhsr.basis_mapping_amplitude[icol_p - 1] * inv_ampl_row))
        return out

    def get_column_iterator(self, icol_r: int):
        """reduced_operator_representation.jl:88-116."""
        rhsr, hsr = self.rhsr, self.rhsr.parent
        if not (1 <= icol_r <= rhsr.dimension):
            raise IndexError(icol_r)
        bcol = int(rhsr.basis_list[icol_r - 1])
        icol_p = hsr.lookup(bcol, 0)
        inv_ampl_col = 1.0 / np.conj(rhsr.basis_mapping_amplitude[icol_p - 1])
        out = []
        for brow, ampl in get_column_iterator(self.operator, bcol):
            ampl = complex(ampl)
            irow_p = hsr.lookup(brow, -1)
            if irow_p <= 0:
                out.append((-1, ampl)); continue
            irow_r = int(rhsr.basis_mapping_index[irow_p - 1])
            if irow_r <= 0:
                out.append((-1, ampl)); continue
            out.append((irow_r, ampl * np.conj(rhsr.basis_mapping_amplitude[irow_p - 1]) * inv_ampl_col))
        return out

    def get_element(self, irow_r, icol_r):
        """reduced_operator_representation.jl:120-138."""
        dim = self.dimension
        if irow_r <= 0 or irow_r > dim or icol_r <= 0 or icol_r > dim:
            raise IndexError((irow_r, icol_r))
        return sum((a for i, a in self.get_column_iterator(icol_r) if i == irow_r), 0j)


def represent_operator(space, op):
    """operator_representation.jl:34-36 / reduced_operator_representation.jl:40-42."""
    if isinstance(space, ReducedHilbertSpaceRepresentation):
        return ReducedOperatorRepresentation(space, op)
    return OperatorRepresentation(space, op)


# --------------------------------------------------------------------------- on-the-fly reduced path
def reduced_representatives_on_the_fly(hs, words: Iterable[int], symops_and_amplitudes, tol=RTOL_DEFAULT):
    """Representative filter of symmetry_reduce_generic.jl:51-72 applied to a stream of
    parent words (no visited[] shortcut, no parent lookup): a word is kept iff it is its
    orbit's minimum and every stabilising element has chi*sgn ~ 1.  Returns (reps, orbit sizes)."""
    reps, sizes = [], []
    G = len(symops_and_amplitudes)
    for bvec in words:
        ok = True
        images = {bvec}
        for i in range(1, G):
            symop, ampl = symops_and_amplitudes[i]
            bp, sgn = symmetry_apply(hs, symop, bvec)
            if bp < bvec or (bp == bvec and not (abs(ampl * sgn - 1.0) <= tol)):
                ok = False
                break
            images.add(bp)
        if ok:
            reps.append(bvec)
            sizes.append(len(images))
    return reps, sizes


# --------------------------------------------------------------------------- Lanczos (not in the reference)
def lanczos(matvec, v0: np.ndarray, n_steps: int):
    """Plain three-term Lanczos used as the yardstick for the engine's driver (the reference
    delegates to Arpack, docs/src/examples/spinhalf.md:26).  Returns (alpha, beta)."""
    v = v0 / np.linalg.norm(v0)
    v_prev = np.zeros_like(v)
    beta_prev = 0.0
    alphas, betas = [], []
    for _ in range(n_steps):
        w = matvec(v)
        a = np.vdot(v, w).real
        w = w - a * v - beta_prev * v_prev
        b = float(np.linalg.norm(w))
        alphas.append(a); betas.append(b)
        if b == 0.0:
            break
        v_prev, v, beta_prev = v, w / b, b
    return np.array(alphas), np.array(betas)


def tridiag_eigvals(alpha, beta):
    from scipy.linalg import eigvalsh_tridiagonal
    k = len(alpha)
    return eigvalsh_tridiagonal(np.asarray(alpha), np.asarray(beta[: k - 1]))
