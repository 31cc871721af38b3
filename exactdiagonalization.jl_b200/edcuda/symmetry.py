"""Symmetry operations and symmetry-reduced representations backed by libedcuda.

Mirrors (reference, /root/reference/src):
  SitePermutation / DirectProductOperation (LatticeTools, consumed at symmetry_apply.jl:65-92)
  GlobalBitFlip                                   Symmetry/bitflipsymmetry.jl:8-35
  symmetry_apply(hs, op, bitrep)                  Symmetry/symmetry_apply.jl:65-92
  symmetry_reduce(hsr, symops_and_amplitudes)     Symmetry/symmetry_reduce_generic.jl:7-14
  ReducedHilbertSpaceRepresentation               Symmetry/reduced_hilbert_space_representation.jl:13-38
  symmetry_reduce(rhsr, vec) / symmetry_unreduce  Symmetry/symmetry_reduce.jl:29-153,208-225
  ReducedOperatorRepresentation                   Symmetry/reduced_operator_representation.jl:16-138
Site indices are 0-based: SitePermutation([1,2,3,0]) is the reference's SitePermutation([2,3,4,1]).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._lib import ED_C128, ED_F64, ED_SIDE_LEFT, ED_SIDE_RIGHT, DimensionMismatch, check, lib
from .operators import Operator
from .representation import (HilbertSpaceRepresentation, _AbstractOperatorRepresentation, _vec_info)


class SitePermutation:
    """map[i] = j : the state of site i moves to site j."""

    def __init__(self, mapping: Sequence[int]):
        self.map = [int(x) for x in mapping]
        if sorted(self.map) != list(range(len(self.map))):
            raise ValueError("not a permutation")

    def __mul__(self, rhs):
        if isinstance(rhs, SitePermutation):  # (A*B)(psi) = A(B(psi))
            return SitePermutation([self.map[rhs.map[i]] for i in range(len(self.map))])
        if isinstance(rhs, GlobalBitFlip):
            return DirectProductOperation([self, rhs])
        return NotImplemented

    def __eq__(self, other):
        return isinstance(other, SitePermutation) and self.map == other.map

    def inverse(self):
        inv = [0] * len(self.map)
        for i, j in enumerate(self.map):
            inv[j] = i
        return SitePermutation(inv)


class GlobalBitFlip:
    """bitflipsymmetry.jl:8-12."""

    def __init__(self, value: bool = False):
        self.value = bool(value)

    def __mul__(self, rhs):
        if isinstance(rhs, GlobalBitFlip):
            return GlobalBitFlip(self.value ^ rhs.value)
        if isinstance(rhs, SitePermutation):
            return DirectProductOperation([self, rhs])
        return NotImplemented

    def __eq__(self, other):
        return isinstance(other, GlobalBitFlip) and self.value == other.value


class DirectProductOperation:
    """Product of commuting operations, applied right to left (symmetry_apply.jl:65-77)."""

    def __init__(self, operations):
        self.operations = list(operations)

    def __mul__(self, rhs):
        ops = rhs.operations if isinstance(rhs, DirectProductOperation) else [rhs]
        return DirectProductOperation(self.operations + ops)


def _flatten(op, n_sites: int) -> Tuple[List[int], int]:
    """Any operation -> (site permutation map, flip flag); flips commute with site permutations."""
    if isinstance(op, SitePermutation):
        if len(op.map) != n_sites:
            raise ValueError("permutation length differs from the number of sites")
        return list(op.map), 0
    if isinstance(op, GlobalBitFlip):
        return list(range(n_sites)), int(op.value)
    if isinstance(op, DirectProductOperation):
        perm, flip = list(range(n_sites)), 0
        for o in reversed(op.operations):  # (ABC)(psi) = A(B(C(psi)))
            p, f = _flatten(o, n_sites)
            perm = [p[perm[i]] for i in range(n_sites)]
            flip ^= f
        return perm, flip
    raise TypeError(f"unsupported symmetry operation {op!r}")


class SymmetryHandle:
    def __init__(self, n_sites: int, symops_and_amplitudes):
        perms, flips, chis = [], [], []
        for op, chi in symops_and_amplitudes:
            p, f = _flatten(op, n_sites)
            perms.append(p)
            flips.append(f)
            chis.append([complex(chi).real, complex(chi).imag])
        self.n_ops = len(perms)
        perm = np.ascontiguousarray(np.array(perms, dtype=np.int32).reshape(self.n_ops, n_sites))
        flip = np.ascontiguousarray(np.array(flips, dtype=np.uint8))
        chi = np.ascontiguousarray(np.array(chis, dtype=np.float64))
        h = C.c_void_p()
        check(lib.ed_symmetry_create(self.n_ops, n_sites, perm.ctypes.data, flip.ctypes.data, chi.ctypes.data, C.byref(h)))
        self._handle = h

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and lib is not None:
            try:
                lib.ed_symmetry_destroy(h)
            except Exception:
                pass


def symmetry_apply(hs, symop, bitrep):
    """symmetry_apply(hs, op, bitrep) -> (image, sign)  (symmetry_apply.jl:82-92; sign is always 1).
    `bitrep` may be an int or an array of words; evaluated on the device."""
    base = hs.basespace()
    scalar = np.isscalar(bitrep)
    words = np.ascontiguousarray(np.atleast_1d(np.asarray(bitrep, dtype=np.uint64)))
    sh = SymmetryHandle(len(base.sites), [(SitePermutation(range(len(base.sites))), 1.0), (symop, 1.0)])
    out = np.empty_like(words)
    check(lib.ed_symmetry_apply(base.handle(), sh._handle, 1, words.ctypes.data, words.size, out.ctypes.data))
    if scalar:
        return int(out[0]), 1
    return out, 1


def symmetry_apply_operator(hs, symop, op: Operator) -> Operator:
    """symmetry_apply on operators (symmetry_apply.jl:56-63,96-106)."""
    if not op.terms:
        return op
    m, r, c, _ = op.arrays()
    allw = np.concatenate([m, r, c])
    img, _ = symmetry_apply(hs, symop, allw)
    n = len(op.terms)
    return Operator([(int(img[i]), int(img[n + i]), int(img[2 * n + i]), op.terms[i][3]) for i in range(n)])


def isinvariant(hs, symop, op: Operator) -> bool:
    """symmetry_apply.jl:110-116."""
    from .operators import simplify
    return len(simplify(op - symmetry_apply_operator(hs, symop, op)).terms) == 0


def isinvariant_all(hs, symops_and_amplitudes, op: Operator, tol: Optional[float] = None):
    """isinvariant for every element of a symmetry at once (ed_operator_isinvariant: matrix elements compared on sample
    words, host only, no GPU needed) -> (True, -1) or (False, index of the first violating element).
    represent(rhsr, op) runs the same check inside the library and raises ValueError (ArgumentError)."""
    base = hs.basespace()
    sym = SymmetryHandle(len(base.sites), list(symops_and_amplitudes))
    inv, bad = C.c_int32(), C.c_int32()
    check(lib.ed_operator_isinvariant(base.handle(), sym._handle, op.handle(), -1.0 if tol is None else float(tol), C.byref(inv), C.byref(bad)))
    return bool(inv.value), bad.value


class ReducedHilbertSpaceRepresentation:
    """reduced_hilbert_space_representation.jl:13-22.  `basis_mapping_index` / `basis_mapping_amplitude`
    (length = parent dimension in the reference) are computed on demand by the device."""

    def __init__(self, parent: HilbertSpaceRepresentation, handle, sym: SymmetryHandle):
        self.parent = parent
        self._handle = handle
        self._sym = sym
        d = C.c_int64()
        check(lib.ed_rbasis_dim(handle, C.byref(d)))
        self._dim = d.value
        self._basis_list = None
        self._map = None

    @property
    def dimension(self) -> int:
        return self._dim

    @property
    def bitwidth(self) -> int:
        return self.parent.bitwidth

    @property
    def basis_list(self) -> np.ndarray:
        if self._basis_list is None:
            out = np.empty(self._dim, dtype=np.uint64)
            check(lib.ed_rbasis_download(self._handle, 0, self._dim, out.ctypes.data))
            self._basis_list = out
        return self._basis_list

    def orbit_sizes(self) -> np.ndarray:
        out = np.empty(self._dim, dtype=np.int32)
        check(lib.ed_rbasis_orbit_sizes(self._handle, 0, self._dim, out.ctypes.data))
        return out

    def mapping(self, parent_words) -> Tuple[np.ndarray, np.ndarray]:
        w = np.ascontiguousarray(np.asarray(parent_words, dtype=np.uint64))
        idx = np.empty(w.size, dtype=np.int64)
        amp = np.empty(w.size, dtype=np.complex128)
        check(lib.ed_rbasis_mapping(self._handle, w.ctypes.data, w.size, idx.ctypes.data, amp.ctypes.data))
        return idx, amp

    def save(self, path: str):
        """checkpoint (ed_rbasis_save): representatives, orbit sizes, stabiliser marks; bound to the symmetry by a hash"""
        check(lib.ed_rbasis_save(self._handle, str(path).encode()))

    @classmethod
    def load(cls, hsr: HilbertSpaceRepresentation, symops_and_amplitudes, path: str, tol: Optional[float] = None):
        """skips the filter pass over the parent space; fails (ValueError) for a different space / symmetry / tolerance"""
        sym = SymmetryHandle(len(hsr.hilbert_space.sites), list(symops_and_amplitudes))
        h = C.c_void_p()
        check(lib.ed_rbasis_load(hsr._handle, sym._handle, -1.0 if tol is None else float(tol), str(path).encode(), C.byref(h)))
        return cls(hsr, h, sym)

    def _materialise_mapping(self):
        if self._map is None:
            n = self.parent.dimension
            idx = np.empty(n, dtype=np.int64)
            amp = np.empty(n, dtype=np.complex128)
            check(lib.ed_rbasis_mapping_rows(self._handle, 0, n, idx.ctypes.data, amp.ctypes.data))
            self._map = (idx, amp)
        return self._map

    @property
    def basis_mapping_index(self) -> np.ndarray:
        return self._materialise_mapping()[0]

    @property
    def basis_mapping_amplitude(self) -> np.ndarray:
        return self._materialise_mapping()[1]

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and lib is not None:
            try:
                lib.ed_rbasis_destroy(h)
            except Exception:
                pass


def symmetry_reduce(first, second, tol: Optional[float] = None):
    """symmetry_reduce(hsr, symops_and_amplitudes; tol) -> ReducedHilbertSpaceRepresentation
       symmetry_reduce(rhsr, large_vector)              -> small vector."""
    if isinstance(first, ReducedHilbertSpaceRepresentation):
        return symmetry_reduce_vector(first, second)
    hsr: HilbertSpaceRepresentation = first
    sym = SymmetryHandle(len(hsr.hilbert_space.sites), list(second))
    h = C.c_void_p()
    check(lib.ed_symmetry_reduce(hsr._handle, sym._handle, -1.0 if tol is None else float(tol), C.byref(h)))
    return ReducedHilbertSpaceRepresentation(hsr, h, sym)


symmetry_reduce_serial = symmetry_reduce
symmetry_reduce_parallel = symmetry_reduce


def symmetry_reduce_vector(rhsr: ReducedHilbertSpaceRepresentation, large_vector, out=None):
    """symmetry_reduce(rhsr, large) / symmetry_reduce!(out, rhsr, large) (adds when `out` is given)."""
    large = np.ascontiguousarray(large_vector)
    if large.dtype not in (np.float64, np.complex128):
        large = large.astype(np.complex128 if np.iscomplexobj(large) else np.float64)
    accumulate = out is not None
    if out is None:
        out = np.zeros(rhsr.dimension, dtype=np.complex128)
    elif out.dtype != np.complex128:
        raise TypeError("out must be complex128")
    code = ED_C128 if large.dtype == np.complex128 else ED_F64
    check(lib.ed_vector_reduce(rhsr._handle, out.ctypes.data, out.size, large.ctypes.data, large.size, code, int(accumulate)))
    return out


def symmetry_reduce_b(out, rhsr, large_vector):
    return symmetry_reduce_vector(rhsr, large_vector, out=out)


def symmetry_unreduce(rhsr: ReducedHilbertSpaceRepresentation, small_vector):
    """symmetry_unreduce(rhsr, small) (symmetry_reduce.jl:208-225)."""
    small = np.ascontiguousarray(small_vector)
    if small.dtype not in (np.float64, np.complex128):
        small = small.astype(np.complex128 if np.iscomplexobj(small) else np.float64)
    out = np.zeros(rhsr.parent.dimension, dtype=np.complex128)
    code = ED_C128 if small.dtype == np.complex128 else ED_F64
    check(lib.ed_vector_unreduce(rhsr._handle, out.ctypes.data, out.size, small.ctypes.data, small.size, code))
    return out


class ReducedOperatorRepresentation(_AbstractOperatorRepresentation):
    """reduced_operator_representation.jl:16-31 (always ComplexF64)."""

    def __init__(self, rhsr: ReducedHilbertSpaceRepresentation, op: Operator):
        self.reduced_hilbert_space_representation = rhsr
        self.operator = op
        h = C.c_void_p()
        check(lib.ed_oprep_create_reduced(rhsr._handle, op.handle(), C.byref(h)))
        self._handle = h

    @property
    def space(self):
        return self.reduced_hilbert_space_representation

    get_space = space
