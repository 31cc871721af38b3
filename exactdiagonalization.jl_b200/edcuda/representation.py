"""HilbertSpaceRepresentation / OperatorRepresentation backed by libedcuda.

Mirrors the reference's method surface (same names, argument meaning and error behaviour):
  represent(hs [,BR]) / represent(hs, basis_list)      Representation/hilbert_space_representation.jl:215-256
  HilbertSpaceRepresentation: .hilbert_space .basis_list .basis_lookup, dimension     :16-90
  represent(hsr, op) -> OperatorRepresentation          Representation/operator_representation.jl:34-36
  apply!(out, opr, x), apply!(out, x, opr), mul!(out, opr, x), opr*x, x*opr            abstract_operator_representation.jl:110-118,260-409
  sparse(opr; tol), Matrix(opr), get_row, get_column, get_element, iterators           :121-249
Python has no `!` in identifiers: apply!/mul! are `apply_b`/`mul_b` (b for "bang").
All compute goes through the C ABI; vectors may be numpy arrays (host) or torch CUDA tensors (device).
Indices are 1-based exactly as in the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ED_C128, ED_F64, ED_SIDE_LEFT, ED_SIDE_RIGHT, check, lib
from .hilbert import HilbertSpace, HilbertSpaceSector
from .operators import Operator

_BR_BITS = {np.uint8: 8, np.uint16: 16, np.uint32: 32, np.uint64: 64, "UInt8": 8, "UInt16": 16, "UInt32": 32,
            "UInt64": 64, "UInt128": 128, None: 64, int: 64}


def _br_bits(binary_type) -> int:
    if isinstance(binary_type, (int, np.integer)) and not isinstance(binary_type, bool):
        return int(binary_type)
    if binary_type in _BR_BITS:
        return _BR_BITS[binary_type]
    return np.dtype(binary_type).itemsize * 8


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _vec_info(v):
    """-> (pointer, length, dtype code, keepalive)."""
    if _is_torch(v):
        import torch
        if not v.is_contiguous():
            raise ValueError("torch vectors must be contiguous")
        if v.dtype == torch.float64:
            code = ED_F64
        elif v.dtype == torch.complex128:
            code = ED_C128
        else:
            raise TypeError("vectors must be float64 or complex128")
        return v.data_ptr(), v.numel(), code, v
    if not isinstance(v, np.ndarray):
        raise TypeError("vectors must be numpy arrays or torch tensors")
    if v.dtype == np.float64:
        code = ED_F64
    elif v.dtype == np.complex128:
        code = ED_C128
    else:
        raise TypeError("vectors must be float64 or complex128")
    if not v.flags["C_CONTIGUOUS"]:
        raise ValueError("numpy vectors must be contiguous")
    return v.ctypes.data, v.size, code, v


class BasisLookup:
    """Stand-in for hsr.basis_lookup (FrozenSortedArrayIndex, src/frozensortedarray.jl:11-48)."""

    def __init__(self, hsr: "HilbertSpaceRepresentation"):
        self._hsr = hsr

    def get(self, key: int, default: int = -1) -> int:
        idx = self._hsr.lookup([key])[0]
        return int(idx) if idx > 0 else default

    def __getitem__(self, key: int) -> int:
        idx = self.get(key, 0)
        if idx <= 0:
            raise KeyError(f"key {key} not found")
        return idx

    def __contains__(self, key: int) -> bool:
        return self.get(key, 0) > 0

    def __len__(self):
        return self._hsr.dimension


class HilbertSpaceRepresentation:
    """hilbert_space_representation.jl:16-73.  `basis_list` is fetched from the device on first use."""

    def __init__(self, hilbert_space, handle, br_bits: int = 64):
        self.hilbert_space: HilbertSpace = hilbert_space.basespace()
        self._handle = handle
        self.br_bits = br_bits
        d = C.c_int64()
        check(lib.ed_basis_dim(handle, C.byref(d)))
        self._dim = d.value
        self._basis_list: Optional[np.ndarray] = None
        self.basis_lookup = BasisLookup(self)

    @property
    def dimension(self) -> int:
        return self._dim

    @property
    def bitwidth(self) -> int:
        return self.hilbert_space.bitwidth

    @property
    def kind(self) -> int:
        k = C.c_int32()
        check(lib.ed_basis_kind(self._handle, C.byref(k)))
        return k.value

    @property
    def basis_list(self) -> np.ndarray:
        if self._basis_list is None:
            self._basis_list = self.download(0, self._dim)
        return self._basis_list

    def download(self, lo: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        check(lib.ed_basis_download(self._handle, lo, n, out.ctypes.data))
        return out

    def lookup(self, keys: Sequence[int]) -> np.ndarray:
        """get(basis_lookup, key, -1) vectorised: 1-based indices, -1 for misses."""
        k = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64))
        out = np.empty(k.size, dtype=np.int64)
        check(lib.ed_basis_lookup(self._handle, k.ctypes.data, k.size, out.ctypes.data))
        return out

    def save(self, path: str):
        """checkpoint (ed_basis_save): sector bases store how they were generated, user lists their words"""
        check(lib.ed_basis_save(self._handle, str(path).encode()))

    @classmethod
    def load(cls, hilbert_space, path: str):
        base = hilbert_space.basespace()
        h = C.c_void_p()
        check(lib.ed_basis_load(base.handle(), str(path).encode(), C.byref(h)))
        return cls(base, h)

    def __eq__(self, other):
        return (isinstance(other, HilbertSpaceRepresentation) and self.hilbert_space == other.hilbert_space
                and self._dim == other._dim and np.array_equal(self.basis_list, other.basis_list))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and lib is not None:
            try:
                lib.ed_basis_destroy(h)
            except Exception:
                pass


def represent(space, second=None, binary_type=None):
    """represent(hs [,BR]) | represent(hs, basis_list) | represent(hsr, op) | represent(rhsr, op)."""
    from .symmetry import ReducedHilbertSpaceRepresentation, ReducedOperatorRepresentation
    if isinstance(space, HilbertSpaceRepresentation):
        return OperatorRepresentation(space, second)
    if isinstance(space, ReducedHilbertSpaceRepresentation):
        return ReducedOperatorRepresentation(space, second)
    if second is not None and not isinstance(second, (type, str)) and np.ndim(second) == 1:
        return represent_array(space, second, binary_type)
    if second is not None and binary_type is None:
        binary_type = second
    return represent_array(space, None, binary_type)


def represent_array(space, basis_list=None, binary_type=None) -> HilbertSpaceRepresentation:
    """hilbert_space_representation.jl:226-256."""
    base = space.basespace()
    h = C.c_void_p()
    if basis_list is not None:
        arr = np.asarray(basis_list)
        bits = _br_bits(binary_type) if binary_type is not None else (arr.dtype.itemsize * 8 if arr.dtype.kind == "u" else 64)
        words = np.ascontiguousarray(arr.astype(np.uint64))
        check(lib.ed_basis_from_list(base.handle(), words.ctypes.data if words.size else None, words.size, bits, C.byref(h)))
        return HilbertSpaceRepresentation(base, h, bits)
    bits = _br_bits(binary_type)
    if isinstance(space, HilbertSpaceSector):
        allowed = np.array(sorted(space.allowed_quantum_numbers), dtype=np.int64).reshape(-1, max(base.n_qn, 1))
        if base.n_qn == 0:
            allowed = np.zeros((len(space.allowed_quantum_numbers), 0), dtype=np.int64)
        allowed = np.ascontiguousarray(allowed)
        check(lib.ed_basis_generate(base.handle(), allowed.ctypes.data if allowed.size else None, allowed.shape[0], bits, C.byref(h)))
    else:
        check(lib.ed_basis_generate(base.handle(), None, -1, bits, C.byref(h)))
    return HilbertSpaceRepresentation(base, h, bits)


represent_dict = represent_array  # the Dict-backed variant (:265-291) has the same observable behaviour


class _AbstractOperatorRepresentation:
    """Shared behaviour of plain and reduced representations (abstract_operator_representation.jl)."""

    _handle = None
    operator: Operator
    __array_ufunc__ = None   # numpy defers to __rmul__: `state * opr` works like the reference's Base.:(*)(state, opr)

    # -- traits
    @property
    def dimension(self) -> int:
        d = C.c_int64()
        check(lib.ed_oprep_dim(self._handle, C.byref(d)))
        return d.value

    @property
    def shape(self):
        d = self.dimension
        return (d, d)

    def size(self, i: Optional[int] = None):
        if i is None:
            return self.shape
        if i not in (1, 2):
            raise IndexError(i)
        return self.shape[i - 1]

    @property
    def is_complex(self) -> bool:
        t = C.c_int32()
        check(lib.ed_oprep_dtype(self._handle, C.byref(t)))
        return t.value == ED_C128

    @property
    def dtype(self):
        return np.complex128 if self.is_complex else np.float64

    def set_rows(self, lo: int, hi: int):
        """Row-shard this representation: this process owns output rows [lo, hi) (0-based)."""
        check(lib.ed_oprep_set_rows(self._handle, lo, hi))
        self._rows = (lo, hi)
        return self

    def cache_matrix(self, side: int = ED_SIDE_LEFT) -> int:
        """Assemble the owned rows once and keep them on device: later apply!/mul!/Lanczos calls become an SpMV.
        Returns the number of stored entries."""
        nnz = C.c_int64()
        check(lib.ed_oprep_cache_matrix(self._handle, side, C.byref(nnz)))
        return nnz.value

    def drop_cache(self):
        check(lib.ed_oprep_drop_cache(self._handle))
        return self

    def set_kernel(self, which: int):
        check(lib.ed_oprep_set_kernel(self._handle, which))
        return self

    # -- apply!/mul!
    def _apply(self, out, x, side: int, accumulate: int):
        po, no, co, _ko = _vec_info(out)
        px, nx, cx, _kx = _vec_info(x)
        if co != cx:
            raise TypeError("out and state must have the same element type")
        if _is_torch(out) != _is_torch(x):
            raise TypeError("out and state must both be host (numpy) or both be device (torch) vectors")
        if _is_torch(out):
            import torch
            check(lib.ed_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream), 1))
        try:
            check(lib.ed_apply(self._handle, po, no, px, nx, co, side, accumulate))
        finally:
            if _is_torch(out):
                lib.ed_set_stream(None, 0)
        return out

    def apply_b(self, out, a, b=None):
        raise NotImplementedError

    def __matmul__(self, x):
        return self.__mul__(x)

    def __mul__(self, x):
        """opr * state (operator_representation.jl:122-130)."""
        if isinstance(x, (int, float, complex)):
            from .operators import simplify
            return type(self)(self.space, simplify(self.operator * x))
        x = np.asarray(x)
        t = np.complex128 if (self.is_complex or np.iscomplexobj(x)) else np.float64
        out = np.zeros(self.dimension, dtype=t)
        return self._apply(out, np.ascontiguousarray(x.astype(t)), ED_SIDE_LEFT, 1)

    def __rmul__(self, x):
        """state * opr (operator_representation.jl:132-139)."""
        x = np.asarray(x)
        t = np.complex128 if (self.is_complex or np.iscomplexobj(x)) else np.float64
        out = np.zeros(self.dimension, dtype=t)
        return self._apply(out, np.ascontiguousarray(x.astype(t)), ED_SIDE_RIGHT, 1)

    # -- iterators / elements
    def _iterator(self, i: int, side: int):
        cap = max(len(self.operator.terms), 1)
        idx = np.empty(cap, dtype=np.int64)
        amp = np.empty(cap * 2, dtype=np.float64)
        n = C.c_int64()
        check(lib.ed_oprep_row_iterator(self._handle, int(i), side, cap, idx.ctypes.data, amp.ctypes.data, C.byref(n)))
        k = n.value
        if self.is_complex:
            vals = amp[: 2 * k].view(np.complex128)
        else:
            vals = amp[:k]
        return [(int(idx[j]), vals[j].item()) for j in range(k)]

    def get_row_iterator(self, irow: int):
        return self._iterator(irow, ED_SIDE_LEFT)

    def get_column_iterator(self, icol: int):
        return self._iterator(icol, ED_SIDE_RIGHT)

    def get_element(self, irow: int, icol: int):
        v = np.zeros(2, dtype=np.float64)
        check(lib.ed_oprep_get_element(self._handle, int(irow), int(icol), v.ctypes.data))
        return complex(v[0], v[1]) if self.is_complex else float(v[0])

    def _get_line(self, i: int, side: int):
        """get_row / get_column (abstract_operator_representation.jl:207-232): merged, chopped sparse vector as dict."""
        from .operators import RTOL
        items = {}
        for j, v in self._iterator(i, side):
            if 1 <= j <= self.dimension:
                items[j] = items.get(j, 0) + v
        return {j: v for j, v in items.items() if not abs(v) < RTOL}

    def get_row(self, irow: int):
        return self._get_line(irow, ED_SIDE_LEFT)

    def get_column(self, icol: int):
        return self._get_line(icol, ED_SIDE_RIGHT)

    def __getitem__(self, key):
        i, j = key
        full = slice(None)
        if i == full and j == full:
            return self.sparse()
        if i == full:
            return self.get_column(j)
        if j == full:
            return self.get_row(i)
        return self.get_element(i, j)

    # -- sparse / dense
    def sparse_csc(self, tol: Optional[float] = None):
        """sparse(opr; tol) -> (colptr, rowval, nzval), 1-based Int64 like SparseMatrixCSC."""
        nnz = C.c_int64()
        check(lib.ed_sparse_count(self._handle, -1.0 if tol is None else float(tol), C.byref(nnz)))
        dim = self.dimension
        colptr = np.empty(dim + 1, dtype=np.int64)
        rowval = np.empty(nnz.value, dtype=np.int64)
        nzval = np.empty(nnz.value, dtype=self.dtype)
        check(lib.ed_sparse_fetch(self._handle, colptr.ctypes.data, rowval.ctypes.data if nnz.value else None,
                                  nzval.ctypes.data if nnz.value else None))
        return colptr, rowval, nzval

    def sparse(self, tol: Optional[float] = None):
        import scipy.sparse as sp
        colptr, rowval, nzval = self.sparse_csc(tol)
        d = self.dimension
        return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(d, d))

    sparse_serial = sparse
    sparse_parallel = sparse

    def matrix(self) -> np.ndarray:
        """Matrix(opr) (abstract_operator_representation.jl:121-132)."""
        d = self.dimension
        out = np.zeros((d, d), dtype=self.dtype, order="F")
        check(lib.ed_dense(self._handle, out.ctypes.data))
        return out

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and lib is not None:
            try:
                lib.ed_oprep_destroy(h)
            except Exception:
                pass


class OperatorRepresentation(_AbstractOperatorRepresentation):
    """operator_representation.jl:13-26."""

    def __init__(self, hsr: HilbertSpaceRepresentation, op: Operator):
        self.hilbert_space_representation = hsr
        self.operator = op
        h = C.c_void_p()
        check(lib.ed_oprep_create(hsr._handle, op.handle(), C.byref(h)))
        self._handle = h

    @property
    def space(self):
        return self.hilbert_space_representation

    get_space = space


def apply_b(out, a, b):
    """apply!(out, opr, state) [out += opr*state]  or  apply!(out, state, opr) [out += state*opr]."""
    if isinstance(a, _AbstractOperatorRepresentation):
        return a._apply(out, b, ED_SIDE_LEFT, 1)
    if isinstance(b, _AbstractOperatorRepresentation):
        return b._apply(out, a, ED_SIDE_RIGHT, 1)
    raise TypeError("one of the arguments must be an operator representation")


apply_serial_b = apply_b
apply_parallel_b = apply_b


def mul_b(out, opr, state):
    """LinearAlgebra.mul!(out, opr, state): out = opr*state (abstract_operator_representation.jl:110-118)."""
    return opr._apply(out, state, ED_SIDE_LEFT, 0)


def sparse(opr, tol: Optional[float] = None):
    return opr.sparse(tol)


def dimension(x) -> int:
    return x.dimension
