"""CPU ORACLE, vectorised twin -- TEST INFRASTRUCTURE ONLY (tests/ may import it; the product never does).

numpy restatement of the loops of oracle/ed_oracle.py for spaces of 1-bit sites, so that the reference's algorithm can be
followed at sizes where the pure-Python loops take minutes (all 16 momentum sectors of the 4x4 square lattice, the
12,870-column CSC of the L=16 chain).  Every function names the reference lines (ExactDiagonalization.jl v0.14.3,
/root/reference/src) it restates and is pinned in tests/test_oracle_np.py against the loop oracle, which is itself pinned
to the reference's golden vectors.  Parity status: pinned (transitively).

The arithmetic is the same; only the loop nest is turned inside out (terms / group elements outermost, states as array
lanes).  Where the reference's result depends on iteration ORDER the order is kept:
  * "later duplicates overwrite" of basis_mapping_amplitude (symmetry_reduce_generic.jl:74-78): elements are visited in
    ascending order and every element writes all of its images at once (one element never hits a word twice);
  * per-column accumulation of sparse() in term order (abstract_operator_representation.jl:150-160): the duplicate merge
    adds the hits of a (row, column) pair in term order (np.add.at over term-major triplets).
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

RTOL_DEFAULT = math.sqrt(np.finfo(np.float64).eps)


def lookup(basis: np.ndarray, keys: np.ndarray) -> np.ndarray:
    """frozensortedarray.jl:29-48, vectorised: 1-based index or -1."""
    pos = np.searchsorted(basis, keys)
    pos_c = np.minimum(pos, len(basis) - 1)
    hit = (pos < len(basis)) & (basis[pos_c] == keys)
    return np.where(hit, pos + 1, -1).astype(np.int64)


def perm_apply(perm: Sequence[int], words: np.ndarray) -> np.ndarray:
    """symmetry_apply.jl:82-92 for 1-bit sites: bit i of the word moves to bit perm[i]."""
    out = np.zeros_like(words)
    one = np.uint64(1)
    for i, j in enumerate(perm):
        out |= ((words >> np.uint64(i)) & one) << np.uint64(j)
    return out


def symmetry_reduce(basis: np.ndarray, perms: Sequence[Sequence[int]], chis: Sequence[complex], tol: float = RTOL_DEFAULT):
    """symmetry_reduce_generic.jl:22-106.  basis ascending uint64; perms[0] must be the identity.
    Returns (reduced basis, basis_mapping_index [1-based / -1], basis_mapping_amplitude)."""
    chis = np.asarray(chis, dtype=np.complex128)
    if not np.all(np.abs(np.abs(chis) - 1.0) <= RTOL_DEFAULT):
        raise ValueError("all amplitudes need to have norm 1")
    n = len(basis)
    G = len(perms)
    images = np.empty((G, n), dtype=np.uint64)
    images[0] = basis
    keep = np.ones(n, dtype=bool)
    for g in range(1, G):
        images[g] = perm_apply(perms[g], basis)
        keep &= images[g] >= basis                                                  # :59-62 bp < bvec -> incompatible
        keep &= ~((images[g] == basis) & ~(np.abs(chis[g] - 1.0) <= tol))          # :63-66 stabiliser character
    reps = np.nonzero(keep)[0]
    reduced = basis[reps]
    # orbit size = number of distinct images (length of the Dict, :74-79)
    img_r = images[:, reps]
    srt = np.sort(img_r, axis=0)
    n_orbit = 1 + np.count_nonzero(srt[1:] != srt[:-1], axis=0)
    inv_norm = 1.0 / np.sqrt(n_orbit.astype(np.float64))
    index = np.full(n, -1, dtype=np.int64)
    amp = np.zeros(n, dtype=np.complex128)
    for g in range(G):                                                              # ascending: later duplicates overwrite
        ip = lookup(basis, img_r[g])
        if np.any(ip <= 0):
            raise KeyError("image outside the basis")                              # :81 basis_lookup[bvec_prime] throws
        phase = 1.0 + 0j if g == 0 else np.conj(chis[g])
        index[ip - 1] = np.arange(1, len(reps) + 1)
        amp[ip - 1] = phase * inv_norm
    return reduced, index, amp


def _walk(basis_words: np.ndarray, terms, column: bool):
    """operator_iterator.jl:48-63 for every word at once: yields (t, lanes that match, emitted words) in term order."""
    m, r, c, a = terms
    for t in range(len(m)):
        mt, match, emit = np.uint64(m[t]), np.uint64(c[t] if column else r[t]), np.uint64(r[t] if column else c[t])
        sel = np.nonzero((basis_words & mt) == match)[0]
        if sel.size:
            yield t, sel, (basis_words[sel] & ~mt) | emit


def _assemble_csc(n: int, cols: np.ndarray, rows: np.ndarray, vals: np.ndarray, tol: float, dtype):
    """abstract_operator_representation.jl:145-169 after the walk: merge duplicates (in the order given), chop |v| < tol
    (util.jl:88-94, strict), rows ascending per column; 1-based Int64 colptr/rowval like SparseMatrixCSC."""
    key = cols.astype(np.int64) * np.int64(n) + rows.astype(np.int64)
    uniq, inv = np.unique(key, return_inverse=True)
    acc = np.zeros(len(uniq), dtype=dtype)
    np.add.at(acc, inv, vals)                       # sequential: hits are added in the order they were emitted
    ok = ~(np.abs(acc) < tol)
    uniq, acc = uniq[ok], acc[ok]
    ucol, urow = uniq // n, uniq % n
    colptr = np.ones(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(ucol, minlength=n), out=colptr[1:])
    colptr[1:] += 1
    return colptr, (urow + 1).astype(np.int64), acc


def sparse_plain(basis: np.ndarray, terms, tol: float = RTOL_DEFAULT):
    """sparse(opr) of a plain OperatorRepresentation: column walk (operator_representation.jl:90-103) + assembly."""
    m, r, c, a = terms
    dtype = np.complex128 if np.iscomplexobj(a) else np.float64
    cols, rows, vals = [], [], []
    for t, sel, words in _walk(basis, terms, column=True):
        ir = lookup(basis, words)
        ok = ir > 0
        cols.append(sel[ok]); rows.append(ir[ok] - 1); vals.append(np.full(int(ok.sum()), a[t], dtype=dtype))
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dtype=dt)
    return _assemble_csc(len(basis), cat(cols, np.int64), cat(rows, np.int64), cat(vals, dtype), tol, dtype)


def sparse_reduced(basis: np.ndarray, reduced: np.ndarray, index: np.ndarray, amp: np.ndarray, terms, tol: float = RTOL_DEFAULT):
    """sparse(ropr) of a ReducedOperatorRepresentation: column iterator reduced_operator_representation.jl:88-116
    (value = a * conj(amp[row_p]) / conj(amp[col_p]), misses dropped) + the generic assembly (always ComplexF64, :26)."""
    m, r, c, a = terms
    icol_p = lookup(basis, reduced)
    inv_col = 1.0 / np.conj(amp[icol_p - 1])
    cols, rows, vals = [], [], []
    for t, sel, words in _walk(reduced, terms, column=True):
        irow_p = lookup(basis, words)
        ok = irow_p > 0
        irow_r = np.where(ok, index[np.maximum(irow_p, 1) - 1], -1)
        ok &= irow_r > 0
        cols.append(sel[ok]); rows.append(irow_r[ok] - 1)
        vals.append(complex(a[t]) * np.conj(amp[irow_p[ok] - 1]) * inv_col[sel[ok]])
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dtype=dt)
    return _assemble_csc(len(reduced), cat(cols, np.int64), cat(rows, np.int64), cat(vals, np.complex128), tol, np.complex128)


def apply_reduced(basis, reduced, index, amp, terms, x: np.ndarray) -> np.ndarray:
    """out = H_r x through the row iterator reduced_operator_representation.jl:57-85
    (value = a * amp[col_p] / amp[row_p]) and apply_serial! (abstract_operator_representation.jl:296-316)."""
    m, r, c, a = terms
    irow_p = lookup(basis, reduced)
    inv_row = 1.0 / amp[irow_p - 1]
    out = np.zeros(len(reduced), dtype=np.complex128)
    for t, sel, words in _walk(reduced, terms, column=False):
        icol_p = lookup(basis, words)
        ok = icol_p > 0
        icol_r = np.where(ok, index[np.maximum(icol_p, 1) - 1], -1)
        ok &= icol_r > 0
        out[sel[ok]] += complex(a[t]) * amp[icol_p[ok] - 1] * inv_row[sel[ok]] * x[icol_r[ok] - 1]
    return out


def vector_reduce(index: np.ndarray, amp: np.ndarray, n_small: int, large: np.ndarray) -> np.ndarray:
    """symmetry_reduce(rhsr, large) symmetry_reduce.jl:38-56: small[idx[p]] += conj(amp[p]) * large[p], p ascending."""
    small = np.zeros(n_small, dtype=np.complex128)
    sel = np.nonzero(index > 0)[0]
    np.add.at(small, index[sel] - 1, np.conj(amp[sel]) * large[sel])
    return small


def vector_unreduce(index: np.ndarray, amp: np.ndarray, small: np.ndarray) -> np.ndarray:
    """symmetry_unreduce symmetry_reduce.jl:208-225: large[p] = amp[p] * small[idx[p]]."""
    large = np.zeros(len(index), dtype=np.complex128)
    sel = np.nonzero(index > 0)[0]
    large[sel] = amp[sel] * small[index[sel] - 1]
    return large
