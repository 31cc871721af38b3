"""ctypes wrapper of oracle/libed_oracle.so (the oracle's C/OpenMP twin).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libed_oracle.so")
BUILD = "portable (-O3 -fopenmp, built in the build container)"


def _native_build():
    """bench.py's CPU arms set ED_ORACLE_NATIVE=1: rebuild the twin on the machine that runs it with -march=native
    (BASELINE.md section 3 promises that for the timed CPU baseline); the portable prebuilt library is the fallback."""
    import subprocess
    out_dir = os.path.join(_HERE, "_native")
    out = os.path.join(out_dir, "libed_oracle_native.so")
    try:
        os.makedirs(out_dir, exist_ok=True)
        src = os.path.join(_HERE, "ed_oracle_c.c")
        if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", out + ".tmp%d" % os.getpid(), src],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            os.replace(out + ".tmp%d" % os.getpid(), out)
        return out
    except Exception:
        return None


if os.environ.get("ED_ORACLE_NATIVE") == "1":
    _n = _native_build()
    if _n:
        _PATH, BUILD = _n, "native (-O3 -march=native -fopenmp, built on this host)"
if not os.path.exists(_PATH):
    raise ImportError(f"{_PATH} missing: run `make -C oracle`")
_lib = C.CDLL(_PATH)
vp, i64 = C.c_void_p, C.c_int64
_lib.oc_num_threads.restype = C.c_int
_lib.oc_set_num_threads.argtypes = [C.c_int]
_lib.oc_basis_fixed_popcount.restype = i64
_lib.oc_basis_fixed_popcount.argtypes = [C.c_int, C.c_int, vp, i64]
_lib.oc_sector_basis_dp.restype = i64
_lib.oc_sector_basis_dp.argtypes = [C.c_int, vp, vp, vp, C.c_int, i64, vp, i64]
_lib.oc_apply.restype = None
_lib.oc_apply.argtypes = [vp, i64, i64, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, i64, i64, C.c_int]
_lib.oc_count_hits.restype = i64
_lib.oc_count_hits.argtypes = [vp, i64, i64, vp, vp, vp, i64, i64]
_lib.oc_apply_reduced_onthefly.restype = None
_lib.oc_apply_reduced_onthefly.argtypes = [vp, vp, i64, C.c_int, C.c_int, vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, i64]


def num_threads() -> int:
    return int(_lib.oc_num_threads())


def set_num_threads(n: int) -> None:
    _lib.oc_set_num_threads(int(n))


def basis_fixed_popcount(n_bits: int, n_set: int) -> np.ndarray:
    import math
    n = math.comb(n_bits, n_set)
    out = np.empty(n, dtype=np.uint64)
    got = _lib.oc_basis_fixed_popcount(n_bits, n_set, out.ctypes.data, n)
    assert got == n
    return out


def sector_basis_dp(n_states, qn, target: int) -> np.ndarray:
    """qn: list of per-site lists of integer quantum numbers."""
    n_sites = len(n_states)
    widths = [int(np.ceil(np.log2(k))) for k in n_states]
    offs = np.array([sum(widths[:i]) for i in range(n_sites)], dtype=np.int32)
    ms = max(n_states)
    q = np.zeros((n_sites, ms), dtype=np.int64)
    for i, row in enumerate(qn):
        q[i, : len(row)] = row
    ns = np.array(n_states, dtype=np.int32)
    cap = int(np.prod(n_states))
    out = np.empty(cap, dtype=np.uint64)
    n = _lib.oc_sector_basis_dp(n_sites, ns.ctypes.data, offs.ctypes.data, q.ctypes.data, ms, int(target), out.ctypes.data, cap)
    return out[:n].copy()


def _terms(terms):
    m, r, c, a = terms
    m = np.ascontiguousarray(m, dtype=np.uint64); r = np.ascontiguousarray(r, dtype=np.uint64)
    c = np.ascontiguousarray(c, dtype=np.uint64)
    cplx = np.iscomplexobj(a)
    a = np.ascontiguousarray(a, dtype=np.complex128 if cplx else np.float64)
    return m, r, c, a, int(cplx)


def apply(basis: np.ndarray, terms, x: np.ndarray, out: np.ndarray, row_lo: int = 0, row_hi=None, side: int = 0):
    """out[row_lo:row_hi] += (H x)[row_lo:row_hi]; `terms` = (mask, row, col, amp) arrays; out has row_hi-row_lo entries."""
    m, r, c, a, cplx = _terms(terms)
    dim = len(basis)
    row_hi = dim if row_hi is None else row_hi
    vec_c = int(np.iscomplexobj(x))
    assert out.dtype == x.dtype and len(out) == row_hi - row_lo and len(x) == dim
    assert vec_c or not cplx
    _lib.oc_apply(basis.ctypes.data, dim, len(m), m.ctypes.data, r.ctypes.data, c.ctypes.data, a.ctypes.data, cplx,
                  x.ctypes.data, out.ctypes.data, vec_c, row_lo, row_hi, side)
    return out


def count_hits(basis, terms, row_lo=0, row_hi=None) -> int:
    m, r, c, a, _ = _terms(terms)
    row_hi = len(basis) if row_hi is None else row_hi
    return int(_lib.oc_count_hits(basis.ctypes.data, len(basis), len(m), m.ctypes.data, r.ctypes.data, c.ctypes.data, row_lo, row_hi))


def apply_reduced_onthefly(rbasis, orbit, perms, chi, terms, x, out, row_lo=0, row_hi=None):
    m, r, c, a, cplx = _terms(terms)
    assert not cplx
    perms = np.ascontiguousarray(perms, dtype=np.int32)
    chi = np.ascontiguousarray(np.asarray(chi, dtype=np.complex128))
    orbit = np.ascontiguousarray(orbit, dtype=np.int32)
    row_hi = len(rbasis) if row_hi is None else row_hi
    _lib.oc_apply_reduced_onthefly(rbasis.ctypes.data, orbit.ctypes.data, len(rbasis), perms.shape[1], perms.shape[0],
                                   perms.ctypes.data, chi.ctypes.data, len(m), m.ctypes.data, r.ctypes.data, c.ctypes.data,
                                   a.ctypes.data, x.ctypes.data, out.ctypes.data, row_lo, row_hi)
    return out
