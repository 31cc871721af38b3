"""ctypes binding of libedcuda.so (the C ABI declared in include/edcuda.h).

The library is the product; this module only loads it and converts its status codes to
Python exceptions mirroring the reference's Julia exceptions:
    ArgumentError -> ValueError, DimensionMismatch -> DimensionMismatch(ValueError),
    BoundsError -> IndexError, KeyError -> KeyError.
There is no fallback: if the shared library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

ED_OK = 0
ED_ERR_ARGUMENT = 1
ED_ERR_DIMENSION_MISMATCH = 2
ED_ERR_BOUNDS = 3
ED_ERR_KEY = 4
ED_ERR_CUDA = 5
ED_ERR_UNSUPPORTED = 6
ED_ERR_INTERNAL = 7

ED_F64, ED_C128 = 0, 1
ED_SIDE_LEFT, ED_SIDE_RIGHT = 0, 1
ED_BASIS_LIST, ED_BASIS_FULL, ED_BASIS_COMBINADIC, ED_BASIS_DPRANK = 0, 1, 2, 3


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch."""


class CudaError(RuntimeError):
    """CUDA failure / no device (the engine has no CPU fallback)."""


class UnsupportedError(NotImplementedError):
    """Valid in the reference but outside the engine (e.g. BR wider than 64 bits)."""


_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("EDCUDA_LIB", os.path.join(_PKG_DIR, "libedcuda.so"))

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C exactdiagonalization.jl_b200/csrc`). The engine has no CPU fallback.")

# The multi-GPU context binds NCCL at run time (dlopen of libnccl.so.2).  When PyTorch's bundled copy exists, point the
# library at it: whichever copy is loaded first under that SONAME is the one PyTorch gets too, and an older system NCCL
# lacks symbols libtorch_cuda.so needs (ImportError on a later `import torch`).
if "EDCUDA_NCCL_LIB" not in os.environ:
    try:
        import importlib.util as _ilu
        _spec = _ilu.find_spec("nvidia.nccl")
        for _loc in (list(_spec.submodule_search_locations) if _spec and _spec.submodule_search_locations else []):
            _cand = os.path.join(_loc, "lib", "libnccl.so.2")
            if os.path.exists(_cand):
                os.environ["EDCUDA_NCCL_LIB"] = _cand
                break
    except Exception:
        pass

lib = C.CDLL(LIB_PATH)

vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
P = C.POINTER

# name -> (restype, argtypes); every symbol declared in include/edcuda.h
SIGNATURES = {
    "ed_last_error": (C.c_char_p, []),
    "ed_version": (C.c_char_p, []),
    "ed_device_count": (C.c_int, []),
    "ed_set_device": (C.c_int, [C.c_int]),
    "ed_set_stream": (C.c_int, [vp, i32]),
    "ed_kernel_launch_count": (i64, []),
    "ed_release_staging": (C.c_int, []),
    "ed_device_malloc": (C.c_int, [i64, P(vp)]),
    "ed_device_free": (C.c_int, [vp]),
    "ed_space_create": (C.c_int, [i32, vp, vp, i32, P(vp)]),
    "ed_space_destroy": (C.c_int, [vp]),
    "ed_space_bitwidth": (C.c_int, [vp, P(i32)]),
    "ed_basis_generate": (C.c_int, [vp, vp, i64, i32, P(vp)]),
    "ed_basis_from_list": (C.c_int, [vp, vp, i64, i32, P(vp)]),
    "ed_basis_destroy": (C.c_int, [vp]),
    "ed_basis_dim": (C.c_int, [vp, P(i64)]),
    "ed_basis_kind": (C.c_int, [vp, P(i32)]),
    "ed_basis_download": (C.c_int, [vp, i64, i64, vp]),
    "ed_basis_lookup": (C.c_int, [vp, vp, i64, vp]),
    "ed_basis_device_words": (C.c_int, [vp, P(vp)]),
    "ed_operator_create": (C.c_int, [i64, vp, vp, vp, vp, i32, P(vp)]),
    "ed_operator_destroy": (C.c_int, [vp]),
    "ed_symmetry_create": (C.c_int, [i32, i32, vp, vp, vp, P(vp)]),
    "ed_symmetry_destroy": (C.c_int, [vp]),
    "ed_symmetry_apply": (C.c_int, [vp, vp, i32, vp, i64, vp]),
    "ed_symmetry_reduce": (C.c_int, [vp, vp, dbl, P(vp)]),
    "ed_rbasis_destroy": (C.c_int, [vp]),
    "ed_rbasis_dim": (C.c_int, [vp, P(i64)]),
    "ed_rbasis_download": (C.c_int, [vp, i64, i64, vp]),
    "ed_rbasis_orbit_sizes": (C.c_int, [vp, i64, i64, vp]),
    "ed_rbasis_mapping": (C.c_int, [vp, vp, i64, vp, vp]),
    "ed_rbasis_mapping_rows": (C.c_int, [vp, i64, i64, vp, vp]),
    "ed_vector_reduce": (C.c_int, [vp, vp, i64, vp, i64, i32, i32]),
    "ed_vector_unreduce": (C.c_int, [vp, vp, i64, vp, i64, i32]),
    "ed_oprep_create": (C.c_int, [vp, vp, P(vp)]),
    "ed_oprep_create_reduced": (C.c_int, [vp, vp, P(vp)]),
    "ed_oprep_destroy": (C.c_int, [vp]),
    "ed_oprep_dim": (C.c_int, [vp, P(i64)]),
    "ed_oprep_dtype": (C.c_int, [vp, P(i32)]),
    "ed_oprep_set_rows": (C.c_int, [vp, i64, i64]),
    "ed_oprep_suggest_rows": (C.c_int, [vp, i32, i32, i32, P(i64), P(i64)]),
    "ed_oprep_set_kernel": (C.c_int, [vp, i32]),
    "ed_apply": (C.c_int, [vp, vp, i64, vp, i64, i32, i32, i32]),
    "ed_apply_async": (C.c_int, [vp, vp, vp, i32, i32, i32, vp]),
    "ed_oprep_row_iterator": (C.c_int, [vp, i64, i32, i64, vp, vp, P(i64)]),
    "ed_oprep_get_element": (C.c_int, [vp, i64, i64, vp]),
    "ed_sparse_count": (C.c_int, [vp, dbl, P(i64)]),
    "ed_sparse_fetch": (C.c_int, [vp, vp, vp, vp]),
    "ed_oprep_cache_matrix": (C.c_int, [vp, i32, P(i64)]),
    "ed_oprep_drop_cache": (C.c_int, [vp]),
    "ed_dense": (C.c_int, [vp, vp]),
    "ed_lanczos": (C.c_int, [vp, i32, vp, i32, u64, vp, vp, vp, i32, P(i32)]),
    "ed_lanczos_update_async": (C.c_int, [vp, vp, vp, i64, i32, vp, vp, vp, vp]),
    "ed_vector_norm2_async": (C.c_int, [vp, i64, i32, vp]),
    "ed_vector_randn_async": (C.c_int, [vp, i64, i32, u64, i64]),
    "ed_tridiag_eigvals": (C.c_int, [vp, vp, i32, vp]),
    "ed_vector_scale_async": (C.c_int, [vp, i64, i32, dbl]),
    "ed_operator_isinvariant": (C.c_int, [vp, vp, vp, dbl, P(i32), P(i32)]),
    "ed_basis_save": (C.c_int, [vp, C.c_char_p]),
    "ed_basis_load": (C.c_int, [vp, C.c_char_p, P(vp)]),
    "ed_rbasis_save": (C.c_int, [vp, C.c_char_p]),
    "ed_rbasis_load": (C.c_int, [vp, vp, dbl, C.c_char_p, P(vp)]),
    "ed_lanczos_state_create": (C.c_int, [vp, i32, vp, u64, P(vp)]),
    "ed_lanczos_state_destroy": (C.c_int, [vp]),
    "ed_lanczos_state_step": (C.c_int, [vp, i32]),
    "ed_lanczos_state_result": (C.c_int, [vp, i32, vp, vp, vp, i32, P(i32)]),
    "ed_lanczos_state_save": (C.c_int, [vp, C.c_char_p]),
    "ed_lanczos_state_load": (C.c_int, [vp, C.c_char_p, P(vp)]),
    "ed_ctx_unique_id": (C.c_int, [vp]),
    "ed_ctx_create": (C.c_int, [i32, vp, P(vp)]),
    "ed_ctx_create_rank": (C.c_int, [i32, i32, i32, vp, P(vp)]),
    "ed_ctx_destroy": (C.c_int, [vp]),
    "ed_ctx_info": (C.c_int, [vp, P(i32), P(i32), P(i32), P(i32)]),
    "ed_ctx_device_stream": (C.c_int, [vp, i32, P(i32), P(vp)]),
    "ed_ctx_sync": (C.c_int, [vp]),
    "ed_ctx_barrier": (C.c_int, [vp]),
    "ed_ctx_timer_record": (C.c_int, [vp, i32]),
    "ed_ctx_timer_elapsed": (C.c_int, [vp, i32, i32, P(dbl)]),
    "ed_ctx_allreduce_host": (C.c_int, [vp, vp, i32, i32]),
    "ed_sharded_create": (C.c_int, [vp, vp, i32, i32, i32, P(vp)]),
    "ed_sharded_destroy": (C.c_int, [vp]),
    "ed_sharded_info": (C.c_int, [vp, i32, P(i64), P(i64), P(i32), P(i32), P(i32), P(i32)]),
    "ed_sharded_ranges": (C.c_int, [vp, i32, vp, vp]),
    "ed_dvec_create": (C.c_int, [vp, P(vp)]),
    "ed_dvec_destroy": (C.c_int, [vp]),
    "ed_dvec_local": (C.c_int, [vp, i32, P(vp), P(i64)]),
    "ed_dvec_randn": (C.c_int, [vp, u64, dbl]),
    "ed_dvec_upload": (C.c_int, [vp, vp]),
    "ed_dvec_download": (C.c_int, [vp, vp]),
    "ed_apply_sharded": (C.c_int, [vp, vp, vp, i32, vp]),
    "ed_sharded_profile": (C.c_int, [vp, vp, vp, vp]),
    "ed_lanczos_sharded": (C.c_int, [vp, i32, u64, vp, vp, vp, vp, i32, P(i32), P(dbl)]),
    "ed_shard_plan_describe": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.ed_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    if status == ED_OK:
        return
    msg = last_error()
    if status == ED_ERR_ARGUMENT:
        raise ValueError(msg)
    if status == ED_ERR_DIMENSION_MISMATCH:
        raise DimensionMismatch(msg)
    if status == ED_ERR_BOUNDS:
        raise IndexError(msg)
    if status == ED_ERR_KEY:
        raise KeyError(msg)
    if status == ED_ERR_CUDA:
        raise CudaError(msg)
    if status == ED_ERR_UNSUPPORTED:
        raise UnsupportedError(msg)
    raise RuntimeError(f"edcuda internal error {status}: {msg}")


def device_count() -> int:
    return int(lib.ed_device_count())


def kernel_launch_count() -> int:
    return int(lib.ed_kernel_launch_count())


def version() -> str:
    return lib.ed_version().decode()
