"""edcuda -- host-side mirror of ExactDiagonalization.jl's Hamiltonian-application API on top of
libedcuda.so (hand-written sm_100a kernels behind a C ABI, include/edcuda.h).  No CPU fallback."""
from ._lib import (DimensionMismatch, CudaError, UnsupportedError, device_count, kernel_launch_count, version,
                   ED_F64, ED_C128, ED_SIDE_LEFT, ED_SIDE_RIGHT, ED_BASIS_LIST, ED_BASIS_FULL,
                   ED_BASIS_COMBINADIC, ED_BASIS_DPRANK, LIB_PATH)
from .hilbert import State, Site, HilbertSpace, HilbertSpaceSector, basespace
from .operators import (Operator, NullOperator, simplify, pure_operator, pauli_matrix, spin_half_system,
                        get_row_iterator, get_column_iterator, get_element)
from .representation import (HilbertSpaceRepresentation, OperatorRepresentation, represent, represent_array,
                             represent_dict, apply_b, apply_serial_b, apply_parallel_b, mul_b, sparse, dimension)
from .symmetry import (SitePermutation, GlobalBitFlip, DirectProductOperation, symmetry_apply,
                       symmetry_apply_operator, isinvariant, isinvariant_all, symmetry_reduce, symmetry_reduce_serial,
                       symmetry_reduce_parallel, symmetry_reduce_b, symmetry_unreduce,
                       ReducedHilbertSpaceRepresentation, ReducedOperatorRepresentation)
from . import lattices, models
