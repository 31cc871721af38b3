"""Host-side descriptors mirroring the reference's abstract Hilbert-space layer.

Reference: src/HilbertSpace/site.jl:29-93 (State, Site, bitwidth), hilbert_space.jl:25-129
(HilbertSpace: bit layout, bitmasks, quantum_number_sectors), hilbert_space_sector.jl:13-64.
These objects only DESCRIBE the space (they are what the Julia shim would hand over); the basis
itself is generated on the device by libedcuda (`represent`).  Site indices are 0-based here
(Julia site i = Python site i-1); local state indices likewise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check, lib


def _as_qn(q) -> Tuple[int, ...]:
    if isinstance(q, (int, np.integer)):
        return (int(q),)
    return tuple(int(x) for x in q)


class State:
    """State(name, quantum_number) -- site.jl:29-41."""

    def __init__(self, name: str, quantum_number=0):
        self.name = name
        self.quantum_number = _as_qn(quantum_number)

    def __eq__(self, other):
        return isinstance(other, State) and (self.name, self.quantum_number) == (other.name, other.quantum_number)

    def __repr__(self):
        return f"State({self.name!r}, {self.quantum_number})"


class Site:
    """Site(states) -- site.jl:69-72; bitwidth = ceil(log2(#states)) (:93)."""

    def __init__(self, states: Sequence[State]):
        self.states = list(states)
        if not self.states:
            raise ValueError("a site needs at least one state")

    @property
    def bitwidth(self) -> int:
        return int(math.ceil(math.log2(len(self.states))))

    @property
    def dimension(self) -> int:
        return len(self.states)

    def __eq__(self, other):
        return isinstance(other, Site) and self.states == other.states


class HilbertSpace:
    """HilbertSpace(sites) -- hilbert_space.jl:25-41."""

    def __init__(self, sites: Sequence[Site]):
        self.sites = list(sites)
        self.bitwidths = [s.bitwidth for s in self.sites]
        self.bitoffsets = [0]
        for w in self.bitwidths:
            self.bitoffsets.append(self.bitoffsets[-1] + w)
        nq = {len(st.quantum_number) for s in self.sites for st in s.states}
        if len(nq) > 1:
            raise ValueError("all states need quantum numbers of the same length")
        self.n_qn = nq.pop() if nq else 0
        self._handle = None

    # -- reference accessors
    @property
    def bitwidth(self) -> int:
        return self.bitoffsets[-1]

    @property
    def parent(self):
        return self

    def basespace(self):
        return self

    def get_bitmask(self, isite: int | None = None) -> int:
        """hilbert_space.jl:107-113."""
        if isite is None:
            return (1 << self.bitwidth) - 1
        return ((1 << self.bitoffsets[isite + 1]) - 1) ^ ((1 << self.bitoffsets[isite]) - 1)

    def quantum_number_sectors(self) -> List[Tuple[int, ...]]:
        """hilbert_space.jl:119-129."""
        cur = {tuple([0] * self.n_qn)}
        for site in self.sites:
            cur = {tuple(a + b for a, b in zip(q, st.quantum_number)) for st in site.states for q in cur}
        return sorted(cur)

    def compress(self, indices: Iterable[int]) -> int:
        """hilbert_space.jl:199-214 (0-based local indices)."""
        out = 0
        for isite, idx in enumerate(indices):
            if not (0 <= idx < self.sites[isite].dimension):
                raise IndexError(idx)
            out |= int(idx) << self.bitoffsets[isite]
        return out

    def extract(self, binrep: int) -> Tuple[int, ...]:
        """hilbert_space.jl:172-185 (0-based local indices)."""
        out = []
        for isite, site in enumerate(self.sites):
            idx = (binrep >> self.bitoffsets[isite]) & ((1 << self.bitwidths[isite]) - 1)
            if idx >= site.dimension:
                raise IndexError(idx)
            out.append(idx)
        return tuple(out)

    def get_quantum_number(self, binrep: int) -> Tuple[int, ...]:
        """hilbert_space.jl:135-145."""
        q = [0] * self.n_qn
        for isite, idx in enumerate(self.extract(binrep)):
            q = [a + b for a, b in zip(q, self.sites[isite].states[idx].quantum_number)]
        return tuple(q)

    def __eq__(self, other):
        return isinstance(other, HilbertSpace) and self.sites == other.sites

    # -- engine handle
    def handle(self):
        if self._handle is None:
            n_states = np.array([s.dimension for s in self.sites], dtype=np.int32)
            qn = np.array([q for s in self.sites for st in s.states for q in st.quantum_number], dtype=np.int64)
            h = C.c_void_p()
            check(lib.ed_space_create(len(self.sites), n_states.ctypes.data, qn.ctypes.data if qn.size else None,
                                      self.n_qn, C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and lib is not None:
            try:
                lib.ed_space_destroy(h)
            except Exception:
                pass


class HilbertSpaceSector:
    """HilbertSpaceSector(parent, allowed) -- hilbert_space_sector.jl:13-64: the allowed set is
    intersected with the sectors the space can reach."""

    def __init__(self, parent: HilbertSpace, allowed=None):
        self.parent = parent
        sectors = set(parent.quantum_number_sectors())
        if allowed is None:
            req = sectors
        elif isinstance(allowed, (int, np.integer)):
            req = {(int(allowed),)}
        elif isinstance(allowed, tuple) and all(isinstance(a, (int, np.integer)) for a in allowed):
            req = {_as_qn(allowed)}
        else:
            req = {_as_qn(a) for a in allowed}
        self.allowed_quantum_numbers = sectors & req

    @property
    def bitwidth(self) -> int:
        return self.parent.bitwidth

    def basespace(self):
        return self.parent

    def __eq__(self, other):
        return (isinstance(other, HilbertSpaceSector) and self.parent == other.parent
                and self.allowed_quantum_numbers == other.allowed_quantum_numbers)


def basespace(hs):
    return hs.basespace()
