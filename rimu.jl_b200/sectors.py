"""Dense-indexed deterministic vectors over a complete particle-number sector (BASELINE config 3).

`SectorBasis(ham)` numbers every address of the Hamiltonian's address type by its combinadic rank (the role of the `basis`
of Rimu's `BasisSetRepresentation`, ExactDiagonalization/basis_set_representation.jl:35-60, in a fixed analytic order instead
of BFS order); `DenseSectorVec` is a coefficient vector over it living in HBM.  `mul(y, H, x)` on such vectors is the
matrix-free gather of csrc/sector.cuh.  The class offers the same methods the Lanczos driver uses on `GPUDVec`
(`norm`, `scale_`, `dot`, `add_`, `similar`, `copy`), so `eigsolve_lanczos` runs on either container.
Nothing is computed on the host."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .hamiltonians import AbstractHamiltonian, Context, get_context


class SectorBasis:
    def __init__(self, ham: AbstractHamiltonian, ctx: Context | None = None):
        self.ham = ham
        self.ctx = ctx or get_context(ham.address.address_type.words)
        h = C.c_void_p()
        _lib.check(_lib.lib().rimu_sector_create(self.ctx.handle, ham.handle, C.byref(h)))
        self.handle = h
        d = C.c_uint64()
        _lib.check(_lib.lib().rimu_sector_dim(self.handle, C.byref(d)))
        self.dim = int(d.value)

    def __len__(self):
        return self.dim

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().rimu_sector_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def rank(self, keys) -> np.ndarray:
        keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64).reshape(-1))
        out = np.zeros(len(keys), dtype=np.int64)
        _lib.check(_lib.lib().rimu_sector_rank(self.handle, keys.ctypes.data_as(_lib._u64p), len(keys), out.ctypes.data_as(_lib._i64p)))
        return out

    def keys(self, first=0, count=None) -> np.ndarray:
        count = self.dim - first if count is None else count
        out = np.zeros(count, dtype=np.uint64)
        _lib.check(_lib.lib().rimu_sector_keys(self.handle, first, count, out.ctypes.data_as(_lib._u64p)))
        return out

    def zeros(self) -> "DenseSectorVec":
        return DenseSectorVec(self)

    def vector(self, pairs) -> "DenseSectorVec":
        """DVec(address => value, ...) in the dense layout."""
        pairs = list(pairs.items()) if isinstance(pairs, dict) else list(pairs)
        idx = self.rank([a.key()[0] if hasattr(a, "key") else a for a, _ in pairs])
        return DenseSectorVec(self).set(idx, [v for _, v in pairs])


class DenseSectorVec:
    def __init__(self, basis: SectorBasis):
        self.basis = basis
        p = C.c_void_p()
        _lib.check(_lib.lib().rimu_sector_vec_create(basis.handle, C.byref(p)))
        self.ptr = p
        self.last_mul_ms = None

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                _lib.lib().rimu_sector_vec_destroy(self.basis.handle, self.ptr)
                self.ptr = None
        except Exception:
            pass

    def __len__(self):
        return self.basis.dim

    # ---- host <-> device
    def set(self, index, vals):
        index = np.ascontiguousarray(np.asarray(index, dtype=np.int64))
        vals = np.ascontiguousarray(np.asarray(vals, dtype=np.float64))
        _lib.check(_lib.lib().rimu_sector_vec_set(self.basis.handle, self.ptr, index.ctypes.data_as(_lib._i64p),
                                                  vals.ctypes.data_as(_lib._f64p), len(vals)))
        return self

    def get(self, first=0, count=None) -> np.ndarray:
        count = self.basis.dim - first if count is None else count
        out = np.zeros(count, dtype=np.float64)
        _lib.check(_lib.lib().rimu_sector_vec_get(self.basis.handle, self.ptr, first, count, out.ctypes.data_as(_lib._f64p)))
        return out

    def gather(self, index) -> np.ndarray:
        index = np.ascontiguousarray(np.asarray(index, dtype=np.int64))
        out = np.zeros(len(index), dtype=np.float64)
        _lib.check(_lib.lib().rimu_sector_vec_gather(self.basis.handle, self.ptr, index.ctypes.data_as(_lib._i64p), len(index),
                                                     out.ctypes.data_as(_lib._f64p)))
        return out

    def from_dvec(self, v):
        """copy!(dense, dictionary vector)"""
        _lib.check(_lib.lib().rimu_sector_from_vec(self.basis.handle, v.handle, self.ptr))
        return self

    def to_dvec(self, out=None):
        """dictionary vector of the non-zero entries"""
        from .dictvectors import GPUDVec
        from .stochasticstyles import IsDeterministic
        if out is None:
            out = GPUDVec(style=IsDeterministic(), address_type=self.basis.ham.address.address_type, ctx=self.basis.ctx)
        _lib.check(_lib.lib().rimu_sector_to_vec(self.basis.handle, self.ptr, out.handle))
        return out

    # ---- the vector operations of the Krylov driver (same names as GPUDVec)
    def similar(self, style=None):
        return DenseSectorVec(self.basis)

    zerovector = similar

    def copy(self):
        out = DenseSectorVec(self.basis)
        _lib.check(_lib.lib().rimu_sector_axpby(self.basis.handle, 1.0, self.ptr, 0.0, out.ptr))
        return out

    def copy_from(self, other):
        _lib.check(_lib.lib().rimu_sector_axpby(self.basis.handle, 1.0, other.ptr, 0.0, self.ptr))
        return self

    def dot(self, other) -> float:
        out = C.c_double()
        _lib.check(_lib.lib().rimu_sector_dot(self.basis.handle, self.ptr, other.ptr, C.byref(out)))
        return out.value

    def norm(self, p=2) -> float:
        if p != 2:
            raise ValueError("dense sector vectors offer the 2-norm")
        return math.sqrt(self.dot(self))

    def scale_(self, alpha):
        _lib.check(_lib.lib().rimu_sector_axpby(self.basis.handle, 0.0, self.ptr, float(alpha), self.ptr))
        return self

    def add_(self, other, alpha=1.0):
        """self += alpha * other"""
        _lib.check(_lib.lib().rimu_sector_axpby(self.basis.handle, float(alpha), other.ptr, 1.0, self.ptr))
        return self

    def mul_from(self, ham, x):
        """self = ham * x (mul!)"""
        if ham is not self.basis.ham:
            raise ValueError("the sector was built for another Hamiltonian")
        ms = C.c_float()
        _lib.check(_lib.lib().rimu_sector_mul(self.basis.handle, x.ptr, self.ptr, C.byref(ms)))
        self.last_mul_ms = ms.value
        return self
