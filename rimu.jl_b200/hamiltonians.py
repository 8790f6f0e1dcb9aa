"""Hamiltonians: host mirror of Rimu's `AbstractHamiltonian` interface for the four device models.

Classes keep the reference's constructor signatures
  HubbardReal1D(address; u, t)                         Hamiltonians/HubbardReal1D.jl:33-36
  HubbardMom1D(address; u, t, dispersion)              Hamiltonians/HubbardMom1D.jl:49-65
  HubbardRealSpace(address; geometry, t, u, v)         Hamiltonians/HubbardRealSpace.jl:176-241
  Transcorrelated1D(address; t, v, v_ho, cutoff, three_body_term)  Transcorrelated1D.jl:72-89
and the interface functions `starting_address`, `diagonal_element`, `num_offdiagonals`,
`get_offdiagonal`, `offdiagonals`, `random_offdiagonal` (Interfaces/hamiltonians.jl:143-370).
All matrix elements are evaluated BY THE DEVICE CODE through the C-ABI hooks
(rimu_ham_diagonal / rimu_ham_offdiagonals); the host only precomputes the same constant
tables the Julia constructors do.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Sequence

import numpy as np

from . import _lib
from .addresses import AddressType, BoseFS, CompositeFS, FermiFS

# --------------------------------------------------------------------------- device contexts
_contexts = {}


class Context:
    """One GPU context (stream + working table [+ NCCL communicator]) per address width."""

    def __init__(self, words: int, device: int | None = None, table_slots: int | None = None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        if table_slots is None:
            table_slots = int(os.environ.get("RIMU_B200_TABLE_SLOTS", str(1 << 22)))
        h = C.c_void_p()
        _lib.check(_lib.lib().rimu_ctx_create(device, words, table_slots, C.byref(h)))
        self.handle, self.words, self.device = h, words, device
        self.rank, self.nranks = 0, 1

    @property
    def table_slots(self):
        out = C.c_uint64()
        _lib.check(_lib.lib().rimu_ctx_table_slots(self.handle, C.byref(out)))
        return out.value

    def resize_table(self, slots: int):
        _lib.check(_lib.lib().rimu_ctx_resize_table(self.handle, slots))

    def synchronize(self):
        _lib.check(_lib.lib().rimu_ctx_synchronize(self.handle))

    def stream(self) -> int:
        s = C.c_void_p()
        _lib.check(_lib.lib().rimu_ctx_stream(self.handle, C.byref(s)))
        return s.value or 0

    def attach_comm(self, unique_id: bytes, rank: int, nranks: int, records_per_peer: int):
        buf = C.create_string_buffer(unique_id, 128)
        _lib.check(_lib.lib().rimu_comm_init(self.handle, buf, rank, nranks, records_per_peer))
        self.rank, self.nranks = rank, nranks

    def allreduce(self, values: Sequence[float]):
        arr = (C.c_double * len(values))(*values)
        _lib.check(_lib.lib().rimu_comm_allreduce_f64(self.handle, arr, len(values)))
        return list(arr)

    def detach(self):
        """Collective (every rank, same order): close the peer mappings before anybody frees the buffers behind them."""
        if self.handle and self.nranks > 1:
            _lib.check(_lib.lib().rimu_comm_detach(self.handle))

    def close(self, collective: bool = False):
        """Destroy the context.  `collective=True` (multi-GPU: called by every rank at the same point) detaches from the peers
        first; the default is a local, best-effort teardown (finalisers, process exit)."""
        if self.handle:
            if collective:
                self.detach()
            _lib.lib().rimu_ctx_destroy(self.handle)
            self.handle = None


def get_context(words: int) -> Context:
    ctx = _contexts.get(words)
    if ctx is None:
        ctx = _contexts[words] = Context(words)
    return ctx


def reset_contexts():
    for c in _contexts.values():
        c.close()
    _contexts.clear()


# --------------------------------------------------------------------------- geometry (geometry.jl:45-125)
class CubicGrid:
    def __init__(self, dims, fold=None):
        dims = tuple(int(d) for d in dims)
        if any(d <= 1 for d in dims):
            raise ValueError("All dimensions must be at least 2 in size")
        self.dims = dims
        self.fold = tuple(bool(f) for f in fold) if fold is not None else (True,) * len(dims)

    def __len__(self):
        return int(np.prod(self.dims))

    def __repr__(self):
        return f"CubicGrid({self.dims}, {self.fold})"


def PeriodicBoundaries(*dims):
    dims = dims[0] if len(dims) == 1 and hasattr(dims[0], "__iter__") else dims
    return CubicGrid(dims, (True,) * len(dims))


def HardwallBoundaries(*dims):
    dims = dims[0] if len(dims) == 1 and hasattr(dims[0], "__iter__") else dims
    return CubicGrid(dims, (False,) * len(dims))


def LadderBoundaries(*dims):
    dims = dims[0] if len(dims) == 1 and hasattr(dims[0], "__iter__") else dims
    return CubicGrid(dims, tuple(i > 0 for i in range(len(dims))))


def hubbard_dispersion(t, k):
    return -2 * t * np.cos(k)


def continuum_dispersion(t, k):
    return t * k ** 2


# --------------------------------------------------------------------------- base class
class AbstractHamiltonian:
    """Device-backed Hamiltonian.  Subclasses fill `self.desc` (a `_lib.HamDesc`)."""

    address = None
    desc: _lib.HamDesc
    hermitian = True  # LOStructure: IsHermitian vs AdjointUnknown

    def _finish(self, address):
        self.address = address
        self.address_type: AddressType = address.address_type
        at = self.address_type
        d = self.desc
        d.addr_kind, d.num_modes, d.num_components = at.kind, at.num_modes, at.num_components
        if at.kind == _lib.ADDR_COMPOSITE:
            if d.model != _lib.HUBBARD_REAL_SPACE:
                raise TypeError(f"{type(self).__name__} is defined for BoseFS or two-component fermionic (FermiFS2C, at most "
                                "32 modes) addresses; general CompositeFS addresses are supported by HubbardRealSpace")
            if at.num_components > _lib.MAX_COMPONENTS:
                raise ValueError(f"at most {_lib.MAX_COMPONENTS} components are supported on the device path")
            if at.bits > 127:
                raise ValueError(f"the address needs {at.bits} bits; at most 127 are supported on the device path")
            for c, (k, n) in enumerate(zip(at.comp_kinds, at.num_particles)):
                d.comp_kind[c], d.comp_particles[c] = k, n
        else:
            for c, n in enumerate(at.num_particles):
                d.num_particles[c] = n
        self._handle = None
        self._ctx = None

    # lazily created device handle
    @property
    def words(self):
        return self.address_type.words

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = get_context(self.words)
        return self._ctx

    @property
    def handle(self):
        if self._handle is None:
            ctx = self.ctx
            _lib.check(_lib.lib().rimu_ctx_make_current(ctx.handle))  # the tables go to the context's GPU
            h = C.c_void_p()
            _lib.check(_lib.lib().rimu_ham_create(C.byref(self.desc), C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().rimu_ham_destroy(self._handle)
        except Exception:
            pass

    # ---- key-level vectorised hooks
    def _keys_array(self, keys):
        a = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64).reshape(-1, self.words))
        return a

    def diagonal_elements(self, keys) -> np.ndarray:
        a = self._keys_array(keys)
        out = np.zeros(a.shape[0], dtype=np.float64)
        _lib.check(_lib.lib().rimu_ham_diagonal(self.ctx.handle, self.handle, a.ctypes.data_as(_lib._u64p), a.shape[0],
                                                out.ctypes.data_as(_lib._f64p)))
        return out

    def nums_offdiagonals(self, keys) -> np.ndarray:
        a = self._keys_array(keys)
        out = np.zeros(a.shape[0], dtype=np.int64)
        _lib.check(_lib.lib().rimu_ham_num_offdiagonals(self.ctx.handle, self.handle, a.ctypes.data_as(_lib._u64p),
                                                        a.shape[0], out.ctypes.data_as(_lib._i64p)))
        return out

    def offdiagonals_of_key(self, key, first=1, count=None):
        """(keys[count, W], values[count]) of get_offdiagonal(h, key, first..first+count-1)."""
        a = self._keys_array(key)
        if count is None:
            count = int(self.nums_offdiagonals(a)[0]) - first + 1
        ko = np.zeros((max(count, 0), self.words), dtype=np.uint64)
        vo = np.zeros(max(count, 0), dtype=np.float64)
        if count > 0:
            _lib.check(_lib.lib().rimu_ham_offdiagonals(self.ctx.handle, self.handle, a.ctypes.data_as(_lib._u64p), first, count,
                                                        ko.ctypes.data_as(_lib._u64p), vo.ctypes.data_as(_lib._f64p)))
        return ko, vo


# ---- the reference's generic functions
def starting_address(h):
    return h.address


def diagonal_element(h, addr):
    return float(h.diagonal_elements([addr.key()])[0])


def num_offdiagonals(h, addr):
    return int(h.nums_offdiagonals([addr.key()])[0])


def get_offdiagonal(h, addr, chosen):
    n = num_offdiagonals(h, addr)
    if not 1 <= chosen <= n:
        raise IndexError(f"off-diagonal index {chosen} out of range 1:{n}")  # BoundsError in the reference
    k, v = h.offdiagonals_of_key(addr.key(), chosen, 1)
    return h.address_type.from_key(k[0]), float(v[0])


def offdiagonals(h, addr):
    """AbstractVector{Tuple{A,T}} of all off-diagonals of `addr` in the reference's order."""
    k, v = h.offdiagonals_of_key(addr.key())
    return [(h.address_type.from_key(k[i]), float(v[i])) for i in range(len(v))]


def random_offdiagonal(h, addr, rng=None):
    """(new_address, probability, value) with uniform choice (Interfaces/hamiltonians.jl:361-370)."""
    rng = np.random.default_rng() if rng is None else rng
    n = num_offdiagonals(h, addr)
    i = int(rng.integers(1, n + 1))
    a, v = get_offdiagonal(h, addr, i)
    return a, 1.0 / n, v


def dimension(h):
    """number_conserving_dimension of the starting address (Hamiltonians/abstract.jl:68-142)."""
    at = h.address_type
    M = at.num_modes
    if at.kind == _lib.ADDR_BOSE:
        return math.comb(at.num_particles[0] + M - 1, at.num_particles[0])
    out = 1
    kinds = at.comp_kinds if at.kind == _lib.ADDR_COMPOSITE else (_lib.ADDR_FERMI,) * at.num_components
    for k, n in zip(kinds, at.num_particles):
        out *= math.comb(n + M - 1, n) if k == _lib.ADDR_BOSE else math.comb(M, n)
    return out


# --------------------------------------------------------------------------- models
class HubbardReal1D(AbstractHamiltonian):
    def __init__(self, address, u=1.0, t=1.0):
        if not isinstance(address, BoseFS):
            raise TypeError("HubbardReal1D on the device path needs a BoseFS address")
        self.u, self.t = float(u), float(t)
        self.desc = _lib.HamDesc()
        self.desc.model, self.desc.u, self.desc.t = _lib.HUBBARD_REAL_1D, self.u, self.t
        self._finish(address)

    def __repr__(self):
        return f"HubbardReal1D({self.address}; u={self.u}, t={self.t})"


def shift_lattice(is_):
    """shift_lattice(is) = circshift(is, cld(length(is), 2)) (HubbardReal1DEP.jl:9)."""
    is_ = list(is_)
    k = -(-len(is_) // 2)
    return is_[-k:] + is_[:-k]


def shift_lattice_inv(js):
    """shift_lattice_inv(js) = circshift(js, fld(length(js), 2)) (HubbardReal1DEP.jl:19)."""
    js = list(js)
    k = len(js) // 2
    return js[-k:] + js[:-k] if k else js


class HubbardReal1DEP(AbstractHamiltonian):
    """HubbardReal1DEP(address; u, t, v_ho) (Hamiltonians/HubbardReal1DEP.jl:47-92): Bose-Hubbard chain with the
    harmonic potential eps_i = v_ho * j_i^2, j = shift_lattice(-M÷2 : M÷2-1)."""

    def __init__(self, address, u=1.0, t=1.0, v_ho=1.0):
        if not isinstance(address, BoseFS):
            raise TypeError("HubbardReal1DEP on the device path needs a BoseFS address")
        self.u, self.t, self.v_ho = float(u), float(t), float(v_ho)
        M = address.num_modes
        js = shift_lattice(range(-(M // 2), -(M // 2) + M))
        self.ep = np.array([self.v_ho * j ** 2 for j in js], dtype=float)
        self.desc = _lib.HamDesc()
        self.desc.model, self.desc.u, self.desc.t, self.desc.has_potential = _lib.HUBBARD_REAL_1D_EP, self.u, self.t, 1
        for i in range(M):
            self.desc.potential[i] = self.ep[i]
        self._finish(address)

    def __repr__(self):
        return f"HubbardReal1DEP({self.address}; u={self.u}, t={self.t}, v_ho={self.v_ho})"


class ExtendedHubbardReal1D(AbstractHamiltonian):
    """ExtendedHubbardReal1D(address; u, v, t, boundary_condition) (Hamiltonians/ExtendedHubbardReal1D.jl:30-135) for
    bosons with the real boundary conditions :periodic, :hard_wall, :twisted (a complex twist angle makes the
    Hamiltonian complex; there is no device path for that and no CPU fallback)."""

    BCS = {"periodic": _lib.BC_PERIODIC, "hard_wall": _lib.BC_HARD_WALL, "twisted": _lib.BC_TWISTED}

    def __init__(self, address, u=1.0, v=1.0, t=1.0, boundary_condition="periodic"):
        if not isinstance(address, BoseFS):
            raise TypeError("ExtendedHubbardReal1D on the device path needs a BoseFS address")
        if boundary_condition not in self.BCS:
            raise ValueError("invalid boundary condition")  # ArgumentError (ExtendedHubbardReal1D.jl:64)
        self.u, self.v, self.t, self.boundary_condition = float(u), float(v), float(t), boundary_condition
        self.desc = _lib.HamDesc()
        self.desc.model, self.desc.u, self.desc.v, self.desc.t = _lib.EXTENDED_HUBBARD_REAL_1D, self.u, self.v, self.t
        self.desc.boundary_condition = self.BCS[boundary_condition]
        self._finish(address)

    def __repr__(self):
        return f"ExtendedHubbardReal1D({self.address}; u={self.u}, v={self.v}, t={self.t}, boundary_condition=:{self.boundary_condition})"


class HubbardMom1D(AbstractHamiltonian):
    def __init__(self, address, u=1.0, t=1.0, dispersion=hubbard_dispersion):
        if not isinstance(address, (BoseFS, CompositeFS)):
            raise TypeError("HubbardMom1D needs a BoseFS or FermiFS2C address")
        self.u, self.t = float(u), float(t)
        M = address.num_modes
        if M > _lib.MAX_TABLE_MODES:
            raise ValueError(f"momentum-space models support at most {_lib.MAX_TABLE_MODES} modes")
        step = 2 * math.pi / M
        start = -math.pi * (1 + 1 / M) + step if M % 2 else -math.pi + step
        self.ks = np.array([start + i * step for i in range(M)])
        self.kes = np.asarray(dispersion(self.t, self.ks), dtype=float)
        self.desc = _lib.HamDesc()
        self.desc.model, self.desc.u, self.desc.t = _lib.HUBBARD_MOM_1D, self.u, self.t
        for i in range(M):
            self.desc.kes[i] = self.kes[i]
        self._finish(address)

    def __repr__(self):
        return f"HubbardMom1D({self.address}; u={self.u}, t={self.t})"


def _momentum_grid(M, t, dispersion):
    """ks and kes of the momentum-space models (HubbardMom1D.jl:55-63)"""
    step = 2 * math.pi / M
    start = -math.pi * (1 + 1 / M) + step if M % 2 else -math.pi + step
    ks = np.array([start + i * step for i in range(M)])
    return ks, np.asarray(dispersion(t, ks), dtype=float)


class ExtendedHubbardMom1D(AbstractHamiltonian):
    """ExtendedHubbardMom1D(address; u, v, t, dispersion) (Hamiltonians/ExtendedHubbardMom1D.jl:37-117): the t-V model in
    momentum space, bosons, boundary_condition = 0.  The cosines the reference evaluates per element -- cos(q * 2pi / M) in
    get_offdiagonal, cos((mode_j - mode_i) * (2pi / M)) in the diagonal -- are tabulated by the host with the same expressions
    and passed through the descriptor, like kes."""

    def __init__(self, address, u=1.0, v=1.0, t=1.0, dispersion=hubbard_dispersion, boundary_condition=0.0):
        if not isinstance(address, BoseFS):
            raise TypeError("ExtendedHubbardMom1D on the device path needs a BoseFS address")
        if boundary_condition != 0.0:
            raise NotImplementedError("a twisted boundary condition has no device path")
        self.u, self.v, self.t, self.boundary_condition = float(u), float(v), float(t), 0.0
        M = address.num_modes
        if M > _lib.MAX_TABLE_MODES:
            raise ValueError(f"momentum-space models support at most {_lib.MAX_TABLE_MODES} modes")
        self.ks, self.kes = _momentum_grid(M, self.t, dispersion)
        self.desc = _lib.HamDesc()
        self.desc.model, self.desc.u, self.desc.v, self.desc.t = _lib.EXTENDED_HUBBARD_MOM_1D, self.u, self.v, self.t
        step = 2 * math.pi / M
        for i in range(M):
            self.desc.kes[i] = self.kes[i]
            self.desc.ws[i] = math.cos(i * 2 * math.pi / M)
            self.desc.us[i] = math.cos(i * step)
        self._finish(address)

    def __repr__(self):
        return f"ExtendedHubbardMom1D({self.address}; u={self.u}, v={self.v}, t={self.t}, boundary_condition={self.boundary_condition})"


def momentum_space_harmonic_potential(M, v):
    """HubbardMom1DEP.jl:14-31: 1/M * real(fft(v j^2)) over the shifted lattice, symmetrised"""
    js = shift_lattice(range(-(M // 2), -(M // 2) + M))
    mom = np.fft.fft(np.array([v * j * j for j in js], dtype=float))
    for i in range(1, M // 2 + 1):
        mom[M - i] = mom[i]
    return (1 / M) * np.real(mom)


class HubbardMom1DEP(AbstractHamiltonian):
    """HubbardMom1DEP(address; u, t, v_ho, dispersion) (Hamiltonians/HubbardMom1DEP.jl:68-257): Hubbard chain in momentum space
    with a harmonic trap; BoseFS or two FermiFS components.  The off-diagonals are the momentum-transfer block of
    HubbardMom1D followed by the one-body block of the potential (excitations.jl:257-267)."""

    def __init__(self, address, u=1.0, t=1.0, v_ho=1.0, dispersion=hubbard_dispersion):
        if not isinstance(address, (BoseFS, CompositeFS)):
            raise TypeError("HubbardMom1DEP needs a BoseFS or FermiFS2C address")
        self.u, self.t, self.v_ho = float(u), float(t), float(v_ho)
        M = address.num_modes
        if M > _lib.MAX_TABLE_MODES:
            raise ValueError(f"momentum-space models support at most {_lib.MAX_TABLE_MODES} modes")
        self.ks, self.kes = _momentum_grid(M, self.t, dispersion)
        self.ep = momentum_space_harmonic_potential(M, self.v_ho)
        self.desc = _lib.HamDesc()
        self.desc.model, self.desc.u, self.desc.t, self.desc.has_potential = _lib.HUBBARD_MOM_1D_EP, self.u, self.t, 1
        for i in range(M):
            self.desc.kes[i] = self.kes[i]
            self.desc.potential[i] = self.ep[i]
        self._finish(address)

    def __repr__(self):
        return f"HubbardMom1DEP({self.address}; u={self.u}, t={self.t}, v_ho={self.v_ho})"


class HubbardRealSpace(AbstractHamiltonian):
    def __init__(self, address, geometry=None, t=None, u=None, v=None):
        C_ = 1 if not isinstance(address, CompositeFS) else len(address.components)
        if isinstance(address, CompositeFS) and C_ < 2:
            raise TypeError("a CompositeFS address needs at least two components")  # (MethodError in the reference)
        general = isinstance(address, CompositeFS) and not address.is_fermi2c
        if C_ > _lib.MAX_COMPONENTS:
            raise ValueError(f"at most {_lib.MAX_COMPONENTS} components are supported on the device path")
        M = address.num_modes
        geometry = PeriodicBoundaries(M) if geometry is None else geometry
        D = len(geometry.dims)
        if len(geometry) != M:
            raise ValueError("`geometry` does not have the correct number of sites")
        t = np.ones(C_) if t is None else np.asarray(t, dtype=float).reshape(-1)
        u = np.ones((C_, C_)) if u is None else np.asarray(u, dtype=float).reshape(C_, C_)
        v = np.zeros((C_, D)) if v is None else np.asarray(v, dtype=float).reshape(C_, D)
        if t.shape != (C_,):
            raise ValueError(f"`t` must be a vector of length {C_}")
        if not np.array_equal(u, u.T):
            raise ValueError("`u` must be symmetric")
        if D > 3:
            raise ValueError("at most 3 lattice dimensions are supported on the device path")
        self.geometry, self.t, self.u, self.v = geometry, t, u, v
        d = self.desc = _lib.HamDesc()
        d.model, d.ndim = _lib.HUBBARD_REAL_SPACE, D
        for k in range(D):
            d.dims[k], d.fold[k] = geometry.dims[k], int(geometry.fold[k])
        for c in range(C_):
            if general:
                d.comp_t[c] = t[c]
            else:
                d.t_comp[c] = t[c]
            for c2 in range(C_):
                if general:
                    d.comp_u[c + C_ * c2] = u[c, c2]
                else:
                    d.u_mat[c + 2 * c2] = u[c, c2]
        if np.any(v != 0):  # HubbardRealSpace.jl:214-227
            d.has_potential = 1
            for site in range(M):
                idx, x2 = site, []
                for dim in geometry.dims:
                    x = idx % dim - dim // 2
                    idx //= dim
                    x2.append(x * x)
                for c in range(C_):
                    d.potential[c * M + site] = sum(v[c, k] * x2[k] for k in range(D))
        self._finish(address)

    def __repr__(self):
        return f"HubbardRealSpace({self.address}, geometry={self.geometry}, t={self.t}, u={self.u})"


class Transcorrelated1D(AbstractHamiltonian):
    hermitian = False  # LOStructure = AdjointUnknown (Transcorrelated1D.jl:103)

    def __init__(self, address, t=1.0, v=1.0, v_ho=0.0, cutoff=1, three_body_term=True):
        if not isinstance(address, CompositeFS):
            raise TypeError("Transcorrelated1D needs a two-component fermionic address")
        if cutoff < 1:
            raise ValueError("`cutoff` must be a positive integer")
        if v_ho != 0:
            raise NotImplementedError("v_ho != 0 (momentum-space harmonic trap) is outside the device path")
        M = address.num_modes
        if M > _lib.MAX_TABLE_MODES:
            raise ValueError(f"momentum-space models support at most {_lib.MAX_TABLE_MODES} modes")
        self.t, self.v, self.cutoff, self.three_body_term = float(t), float(v), int(cutoff), bool(three_body_term)
        i_to_n = lambda i: i - M // 2 - (M % 2)
        self.ks = np.array([i_to_n(i) * 2 * math.pi / M for i in range(1, M + 1)])
        self.kes = self.t * self.ks ** 2
        self.ws = np.array([_w_function(n, self.cutoff) for n in range(M)])
        self.us = np.array([(-1 / (2 * (n * 2 * math.pi / M))) if abs(n) >= self.cutoff else 0.0 for n in range(1, M + 1)])
        d = self.desc = _lib.HamDesc()
        d.model, d.t, d.v, d.cutoff, d.three_body_term = _lib.TRANSCORRELATED_1D, self.t, self.v, self.cutoff, int(self.three_body_term)
        for i in range(M):
            d.kes[i], d.ws[i], d.us[i] = self.kes[i], self.ws[i], self.us[i]
        self._finish(address)

    def __repr__(self):
        return f"Transcorrelated1D({self.address}, t={self.t}, v={self.v})"


def _w_function(n, nc):  # Transcorrelated1D.jl:164-178
    prefactor = -1 / (8 * math.pi ** 2)
    n = abs(n)
    if n == 0:
        x = math.pi ** 2 / 6 - sum(1 / (q * q) for q in range(1, nc))
    elif 2 * nc > n > 0:
        x = 1 / n * sum(1 / q for q in range(nc, n + nc))
    else:
        x = 1 / n * sum(1 / q for q in range(nc, n + nc)) - 0.5 * sum(1 / (q * (n - q)) for q in range(nc, n - nc + 1))
    return prefactor * x
