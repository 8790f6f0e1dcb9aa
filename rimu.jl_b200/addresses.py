"""Fock-state addresses: host mirror of Rimu's BitStringAddresses for the device key layout.

Mirrors `BoseFS{N,M}` (BitStringAddresses/bosefs.jl:54-75), `FermiFS{N,M}` (fermifs.jl:54-56),
`CompositeFS` / `FermiFS2C` (multicomponent.jl:10-19,125-126), `near_uniform` (bosefs.jl:151-168)
and `onr`.  The packed form is the W x uint64 interchange format documented in
include/rimu_b200.h (same bit order as bitstring.jl:464-472 / :713-723).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

from . import _lib


@dataclass(frozen=True)
class AddressType:
    """Static type information of an address (the Julia type parameters)."""
    kind: int                      # _lib.ADDR_*
    num_particles: Tuple[int, ...]
    num_modes: int
    comp_kinds: Tuple[int, ...] = ()   # ADDR_COMPOSITE: ADDR_BOSE / ADDR_FERMI per component

    @property
    def num_components(self):
        return len(self.num_particles)

    @property
    def comp_bits(self):
        """Bits of every component: N + M - 1 for a BoseFS component, M for a FermiFS component."""
        M = self.num_modes
        if self.kind == _lib.ADDR_BOSE:
            return (self.num_particles[0] + M - 1,)
        if self.kind == _lib.ADDR_COMPOSITE:
            return tuple(n + M - 1 if k == _lib.ADDR_BOSE else M for k, n in zip(self.comp_kinds, self.num_particles))
        return (M,) * self.num_components

    @property
    def bits(self):
        return sum(self.comp_bits)

    @property
    def words(self):
        b = self.bits + 1 if self.kind in (_lib.ADDR_BOSE, _lib.ADDR_COMPOSITE) else self.bits
        return (b + 63) // 64

    def from_key(self, key):
        """Decode W uint64 words into an address object."""
        x = 0
        for j, w in enumerate(key):
            x |= int(w) << (64 * j)
        M = self.num_modes
        if self.kind == _lib.ADDR_BOSE:
            onr, mode, n = [0] * M, 0, 0
            for pos in range(self.bits):
                if (x >> pos) & 1:
                    onr[mode] += 1
                else:
                    mode += 1
            return BoseFS(tuple(onr))
        if self.kind == _lib.ADDR_COMPOSITE:
            comps, off = [], 0
            for k, n, b in zip(self.comp_kinds, self.num_particles, self.comp_bits):
                sub = AddressType(k, (n,), M)
                xc = (x >> off) & ((1 << b) - 1)
                comps.append(sub.from_key([(xc >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(max(sub.words, 1))]))
                off += b
            return CompositeFS(*comps)
        comps = []
        for c in range(self.num_components):
            comps.append(FermiFS(tuple((x >> (c * M + m)) & 1 for m in range(M))))
        return comps[0] if self.kind == _lib.ADDR_FERMI else CompositeFS(*comps)


def _as_onr(args):
    if len(args) == 1 and hasattr(args[0], "__iter__"):
        return tuple(int(v) for v in args[0])
    return tuple(int(v) for v in args)


class _SingleComponent:
    onr: Tuple[int, ...]

    @property
    def num_modes(self):
        return len(self.onr)

    @property
    def num_particles(self):
        return sum(self.onr)

    def __eq__(self, other):
        return type(self) is type(other) and self.onr == other.onr

    def __hash__(self):
        return hash((type(self).__name__, self.onr))

    def __repr__(self):
        return f"{type(self).__name__}{{{self.num_particles},{self.num_modes}}}{self.onr}"

    def occupied_modes(self):
        """(occnum, mode) for occupied modes, ascending (OccupiedModeMap, fockaddress.jl:258-275)."""
        return [(n, i + 1) for i, n in enumerate(self.onr) if n > 0]

    def key(self):
        x = self._bits()
        W = self.address_type.words
        return tuple((x >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(W))


class BoseFS(_SingleComponent):
    """BoseFS(onr...) or BoseFS.from_pairs(M, {mode: n}) (Julia: BoseFS(M, mode => n))."""

    def __init__(self, *args):
        self.onr = _as_onr(args)
        if any(n < 0 for n in self.onr):
            raise ValueError("occupation numbers must be non-negative")

    @classmethod
    def from_pairs(cls, M, pairs):
        onr = [0] * M
        for mode, n in dict(pairs).items():
            onr[mode - 1] += n
        return cls(tuple(onr))

    @property
    def address_type(self):
        return AddressType(_lib.ADDR_BOSE, (self.num_particles,), self.num_modes)

    def _bits(self):
        x, pos = 0, 0
        for n in self.onr:
            x |= ((1 << n) - 1) << pos
            pos += n + 1
        return x


class FermiFS(_SingleComponent):
    """FermiFS(onr...) with occupation numbers 0/1, or FermiFS.from_modes(M, modes)."""

    def __init__(self, *args):
        self.onr = _as_onr(args)
        if any(n not in (0, 1) for n in self.onr):
            raise ValueError("fermionic occupation numbers must be 0 or 1")

    @classmethod
    def from_modes(cls, M, modes):
        onr = [0] * M
        for m in modes:
            onr[m - 1] = 1
        return cls(tuple(onr))

    @property
    def address_type(self):
        return AddressType(_lib.ADDR_FERMI, (self.num_particles,), self.num_modes)

    def _bits(self):
        x = 0
        for m, n in enumerate(self.onr):
            x |= n << m
        return x


class CompositeFS:
    """CompositeFS(components...) (multicomponent.jl:10-34): BoseFS / FermiFS components over the same modes.

    Two FermiFS components of at most 32 modes are the one-word `FermiFS2C` layout every two-component model accepts;
    anything else (bosonic or mixed components, more than two components, wider fermions) is the general packed layout
    (`_lib.ADDR_COMPOSITE`, components side by side from the low bits) that `HubbardRealSpace` accepts."""

    def __init__(self, *components):
        if len(components) < 1 or not all(isinstance(c, (BoseFS, FermiFS)) for c in components):
            raise TypeError("the components of a CompositeFS must be BoseFS or FermiFS addresses")
        if len({c.num_modes for c in components}) != 1:
            raise ValueError("all addresses must have the same number of modes")
        self.components = tuple(components)

    @property
    def num_modes(self):
        return self.components[0].num_modes

    @property
    def num_particles(self):
        return sum(c.num_particles for c in self.components)

    @property
    def onr(self):
        return tuple(c.onr for c in self.components)

    @property
    def is_fermi2c(self):
        return (len(self.components) == 2 and all(isinstance(c, FermiFS) for c in self.components)
                and self.num_modes <= 32)

    @property
    def address_type(self):
        nps = tuple(c.num_particles for c in self.components)
        if self.is_fermi2c:
            return AddressType(_lib.ADDR_FERMI2C, nps, self.num_modes)
        kinds = tuple(_lib.ADDR_BOSE if isinstance(c, BoseFS) else _lib.ADDR_FERMI for c in self.components)
        return AddressType(_lib.ADDR_COMPOSITE, nps, self.num_modes, kinds)

    def key(self):
        at = self.address_type
        x, off = 0, 0
        for c, b in zip(self.components, at.comp_bits):
            x |= c._bits() << off
            off += b
        return tuple((x >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(at.words))

    def __eq__(self, other):
        return isinstance(other, CompositeFS) and self.components == other.components

    def __hash__(self):
        return hash(self.components)

    def __repr__(self):
        return "CompositeFS(" + ", ".join(repr(c) for c in self.components) + ")"


def FermiFS2C(onr_a, onr_b):
    """FermiFS2C(onr_up, onr_down) (multicomponent.jl:125-126)."""
    return CompositeFS(FermiFS(tuple(onr_a)), FermiFS(tuple(onr_b)))


def near_uniform_onr(n, m):
    fill, extras = divmod(n, m)
    return tuple(fill + (1 if i < extras else 0) for i in range(m))


def near_uniform(cls, n, m):
    """near_uniform(BoseFS{N,M}) (bosefs.jl:151-168): call as near_uniform(BoseFS, N, M)."""
    if cls is BoseFS:
        return BoseFS(near_uniform_onr(n, m))
    if cls is FermiFS:
        return FermiFS(tuple(1 if i < n else 0 for i in range(m)))
    raise TypeError("near_uniform is defined for BoseFS and FermiFS")


def onr(addr):
    return addr.onr


def num_modes(addr):
    return addr.num_modes


def num_particles(addr):
    return addr.num_particles
