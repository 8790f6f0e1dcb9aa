"""GPU-resident walker vectors: host mirror of Rimu's DictVectors for the device path.

`GPUDVec` is the `AbstractDVec` the FCIQMC driver sees (Interfaces/dictvectors.jl:22-56).  It
fills the roles of `DVec` (DictVectors/dvec.jl:44-47) and `PDVec` (pdvec.jl:156-163): both names
are exported as aliases.  Storage is a dense (keys, values) pair of HBM arrays owned by
librimu_b200.so; every operation below is a C-ABI call -- nothing is computed on the host.

Julia name -> Python name:  apply_operator! -> apply_operator,  mul! -> mul,  scale! -> scale_,
add!/axpy! -> add_,  zerovector -> zerovector,  walkernumber_and_length -> walkernumber_and_length.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .addresses import AddressType
from .hamiltonians import AbstractHamiltonian, Context, get_context
from .stochasticstyles import (IsDeterministic, IsStochasticInteger, StochasticStyle, ThresholdCompression,
                               as_initiator_rule, default_style, step_stats)


class FirstOrderTransitionOperator:
    """T = 1 + dτ (S - H)  (fciqmc.jl:78-112)."""

    def __init__(self, hamiltonian, shift, time_step):
        self.hamiltonian, self.shift, self.time_step = hamiltonian, float(shift), float(time_step)


class GPUDVec:
    """Dictionary-semantics vector (missing -> 0, zeros never stored) living on the GPU."""

    def __init__(self, pairs=None, *, style: StochasticStyle | None = None, address_type: AddressType | None = None,
                 capacity: int = 1 << 12, ctx: Context | None = None, initiator=None, initiator_threshold=None):
        items = list(pairs.items()) if isinstance(pairs, dict) else list(pairs or [])
        if address_type is None:
            if not items:
                raise ValueError("an empty GPUDVec needs an explicit address_type")
            address_type = items[0][0].address_type
        if style is None:
            style = default_style(int if items and all(isinstance(v, (int, np.integer)) for _, v in items) else float)
        self.style, self.address_type = style, address_type
        self.initiator = as_initiator_rule(initiator, initiator_threshold)  # PDVec(...; initiator=...) pdvec.jl:181-199
        self.ctx = ctx or get_context(address_type.words)
        h = C.c_void_p()
        _lib.check(_lib.lib().rimu_vec_create(self.ctx.handle, style.val_type, max(capacity, len(items)), C.byref(h)))
        self.handle = h
        if items:
            keys = np.array([a.key() for a, _ in items], dtype=np.uint64).reshape(-1, self.words)
            vals = np.array([v for _, v in items], dtype=self.dtype)
            self.upload(keys, vals)

    # ---- basics
    @property
    def words(self):
        return self.address_type.words

    @property
    def dtype(self):
        return np.int64 if self.style.val_type == _lib.VAL_I64 else np.float64

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().rimu_vec_destroy(self.handle)
        except Exception:
            pass

    def __len__(self):
        """Number of stored (non-zero) entries on this rank: length(localpart(v))."""
        n = C.c_int64()
        _lib.check(_lib.lib().rimu_vec_length(self.handle, C.byref(n)))
        return n.value

    def similar(self, style=None):
        return GPUDVec(style=style or self.style, address_type=self.address_type, capacity=max(len(self), 256), ctx=self.ctx,
                       initiator=self.initiator)

    zerovector = similar
    empty = similar

    def copy(self):
        out = self.similar()
        _lib.check(_lib.lib().rimu_vec_copy(out.handle, self.handle))
        return out

    def copy_from(self, other):
        _lib.check(_lib.lib().rimu_vec_copy(self.handle, other.handle))
        return self

    def clear(self):
        _lib.check(_lib.lib().rimu_vec_clear(self.handle))
        return self

    # ---- host <-> device
    def _with_table_retry(self, fn, max_retries=6):
        """RIMU_ERR_TABLE_FULL means the global HBM table of the table method was too small: grow it (x4) and repeat, a
        bounded number of times; every other status is raised as it is."""
        for attempt in range(max_retries + 1):
            try:
                return fn()
            except _lib.RimuB200Error as e:
                if e.status != _lib.ERR_TABLE_FULL or attempt == max_retries:
                    raise
                self.ctx.resize_table(self.ctx.table_slots * 4)

    def upload(self, keys, vals):
        """DVec(pairs...): duplicates summed, zeros dropped, non-local keys dropped."""
        keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64).reshape(-1, self.words))
        vals = np.ascontiguousarray(np.asarray(vals, dtype=self.dtype))
        self._with_table_retry(lambda: _lib.check(_lib.lib().rimu_vec_upload(
            self.handle, keys.ctypes.data_as(_lib._u64p), vals.ctypes.data_as(C.c_void_p), keys.shape[0])))
        return self

    def assign(self, keys, vals):
        """copyto! of distinct non-zero pairs without deduplication."""
        keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64).reshape(-1, self.words))
        vals = np.ascontiguousarray(np.asarray(vals, dtype=self.dtype))
        _lib.check(_lib.lib().rimu_vec_assign(self.handle, keys.ctypes.data_as(_lib._u64p), vals.ctypes.data_as(C.c_void_p), keys.shape[0]))
        return self

    def download(self):
        """(keys[n, W], vals[n]) of the local part, unspecified order."""
        n = len(self)
        keys = np.zeros((n, self.words), dtype=np.uint64)
        vals = np.zeros(n, dtype=self.dtype)
        m = C.c_int64()
        _lib.check(_lib.lib().rimu_vec_download(self.handle, keys.ctypes.data_as(_lib._u64p), vals.ctypes.data_as(C.c_void_p), n, C.byref(m)))
        return keys, vals

    def download_sorted(self):
        keys, vals = self.download()
        order = np.lexsort(tuple(keys[:, j] for j in range(self.words)))
        return keys[order], vals[order]

    def pairs(self):
        keys, vals = self.download_sorted()
        return [(self.address_type.from_key(k), v.item()) for k, v in zip(keys, vals)]

    def to_dict(self):
        return dict(self.pairs())

    def __getitem__(self, addr):
        key = np.array(addr.key(), dtype=np.uint64)
        out = np.zeros(1, dtype=self.dtype)
        _lib.check(_lib.lib().rimu_vec_get(self.handle, key.ctypes.data_as(_lib._u64p), out.ctypes.data_as(C.c_void_p)))
        return out[0].item()

    def __setitem__(self, addr, value):
        delta = value - self[addr]
        if delta != 0:
            self.add_(GPUDVec([(addr, delta)], style=self.style, address_type=self.address_type, ctx=self.ctx))

    def deposit(self, addr, value):
        """deposit!(w, add, val, parent) (Interfaces/dictvectors.jl:49-51)."""
        self[addr] = self[addr] + value

    # ---- linear algebra (VectorInterface subset used by the drivers)
    def norm(self, p=2):
        out = C.c_double()
        code = 0 if p in (math.inf, "inf", 0) else int(p)
        _lib.check(_lib.lib().rimu_vec_norm(self.handle, code, C.byref(out)))
        return out.value

    def walkernumber(self):
        """Norm1ProjectorPPop ⋅ v (abstractdvec.jl:258-260)."""
        return self.norm(1)

    def dot(self, other):
        if isinstance(other, FrozenDVec):
            return other.dot(self)
        out = C.c_double()
        self._with_table_retry(lambda: _lib.check(_lib.lib().rimu_vec_dot(self.handle, other.handle, C.byref(out))))
        return out.value

    def scale_(self, alpha):
        _lib.check(_lib.lib().rimu_vec_scale(self.handle, float(alpha)))
        return self

    def add_(self, other, alpha=1.0):
        """add!(self, other, alpha): self += alpha * other."""
        self._with_table_retry(lambda: _lib.check(_lib.lib().rimu_vec_axpby(float(alpha), other.handle, 1.0, self.handle, self.handle)))
        return self

    def axpby_(self, alpha, x, beta):
        """self = alpha*x + beta*self."""
        self._with_table_retry(lambda: _lib.check(_lib.lib().rimu_vec_axpby(float(alpha), x.handle, float(beta), self.handle, self.handle)))
        return self

    def __mul__(self, alpha):
        return self.copy().scale_(alpha)

    __rmul__ = __mul__

    def __add__(self, other):
        return self.copy().add_(other)

    def __sub__(self, other):
        return self.copy().add_(other, -1.0)

    def __neg__(self):
        return self.copy().scale_(-1.0)

    def __truediv__(self, alpha):
        return self.copy().scale_(1.0 / alpha)

    def normalize_(self, p=2):
        """normalize!(v, p) (abstractdvec.jl:200-256)"""
        n = self.norm(p)
        return self.scale_(1.0 / n) if n != 0 else self

    def normalize(self, p=2):
        return self.copy().normalize_(p)

    # ---- iteration and reductions over the local part (test/DictVectors.jl:193-249).  These take arbitrary host
    # callables, so they work on a downloaded copy of the (address, value) pairs; nothing on the step path uses them.
    def keys(self):
        return [a for a, _ in self.pairs()]

    def values(self):
        return [x for _, x in self.pairs()]

    def __iter__(self):
        return iter(self.pairs())

    def __contains__(self, addr):
        return self[addr] != 0

    def get(self, addr, default=0):
        x = self[addr]
        return x if x != 0 else default

    def mapreduce(self, f, op, init=None, over="values"):
        """mapreduce(f, op, values(v) | keys(v) | pairs(v); init) on the local part"""
        items = {"values": self.values, "keys": self.keys, "pairs": self.pairs}[over]()
        import functools
        mapped = [f(x) for x in items]
        return functools.reduce(op, mapped) if init is None else functools.reduce(op, mapped, init)

    def sum(self, f=None, over="values"):
        return self.mapreduce(f or (lambda x: x), lambda a, b: a + b, init=0, over=over)

    def all(self, f, over="values"):
        return all(f(x) for x in {"values": self.values, "keys": self.keys, "pairs": self.pairs}[over]())

    def any(self, f, over="values"):
        return any(f(x) for x in {"values": self.values, "keys": self.keys, "pairs": self.pairs}[over]())

    def __eq__(self, other):
        return isinstance(other, GPUDVec) and self.to_dict() == other.to_dict()

    __hash__ = object.__hash__

    def __repr__(self):
        return f"GPUDVec{{{type(self.style).__name__}}} with {len(self)} entries on the device"

    def freeze(self):
        """freeze(v) (projectors.jl:164): an immutable list of (address, value) pairs on the HOST; its dot with a device
        vector looks every address up in its bucket segment (rimu_vec_dot_sparse) instead of building a hash table."""
        keys, vals = self.download()
        return FrozenDVec(keys, np.asarray(vals, dtype=np.float64), self.address_type, self.ctx)


class FrozenDVec:
    """FrozenDVec (DictVectors/projectors.jl:141-176): used as the projector of ProjectedEnergy / Projector."""

    def __init__(self, keys, vals, address_type, ctx=None):
        self.address_type = address_type
        self.keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64).reshape(-1, address_type.words))
        self.vals = np.ascontiguousarray(np.asarray(vals, dtype=np.float64))
        self.ctx = ctx

    def __len__(self):
        return len(self.vals)

    def dot(self, v: "GPUDVec") -> float:
        """dot(::FrozenDVec, ::PDVec) (pdvec.jl:773-779); global over ranks."""
        out = C.c_double()
        _lib.check(_lib.lib().rimu_vec_dot_sparse(v.handle, self.keys.ctypes.data_as(_lib._u64p),
                                                  self.vals.ctypes.data_as(_lib._f64p), len(self.vals), C.byref(out)))
        return out.value


DVec = GPUDVec
PDVec = GPUDVec


def InitiatorDVec(pairs=None, *, initiator=None, **kw):
    """InitiatorDVec(pairs...; initiator=Initiator(1), style, ...) (DictVectors/initiatordvec.jl:42-77)."""
    from .stochasticstyles import Initiator
    return GPUDVec(pairs, initiator=Initiator(1.0) if initiator is None else initiator, **kw)


class WorkingMemory:
    """working_memory(v) (Interfaces/dictvectors.jl:87, pdworkingmemory.jl:295): the context's
    HBM working table plus the Philox stream position (seed, call counter)."""

    def __init__(self, v: GPUDVec, seed: int = 0, ordered: bool = False):
        self.ctx, self.style, self.seed, self.counter = v.ctx, v.style, int(seed) & 0xFFFFFFFFFFFFFFFF, 0
        self.ordered = bool(ordered)  # order-deterministic Float64 summation (rimu_step_params.ordered): bit-reproducible steps
        self.initiator = v.initiator  # PDWorkingMemory(t).initiator = t.initiator (pdworkingmemory.jl:104-108)
        self.last_stats: _lib.StepStats | None = None


def working_memory(v: GPUDVec, seed: int = 0, ordered: bool = False) -> WorkingMemory:
    return WorkingMemory(v, seed, ordered)


def walkernumber_and_length(v: GPUDVec):
    return v.walkernumber(), len(v)


def apply_operator(wm: WorkingMemory, target: GPUDVec, source: GPUDVec, op, boost=1.0, table_slots=0):
    """apply_operator!(working_memory, target, source, operator, boost)
    -> (stat_names, stats, working_memory, target)   (Interfaces/dictvectors.jl:90-140,
    PDVec method pdworkingmemory.jl:297-309).  `source` must not alias `target`."""
    if target is source:
        raise ValueError("source and target must not alias")
    style = wm.style
    p = _lib.StepParams()
    style.fill(p)
    if isinstance(op, FirstOrderTransitionOperator):
        ham, p.plain_h, p.shift, p.time_step = op.hamiltonian, 0, op.shift, op.time_step
    elif isinstance(op, AbstractHamiltonian):
        ham, p.plain_h, p.shift, p.time_step = op, 1, 0.0, 0.0
    else:
        raise TypeError("operator must be one of the device Hamiltonians or a FirstOrderTransitionOperator "
                        "(custom Julia/Python operators cannot run on the GPU; there is no CPU fallback)")
    p.boost, p.seed, p.step, p.table_slots = float(boost), wm.seed, wm.counter, int(table_slots)
    p.ordered = int(getattr(wm, "ordered", False))
    rule = getattr(wm, "initiator", None)
    if rule is not None and rule.rule_id:
        p.initiator_rule, p.initiator_threshold = rule.rule_id, float(rule.threshold)
    stats = _lib.StepStats()
    ctx = wm.ctx
    table_retries = 0
    while True:
        st = _lib.lib().rimu_step(ctx.handle, ham.handle, C.byref(p), source.handle, target.handle, C.byref(stats))
        if st == _lib.ERR_TABLE_FULL and table_retries < 6:
            # the table method ran out of slots: grow the global table and repeat.  (Failures of the partitioned method's
            # record streams come back as ERR_WORKMEM and are raised below: a bigger table would not help them.)
            table_retries += 1
            ctx.resize_table(ctx.table_slots * 4)
            continue
        if st == _lib.ERR_EXCHANGE_FULL:  # same decision on every rank: grow the per-peer buffers and repeat
            cap, need = C.c_uint64(), C.c_uint64()
            _lib.check(_lib.lib().rimu_comm_capacity(ctx.handle, C.byref(cap), C.byref(need)))
            _lib.check(_lib.lib().rimu_comm_reserve(ctx.handle, max(2 * cap.value, int(1.5 * need.value))))
            continue
        _lib.check(st)
        break
    wm.counter += 1
    wm.last_stats = stats
    names, values = style.stat_names, style.stats(stats)
    if isinstance(getattr(style, "compression", None), ThresholdCompression):
        names, values = names + ("len_before",), values + (stats.len_before,)
    return names, values, wm, target


def advance(wm: WorkingMemory, v: GPUDVec, pv: GPUDVec, hamiltonian, shift_params, strategy_id, *, target_walkers=0.0, zeta=0.0,
            xi=0.0, nsteps=1, max_length=0, boost=1.0, projectors=()):
    """`nsteps` x { apply_operator!(wm, pv, v, FirstOrderTransitionOperator(H, shift, dt)); v, pv = pv, v;
    update_shift_parameters! } -- the body of advance! (fciqmc.jl:126-181) -- in ONE call: the shift update and the abort
    rules run on the device, so no step waits for the host (rimu_advance, include/rimu_b200.h).
    `shift_params` (fciqmc.ShiftParameters) is updated in place.
    `projectors`: FrozenDVecs whose dot with the vector is evaluated on the device after every step (the reports of
    ProjectedEnergy / Projector, poststepstrategy.jl:50-121).
    -> (v, pv, stats [StepStats per step taken], shifts [shift after each step's update], steps_done) and, with projectors,
    a sixth element: array [steps_done, len(projectors)] of dots.  steps_done < nsteps only when the run ended (dead
    population, max_length, DontUpdate target reached)."""
    style = wm.style
    p = _lib.StepParams()
    style.fill(p)
    p.plain_h, p.shift, p.time_step = 0, shift_params.shift, shift_params.time_step
    p.boost, p.seed, p.step = float(boost), wm.seed, wm.counter
    p.ordered = int(getattr(wm, "ordered", False))
    rule = getattr(wm, "initiator", None)
    if rule is not None and rule.rule_id:
        p.initiator_rule, p.initiator_threshold = rule.rule_id, float(rule.threshold)
    sp = _lib.ShiftParams()
    sp.strategy, sp.shift_mode = int(strategy_id), int(bool(shift_params.shift_mode))
    sp.target_walkers, sp.zeta, sp.xi = float(target_walkers), float(zeta), float(xi)
    sp.shift, sp.pnorm, sp.max_length = float(shift_params.shift), float(shift_params.pnorm), int(max_length)
    stats = (_lib.StepStats * nsteps)()
    shifts = (C.c_double * nsteps)()
    done, in_w = C.c_int64(0), C.c_int32(0)
    projectors = list(projectors)
    nproj = len(projectors)
    parr = (_lib.Projector * max(nproj, 1))()
    for j, fr in enumerate(projectors):
        parr[j].keys, parr[j].values, parr[j].n = fr.keys.ctypes.data_as(_lib._u64p), fr.vals.ctypes.data_as(_lib._f64p), len(fr.vals)
    pout = np.zeros((nsteps, max(nproj, 1)), dtype=np.float64)
    status = _lib.lib().rimu_advance(wm.ctx.handle, hamiltonian.handle, C.byref(p), C.byref(sp), v.handle, pv.handle, nsteps,
                                     parr if nproj else None, nproj, stats, shifts, pout.ctypes.data_as(_lib._f64p) if nproj else None,
                                     C.byref(done), C.byref(in_w))
    # the steps taken so far stay taken, also when a later one failed: bring the host's bookkeeping up to date before raising
    wm.counter += done.value
    shift_params.shift, shift_params.pnorm, shift_params.shift_mode = sp.shift, sp.pnorm, bool(sp.shift_mode)
    if done.value:
        wm.last_stats = stats[done.value - 1]
    if status != 0 and in_w.value:
        v.handle, pv.handle = pv.handle, v.handle  # no return value on this path: the caller's `v` object stays the current vector
    _lib.check(status)
    if in_w.value:
        v, pv = pv, v
    out = (v, pv, [stats[k] for k in range(done.value)], [shifts[k] for k in range(done.value)], done.value)
    return out + (pout[:done.value, :nproj],) if nproj else out


def mul(y: GPUDVec, op, x: GPUDVec, wm: WorkingMemory | None = None):
    """mul!(y, op, x, w) (pdvec.jl:810-822): deterministic y = op * x.  Dense sector vectors (sectors.py) take the gather path."""
    if hasattr(y, "mul_from"):
        return y.mul_from(op, x)
    wm = wm or WorkingMemory(GPUDVec(style=IsDeterministic(), address_type=x.address_type, ctx=x.ctx))
    if not isinstance(wm.style, IsDeterministic):
        raise ValueError("Attempted to use `mul!` with non-deterministic working memory. "
                         "Use `apply_operator!` instead.")  # ArgumentError pdvec.jl:814-819
    if not isinstance(y.style, IsDeterministic) or not isinstance(x.style, IsDeterministic):
        raise ValueError("mul! needs IsDeterministic (Float64) vectors")
    apply_operator(wm, y, x, op)
    return y


def dot(x: GPUDVec, *args):
    """dot(x, y) or dot(x, op, y) (abstractdvec.jl:286-324, pdvec.jl:833-894)."""
    if len(args) == 1:
        return x.dot(args[0])  # either side may be a FrozenDVec
    op, y = args
    tmp = y.similar(style=IsDeterministic())
    yy = y if isinstance(y.style, IsDeterministic) else GPUDVec(style=IsDeterministic(), address_type=y.address_type, ctx=y.ctx).copy_from(y)
    mul(tmp, op, yy)
    xx = x if x.style.val_type == _lib.VAL_F64 else GPUDVec(style=IsDeterministic(), address_type=x.address_type, ctx=x.ctx).copy_from(x)
    return xx.dot(tmp)
