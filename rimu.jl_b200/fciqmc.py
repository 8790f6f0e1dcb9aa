"""FCIQMC driver: host mirror of ProjectorMonteCarloProblem / init / step! / solve.

This is the thin host loop that CALLS the drop-in boundary (`apply_operator`), restated in
Python because Julia is not available in this image; in a Julia deployment Rimu's own
`advance!` (fciqmc.jl:126-181) does this job unchanged.  Mirrors
  ProjectorMonteCarloProblem kwargs/defaults   projector_monte_carlo_problem.jl:150-290
  PMCSimulation / init / step! / solve!         pmc_simulation.jl:88-174,265-452
  advance!(::FCIQMC)                            fciqmc.jl:126-181
  shift strategies                              strategies_and_params/shiftstrategy.jl:100-230
  ProjectedEnergy / Projector post-steps        strategies_and_params/poststepstrategy.jl:50-121
  default_starting_vector                       qmc_states.jl:236-260
Only O(10) scalars per step cross the C ABI; all vector work runs on the GPU.
"""
from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .dictvectors import (FirstOrderTransitionOperator, GPUDVec, WorkingMemory, advance, apply_operator, dot, mul,
                          walkernumber_and_length)
from .hamiltonians import AbstractHamiltonian, starting_address
from .stochasticstyles import IsDeterministic, IsDynamicSemistochastic, StochasticStyle, default_style


# --------------------------------------------------------------------------- shift strategies
@dataclass
class ShiftParameters:
    """DefaultShiftParameters (shiftstrategy.jl:32-38)"""
    shift: float
    pnorm: float
    time_step: float
    counter: int = 0
    shift_mode: bool = False


@dataclass
class DontUpdate:
    """shiftstrategy.jl:77-91: the shift stays fixed; the run stops once the walker number reaches `target_walkers`
    (`proceed = tnorm < target_walkers`).  `pnorm` is left alone, as in the reference."""
    target_walkers: float = 1000

    def update(self, sp, tnorm):
        return {"shift": sp.shift, "norm": tnorm}, tnorm < self.target_walkers


@dataclass
class LogUpdate:
    zeta: float = 0.08

    def update(self, sp, tnorm):
        sp.shift -= self.zeta / sp.time_step * math.log(tnorm / sp.pnorm)
        sp.pnorm = tnorm
        return {"shift": sp.shift, "norm": tnorm}, True


@dataclass
class LogUpdateAfterTargetWalkers:
    """shiftstrategy.jl:100-122: LogUpdate that switches on once the walker number exceeds the target."""
    target_walkers: float = 1000
    zeta: float = 0.08

    def update(self, sp, tnorm):
        if sp.shift_mode or tnorm > self.target_walkers:
            sp.shift_mode = True
            sp.shift -= self.zeta / sp.time_step * math.log(tnorm / sp.pnorm)
        sp.pnorm = tnorm
        return {"shift": sp.shift, "norm": tnorm, "shift_mode": sp.shift_mode}, True


@dataclass
class DoubleLogUpdate:
    """shiftstrategy.jl:160-181"""
    target_walkers: float = 1000
    zeta: float = 0.08
    xi: float | None = None

    def __post_init__(self):
        if self.xi is None:
            self.xi = self.zeta ** 2 / 4

    def update(self, sp, tnorm):
        dt = sp.time_step
        sp.shift -= self.xi / dt * math.log(tnorm / self.target_walkers) + self.zeta / dt * math.log(tnorm / sp.pnorm)
        sp.pnorm = tnorm
        return {"shift": sp.shift, "norm": tnorm}, True


@dataclass
class DoubleLogUpdateAfterTargetWalkers:
    """shiftstrategy.jl:190-213"""
    target_walkers: float = 1000
    zeta: float = 0.08
    xi: float | None = None

    def __post_init__(self):
        if self.xi is None:
            self.xi = self.zeta ** 2 / 4

    def update(self, sp, tnorm):
        if sp.shift_mode or tnorm > self.target_walkers:
            sp.shift_mode = True
            dt = sp.time_step
            sp.shift -= self.xi / dt * math.log(tnorm / self.target_walkers) + self.zeta / dt * math.log(tnorm / sp.pnorm)
        sp.pnorm = tnorm
        return {"shift": sp.shift, "norm": tnorm, "shift_mode": sp.shift_mode}, True


def _device_strategy(strategy):
    """(RIMU_SHIFT_* id, target_walkers, zeta, xi) of a shift strategy the device controller implements, or None"""
    t = type(strategy)
    if t is DontUpdate:
        return _lib.SHIFT_DONT_UPDATE, strategy.target_walkers, 0.0, 0.0
    if t is LogUpdate:
        return _lib.SHIFT_LOG_UPDATE, 0.0, strategy.zeta, 0.0
    if t is LogUpdateAfterTargetWalkers:
        return _lib.SHIFT_LOG_UPDATE_AFTER_TARGET, strategy.target_walkers, strategy.zeta, 0.0
    if t is DoubleLogUpdate:
        return _lib.SHIFT_DOUBLE_LOG_UPDATE, strategy.target_walkers, strategy.zeta, strategy.xi
    if t is DoubleLogUpdateAfterTargetWalkers:
        return _lib.SHIFT_DOUBLE_LOG_UPDATE_AFTER_TARGET, strategy.target_walkers, strategy.zeta, strategy.xi
    return None


# --------------------------------------------------------------------------- post-step strategies
FREEZE_LIMIT = 1 << 16  # projectors up to this many entries are frozen (host pairs + per-key segment lookups)


def _frozen_or_device(vec):
    """(multi-GPU: every rank freezes its local part; both kinds of dot end in exactly one all-reduce, so ranks may differ)"""
    return vec.freeze() if len(vec) <= FREEZE_LIMIT else vec


class ProjectedEnergy:
    """poststepstrategy.jl:82-121: reports vproj = projector⋅v and hproj = (H' projector)⋅v, or
    dot(projector, H, v) when the adjoint is unknown (Transcorrelated1D).  Both projectors are frozen
    (`freeze`, projectors.jl:164), so a report costs a few bucket-segment lookups, not a pass over the vector."""

    def __init__(self, hamiltonian, projector: GPUDVec, vproj="vproj", hproj="hproj"):
        self.ham, self.vproj_name, self.hproj_name = hamiltonian, vproj, hproj
        det = IsDeterministic()
        self._vdev = GPUDVec(style=det, address_type=projector.address_type, ctx=projector.ctx).copy_from(projector)
        self.vproj = _frozen_or_device(self._vdev)
        if hamiltonian.hermitian:
            self.hproj = _frozen_or_device(mul(self._vdev.similar(), hamiltonian, self._vdev))  # H' = H
        else:
            self.hproj = None

    def __call__(self, state, step):
        v = state.v
        out = {self.vproj_name: _pdot(self.vproj, v)}
        if self.hproj is not None:
            out[self.hproj_name] = _pdot(self.hproj, v)
        else:
            vf = v if v.style.val_type == _lib.VAL_F64 else GPUDVec(style=IsDeterministic(), address_type=v.address_type, ctx=v.ctx).copy_from(v)
            out[self.hproj_name] = dot(self._vdev, self.ham, vf)
        return out


def _pdot(projector, v):
    """projector ⋅ v for a frozen (any value type of v) or device-resident projector (Float64 against Float64)."""
    from .dictvectors import FrozenDVec
    if isinstance(projector, FrozenDVec):
        return projector.dot(v)
    vf = v if v.style.val_type == _lib.VAL_F64 else GPUDVec(style=IsDeterministic(), address_type=v.address_type, ctx=v.ctx).copy_from(v)
    return projector.dot(vf)


class Projector:
    def __init__(self, **kw):
        (self.name, proj), = kw.items()
        dev = GPUDVec(style=IsDeterministic(), address_type=proj.address_type, ctx=proj.ctx).copy_from(proj)
        self.projector = _frozen_or_device(dev)

    def __call__(self, state, step):
        return {self.name: _pdot(self.projector, state.v)}


class AllOverlaps:
    """AllOverlaps(n_replicas=2; operator=nothing, vecnorm=true) (replicastrategy.jl:60-183): after every step report
    `c{i}_dot_c{j}` = dot(v_i, v_j) and, for each operator k, `c{i}_Op{k}_c{j}` = dot(v_i, Op_k, v_j) for all replica pairs.
    Operators are device Hamiltonians (three-argument dot = mul! + dot on the GPU).  Transformed Hamiltonians
    (`transform=`) are not supported."""

    def __init__(self, n_replicas=2, operator=None, vecnorm=True):
        if not isinstance(n_replicas, int):
            raise TypeError("n_replicas must be an integer")
        self.n_replicas, self.vecnorm = n_replicas, bool(vecnorm)
        self.operators = () if operator is None else tuple(operator) if isinstance(operator, (tuple, list)) else (operator,)

    def __call__(self, states):
        out = {}
        vecs = []
        for st in states:  # overlaps are Float64 (promote_type of the value types, replicastrategy.jl:165)
            v = st.v
            vecs.append(v if v.style.val_type == _lib.VAL_F64 else
                        GPUDVec(style=IsDeterministic(), address_type=v.address_type, ctx=v.ctx).copy_from(v))
        for i in range(len(vecs)):
            for j in range(i + 1, len(vecs)):
                if self.vecnorm:
                    out[f"c{i + 1}_dot_c{j + 1}"] = vecs[i].dot(vecs[j])
                for k, op in enumerate(self.operators):
                    out[f"c{i + 1}_Op{k + 1}_c{j + 1}"] = dot(vecs[i], op, vecs[j])
        return out


class Timer:
    def __call__(self, state, step):
        return {"time": time.time()}


# --------------------------------------------------------------------------- reporting strategies
@dataclass
class ReportDFAndInfo:
    """reportingstrategy.jl:264-300: the report lives in memory and becomes a DataFrame."""
    reporting_interval: int = 1


class ReportToFile:
    """ReportToFile(; filename, reporting_interval, chunk_size, save_if, return_df, compress) (reportingstrategy.jl:302-434):
    the report is written to an Arrow file in chunks of `chunk_size` reported steps and dropped from memory after every
    chunk; an existing file is never overwritten (`out.arrow` -> `out-1.arrow` -> ...); only the root rank saves."""

    def __init__(self, filename="out.arrow", reporting_interval=1, chunk_size=1000, save_if=None, return_df=False, compress="zstd"):
        if compress not in (None, "zstd", "lz4"):
            raise ValueError("compress must be None, 'zstd' or 'lz4'")  # ArgumentError (reportingstrategy.jl:346-350)
        self.filename, self.reporting_interval, self.chunk_size = str(filename), int(reporting_interval), int(chunk_size)
        self.save_if = (int(os.environ.get("RANK", "0")) == 0) if save_if is None else bool(save_if)  # is_mpi_root()
        self.return_df, self.compress = bool(return_df), compress
        self._writer = self._sink = self._schema = None
        self.chunks_written = 0

    def refine(self):
        """refine_reporting_strategy (reportingstrategy.jl:362-381): pick a file name that does not exist yet"""
        if self.save_if:
            import re
            name = self.filename
            while os.path.isfile(name):
                base, ext = os.path.splitext(name)
                m = re.match(r"(.*)-([0-9]+)$", base)
                name = f"{base}-1{ext}" if m is None else f"{m.group(1)}-{int(m.group(2)) + 1}{ext}"
            self.filename = name
        return self

    def _write(self, report, metadata):
        import pyarrow as pa
        if not report or not len(next(iter(report.values()))):
            return
        table = pa.Table.from_pydict({k: np.asarray(v) for k, v in report.items()})
        if self._writer is None:
            md = {str(k).encode(): str(v).encode() for k, v in (metadata or {}).items()}
            self._schema = table.schema.with_metadata(md)
            self._sink = pa.OSFile(self.filename, "wb")
            self._writer = pa.ipc.new_file(self._sink, self._schema, options=pa.ipc.IpcWriteOptions(compression=self.compress))
        self._writer.write_table(table.cast(self._schema.remove_metadata()).replace_schema_metadata(self._schema.metadata))
        self.chunks_written += 1

    def after_step(self, step, report, metadata):
        """report_after_step! (:386-401): flush a full chunk and empty the in-memory report"""
        if self.save_if and step % (self.chunk_size * self.reporting_interval) == 0:
            self._write(report, metadata)
            for v in report.values():
                v.clear()

    def finalize(self, report, metadata):
        """finalize_report! (:402-427)"""
        if self.save_if:
            self._write(report, metadata)
            if self._writer is not None:
                self._writer.close()
                self._sink.close()
                self._writer = self._sink = None
            for v in report.values():
                v.clear()


def load_df(filename):
    """RimuIO.load_df (RimuIO.jl:27-44): the report file as a DataFrame; the metadata come back in `df.attrs`."""
    import pyarrow as pa
    with pa.OSFile(str(filename), "rb") as src:
        table = pa.ipc.open_file(src).read_all()
    df = table.to_pandas()
    df.attrs.update({k.decode(): v.decode() for k, v in (table.schema.metadata or {}).items()})
    return df


# --------------------------------------------------------------------------- spectral strategies
@dataclass
class GramSchmidt:
    """GramSchmidt(S; orthogonalization_interval=1) (strategies_and_params/spectralstrategy.jl:22-35): S spectral states per
    replica, state i orthogonalised against the states j < i every `orthogonalization_interval` steps (fciqmc.jl:187-202) --
    on the device `u <- u - (u.v / |v|^2) v` is one dot, one norm and one axpby per pair."""
    num_spectral_states: int = 1
    orthogonalization_interval: int = 1

    def orthogonalize(self, states):
        for i in range(len(states)):
            for j in range(i):
                u, v = states[i].v, states[j].v
                vv = v.dot(v)
                if vv != 0.0:
                    u.add_(v, -u.dot(v) / vv)


def build_basis(ham, start_address, minimum_size):
    """build_basis(ham, address; minimum_size) (ExactDiagonalization/basis_breadth_first_search.jl:216-399): breadth-first
    search from the starting address, level by level, until at least `minimum_size` addresses are known."""
    from .hamiltonians import offdiagonals
    basis, seen, frontier = [start_address], {start_address.key()}, [start_address]
    while frontier and len(basis) < minimum_size:
        nxt = []
        for a in frontier:
            for b, val in offdiagonals(ham, a):
                if val != 0.0 and b.key() not in seen:
                    seen.add(b.key())
                    basis.append(b)
                    nxt.append(b)
        frontier = nxt
    return basis


def spectral_starting_vectors(ham, start_address, n_spectral, style, initiator, minimum_size=None):
    """pmc_simulation.jl:48-63: the lowest `n_spectral` eigenvectors (times 10) of H truncated to a small breadth-first basis
    around the starting address are the starting vectors of the spectral states."""
    from .hamiltonians import diagonal_element, offdiagonals
    basis = build_basis(ham, start_address, minimum_size or 2 * n_spectral)
    index = {a.key(): i for i, a in enumerate(basis)}
    mat = np.zeros((len(basis), len(basis)))
    for i, a in enumerate(basis):
        mat[i, i] = diagonal_element(ham, a)
        for b, val in offdiagonals(ham, a):
            j = index.get(b.key())
            if j is not None and j != i:
                mat[j, i] += val
    w, vecs = np.linalg.eig(mat)
    order = np.argsort(w.real)
    out = []
    for s in range(n_spectral):
        col = np.real(vecs[:, order[s]])
        pairs = [(a, 10.0 * c) for a, c in zip(basis, col) if c != 0.0]
        if style.val_type == _lib.VAL_I64:
            pairs = [(a, int(round(c))) for a, c in pairs if int(round(c)) != 0]
        out.append(GPUDVec(pairs, style=style, initiator=initiator))
    return out


# --------------------------------------------------------------------------- problem / simulation
def default_starting_vector(ham_or_address, population=10, style=None, initiator=None):
    """qmc_states.jl:236-260: `address => population` with the given style (and initiator rule)."""
    address = starting_address(ham_or_address) if isinstance(ham_or_address, AbstractHamiltonian) else ham_or_address
    style = IsDynamicSemistochastic() if style is None else style
    val = int(population) if style.val_type == _lib.VAL_I64 else float(population)
    return GPUDVec([(address, val)], style=style, initiator=initiator)


@dataclass
class SingleState:
    """qmc_states.jl:20-29"""
    hamiltonian: AbstractHamiltonian
    v: GPUDVec
    pv: GPUDVec
    wm: WorkingMemory
    shift_parameters: ShiftParameters


class ProjectorMonteCarloProblem:
    """ProjectorMonteCarloProblem(hamiltonian; kwargs...) with the reference's defaults."""

    def __init__(self, hamiltonian, *, start_at=None, shift=None, style=None, time_step=0.01, starting_step=0,
                 last_step=100, wall_time=math.inf, target_walkers=1000, zeta=0.08, xi=None, shift_strategy=None,
                 post_step_strategy=(), max_length=None, random_seed=True, reporting_interval=1, metadata=None,
                 n_replicas=1, initiator=False, replica_strategy=None, spectral_strategy=None, minimum_size=None,
                 reporting_strategy=None, device_steps=None):
        if int(n_replicas) < 1:
            raise ValueError("n_replicas must be at least 1")
        self.n_replicas = int(n_replicas)  # independent copies of the walker vector, advanced side by side (qmc_states.jl:89-140)
        self.replica_strategy = replica_strategy  # e.g. AllOverlaps: it also fixes the number of replicas (pmc problem :205-215)
        self.spectral_strategy = spectral_strategy or GramSchmidt(1)  # projector_monte_carlo_problem.jl:176
        self.minimum_size = minimum_size or 2 * self.spectral_strategy.num_spectral_states  # :177
        if replica_strategy is not None:
            if n_replicas not in (1, replica_strategy.n_replicas):
                raise ValueError("n_replicas conflicts with the replica strategy")
            self.n_replicas = replica_strategy.n_replicas
        # initiator=true -> Initiator(threshold 1) (projector_monte_carlo_problem.jl:156-160); a rule object is taken as is
        from .stochasticstyles import as_initiator_rule
        self.initiator = as_initiator_rule(initiator)
        self.hamiltonian = hamiltonian
        self.style = IsDynamicSemistochastic() if style is None else style
        self.start_at = start_at
        self.shift, self.time_step = shift, float(time_step)
        self.starting_step, self.last_step, self.wall_time = starting_step, last_step, wall_time
        self.shift_strategy = shift_strategy or DoubleLogUpdate(target_walkers, zeta, xi)
        tw = getattr(self.shift_strategy, "target_walkers", target_walkers)
        self.max_length = max_length if max_length is not None else round(2 * abs(tw) + 100)
        self.post_step_strategy = tuple(post_step_strategy) if isinstance(post_step_strategy, (tuple, list)) else (post_step_strategy,)
        if random_seed is True:
            random_seed = int.from_bytes(os.urandom(8), "little")
        elif random_seed is False or random_seed is None:
            random_seed = 0
        self.random_seed = int(random_seed) & 0xFFFFFFFFFFFFFFFF
        # reporting_strategy (ReportDFAndInfo / ReportToFile) carries its own interval (projector_monte_carlo_problem.jl:170-172)
        self.reporting_strategy = reporting_strategy or ReportDFAndInfo(reporting_interval)
        self.reporting_interval = self.reporting_strategy.reporting_interval
        self.metadata = dict(metadata or {})
        # device_steps: how many steps `solve` hands to the device in one call (rimu_advance: shift update and abort rules on
        # the GPU, no host round trip between steps).  Applies to single-state runs without post-step strategies on one GPU;
        # 1 = the plain step-by-step loop.  The report has the same rows either way.
        if device_steps is None:
            device_steps = int(os.environ.get("RIMU_B200_DEVICE_STEPS", "64"))
        self.device_steps = max(1, int(device_steps))


class PMCSimulation:
    """init(problem) (pmc_simulation.jl:88-174)."""

    def __init__(self, problem: ProjectorMonteCarloProblem):
        self.problem = p = problem
        ham = p.hamiltonian
        sa = p.start_at
        if sa is None:
            v = default_starting_vector(ham, style=p.style, initiator=p.initiator)
        elif isinstance(sa, (list, tuple)) and sa and all(isinstance(x, GPUDVec) for x in sa):
            v = sa[0].copy()  # one starting vector per spectral state (the matrix form of start_at, pmc_simulation.jl:36-47)
        elif isinstance(sa, GPUDVec):
            v = sa.copy()  # a vector brings its own style and initiator rule
        elif isinstance(sa, (list, tuple, dict)):
            v = GPUDVec(sa, style=p.style, initiator=p.initiator)
        else:  # an address
            v = default_starting_vector(sa, style=p.style, initiator=p.initiator)
        style = v.style
        if p.shift is None:  # Rayleigh quotient of the starting vector (fciqmc.jl:51-61)
            vf = v if style.val_type == _lib.VAL_F64 else GPUDVec(style=IsDeterministic(), address_type=v.address_type, ctx=v.ctx).copy_from(v)
            vdet = vf if isinstance(vf.style, IsDeterministic) else GPUDVec(style=IsDeterministic(), address_type=v.address_type, ctx=v.ctx).copy_from(vf)
            shift = dot(vdet, ham, vdet) / vdet.dot(vdet)
        else:
            shift = float(p.shift)
        # replicas: independent vectors with their own shift parameters and random streams; report columns get the suffix
        # _1, _2, ... when there is more than one (qmc_states.jl:107-140, pmc_simulation.jl:125-146).  Every replica holds
        # n_spectral spectral states (suffix _s1, _s2, ...; pmc_simulation.jl:133-146), orthogonalised by the spectral
        # strategy before they are advanced (fciqmc.jl:187-202).
        n_spec = p.spectral_strategy.num_spectral_states
        if n_spec > 1:
            if isinstance(sa, (list, tuple)) and len(sa) == n_spec and all(isinstance(x, GPUDVec) for x in sa):
                spec_vectors = [x.copy() for x in sa]  # one starting vector per spectral state
            else:
                start = starting_address(ham) if sa is None or isinstance(sa, (GPUDVec, list, tuple, dict)) else sa
                spec_vectors = spectral_starting_vectors(ham, start, n_spec, style, p.initiator, p.minimum_size)
        else:
            spec_vectors = [v]
        self.states, self.replicas = [], []
        self.suffixes = []
        for r in range(p.n_replicas):
            rep = []
            for s_, sv in enumerate(spec_vectors):
                vr = sv if r == 0 else sv.copy()
                k = r * n_spec + s_
                seed = p.random_seed if k == 0 else (p.random_seed + 0x9E3779B97F4A7C15 * k) & 0xFFFFFFFFFFFFFFFF
                if p.shift is None and n_spec > 1:  # every spectral state starts from its own Rayleigh quotient
                    vd = GPUDVec(style=IsDeterministic(), address_type=vr.address_type, ctx=vr.ctx).copy_from(vr)
                    sh = dot(vd, ham, vd) / vd.dot(vd)
                else:
                    sh = shift
                st = SingleState(ham, vr, vr.zerovector(), WorkingMemory(vr, seed=seed), ShiftParameters(sh, vr.walkernumber(), p.time_step))
                rep.append(st)
                self.states.append(st)
                self.suffixes.append((f"_{r + 1}" if p.n_replicas > 1 else "") + (f"_s{s_ + 1}" if n_spec > 1 else ""))
            self.replicas.append(rep)
        self.state = self.states[0]
        if isinstance(p.reporting_strategy, ReportToFile):
            p.reporting_strategy.refine()
        self.step = p.starting_step
        self.report = {}
        self.aborted = False
        self.success = False
        self.message = ""
        self.elapsed_time = 0.0
        self.modified = False

    # ---- one step: advance!(FCIQMC) (fciqmc.jl:126-181)
    def step_(self):
        if self.aborted or self.success:
            return self
        if self.step >= self.problem.last_step:
            self.success = True
            return self
        self.step += 1
        p = self.problem
        report_now = self.step % p.reporting_interval == 0
        row = {"step": self.step}
        dead = too_long = stop = False
        gs = p.spectral_strategy
        if gs.num_spectral_states > 1 and self.step % gs.orthogonalization_interval == 0:
            for rep in self.replicas:  # Gram-Schmidt inside every replica (fciqmc.jl:189-197)
                gs.orthogonalize(rep)
        for r, st in enumerate(self.states):
            sfx = self.suffixes[r]
            sp = st.shift_parameters
            T = FirstOrderTransitionOperator(st.hamiltonian, sp.shift, sp.time_step)
            names, values, wm, pv = apply_operator(st.wm, st.pv, st.v, T)
            st.v, st.pv = pv, st.v
            stats = wm.last_stats
            is_int = st.v.style.val_type == _lib.VAL_I64
            tnorm = float(stats.inorm1) if is_int else stats.norm1
            length = stats.len
            shift_stats, proceed = p.shift_strategy.update(sp, tnorm) if length > 0 else ({"shift": sp.shift, "norm": tnorm}, True)
            if report_now:
                row["len" + sfx] = length
                row.update({k + sfx: val for k, val in shift_stats.items()})
                row.update({k + sfx: val for k, val in zip(names, values)})
                for ps in p.post_step_strategy:
                    row.update({k + sfx: val for k, val in ps(st, self.step).items()})
            dead |= length == 0
            too_long |= length > p.max_length
            stop |= not proceed
        if report_now:
            if p.replica_strategy is not None and not dead:
                row.update(p.replica_strategy([rep[0] for rep in self.replicas]))  # (first spectral state of every replica)
            for k, val in row.items():
                self.report.setdefault(k, []).append(val)
            if isinstance(p.reporting_strategy, ReportToFile):
                p.reporting_strategy.after_step(self.step, self.report, self._report_metadata())
        if dead:
            self.aborted, self.message = True, f"Aborted in step {self.step}."  # dead population
        elif too_long:
            self.aborted, self.message = True, f"Aborted in step {self.step}."  # max_length reached
        elif stop:  # a shift strategy asked to stop (pmc_simulation.jl:301-304 reports it the same way)
            self.aborted, self.message = True, f"Aborted in step {self.step}."
        elif self.step >= p.last_step:
            self.success = True
        self.modified = True
        return self

    # ---- a batch of steps on the device (rimu_advance): advance! x K without a host round trip
    def _batch_projectors(self):
        """[(report name, FrozenDVec)] when every post-step strategy is a frozen projection (ProjectedEnergy of a Hermitian
        Hamiltonian, Projector) -- those are evaluated inside the device batch; None when some strategy needs the host"""
        from .dictvectors import FrozenDVec
        out = []
        for ps in self.problem.post_step_strategy:
            if isinstance(ps, ProjectedEnergy) and isinstance(ps.vproj, FrozenDVec) and isinstance(ps.hproj, FrozenDVec):
                out += [(ps.vproj_name, ps.vproj), (ps.hproj_name, ps.hproj)]
            elif isinstance(ps, Projector) and isinstance(ps.projector, FrozenDVec):
                out.append((ps.name, ps.projector))
            else:
                return None
        return out if len(out) <= _lib.MAX_PROJECTORS else None

    def _batch_size(self):
        p = self.problem
        if p.device_steps <= 1 or len(self.states) != 1 or p.replica_strategy is not None or self._batch_projectors() is None:
            return 0
        st = self.state
        if not hasattr(st.v, "handle") or getattr(st.v.ctx, "nranks", 1) != 1 or _device_strategy(p.shift_strategy) is None:
            return 0
        return min(p.device_steps, p.last_step - self.step)

    def _advance_batch(self, K):
        p, st = self.problem, self.state
        sp = st.shift_parameters
        sid, target, zeta, xi = _device_strategy(p.shift_strategy)
        mode = sp.shift_mode
        projs = self._batch_projectors()
        res = advance(st.wm, st.v, st.pv, st.hamiltonian, sp, sid, target_walkers=target, zeta=zeta, xi=xi,
                      nsteps=K, max_length=p.max_length, projectors=[fr for _, fr in projs])
        v, pv, stats, shifts, done = res[:5]
        pvals = res[5] if projs else None
        st.v, st.pv = v, pv
        is_int = v.style.val_type == _lib.VAL_I64
        style = v.style
        with_len_before = type(getattr(style, "compression", None)).__name__ == "ThresholdCompression"
        after_target = sid in (_lib.SHIFT_LOG_UPDATE_AFTER_TARGET, _lib.SHIFT_DOUBLE_LOG_UPDATE_AFTER_TARGET)
        for k in range(done):
            self.step += 1
            s = stats[k]
            tnorm = float(s.inorm1) if is_int else s.norm1
            length = s.len
            proceed = True
            shift_stats = {"shift": shifts[k], "norm": tnorm}
            if length > 0:
                if after_target:
                    mode = mode or tnorm > target
                    shift_stats["shift_mode"] = mode
                if sid == _lib.SHIFT_DONT_UPDATE:
                    proceed = tnorm < target
            if self.step % p.reporting_interval == 0:
                row = {"step": self.step, "len": length}
                row.update(shift_stats)
                names, values = style.stat_names, style.stats(s)
                if with_len_before:
                    names, values = names + ("len_before",), values + (s.len_before,)
                row.update(dict(zip(names, values)))
                for j, (pname, _) in enumerate(projs):
                    row[pname] = float(pvals[k, j])
                for key, val in row.items():
                    self.report.setdefault(key, []).append(val)
                if isinstance(p.reporting_strategy, ReportToFile):
                    p.reporting_strategy.after_step(self.step, self.report, self._report_metadata())
            if length == 0 or length > p.max_length or not proceed:
                self.aborted, self.message = True, f"Aborted in step {self.step}."
        if not self.aborted and self.step >= p.last_step:
            self.success = True
        self.modified = True
        return self

    def solve_(self, last_step=None, wall_time=None):
        if last_step is not None:
            self.problem.last_step = last_step
            self.success = self.step >= last_step and not self.aborted
            if self.step < last_step:
                self.success = False
        wt = self.problem.wall_time if wall_time is None else wall_time
        t0 = time.time()
        while not self.aborted and not self.success:
            if time.time() - t0 > wt:
                self.aborted, self.message = True, "Wall time reached."
                break
            K = self._batch_size()
            if K >= 2:
                self._advance_batch(K)
            else:
                self.step_()
        self.elapsed_time += time.time() - t0
        if (self.aborted or self.success) and isinstance(self.problem.reporting_strategy, ReportToFile):
            self.problem.reporting_strategy.finalize(self.report, self._report_metadata())
        return self

    def _report_metadata(self):
        """report_simulation_status_metadata! (pmc_simulation.jl:186-200) + the problem's own metadata"""
        md = dict(self.problem.metadata)
        md.update({"modified": self.modified, "aborted": self.aborted, "success": self.success, "message": self.message,
                   "elapsed_time": self.elapsed_time, "num_replicas": self.problem.n_replicas,
                   "num_spectral_states": self.problem.spectral_strategy.num_spectral_states})
        return md

    def dataframe(self):
        import pandas as pd
        rs = self.problem.reporting_strategy
        if isinstance(rs, ReportToFile) and rs.save_if and rs.chunks_written:
            return load_df(rs.filename)  # (the in-memory report was emptied after every chunk)
        return pd.DataFrame(self.report)

    DataFrame = dataframe


def init(problem):
    return PMCSimulation(problem)


def step_(sim):
    return sim.step_()


def solve_(sim, **kw):
    return sim.solve_(**kw)


def solve(problem, **kw):
    return init(problem).solve_(**kw)


def DataFrame(sim):
    return sim.dataframe()
