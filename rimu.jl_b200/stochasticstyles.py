"""StochasticStyles: host mirror of StochasticStyles/styles.jl and compression.jl.

A style only carries parameters; the spawning/compression arithmetic runs in the CUDA kernels
(kernels.cuh).  `step_stats(style)` returns the reference's stat column names.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

from . import _lib


class StochasticStyle:
    code: int
    val_type: int = _lib.VAL_F64

    def fill(self, p: _lib.StepParams):
        p.style = self.code
        p.proj_threshold = 0.0
        p.rel_threshold = 1.0
        p.abs_threshold = math.inf
        p.compress_threshold = 0.0

    @property
    def eltype(self):
        return int if self.val_type == _lib.VAL_I64 else float


@dataclass(frozen=True)
class NoCompression:
    pass


@dataclass(frozen=True)
class ThresholdCompression:
    """compression.jl:8-14"""
    threshold: float = 1.0


@dataclass(frozen=True)
class IsStochasticInteger(StochasticStyle):
    """styles.jl:11-25"""
    code = _lib.STYLE_INTEGER
    val_type = _lib.VAL_I64
    stat_names = ("spawn_attempts", "spawns", "deaths", "clones", "zombies")

    def stats(self, s: _lib.StepStats):
        return (s.spawn_attempts, s.ispawns, s.ideaths, s.iclones, s.izombies)


@dataclass(frozen=True)
class IsDeterministic(StochasticStyle):
    """styles.jl:76-105"""
    compression: object = NoCompression()
    code = _lib.STYLE_DETERMINISTIC
    stat_names = ("exact_steps",)

    def fill(self, p):
        super().fill(p)
        if isinstance(self.compression, ThresholdCompression):
            p.compress_threshold = float(self.compression.threshold)

    def stats(self, s):
        return (s.exact_steps,)


@dataclass(frozen=True)
class IsStochasticWithThreshold(StochasticStyle):
    """styles.jl:117-130"""
    threshold: float = 1.0
    code = _lib.STYLE_WITH_THRESHOLD
    stat_names = ("spawn_attempts", "spawns")

    def fill(self, p):
        super().fill(p)
        p.proj_threshold = float(self.threshold)

    def stats(self, s):
        return (s.spawn_attempts, s.spawns)


@dataclass(frozen=True)
class IsDynamicSemistochastic(StochasticStyle):
    """styles.jl:175-214 (spawning strategy fixed to WithReplacement, the default)."""
    threshold: float = 1.0
    rel_spawning_threshold: float = 1.0
    abs_spawning_threshold: float = math.inf
    late_compression: bool = True
    code = _lib.STYLE_SEMISTOCHASTIC
    stat_names = ("exact_steps", "inexact_steps", "spawn_attempts", "spawns")

    def fill(self, p):
        super().fill(p)
        p.rel_threshold = float(self.rel_spawning_threshold)
        p.abs_threshold = float(self.abs_spawning_threshold)
        if self.late_compression:
            p.compress_threshold = float(self.threshold)
        else:
            p.proj_threshold = float(self.threshold)

    @property
    def compression(self):
        return ThresholdCompression(self.threshold) if self.late_compression else NoCompression()

    def stats(self, s):
        return (s.exact_steps, s.inexact_steps, s.spawn_attempts, s.spawns)


def default_style(eltype):
    """styles.jl:216-219"""
    return IsStochasticInteger() if eltype is int else IsDeterministic()


def step_stats(style):
    names = style.stat_names
    if isinstance(getattr(style, "compression", None), ThresholdCompression):
        names = names + ("len_before",)
    return names


# --------------------------------------------------------------------------- initiator rules
# Host mirror of DictVectors/initiators.jl:132-236.  A rule only carries its id and threshold; the lane arithmetic
# (to_initiator_value / from_initiator_value) runs in the CUDA kernels (partition.cuh).
@dataclass(frozen=True)
class InitiatorRule:
    threshold: float = 1.0
    rule_id = 0


@dataclass(frozen=True)
class NonInitiator(InitiatorRule):
    """Disables the approximation (the default of PDVec)."""
    rule_id = 0


@dataclass(frozen=True)
class Initiator(InitiatorRule):
    """Initiators (|value| > threshold) spawn anywhere; non-initiators spawn only onto initiators' addresses."""
    rule_id = 1


@dataclass(frozen=True)
class SimpleInitiator(InitiatorRule):
    """Initiators spawn anywhere; non-initiators cannot spawn."""
    rule_id = 2


@dataclass(frozen=True)
class CoherentInitiator(InitiatorRule):
    """As Initiator, plus: non-initiator spawns onto one address count if together they exceed the threshold."""
    rule_id = 3


def as_initiator_rule(initiator=None, initiator_threshold=None) -> InitiatorRule:
    """PDVec / InitiatorDVec keyword handling (pdvec.jl:181-199): `initiator=true` or a positive
    `initiator_threshold` selects Initiator(threshold); a rule object is taken as is."""
    if isinstance(initiator, InitiatorRule):
        return initiator
    if initiator_threshold is not None and initiator_threshold > 0:
        return Initiator(float(initiator_threshold))
    if initiator is True:
        return Initiator(1.0)
    return NonInitiator()
