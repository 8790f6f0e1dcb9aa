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
