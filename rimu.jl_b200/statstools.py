"""Blocking analysis for the energy checks: host mirror of the StatsTools entry points the FCIQMC
acceptance tests use (post-processing of O(steps) scalars; nothing here touches the GPU).

  blocking_analysis   StatsTools/blocking.jl:134-159, blocks_with_m :288-325, mtest :274-286
  ratio_of_means      StatsTools/ratio_of_means.jl:126-153,222-246
  shift_estimator / projected_energy   StatsTools/reweighting.jl:721-743
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np

BlockingResult = namedtuple("BlockingResult", "mean err err_err p_cov k blocks")
RatioBlockingResult = namedtuple("RatioBlockingResult", "samples f sigma_f delta_y k blocks success")


def _blocks_with_m(v):
    v = np.asarray(v, dtype=float)
    n_steps = int(math.floor(math.log2(len(v))))
    rows = []
    for _ in range(n_steps):
        n = len(v)
        mean = v.mean()
        var = v.var(ddof=1)
        d = v - mean
        gamma = float(np.dot(d[:-1], d[1:])) / (n - 1)  # autocovariance(v, 1; corrected) (variances.jl:8-16)
        mj = n * ((n - 1) * var / n ** 2 + gamma) ** 2 / var ** 2 if var > 0 else math.nan
        stderr = math.sqrt(var / n)
        rows.append((n, mean, stderr, stderr / math.sqrt(2 * (n - 1)), float(np.dot(d, d)) / (n - 1) / n, mj))
        m = n // 2
        v = (v[0:2 * m:2] + v[1:2 * m:2]) / 2
    return rows


def _mtest(mj, alpha=0.01):
    from scipy.stats import chi2
    m = np.cumsum(np.asarray(mj)[::-1])[::-1]
    for k in range(1, len(m)):
        if m[k - 1] < chi2.isf(alpha, k):
            return k
    return -1


def blocking_analysis(v, alpha=0.01, skip=0):
    v = np.asarray(v, dtype=float)[skip:]
    if len(v) < 2:
        return BlockingResult(float(v[0]) if len(v) else 0.0, math.nan, math.nan, math.nan, -1, len(v))
    rows = _blocks_with_m(v)
    k = _mtest([r[5] for r in rows], alpha)
    if k < 0:
        return BlockingResult(rows[0][1], math.nan, math.nan, math.nan, -1, rows[0][0])
    n, mean, err, err_err, p_cov, _ = rows[k - 1]
    return BlockingResult(rows[0][1], err, err_err, p_cov, k, n)


def _reblock(v, k):
    v = np.asarray(v, dtype=float)
    for _ in range(max(k - 1, 0)):
        m = len(v) // 2
        v = (v[0:2 * m:2] + v[1:2 * m:2]) / 2
    return v


def ratio_of_means(num, denom, alpha=0.01, skip=0, mc_samples=2000, seed=0):
    num, denom = np.asarray(num, dtype=float)[skip:], np.asarray(denom, dtype=float)[skip:]
    bn, bd = blocking_analysis(num, alpha), blocking_analysis(denom, alpha)
    success = bn.k >= 0 and bd.k >= 0
    k = max(bn.k, bd.k)
    x, y = _reblock(num, k), _reblock(denom, k)
    n = len(x)
    mx, my = x.mean(), y.mean()
    var_x, var_y = x.var(ddof=1) / n, y.var(ddof=1) / n
    rho = float(np.cov(x, y, ddof=1)[0, 1]) / n
    rng = np.random.default_rng(seed)
    s = rng.multivariate_normal([mx, my], [[var_x, rho], [rho, var_y]], size=mc_samples)
    samples = s[:, 0] / s[:, 1]
    arg = (math.sqrt(var_x) / my) ** 2 + (mx * math.sqrt(var_y) / my ** 2) ** 2 - 2 * rho * mx / my ** 3
    sigma_f = math.sqrt(arg) if arg >= 0 else math.nan
    return RatioBlockingResult(samples, bn.mean / bd.mean, sigma_f, math.sqrt(var_y) / my, k, n, success)


def shift_estimator(df, shift="shift", skip=0, **kw):
    return blocking_analysis(np.asarray(df[shift]), skip=skip, **kw)


def projected_energy(df, hproj="hproj", vproj="vproj", skip=0, **kw):
    return ratio_of_means(np.asarray(df[hproj]), np.asarray(df[vproj]), skip=skip, **kw)
