#!/usr/bin/env python
"""How much room do the statistical GPU tests have?  (tests/test_gpu_energies.py, tests/test_gpu_initiators.py)

The CUDA path and the CPU oracle draw from the same Philox function, so the oracle reproduces a GPU run's trajectory up
to floating-point summation order; running the tests' parameter sets through the oracle for several seeds shows the spread
of the shift estimator against the assertion's allowance.  Results of the last run (three seeds each, |shift - E0| vs
5 sigma + 1 % of |E0|):

    BHM example 6/6 (1000 walkers)            dev 0.0001 - 0.004   allowed 0.10 - 0.12
    real1d_10 integer walkers (1e4)           dev 0.02 - 0.06      allowed 0.19 - 0.21
    mom1d_bose semistochastic (2000)          dev 0.002 - 0.006    allowed 0.13 - 0.14
    rs_f2c_4x4 with 3000 walkers              dev 0.12 - 0.14      allowed 0.21 - 0.24   <- margin 1.5x: raised to 2e4 walkers
    rs_f2c_4x4 with 20000 walkers             dev 0.009            allowed 0.125 - 0.128
    initiator bias test (300 walkers, 4 seeds) E_no ~ -100, E_i1 = E_i3 ~ -8.6..-8.7, E_i2 ~ -8.1..-8.2, E0 = -9.25:
                                              every inequality of the test holds by >= 10 naive standard errors

    python tests/tools/energy_margins.py [name ...]      (CPU only; minutes)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402  (test tooling: the oracle is the checker)
from tests.cases import oracle_ham  # noqa: E402
from tests.test_gpu_energies import exact_energy  # noqa: E402

CASES = {  # name: (oracle style, walkers, dtau, steps, start population)
    "mom1d_bose": (orc.STYLE_SEMISTOCHASTIC, 2000, 0.002, 4000),
    "rs_f2c_4x4": (orc.STYLE_SEMISTOCHASTIC, 20000, 0.005, 4000),
    "real1d_10": (orc.STYLE_INTEGER, 10000, 0.002, 4000),
}


def blocking_err(x):
    """Flyvbjerg-Petersen plateau estimate (largest standard error over the reblocking levels with >= 32 blocks)."""
    x = np.asarray(x, dtype=float)
    best = 0.0
    while len(x) >= 32:
        best = max(best, x.std(ddof=1) / math.sqrt(len(x)))
        m = len(x) // 2
        x = (x[0:2 * m:2] + x[1:2 * m:2]) / 2
    return best


def run(name, style, walkers, dtau, steps, seed, zeta=0.08, rule=0):
    oh = oracle_ham(name)
    xi = zeta ** 2 / 4
    is_int = style == orc.STYLE_INTEGER
    keys = np.array([oh.start_key], dtype=np.uint64).reshape(1, -1)
    vals = np.array([10], dtype=np.int64) if is_int else np.array([10.0])
    shift, pnorm, shifts = oh.diagonal_element(oh.start_key), 10.0, []
    for step in range(steps):
        p = orc.make_params(style, shift=shift, dtau=dtau, compress_threshold=0.0 if is_int else 1.0,
                            key=orc.step_key(seed, step), initiator_rule=rule, initiator_threshold=1.0)
        keys, vals, st = oh.step(p, keys, vals, threads=8 if len(vals) > 3000 else 0)
        tn = float(st.inorm1) if is_int else st.norm1
        shift -= xi / dtau * math.log(tn / walkers) + zeta / dtau * math.log(tn / pnorm)
        pnorm = tn
        shifts.append(shift)
    sh = np.array(shifts[steps // 3:])
    return float(sh.mean()), blocking_err(sh)


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        e0 = exact_energy(oracle_ham(name))
        for seed in (5, 11, 12):
            m, err = run(name, *CASES[name], seed)
            print(f"{name} seed {seed}: E0 {e0:.4f} shift {m:.4f} +- {err:.4f}  |dev| {abs(m - e0):.4f}  allowed {5 * err + 0.01 * abs(e0):.4f}", flush=True)
