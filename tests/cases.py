"""Shared model zoo: each case builds the product Hamiltonian (rimu_b200) and the oracle Hamiltonian
(oracle/) from the SAME plain parameters, independently."""
import math

import numpy as np

from oracle import oracle as orc


def _nu(n, m):
    f, e = divmod(n, m)
    return tuple(f + (1 if i < e else 0) for i in range(m))


def _fermi(m, modes):
    return tuple(1 if (i + 1) in modes else 0 for i in range(m))


# name -> (model, kind, onr, params)
SPECS = {
    "real1d_6": ("HubbardReal1D", "bose", _nu(6, 6), dict(u=6.0, t=1.0)),
    "real1d_10": ("HubbardReal1D", "bose", _nu(10, 10), dict(u=6.0, t=1.0)),             # BASELINE config 1
    "real1d_w2": ("HubbardReal1D", "bose", _nu(40, 40), dict(u=2.0, t=1.0)),             # 79 bits -> 2 words
    "real1d_ep": ("HubbardReal1DEP", "bose", (1, 2, 3, 4), dict(u=2.0, t=3.0, v_ho=4.0)),              # test/Hamiltonians.jl:392
    "real1d_ep_w2": ("HubbardReal1DEP", "bose", tuple(1 if i == 0 else 0 for i in range(100)), dict(u=1.0, t=50.0, v_ho=0.005)),  # :1045-1054
    "ext1d": ("ExtendedHubbardReal1D", "bose", (1, 0, 2, 1), dict(u=1.0, v=2.0, t=3.0)),
    "ext1d_twisted": ("ExtendedHubbardReal1D", "bose", (2, 0, 1, 1, 0, 1), dict(u=1.5, v=6.0, t=2.0, boundary_condition="twisted")),
    "ext1d_hw": ("ExtendedHubbardReal1D", "bose", (1, 1, 0, 2, 1), dict(u=1.0, v=6.0, t=2.0, boundary_condition="hard_wall")),
    "mom1d_bose": ("HubbardMom1D", "bose", (0, 0, 0, 6, 0, 0, 0, 0), dict(u=4.0, t=1.0)),
    "mom1d_bose_20": ("HubbardMom1D", "bose", tuple(20 if i == 9 else 0 for i in range(20)), dict(u=6.0, t=1.0)),  # config 2
    "mom1d_odd": ("HubbardMom1D", "bose", (1, 2, 3, 0, 0), dict(u=1.0, t=1.0)),
    "ext_mom1d": ("ExtendedHubbardMom1D", "bose", (0, 1, 2, 0, 1, 0), dict(u=1.5, v=2.5, t=1.0)),
    "ext_mom1d_20": ("ExtendedHubbardMom1D", "bose", tuple(10 if i == 9 else 0 for i in range(20)), dict(u=6.0, v=1.0, t=1.0)),
    "mom1d_ep": ("HubbardMom1DEP", "bose", (0, 0, 3, 1, 0), dict(u=2.0, t=1.0, v_ho=0.5)),
    "mom1d_ep_f2c": ("HubbardMom1DEP", "fermi2c", ((0, 1, 1, 0, 0, 0), (0, 0, 1, 0, 1, 0)), dict(u=3.0, t=1.0, v_ho=0.3)),
    "mom1d_f2c": ("HubbardMom1D", "fermi2c", ((1, 1, 0, 0), (0, 0, 1, 1)), dict(u=4.0, t=4 / math.pi ** 2)),
    "rs_bose_1d": ("HubbardRealSpace", "bose", _nu(5, 5), dict(u=1.0, t=1.0, dims=(5,))),
    "rs_bose_2d": ("HubbardRealSpace", "bose", _nu(6, 9), dict(u=2.0, t=1.5, dims=(3, 3))),
    "rs_bose_2d_hw": ("HubbardRealSpace", "bose", _nu(5, 6), dict(u=2.0, t=1.0, dims=(2, 3), fold=(False, True), trap=((0.3, 0.7),))),
    "rs_bose_3d_w2": ("HubbardRealSpace", "bose", _nu(64, 64), dict(u=1.0, t=1.0, dims=(4, 4, 4))),  # config 4
    "rs_fermi": ("HubbardRealSpace", "fermi", _fermi(12, (1, 2, 3)), dict(t=1.0, dims=(3, 4))),
    "rs_fermi_hw": ("HubbardRealSpace", "fermi", _fermi(12, (1, 6, 12)), dict(t=2.0, dims=(4, 3), fold=(False, False))),
    "rs_f2c_4x4": ("HubbardRealSpace", "fermi2c", (_fermi(16, (1, 6)), _fermi(16, (3, 11))), dict(t=(1.0, 1.0), u=((0.0, 4.0), (4.0, 0.0)), dims=(4, 4))),
    "rs_f2c_3up3dn": ("HubbardRealSpace", "fermi2c", (_fermi(16, (1, 6, 11)), _fermi(16, (2, 8, 13))), dict(t=(1.0, 1.0), u=((0.0, 4.0), (4.0, 0.0)), dims=(4, 4))),  # dim 313 600
    "rs_f2c_half": ("HubbardRealSpace", "fermi2c", (_fermi(16, range(1, 9)), _fermi(16, range(5, 13))), dict(t=(1.0, 1.0), u=((0.0, 1.0), (1.0, 0.0)), dims=(4, 4))),  # config 3
    "rs_f2c_trap": ("HubbardRealSpace", "fermi2c", (_fermi(6, (1, 2, 4, 5)), _fermi(6, (2, 3))), dict(t=(1.0, 2.0), u=((0.0, 0.5), (0.5, 0.0)), dims=(6,), trap=((0.1,), (0.2,)))),
    # general CompositeFS (multicomponent.jl:10-34) on HubbardRealSpace: bosonic, mixed, three components, wide fermions, two words
    "rs_comp_bb": ("HubbardRealSpace", "comp:bb", ((1, 1, 1, 0, 0, 0), (1, 0, 0, 0, 0, 0)), dict(t=(1.0, 4.0), u=((2.0, 3.0), (3.0, 0.0)), dims=(6,))),  # test/Hamiltonians.jl:329-333
    "rs_comp_bf": ("HubbardRealSpace", "comp:bf", ((1, 1, 1, 0, 0, 0), (1, 0, 0, 0, 0, 0)), dict(t=(1.0, 4.0), u=((2.0, 3.0), (3.0, 0.0)), dims=(6,))),  # :335-339
    "rs_comp_fb": ("HubbardRealSpace", "comp:fb", ((1, 0, 0, 0, 0, 0), (1, 1, 1, 0, 0, 0)), dict(t=(4.0, 1.0), u=((0.0, 3.0), (3.0, 2.0)), dims=(6,))),  # :347-351
    "rs_comp_bb_trap": ("HubbardRealSpace", "comp:bb", ((1, 1, 1, 0, 0, 0), (1, 0, 0, 0, 0, 0)), dict(t=(1.0, 1.0), u=((2.0, 3.0), (3.0, 0.0)), dims=(6,), trap=((1.0,), (4.0,)))),  # :398-403
    "rs_comp_ffb": ("HubbardRealSpace", "comp:ffb", ((1, 1, 0, 0), (1, 0, 0, 1), (0, 0, 2, 1)), dict(t=(1.0, 2.0, 0.5), u=((0.0, 1.0, 2.0), (1.0, 0.0, 3.0), (2.0, 3.0, 1.5)), dims=(2, 2), fold=(True, False))),
    "rs_comp_bbb": ("HubbardRealSpace", "comp:bbb", ((1, 0, 0, 0), (1, 0, 0, 0), (1, 0, 0, 0)), dict(t=(1.0, 1.0, 1.0), u=((1.0, 1.0, 1.0), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0)), dims=(4,))),  # :78
    "rs_comp_ff_wide": ("HubbardRealSpace", "comp:ff", (_fermi(36, (1, 8, 15, 22)), _fermi(36, (2, 8, 30))), dict(t=(1.0, 1.0), u=((0.0, 4.0), (4.0, 0.0)), dims=(6, 6))),  # 72 bits: two words
    "rs_comp_bf_w2": ("HubbardRealSpace", "comp:bf", (_nu(30, 27), _fermi(27, (1, 5, 9, 14, 20))), dict(t=(1.0, 0.5), u=((1.0, 2.0), (2.0, 0.0)), dims=(3, 3, 3), trap=((0.1, 0.2, 0.3), (0.0, 0.5, 0.0)))),  # 56 + 27 bits
    "tc_7": ("Transcorrelated1D", "fermi2c", (_fermi(7, (3, 5)), _fermi(7, (4,))), dict(t=24.5, v=7.0, cutoff=1, three_body_term=True)),
    "tc_8_cut2": ("Transcorrelated1D", "fermi2c", (_fermi(8, (3, 4, 6)), _fermi(8, (2, 5))), dict(t=1.0, v=1.5, cutoff=2, three_body_term=True)),
    "tc_12": ("Transcorrelated1D", "fermi2c", (_fermi(12, (5, 6, 7)), _fermi(12, (6, 7))), dict(t=1.0, v=1.0, cutoff=1, three_body_term=True)),  # config 5's model, small enough for ED
    "tc_32": ("Transcorrelated1D", "fermi2c", (_fermi(32, (15, 16, 17)), _fermi(32, (15, 16, 17))), dict(t=1.0, v=1.0, cutoff=1, three_body_term=True)),  # config 5
    "tc_no3b": ("Transcorrelated1D", "fermi2c", (_fermi(6, (2, 3)), _fermi(6, (3, 4))), dict(t=1.0, v=-2.0, cutoff=1, three_body_term=False)),
}


def oracle_ham(name):
    model, kind, onr, p = SPECS[name]
    return orc.OracleHam(model, kind, onr, **p)


def product_ham(name):
    import rimu_b200 as R
    model, kind, onr, p = SPECS[name]
    p = dict(p)
    if kind == "bose":
        addr = R.BoseFS(onr)
    elif kind == "fermi":
        addr = R.FermiFS(onr)
    elif kind.startswith("comp:"):
        addr = R.CompositeFS(*[(R.BoseFS if letter == "b" else R.FermiFS)(c) for letter, c in zip(kind[5:], onr)])
    else:
        addr = R.FermiFS2C(onr[0], onr[1])
    if model == "HubbardReal1D":
        return R.HubbardReal1D(addr, **p)
    if model == "HubbardReal1DEP":
        return R.HubbardReal1DEP(addr, **p)
    if model == "ExtendedHubbardReal1D":
        return R.ExtendedHubbardReal1D(addr, **p)
    if model == "HubbardMom1D":
        return R.HubbardMom1D(addr, **p)
    if model == "ExtendedHubbardMom1D":
        return R.ExtendedHubbardMom1D(addr, **p)
    if model == "HubbardMom1DEP":
        return R.HubbardMom1DEP(addr, **p)
    if model == "Transcorrelated1D":
        return R.Transcorrelated1D(addr, **p)
    dims, fold, trap = p.pop("dims"), p.pop("fold", None), p.pop("trap", None)
    geo = R.CubicGrid(dims, fold)
    return R.HubbardRealSpace(addr, geometry=geo, t=p.get("t"), u=p.get("u"), v=trap)


def sample_keys(oh, n, seed=0):
    """Random walk over non-zero off-diagonals starting from the oracle's start address."""
    rng = np.random.default_rng(seed)
    cur = oh.start_key
    keys = {cur}
    tries = 0
    while len(keys) < n and tries < 50 * n:
        tries += 1
        L = oh.num_offdiagonals(cur)
        if L == 0:
            break
        k, v = oh.get_offdiagonal(cur, int(rng.integers(1, L + 1)))
        if v != 0.0:
            cur = k
            keys.add(k)
    return np.array(sorted(keys), dtype=np.uint64).reshape(-1, oh.W)
