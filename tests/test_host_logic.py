"""Host-side mirror of the reference interface (no GPU): address codec, Hamiltonian descriptors,
style parameters, shift strategies, blocking analysis.  CPU only."""
import math

import numpy as np
import pytest

from oracle import oracle as orc


def test_address_codec_matches_oracle(built):
    import rimu_b200 as R
    rng = np.random.default_rng(1)
    for N, M in ((3, 3), (10, 10), (20, 20), (40, 40), (64, 64), (5, 90)):
        oh = orc.OracleHam("HubbardReal1D", "bose", R.near_uniform_onr(N, M))
        for _ in range(20):
            onr = np.bincount(rng.integers(0, M, size=N), minlength=M)
            a = R.BoseFS(tuple(int(x) for x in onr))
            assert a.key() == oh.pack(tuple(int(x) for x in onr))
            assert a.address_type.words == oh.W
            assert a.address_type.from_key(a.key()) == a
    oh = orc.OracleHam("HubbardRealSpace", "fermi2c", ((1, 0, 1, 0), (0, 1, 1, 0)), dims=(4,))
    a = R.FermiFS2C((1, 0, 1, 0), (0, 1, 1, 0))
    assert a.key() == oh.pack(((1, 0, 1, 0), (0, 1, 1, 0)))
    assert a.address_type.from_key(a.key()) == a
    f = R.FermiFS((1, 1, 0, 0, 1, 1, 1, 1))
    assert f.key() == (0b11110011,)


def test_near_uniform(built):
    """bosefs.jl:151-168: near_uniform(BoseFS{7,5}) = (2,2,1,1,1); FermiFS{3,12} fills the first modes."""
    import rimu_b200 as R
    assert R.near_uniform(R.BoseFS, 7, 5).onr == (2, 2, 1, 1, 1)
    assert R.near_uniform(R.BoseFS, 6, 6).onr == (1,) * 6
    assert R.near_uniform(R.FermiFS, 3, 12).onr == (1, 1, 1) + (0,) * 9


def test_hamiltonian_tables_match_oracle(built):
    """The host constructors precompute the same tables (kes, ws, us, trap potential) as the oracle's
    independent builders (HubbardMom1D.jl:55-63, Transcorrelated1D.jl:72-78, HubbardRealSpace.jl:214-227)."""
    import rimu_b200 as R
    H = R.HubbardMom1D(R.BoseFS((0, 0, 5, 0, 0)), u=6.0, t=1.5)
    _, kes = orc.mom1d_grid(5, 1.5)
    assert np.array_equal(np.array(H.desc.kes[:5]), kes)
    H = R.HubbardMom1D(R.BoseFS((0, 0, 0, 6, 0, 0)), u=6.0, t=0.5, dispersion=R.continuum_dispersion)
    _, kes = orc.mom1d_grid(6, 0.5, "continuum")
    assert np.array_equal(np.array(H.desc.kes[:6]), kes)
    a = R.FermiFS2C((0, 0, 1, 1, 0, 0, 0), (0, 0, 0, 1, 0, 0, 0))
    H = R.Transcorrelated1D(a, t=2.0, v=3.0, cutoff=2)
    _, kes, ws, us = orc.tc_tables(7, 2.0, 2)
    assert np.array_equal(np.array(H.desc.kes[:7]), kes)
    assert np.array_equal(np.array(H.desc.ws[:7]), ws)
    assert np.array_equal(np.array(H.desc.us[:7]), us)
    g = R.CubicGrid((2, 3), (False, True))
    H = R.HubbardRealSpace(R.BoseFS((1, 1, 1, 1, 1, 0)), geometry=g, t=1.0, u=2.0, v=((0.3, 0.7),))
    pot = orc.trap_potential((2, 3), ((0.3, 0.7),))
    assert np.array_equal(np.array(H.desc.potential[:6]), pot[0])
    with pytest.raises(ValueError):
        R.HubbardRealSpace(R.BoseFS((1, 1, 1)), geometry=R.PeriodicBoundaries(2, 2))  # wrong number of sites


def test_style_parameters(built):
    """styles.jl:182-194 defaults of IsDynamicSemistochastic; integer style has no thresholds."""
    import rimu_b200 as R
    from rimu_b200 import _lib
    p = _lib.StepParams()
    R.IsDynamicSemistochastic().fill(p)
    assert (p.style, p.rel_threshold, p.abs_threshold, p.proj_threshold, p.compress_threshold) == \
        (_lib.STYLE_SEMISTOCHASTIC, 1.0, math.inf, 0.0, 1.0)
    p = _lib.StepParams()
    R.IsStochasticInteger().fill(p)
    assert p.style == _lib.STYLE_INTEGER and p.compress_threshold == 0.0
    p = _lib.StepParams()
    R.IsDeterministic().fill(p)
    assert p.style == _lib.STYLE_DETERMINISTIC
    assert R.step_stats(R.IsStochasticInteger()) == ("spawn_attempts", "spawns", "deaths", "clones", "zombies")
    assert R.step_stats(R.IsDynamicSemistochastic()) == ("exact_steps", "inexact_steps", "spawn_attempts", "spawns", "len_before")
    assert R.step_stats(R.IsDeterministic()) == ("exact_steps",)


def test_shift_strategies(built):
    """shiftstrategy.jl:174-181: S <- S - xi/dt ln(N/N_t) - zeta/dt ln(N/N_prev); defaults zeta=0.08,
    xi=zeta^2/4 (projector_monte_carlo_problem.jl:167-168)."""
    import rimu_b200 as R
    s = R.DoubleLogUpdate(target_walkers=1000)
    assert s.zeta == 0.08 and s.xi == 0.08 ** 2 / 4
    sp = R.ShiftParameters(1.0, 100.0, 0.01)
    s.update(sp, 150.0)
    want = 1.0 - s.xi / 0.01 * math.log(150 / 1000) - 0.08 / 0.01 * math.log(150 / 100)
    assert sp.shift == want and sp.pnorm == 150.0
    sp = R.ShiftParameters(1.0, 100.0, 0.01)
    R.LogUpdate(0.1).update(sp, 120.0)
    assert sp.shift == 1.0 - 0.1 / 0.01 * math.log(1.2)
    sp = R.ShiftParameters(1.0, 100.0, 0.01)
    st = R.DoubleLogUpdateAfterTargetWalkers(target_walkers=1000)
    st.update(sp, 500.0)
    assert sp.shift == 1.0 and not sp.shift_mode
    st.update(sp, 1500.0)
    assert sp.shift_mode and sp.shift != 1.0
    # LogUpdateAfterTargetWalkers (shiftstrategy.jl:100-122): frozen shift below the target, LogUpdate afterwards, for good
    sp = R.ShiftParameters(2.0, 100.0, 0.01)
    lt = R.LogUpdateAfterTargetWalkers(target_walkers=1000, zeta=0.05)
    lt.update(sp, 800.0)
    assert sp.shift == 2.0 and sp.pnorm == 800.0 and not sp.shift_mode
    lt.update(sp, 1200.0)
    assert sp.shift_mode and sp.shift == 2.0 - 0.05 / 0.01 * math.log(1200 / 800)
    before = sp.shift
    lt.update(sp, 900.0)  # stays in shift mode below the target
    assert sp.shift == before - 0.05 / 0.01 * math.log(900 / 1200)


def test_blocking_analysis(built):
    """StatsTools/blocking.jl:134-159,274-325 on synthetic data: uncorrelated noise needs no blocking
    (k=1, err = std/sqrt(n)); AR(1) noise needs k>1 and the error covers the true mean."""
    import rimu_b200 as R
    rng = np.random.default_rng(5)
    x = rng.normal(3.0, 2.0, size=4096)
    b = R.blocking_analysis(x)
    assert 1 <= b.k <= 4 and math.isclose(b.err, x.std(ddof=1) / math.sqrt(len(x)), rel_tol=0.15)
    rows = R.statstools._blocks_with_m(x)  # first row = unblocked data: err = std/sqrt(n), Eq. (28) error on it
    assert math.isclose(rows[0][2], x.std(ddof=1) / math.sqrt(len(x)), rel_tol=1e-12)
    assert math.isclose(rows[0][3], rows[0][2] / math.sqrt(2 * (len(x) - 1)), rel_tol=1e-12)
    assert abs(b.mean - 3.0) < 4 * b.err
    y = np.zeros(2 ** 14)
    e = rng.normal(size=len(y))
    for i in range(1, len(y)):
        y[i] = 0.9 * y[i - 1] + e[i]
    b = R.blocking_analysis(y + 1.0)
    assert b.k > 3 and b.err > 3 * (y.std(ddof=1) / math.sqrt(len(y)))
    assert abs(b.mean - 1.0) < 4 * b.err
    r = R.ratio_of_means(2.0 * (y + 5.0) + rng.normal(size=len(y)) * 0.01, y + 5.0)
    assert abs(r.f - 2.0) < 5 * max(r.sigma_f, 1e-4)


def test_problem_defaults(built):
    """projector_monte_carlo_problem.jl:158-168,243-253 defaults."""
    import rimu_b200 as R

    class FakeHam:
        pass
    p = R.ProjectorMonteCarloProblem(FakeHam())
    assert p.time_step == 0.01 and p.last_step == 100 and p.max_length == 2 * 1000 + 100
    assert isinstance(p.shift_strategy, R.DoubleLogUpdate) and p.shift_strategy.target_walkers == 1000
    assert isinstance(p.style, R.IsDynamicSemistochastic)
    assert R.ProjectorMonteCarloProblem(FakeHam(), n_replicas=3).n_replicas == 3
    with pytest.raises(ValueError):
        R.ProjectorMonteCarloProblem(FakeHam(), n_replicas=0)


def test_initiator_rule_keywords(built):
    """PDVec / InitiatorDVec / ProjectorMonteCarloProblem keyword handling for initiator rules
    (pdvec.jl:181-199, initiatordvec.jl:42-77, projector_monte_carlo_problem.jl:156-160) and the StepParams the
    step receives -- host logic only, no device call."""
    import ctypes as C
    import rimu_b200 as R
    from rimu_b200 import _lib
    from rimu_b200.stochasticstyles import as_initiator_rule
    assert as_initiator_rule(None) == R.NonInitiator() and as_initiator_rule(False) == R.NonInitiator()
    assert as_initiator_rule(True) == R.Initiator(1.0)
    assert as_initiator_rule(None, 2.5) == R.Initiator(2.5)
    assert as_initiator_rule(R.CoherentInitiator(3.0)) == R.CoherentInitiator(3.0)
    assert [r.rule_id for r in (R.NonInitiator(), R.Initiator(), R.SimpleInitiator(), R.CoherentInitiator())] == [0, 1, 2, 3]
    assert (_lib.lib().rimu_sizeof_step_params(), _lib.lib().rimu_sizeof_step_stats()) == (C.sizeof(_lib.StepParams), C.sizeof(_lib.StepStats))
    names = [f for f, _ in _lib.StepParams._fields_]
    assert names[-3:] == ["initiator_rule", "ordered", "initiator_threshold"]

    class FakeHam:
        pass
    assert R.ProjectorMonteCarloProblem(FakeHam()).initiator == R.NonInitiator()
    assert R.ProjectorMonteCarloProblem(FakeHam(), initiator=True).initiator == R.Initiator(1.0)
    assert R.ProjectorMonteCarloProblem(FakeHam(), initiator=R.SimpleInitiator(2.0)).initiator == R.SimpleInitiator(2.0)
    # the oracle uses the same rule numbering
    from oracle import oracle as orc
    assert (orc.NON_INITIATOR, orc.INITIATOR, orc.SIMPLE_INITIATOR, orc.COHERENT_INITIATOR) == (0, 1, 2, 3)


def test_general_composite_addresses(built):
    """CompositeFS beyond two small fermion components (multicomponent.jl:10-34): packed layout == the oracle's, round trip,
    address-type bookkeeping, and the errors of the constructor / of models that cannot take such an address."""
    import rimu_b200 as R
    from tests.cases import SPECS, oracle_ham, product_ham
    rng = np.random.default_rng(7)
    for name in [n for n in SPECS if n.startswith("rs_comp_")]:
        oh, ph = oracle_ham(name), product_ham(name)
        at = ph.address_type
        assert at.kind == R._lib.ADDR_COMPOSITE and at.words == oh.W and ph.address.key() == oh.start_key
        assert at.from_key(ph.address.key()) == ph.address
        assert R.dimension(ph) == math.prod(
            math.comb(n + at.num_modes - 1, n) if k == R._lib.ADDR_BOSE else math.comb(at.num_modes, n)
            for k, n in zip(at.comp_kinds, at.num_particles))
        # a few random addresses of the same type
        for _ in range(10):
            comps = []
            for k, n in zip(at.comp_kinds, at.num_particles):
                if k == R._lib.ADDR_BOSE:
                    comps.append(R.BoseFS(tuple(int(x) for x in np.bincount(rng.integers(0, at.num_modes, size=n), minlength=at.num_modes))))
                else:
                    occ = np.zeros(at.num_modes, dtype=int)
                    occ[rng.choice(at.num_modes, size=n, replace=False)] = 1
                    comps.append(R.FermiFS(tuple(int(x) for x in occ)))
            a = R.CompositeFS(*comps)
            assert a.key() == oh.pack(tuple(c.onr for c in comps)) and at.from_key(a.key()) == a
    # two fermion components of at most 32 modes stay the one-word FermiFS2C layout; wider ones use the general layout
    assert R.FermiFS2C((1, 0), (0, 1)).address_type.kind == R._lib.ADDR_FERMI2C
    wide = R.CompositeFS(R.near_uniform(R.FermiFS, 3, 40), R.near_uniform(R.FermiFS, 2, 40))
    assert wide.address_type.kind == R._lib.ADDR_COMPOSITE and wide.address_type.words == 2
    with pytest.raises(ValueError, match="same number of modes"):       # ArgumentError in the reference (multicomponent.jl:27-29)
        R.CompositeFS(R.BoseFS((1, 1)), R.FermiFS((1, 0, 0)))
    with pytest.raises(TypeError):
        R.CompositeFS((1, 0), (0, 1))
    mixed = R.CompositeFS(R.BoseFS((1, 2, 0)), R.FermiFS((0, 1, 0)))
    for model in (R.HubbardMom1D, R.Transcorrelated1D, R.HubbardMom1DEP):
        with pytest.raises(TypeError):
            model(mixed)
    with pytest.raises(ValueError, match="127"):
        R.HubbardRealSpace(R.CompositeFS(R.near_uniform(R.BoseFS, 40, 40), R.near_uniform(R.BoseFS, 40, 40)))
    with pytest.raises(ValueError, match="components"):
        R.HubbardRealSpace(R.CompositeFS(*[R.FermiFS((1, 0, 0))] * 5))
    with pytest.raises(ValueError, match="symmetric"):                   # HubbardRealSpace.jl:188-189
        R.HubbardRealSpace(mixed, u=[[1, 2], [3, 4]])
    with pytest.raises(ValueError, match="length 2"):                    # :192-193
        R.HubbardRealSpace(mixed, t=[1, 2, 3])
