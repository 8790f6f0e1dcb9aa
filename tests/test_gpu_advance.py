"""rimu_advance: batches of FCIQMC steps with the shift update and the abort rules on the device (no host round trip between
steps) against (a) the step-by-step path through rimu_step + the host's shift strategies and (b) the CPU oracle driven with
the shifts the device reports.  Integer walkers bit-exact, Float64 at the summation-order tolerance of the default mode."""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from tests.cases import oracle_ham, product_ham

pytestmark = pytest.mark.gpu


def _strategy(R, sid, target, zeta=0.08):
    from rimu_b200 import _lib
    return {_lib.SHIFT_DONT_UPDATE: R.DontUpdate(target), _lib.SHIFT_LOG_UPDATE: R.LogUpdate(zeta),
            _lib.SHIFT_LOG_UPDATE_AFTER_TARGET: R.LogUpdateAfterTargetWalkers(target, zeta),
            _lib.SHIFT_DOUBLE_LOG_UPDATE: R.DoubleLogUpdate(target, zeta),
            _lib.SHIFT_DOUBLE_LOG_UPDATE_AFTER_TARGET: R.DoubleLogUpdateAfterTargetWalkers(target, zeta)}[sid]


def _step_by_step(R, ph, style, start, seed, dtau, shift0, strategy, nsteps, max_length=10**9, initiator=None):
    """the reference loop: apply_operator!, swap, update_shift_parameters!, abort rules (fciqmc.jl:126-181)"""
    v = R.GPUDVec([(ph.address, start)], style=style, initiator=initiator)
    pv = v.similar()
    wm = R.working_memory(v, seed=seed)
    sp = R.ShiftParameters(shift0, v.walkernumber(), dtau)
    is_int = style.val_type == R._lib.VAL_I64
    rows = []
    for _ in range(nsteps):
        R.apply_operator(wm, pv, v, R.FirstOrderTransitionOperator(ph, sp.shift, dtau))
        v, pv = pv, v
        s = wm.last_stats
        tnorm = float(s.inorm1) if is_int else s.norm1
        proceed = True
        if s.len > 0:
            _, proceed = strategy.update(sp, tnorm)
        rows.append((s.len, tnorm, s.spawn_attempts, sp.shift))
        if s.len == 0 or s.len > max_length or not proceed:
            break
    return v, sp, rows


def _batched(R, ph, style, start, seed, dtau, shift0, sid, target, nsteps, max_length=0, initiator=None, calls=1, zeta=0.08):
    v = R.GPUDVec([(ph.address, start)], style=style, initiator=initiator)
    pv = v.similar()
    wm = R.working_memory(v, seed=seed)
    sp = R.ShiftParameters(shift0, v.walkernumber(), dtau)
    strat = _strategy(R, sid, target, zeta)
    rows, is_int = [], style.val_type == R._lib.VAL_I64
    per_call = -(-nsteps // calls)
    left = nsteps
    while left > 0:
        k = min(per_call, left)
        v, pv, stats, shifts, done = R.advance(wm, v, pv, ph, sp, sid, target_walkers=target, zeta=getattr(strat, "zeta", 0.0),
                                               xi=getattr(strat, "xi", 0.0) or 0.0, nsteps=k, max_length=max_length)
        rows += [(s.len, float(s.inorm1) if is_int else s.norm1, s.spawn_attempts, sh) for s, sh in zip(stats, shifts)]
        left -= k
        if done < k:
            break
    return v, sp, rows


CASES = [("real1d_10", "int", 3), ("real1d_6", "semi", 3), ("mom1d_bose", "semi", 4), ("rs_f2c_4x4", "int", 1), ("tc_7", "semi", 3),
         ("real1d_w2", "int", 3), ("rs_comp_bf", "semi", 2), ("mom1d_bose", "int", 0)]


@pytest.mark.parametrize("name,style_name,sid", CASES)
def test_batch_is_the_step_by_step_trajectory(built, name, style_name, sid):
    """200 steps in two rimu_advance calls (chunks of 128 inside) == 200 x (rimu_step + host shift update)."""
    import rimu_b200 as R
    ph = product_ham(name)
    is_int = style_name == "int"
    style = R.IsStochasticInteger() if is_int else R.IsDynamicSemistochastic()
    start = 50 if is_int else 50.0
    dtau = 0.002 if name.startswith("tc") else 0.005
    shift0 = R.diagonal_element(ph, ph.address)
    target, nsteps = 400.0, 200
    if sid == 0:
        shift0 += 40.0  # DontUpdate: a growing population that hits the target and stops
    va, spa, ra = _step_by_step(R, ph, style, start, 99, dtau, shift0, _strategy(R, sid, target), nsteps)
    vb, spb, rb = _batched(R, ph, style, start, 99, dtau, shift0, sid, target, nsteps, calls=2)
    assert len(ra) == len(rb), (len(ra), len(rb))
    if sid == 0:
        assert len(ra) < nsteps  # the run ended when the walker number reached the target
    for k, (a, b) in enumerate(zip(ra, rb)):
        assert a[0] == b[0] and a[2] == b[2], (name, k, a, b)                       # length, spawn attempts
        assert (a[1] == b[1]) if is_int else math.isclose(a[1], b[1], rel_tol=1e-9), (name, k, a, b)
        assert math.isclose(a[3], b[3], rel_tol=1e-11, abs_tol=1e-11), (name, k, a, b)  # shift after the update
    assert math.isclose(spa.shift, spb.shift, rel_tol=1e-11, abs_tol=1e-11) and math.isclose(spa.pnorm, spb.pnorm, rel_tol=1e-9)
    assert spa.shift_mode == spb.shift_mode
    ka, xa = va.download_sorted()
    kb, xb = vb.download_sorted()
    assert np.array_equal(ka, kb)
    assert np.array_equal(xa, xb) if is_int else np.allclose(xa, xb, rtol=1e-9, atol=1e-9 * np.abs(xa).max())
    assert vb.walkernumber() == pytest.approx(rb[-1][1], rel=1e-9) and len(vb) == rb[-1][0]


def test_batch_against_the_oracle(built):
    """integer walkers, 60 steps in one call: the oracle stepped with the shifts the device reports reproduces every step's
    statistics and the final vector bit for bit; the reported shifts obey DoubleLogUpdate (shiftstrategy.jl:160-181)."""
    import rimu_b200 as R
    from rimu_b200 import _lib
    name = "real1d_10"
    oh, ph = oracle_ham(name), product_ham(name)
    seed, dtau, target, zeta = 4242, 0.005, 300.0, 0.08
    xi = zeta * zeta / 4
    shift0 = oh.diagonal_element(oh.start_key)
    vb, spb, rows = _batched(R, ph, R.IsStochasticInteger(), 40, seed, dtau, shift0, _lib.SHIFT_DOUBLE_LOG_UPDATE, target, 60)
    ok, ov = np.array([oh.start_key], dtype=np.uint64), np.array([40], dtype=np.int64)
    shift, pnorm = shift0, 40.0
    for k, (length, tnorm, attempts, shift_after) in enumerate(rows):
        pp = orc.make_params(orc.STYLE_INTEGER, shift=shift, dtau=dtau, key=orc.step_key(seed, k))
        ok, ov, st = oh.step(pp, ok, ov)
        assert (st.len_after, float(st.inorm1), st.spawn_attempts) == (length, tnorm, attempts), k
        want = shift - xi / dtau * math.log(tnorm / target) - zeta / dtau * math.log(tnorm / pnorm)
        assert math.isclose(shift_after, want, rel_tol=1e-13, abs_tol=1e-13), (k, shift_after, want)
        shift, pnorm = shift_after, tnorm  # (the device's own shift drives the next oracle step)
    gk, gv = vb.download_sorted()
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    assert len(rows) == 60


def test_batch_rolls_back_when_a_chunk_outgrows_its_memory(built):
    """a population that explodes inside a chunk (shift far above the ground state) outgrows the vectors / the bucket count the
    chunk was sized for: the chunk is rolled back to its snapshot and repeated step by step -- same trajectory."""
    import rimu_b200 as R
    from rimu_b200 import _lib
    ph = product_ham("real1d_10")
    style = R.IsStochasticInteger()
    shift0 = R.diagonal_element(ph, ph.address) + 25.0
    va, spa, ra = _step_by_step(R, ph, style, 20, 7, 0.01, shift0, R.LogUpdate(0.0), 40)
    vb, spb, rb = _batched(R, ph, style, 20, 7, 0.01, shift0, _lib.SHIFT_LOG_UPDATE, 0.0, 40, zeta=0.0)
    assert ra[-1][0] > 20000  # it did explode
    assert [(a[0], a[1], a[2]) for a in ra] == [(b[0], b[1], b[2]) for b in rb]
    ka, xa = va.download_sorted()
    kb, xb = vb.download_sorted()
    assert np.array_equal(ka, kb) and np.array_equal(xa, xb)


def test_batch_abort_rules(built):
    """max_length ends the run at the step that exceeds it, and the state is that step's (advance!, fciqmc.jl:171-178)"""
    import rimu_b200 as R
    from rimu_b200 import _lib
    ph = product_ham("real1d_10")
    style = R.IsStochasticInteger()
    shift0 = R.diagonal_element(ph, ph.address) + 8.0
    va, spa, ra = _step_by_step(R, ph, style, 30, 11, 0.005, shift0, R.LogUpdate(0.0), 300, max_length=600)
    vb, spb, rb = _batched(R, ph, style, 30, 11, 0.005, shift0, _lib.SHIFT_LOG_UPDATE, 0.0, 300, max_length=600, zeta=0.0)
    assert len(ra) < 300 and ra[-1][0] > 600
    assert [(a[0], a[1], a[2]) for a in ra] == [(b[0], b[1], b[2]) for b in rb]
    ka, xa = va.download_sorted()
    kb, xb = vb.download_sorted()
    assert np.array_equal(ka, kb) and np.array_equal(xa, xb)


def test_solve_reports_the_same_rows_with_and_without_device_batches(built):
    """ProjectorMonteCarloProblem(...; device_steps=K): same report columns and values as the step-by-step driver, incl. the
    initiator rule, ThresholdCompression's len_before and a reporting interval."""
    import rimu_b200 as R
    ph = product_ham("mom1d_bose")
    dfs = []
    for ds in (1, 50):
        prob = R.ProjectorMonteCarloProblem(ph, style=R.IsDynamicSemistochastic(), time_step=0.002, last_step=330, target_walkers=500,
                                            random_seed=5, initiator=True, reporting_interval=3, device_steps=ds)
        sim = R.solve(prob)
        assert sim.success and sim.step == 330
        dfs.append(sim.dataframe())
    a, b = dfs
    assert list(a.columns) == list(b.columns) and len(a) == len(b) == 110
    for col in a.columns:
        if a[col].dtype.kind in "iub":
            assert (a[col] == b[col]).all(), col
        else:
            assert np.allclose(a[col], b[col], rtol=1e-9, atol=1e-9), col
    # integer walkers + an *AfterTargetWalkers strategy: the shift_mode column switches at the same step
    ph = product_ham("real1d_6")
    dfs = []
    for ds in (1, 64):
        prob = R.ProjectorMonteCarloProblem(ph, style=R.IsStochasticInteger(), time_step=0.005, last_step=200, random_seed=3,
                                            shift_strategy=R.DoubleLogUpdateAfterTargetWalkers(target_walkers=300), device_steps=ds)
        dfs.append(R.solve(prob).dataframe())
    a, b = dfs
    assert list(a.columns) == list(b.columns) and "shift_mode" in a.columns
    assert (a["shift_mode"] == b["shift_mode"]).all() and a["shift_mode"].any() and not a["shift_mode"].all()
    assert (a["len"] == b["len"]).all() and (a["norm"] == b["norm"]).all() and np.allclose(a["shift"], b["shift"], rtol=1e-11)


def test_frozen_projections_inside_a_batch(built):
    """dot(::FrozenDVec, v) after every step of a batch == the host-driven dot after every rimu_step (pdvec.jl:773-779), and
    `solve` with a ProjectedEnergy report (poststepstrategy.jl:82-121) gives the same vproj / hproj columns either way."""
    import rimu_b200 as R
    from rimu_b200 import _lib
    ph = product_ham("real1d_6")
    style = R.IsDynamicSemistochastic()
    ref = R.GPUDVec([(ph.address, 1.0)], style=R.IsDeterministic())
    pe = R.ProjectedEnergy(ph, ref)
    frozen = [pe.vproj, pe.hproj]
    shift0, dtau, target = R.diagonal_element(ph, ph.address), 0.004, 300.0
    # step by step
    v = R.GPUDVec([(ph.address, 30.0)], style=style)
    pv = v.similar()
    wm = R.working_memory(v, seed=21)
    sp = R.ShiftParameters(shift0, v.walkernumber(), dtau)
    strat = R.DoubleLogUpdate(target, 0.08)
    want = []
    for _ in range(150):
        R.apply_operator(wm, pv, v, R.FirstOrderTransitionOperator(ph, sp.shift, dtau))
        v, pv = pv, v
        strat.update(sp, wm.last_stats.norm1)
        want.append([fr.dot(v) for fr in frozen])
    # one call
    v2 = R.GPUDVec([(ph.address, 30.0)], style=style)
    pv2 = v2.similar()
    wm2 = R.working_memory(v2, seed=21)
    sp2 = R.ShiftParameters(shift0, v2.walkernumber(), dtau)
    v2, pv2, stats, shifts, done, dots = R.advance(wm2, v2, pv2, ph, sp2, _lib.SHIFT_DOUBLE_LOG_UPDATE, target_walkers=target, zeta=0.08,
                                                   xi=0.08 ** 2 / 4, nsteps=150, projectors=frozen)
    assert done == 150 and dots.shape == (150, 2)
    assert np.allclose(dots, np.array(want), rtol=1e-9, atol=1e-9)
    assert abs(dots[:, 0]).min() > 0  # the reference determinant stays populated
    # the driver
    dfs = []
    for ds in (1, 40):
        prob = R.ProjectorMonteCarloProblem(ph, style=style, time_step=dtau, last_step=300, target_walkers=300, random_seed=8,
                                            post_step_strategy=(R.ProjectedEnergy(ph, ref), R.Projector(ones=ref)), device_steps=ds)
        sim = R.init(prob)
        assert (sim._batch_size() >= 2) == (ds > 1)
        dfs.append(sim.solve_().dataframe())
    a, b = dfs
    assert list(a.columns) == list(b.columns) and {"vproj", "hproj", "ones"} <= set(a.columns)
    for col in a.columns:
        assert np.allclose(a[col].astype(float), b[col].astype(float), rtol=1e-9, atol=1e-9), col
    e = R.projected_energy(b, skip=100)
    assert math.isclose(e.f, oracle_ham("real1d_6").exact_energy(), rel_tol=0.05)
    # a strategy that needs the host every step (non-Hermitian projected energy: dot(projector, H, v)) is not batched
    tc = product_ham("tc_7")
    tref = R.GPUDVec([(tc.address, 1.0)], style=R.IsDeterministic())
    prob = R.ProjectorMonteCarloProblem(tc, time_step=0.001, last_step=10, post_step_strategy=(R.ProjectedEnergy(tc, tref),))
    assert R.init(prob)._batch_size() == 0
