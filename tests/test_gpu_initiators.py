"""Initiator rules on the GPU (reference: DictVectors/initiators.jl:132-236,
PDWorkingMemory deposit! / move_and_compress! pdworkingmemory.jl:21-31,262-273).

* step-by-step parity with the CPU oracle for the three rules, integer (bit-exact) and semistochastic walkers, one- and
  two-word addresses; the first step runs on an unsegmented source (diagonal deposits travel as initiator-lane records),
  the following ones on segmented sources (parents staged by the merge kernel);
* the hand-computed two-site case of tests/test_oracle_pins.py::test_initiator_rules_hand_computed through the C ABI;
* the reference's own acceptance test test/lomc.jl:612-675 ("Energies below the plateau & initiator bias").
"""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from tests.cases import oracle_ham, product_ham

pytestmark = pytest.mark.gpu

RULES = {"initiator": 1, "simple": 2, "coherent": 3}


def _rule(R, name, thr):
    return {"initiator": R.Initiator, "simple": R.SimpleInitiator, "coherent": R.CoherentInitiator}[name](thr)


@pytest.mark.parametrize("rule", sorted(RULES))
@pytest.mark.parametrize("name,style_name", [("real1d_10", "int"), ("mom1d_bose", "semi"), ("rs_bose_3d_w2", "int"),
                                             ("rs_f2c_4x4", "semi")])
def test_initiator_steps_match_oracle(built, name, style_name, rule):
    import rimu_b200 as R
    oh, ph = oracle_ham(name), product_ham(name)
    seed, dtau, thr = 31, (0.001 if name == "rs_bose_3d_w2" else 0.01), 1.0
    if style_name == "int":
        style, pop, dtype, ostyle, kw = R.IsStochasticInteger(), 40, np.int64, orc.STYLE_INTEGER, {}
    else:
        style, pop, dtype, ostyle, kw = R.IsDynamicSemistochastic(), 30.5, np.float64, orc.STYLE_SEMISTOCHASTIC, dict(compress_threshold=1.0)
    v = R.GPUDVec([(ph.address, pop)], style=style, initiator=_rule(R, rule, thr))
    assert v.similar().initiator == v.initiator
    wm = R.working_memory(v, seed=seed)
    ok = np.array([oh.start_key], dtype=np.uint64).reshape(1, -1)
    ov = np.array([pop], dtype=dtype)
    shift = oh.diagonal_element(oh.start_key) + 2.0
    dropped = 0
    for step in range(7):
        out = v.similar()
        R.apply_operator(wm, out, v, R.FirstOrderTransitionOperator(ph, shift, dtau))
        v = out
        pp = orc.make_params(ostyle, shift=shift, dtau=dtau, key=orc.step_key(seed, step), initiator_rule=RULES[rule],
                             initiator_threshold=thr, **kw)
        ok, ov, st = oh.step(pp, ok, ov)
        gk, gv = v.download_sorted()
        assert np.array_equal(gk.reshape(len(gv), -1), ok.reshape(len(ov), -1)), (name, rule, step, len(gv), len(ov))
        s = wm.last_stats
        assert (s.spawn_attempts, s.len_before, s.len) == (st.spawn_attempts, st.len_before, st.len_after), (name, rule, step)
        dropped += st.len_before - st.len_after
        if style_name == "int":
            assert np.array_equal(gv, ov), (name, rule, step)
            assert (s.ispawns, s.ideaths, s.iclones, s.izombies, s.inorm1) == (st.ispawns, st.ideaths, st.iclones, st.izombies, st.inorm1)
        else:
            assert np.allclose(gv, ov, rtol=1e-10, atol=0), (name, rule, step)
            assert math.isclose(s.norm1, st.norm1, rel_tol=1e-10)
            ok, ov = gk, gv  # feed the GPU values back so that last-bit differences cannot flip branches
    assert dropped > 0, "the rule never suppressed anything: the test does not exercise it"


def test_initiator_hand_computed_two_sites(built):
    """Same numbers as tests/test_oracle_pins.py::test_initiator_rules_hand_computed, through rimu_step (operator = H)."""
    import rimu_b200 as R
    a11, a20, a02 = R.BoseFS((1, 1)), R.BoseFS((2, 0)), R.BoseFS((0, 2))
    H = R.HubbardReal1D(a11, u=1.0, t=1.0)
    r2 = math.sqrt(2.0)
    unsafe = 2 * (-r2 * 0.5)
    cases = [
        ([(a11, 2.0), (a20, 0.5)], {"none": {a02: -4 * r2, a20: -4 * r2 + 0.5, a11: unsafe}, "initiator": {a02: -4 * r2, a20: -4 * r2 + 0.5},
                                     "simple": {a02: -4 * r2, a20: -4 * r2 + 0.5}, "coherent": {a02: -4 * r2, a20: -4 * r2 + 0.5, a11: unsafe}}),
        ([(a20, 3.0), (a11, 0.5)], {"none": {a20: 3.0 + unsafe, a02: unsafe, a11: -6 * r2}, "initiator": {a20: 3.0 + unsafe, a11: -6 * r2},
                                     "simple": {a20: 3.0, a11: -6 * r2}, "coherent": {a20: 3.0 + unsafe, a02: unsafe, a11: -6 * r2}}),
    ]
    for pairs, wants in cases:
        for rname, want in wants.items():
            rule = R.NonInitiator() if rname == "none" else _rule(R, rname, 1.0)
            v = R.GPUDVec(pairs, style=R.IsDeterministic(), initiator=rule)
            out = v.similar()
            R.apply_operator(R.working_memory(v), out, v, H)
            got = {tuple(k): val for k, val in zip(*out.download_sorted())}
            ref = {tuple(np.atleast_1d(a.key())): val for a, val in want.items()}
            assert got.keys() == ref.keys(), (rname, got)
            for k in ref:
                assert math.isclose(got[k], ref[k], rel_tol=1e-14), (rname, k, got[k], ref[k])


def test_initiator_bias_below_the_plateau(built):
    """test/lomc.jl:612-675: HubbardMom1D(BoseFS{10,10} in one mode; u=4), 300 walkers (below the annihilation plateau),
    dtau = 5e-4, 6000 steps, zeta = 0.05.  Without initiators the shift is garbage BELOW the exact energy; every initiator
    rule is biased ABOVE it, SimpleInitiator most; Initiator and CoherentInitiator agree."""
    import rimu_b200 as R
    addr = R.BoseFS((0, 0, 0, 0, 10, 0, 0, 0, 0, 0))
    H = R.HubbardMom1D(addr, u=4.0)
    E0 = -9.251592973178997  # the reference's pinned value (also pinned against the oracle in tests/test_oracle_pins.py)

    def run(initiator, seed):
        prob = R.ProjectorMonteCarloProblem(H, start_at=R.GPUDVec([(addr, 1.0)], style=R.IsDynamicSemistochastic(), initiator=initiator),
                                            time_step=5e-4, last_step=6000, random_seed=seed, max_length=10**6,
                                            shift_strategy=R.DoubleLogUpdate(target_walkers=300, zeta=0.05, xi=0.05 ** 2 / 4))
        sim = R.solve(prob)
        assert sim.success, sim.message
        sh = np.asarray(sim.dataframe()["shift"])[2000:]
        return float(sh.mean()), float(sh.std(ddof=1) / math.sqrt(len(sh)))  # mean_and_se as in the reference test

    E_no, s_no = run(None, 8008)
    E_ni, s_ni = run(R.NonInitiator(), 8008)
    E_i1, s_i1 = run(R.Initiator(1.0), 8008)
    E_i2, s_i2 = run(R.SimpleInitiator(1.0), 8008)
    E_i3, s_i3 = run(R.CoherentInitiator(1.0), 8008)
    assert E_no < E0 and E_ni < E0          # garbage energy from no initiator
    assert abs(E_no - E_ni) <= 3 * s_no      # NonInitiator is the plain step (reference: E_no ≈ E_ni atol=3σ_no)
    assert E_i1 > E0 and E_i2 > E0 and E_i3 > E0   # initiator has a bias
    assert E_i2 > E_i1                        # simple initiator has the largest bias
    # mean_and_se ignores autocorrelation (as the reference's test does); allow 5 naive standard errors
    assert abs(E_i1 - E_i3) < 5 * max(s_i1, s_i3) + 0.02 * abs(E0), (E_i1, E_i3, s_i1, s_i3)
