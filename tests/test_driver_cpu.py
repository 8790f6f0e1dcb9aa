"""The host driver above the C ABI -- ProjectorMonteCarloProblem / init / step! / solve, the shift strategies, the report
assembly and the abort rules (mirror of pmc_simulation.jl:88-174,265-452 and fciqmc.jl:126-181) -- run on the CPU against a
stand-in for the device: a vector class and an `apply_operator` that execute the ORACLE step.  Nothing of this touches the
product's compute path (that is what the GPU tests are for); it keeps the driver logic under test in the CPU suite.
The same runs through the real device are tests/test_gpu_energies.py."""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from tests.cases import oracle_ham, product_ham


class _Ctx:
    nranks = 1


class FakeDVec:
    """host stand-in with the GPUDVec surface the driver uses"""

    def __init__(self, pairs=None, *, style=None, address_type=None, capacity=0, ctx=None, initiator=None, initiator_threshold=None):
        import rimu_b200 as R
        from rimu_b200.stochasticstyles import as_initiator_rule
        items = list(pairs or [])
        self.style = style or R.IsDynamicSemistochastic()
        self.address_type = address_type or items[0][0].address_type
        self.ctx, self.initiator = ctx or _Ctx(), as_initiator_rule(initiator, initiator_threshold)
        self.dtype = np.int64 if self.style.val_type == R._lib.VAL_I64 else np.float64
        W = self.address_type.words
        self.keys = np.array([np.atleast_1d(a.key()) for a, _ in items], dtype=np.uint64).reshape(-1, W)
        self.vals = np.array([v for _, v in items], dtype=self.dtype)

    def __len__(self):
        return len(self.vals)

    def similar(self, style=None):
        return FakeDVec(style=style or self.style, address_type=self.address_type, ctx=self.ctx, initiator=self.initiator)

    zerovector = similar

    def copy(self):
        out = self.similar()
        out.keys, out.vals = self.keys.copy(), self.vals.copy()
        return out

    def copy_from(self, other):
        self.keys, self.vals = other.keys.copy(), other.vals.astype(self.dtype)
        return self

    def norm(self, p=2):
        a = np.abs(self.vals.astype(float))
        return float(a.sum()) if p == 1 else float(np.sqrt((a * a).sum())) if p == 2 else float(a.max(initial=0.0))

    def walkernumber(self):
        return self.norm(1)

    def dot(self, other):
        d = {tuple(k): v for k, v in zip(other.keys.tolist(), other.vals.tolist())}
        return float(sum(v * d.get(tuple(k), 0.0) for k, v in zip(self.keys.tolist(), self.vals.tolist())))

    def freeze(self):
        return self.copy()

    def scale_(self, alpha):
        self.vals = self.vals * alpha
        return self

    def add_(self, other, alpha=1.0):
        d = {tuple(k): v for k, v in zip(self.keys.tolist(), self.vals.tolist())}
        for k, v in zip(other.keys.tolist(), other.vals.tolist()):
            d[tuple(k)] = d.get(tuple(k), 0.0) + alpha * v
        items = sorted((k, v) for k, v in d.items() if v != 0.0)
        self.keys = np.array([k for k, _ in items], dtype=np.uint64).reshape(-1, self.address_type.words)
        self.vals = np.array([v for _, v in items], dtype=self.dtype)
        return self


@pytest.fixture
def cpu_device(monkeypatch):
    """route the driver's vector class and apply_operator to the oracle"""
    import rimu_b200 as R
    from rimu_b200 import _lib, dictvectors, fciqmc
    registry = {}

    def fake_apply_operator(wm, target, source, op, boost=1.0, table_slots=0):
        assert target is not source
        if isinstance(op, R.FirstOrderTransitionOperator):
            ham, plain, shift, dt = op.hamiltonian, False, op.shift, op.time_step
        else:
            ham, plain, shift, dt = op, True, 0.0, 0.0
        oh = registry[id(ham)]
        p = _lib.StepParams()
        wm.style.fill(p)
        rule = wm.initiator
        pp = orc.make_params(p.style, shift=shift, dtau=dt, boost=boost, plain_h=plain, proj_threshold=p.proj_threshold,
                             rel_threshold=p.rel_threshold, abs_threshold=p.abs_threshold, compress_threshold=p.compress_threshold,
                             key=orc.step_key(wm.seed, wm.counter), initiator_rule=rule.rule_id, initiator_threshold=rule.threshold)
        ko, vo, st = oh.step(pp, source.keys, source.vals)
        target.keys, target.vals = ko.reshape(len(vo), -1), vo
        s = _lib.StepStats()
        for f in ("exact_steps", "inexact_steps", "spawn_attempts", "len_before", "spawns", "deaths", "clones", "zombies", "norm1",
                  "ispawns", "ideaths", "iclones", "izombies", "inorm1"):
            setattr(s, f, getattr(st, f))
        s.len = st.len_after
        wm.counter += 1
        wm.last_stats = s
        names, values = wm.style.stat_names, wm.style.stats(s)
        if isinstance(getattr(wm.style, "compression", None), R.ThresholdCompression):
            names, values = names + ("len_before",), values + (s.len_before,)
        return names, values, wm, target

    def fake_dot(x, *args):
        if len(args) == 1:
            return x.dot(args[0])
        op, y = args
        oh = registry[id(op)]
        ko, vo, _ = oh.step(orc.make_params(orc.STYLE_DETERMINISTIC, plain_h=True), y.keys, y.vals.astype(np.float64))
        tmp = y.similar()
        tmp.keys, tmp.vals = ko.reshape(len(vo), -1), vo
        return x.dot(tmp)

    def fake_mul(y, op, x, wm=None):
        oh = registry[id(op)]
        ko, vo, _ = oh.step(orc.make_params(orc.STYLE_DETERMINISTIC, plain_h=True), x.keys, x.vals.astype(np.float64))
        y.keys, y.vals = ko.reshape(len(vo), -1), vo
        return y

    from rimu_b200 import lanczos
    for mod in (lanczos,):
        monkeypatch.setattr(mod, "GPUDVec", FakeDVec)
        monkeypatch.setattr(mod, "mul", fake_mul)
        monkeypatch.setattr(mod, "WorkingMemory", lambda v, seed=0: None)
    for mod in (fciqmc, dictvectors):
        monkeypatch.setattr(mod, "GPUDVec", FakeDVec)
        monkeypatch.setattr(mod, "apply_operator", fake_apply_operator)
        monkeypatch.setattr(mod, "dot", fake_dot)
    monkeypatch.setattr(R, "GPUDVec", FakeDVec)

    def make(name):
        oh, ph = oracle_ham(name), product_ham(name)
        registry[id(ph)] = oh
        return oh, ph
    return make


def test_solve_reproduces_the_bhm_example_on_the_cpu(built, cpu_device):
    """scripts/BHM-example.jl through the Python driver (oracle step underneath): columns, shift mode, energy."""
    import rimu_b200 as R
    oh, ph = cpu_device("real1d_6")
    prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, style=R.IsDynamicSemistochastic(), time_step=0.001, last_step=3000,
                                        target_walkers=1000, random_seed=17)
    sim = R.init(prob)
    assert sim.state.shift_parameters.shift == oh.diagonal_element(oh.start_key)  # Rayleigh quotient of the start vector (fciqmc.jl:51-61)
    assert sim.state.shift_parameters.pnorm == 10.0                              # default_starting_vector: address => 10
    R.step_(sim)
    assert sim.step == 1 and not sim.success
    R.solve_(sim)
    assert sim.success and sim.step == 3000 and not sim.aborted
    df = sim.dataframe()
    assert list(df.columns[:4]) == ["step", "len", "shift", "norm"]
    for col in ("exact_steps", "inexact_steps", "spawn_attempts", "spawns", "len_before"):  # styles.jl:203-209 + compression.jl:16
        assert col in df.columns
    assert len(df) == 3000 and df["step"].iloc[-1] == 3000
    se = R.shift_estimator(df, skip=1000)
    assert abs(se.mean - (-4.0215)) < 5 * se.err + 0.04
    assert abs(np.asarray(df["norm"])[1500:].mean() - 1000) < 100
    # solve! continues when last_step is raised (pmc_simulation.jl:400-452)
    R.solve_(sim, last_step=3010)
    assert sim.step == 3010 and len(sim.dataframe()) == 3010


def test_driver_options_on_the_cpu(built, cpu_device):
    import rimu_b200 as R
    oh, ph = cpu_device("real1d_6")
    # integer walkers: the norm fed to the shift strategy is the exact integer walker number; reporting_interval thins the report
    prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, style=R.IsStochasticInteger(), time_step=0.01, last_step=200,
                                        target_walkers=500, random_seed=3, reporting_interval=10)
    sim = R.solve(prob)
    df = sim.dataframe()
    assert sim.success and len(df) == 20 and list(df["step"][:3]) == [10, 20, 30]
    assert all(float(n).is_integer() for n in df["norm"])
    for col in ("spawn_attempts", "spawns", "deaths", "clones", "zombies"):  # styles.jl:14-20
        assert col in df.columns
    # max_length aborts (fciqmc.jl:172-179)
    prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, time_step=0.01, last_step=500, target_walkers=5000, max_length=20,
                                        random_seed=3)
    sim = R.solve(prob)
    assert sim.aborted and not sim.success and "Aborted in step" in sim.message and sim.step < 500
    # an explicit shift, DontUpdate and the initiator keyword reach the state
    prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, shift=-1.5, shift_strategy=R.DontUpdate(), last_step=5, random_seed=1,
                                        initiator=R.CoherentInitiator(2.0))
    sim = R.solve(prob)
    assert set(sim.dataframe()["shift"]) == {-1.5}
    assert sim.state.v.initiator == R.CoherentInitiator(2.0) and sim.state.wm.initiator == R.CoherentInitiator(2.0)
    # DontUpdate stops the run once the walker number reaches target_walkers (shiftstrategy.jl:89-91) and leaves pnorm alone;
    # the stop is reported like every other `proceed == false` (pmc_simulation.jl:301-304)
    prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, shift=20.0, shift_strategy=R.DontUpdate(target_walkers=200),
                                        time_step=0.01, last_step=5000, random_seed=1)
    sim = R.init(prob)
    pnorm0 = sim.state.shift_parameters.pnorm
    R.solve_(sim)
    df = sim.dataframe()
    assert sim.aborted and not sim.success and sim.step < 5000 and sim.message == f"Aborted in step {sim.step}."
    assert df["norm"].iloc[-1] >= 200 and all(n < 200 for n in df["norm"].iloc[:-1])
    assert sim.state.shift_parameters.pnorm == pnorm0 and set(df["shift"]) == {20.0}
    # wall time
    prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, last_step=10 ** 9, random_seed=1, wall_time=0.2)
    sim = R.solve(prob)
    assert sim.aborted and sim.message == "Wall time reached."


@pytest.mark.parametrize("name", ["real1d_6", "mom1d_bose", "rs_fermi", "ext1d"])
def test_lanczos_driver_on_the_cpu(built, cpu_device, name):
    """eigsolve_lanczos (the KrylovKit stand-in of config 3, ext/KrylovKitExt.jl:23-46) over the oracle's H*v: restarts,
    full reorthogonalisation and the Ritz-vector assembly reproduce the exact-diagonalisation energy."""
    import rimu_b200 as R
    from tests.test_gpu_energies import exact_energy
    oh, ph = cpu_device(name)
    start = FakeDVec([(ph.address, 1.0)], style=R.IsDeterministic())
    vals, vecs, info = R.eigsolve_lanczos(ph, start, krylovdim=40, tol=1e-10, maxiter=40, full_reorth=True)
    assert info["converged"], info
    assert math.isclose(vals[0], exact_energy(oh), rel_tol=1e-9, abs_tol=1e-9)
    assert math.isclose(vecs[0].norm(2), 1.0, rel_tol=1e-6)


def test_projected_energy_post_step_on_the_cpu(built, cpu_device):
    """ProjectedEnergy (poststepstrategy.jl:82-121): vproj = projector . v, hproj = (H projector) . v; their ratio of means
    estimates the energy (StatsTools/ratio_of_means.jl).  Integer and Float64 walkers."""
    import rimu_b200 as R
    oh, ph = cpu_device("real1d_6")
    for style in (R.IsDynamicSemistochastic(), R.IsStochasticInteger()):
        ref = FakeDVec([(ph.address, 1.0)], style=R.IsDeterministic())
        prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, style=style, time_step=0.002, last_step=2500, target_walkers=800,
                                            random_seed=9, post_step_strategy=(R.ProjectedEnergy(ph, ref), R.Projector(overlap=ref)))
        sim = R.solve(prob)
        df = sim.dataframe()
        assert {"vproj", "hproj", "overlap"} <= set(df.columns)
        assert np.array_equal(np.asarray(df["vproj"]), np.asarray(df["overlap"]))  # same projector
        pe = R.projected_energy(df, skip=800)
        assert pe.success and abs(pe.f - (-4.0215)) < 5 * pe.sigma_f + 0.04, (pe.f, pe.sigma_f)


def test_replicas_on_the_cpu(built, cpu_device):
    """n_replicas > 1 (projector_monte_carlo_problem.jl:152,205; qmc_states.jl:89-140): independent vectors and shift
    parameters, columns suffixed _1, _2, ...; one replica keeps the plain column names and the single-replica trajectory."""
    import rimu_b200 as R
    oh, ph = cpu_device("real1d_6")
    kw = dict(start_at=ph.address, style=R.IsDynamicSemistochastic(), time_step=0.002, last_step=1500, target_walkers=400, random_seed=4)
    one = R.solve(R.ProjectorMonteCarloProblem(ph, **kw)).dataframe()
    three = R.solve(R.ProjectorMonteCarloProblem(ph, n_replicas=3, **kw))
    df = three.dataframe()
    assert len(three.states) == 3 and "shift" not in df.columns
    for r in (1, 2, 3):
        for col in ("len", "shift", "norm", "spawn_attempts", "len_before"):
            assert f"{col}_{r}" in df.columns
    assert np.array_equal(np.asarray(df["shift_1"]), np.asarray(one["shift"]))       # replica 1 = the single-replica run (same seed)
    assert not np.array_equal(np.asarray(df["shift_1"]), np.asarray(df["shift_2"]))  # independent random streams
    for r in (1, 2, 3):
        se = R.blocking_analysis(np.asarray(df[f"shift_{r}"]), skip=500)
        assert abs(se.mean - (-4.0215)) < 5 * se.err + 0.06


def test_all_overlaps_on_the_cpu(built, cpu_device):
    """AllOverlaps (replicastrategy.jl:60-183): column names c{i}_dot_c{j} / c{i}_Op{k}_c{j}, values = dot(v_i, v_j) and
    dot(v_i, H, v_j); the replica (variational) energy sum c1.H.c2 / sum c1.c2 estimates E0 without the population-control
    bias of a single vector."""
    import rimu_b200 as R
    oh, ph = cpu_device("real1d_6")
    prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, style=R.IsDynamicSemistochastic(), time_step=0.002, last_step=1500,
                                        target_walkers=400, random_seed=6, replica_strategy=R.AllOverlaps(3, operator=ph))
    sim = R.solve(prob)
    df = sim.dataframe()
    assert len(sim.states) == 3
    pairs = [(1, 2), (1, 3), (2, 3)]
    for i, j in pairs:
        assert f"c{i}_dot_c{j}" in df.columns and f"c{i}_Op1_c{j}" in df.columns
    # the last row against the final vectors
    v1, v2 = sim.states[0].v, sim.states[1].v
    assert math.isclose(df["c1_dot_c2"].iloc[-1], v1.dot(v2), rel_tol=1e-12)
    num = sum(np.asarray(df[f"c{i}_Op1_c{j}"])[500:].sum() for i, j in pairs)
    den = sum(np.asarray(df[f"c{i}_dot_c{j}"])[500:].sum() for i, j in pairs)
    assert abs(num / den - (-4.0215)) < 0.05
    with pytest.raises(ValueError):
        R.ProjectorMonteCarloProblem(ph, n_replicas=2, replica_strategy=R.AllOverlaps(3))
    only = R.AllOverlaps(2, vecnorm=False, operator=(ph, ph))
    prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, last_step=3, random_seed=1, replica_strategy=only)
    cols = set(R.solve(prob).dataframe().columns)
    assert {"c1_Op1_c2", "c1_Op2_c2"} <= cols and "c1_dot_c2" not in cols


def test_gram_schmidt_spectral_states_on_the_cpu(built, cpu_device):
    """spectral_strategy=GramSchmidt(2) (fciqmc.jl:187-202, spectralstrategy.jl:22-35): the second spectral state is
    orthogonalised against the first before every step, so a deterministic run drives its shift to the lowest eigenvalue of the
    complement that its starting vector overlaps; report columns carry the _s1/_s2 suffixes (pmc_simulation.jl:133-146)."""
    import rimu_b200 as R
    oh, ph = cpu_device("real1d_6")
    basis = oh.bfs_basis()
    H = oh.sparse_matrix(basis).toarray()
    w, vecs = np.linalg.eigh(H)
    index = {int(k[0]): i for i, k in enumerate(basis)}
    a0 = ph.address
    k1, _ = oh.get_offdiagonal(oh.start_key, 1)
    a1 = ph.address_type.from_key(np.atleast_1d(k1))
    det = R.IsDeterministic()
    starts = [FakeDVec([(a0, 10.0)], style=det), FakeDVec([(a0, 3.0), (a1, 10.0)], style=det)]
    prob = R.ProjectorMonteCarloProblem(ph, start_at=starts, style=det, time_step=0.01, last_step=4000, target_walkers=100,
                                        spectral_strategy=R.GramSchmidt(2), random_seed=1, max_length=10 ** 6)
    sim = R.solve(prob)
    assert sim.success, sim.message
    df = sim.dataframe()
    for col in ("shift_s1", "shift_s2", "norm_s1", "norm_s2", "len_s1", "len_s2"):
        assert col in df.columns, df.columns
    # expected: E0 for the first state; for the second the lowest eigenvalue above E0 that start 2 overlaps
    s2 = np.zeros(len(basis))
    s2[index[int(np.atleast_1d(a0.key())[0])]], s2[index[int(np.atleast_1d(a1.key())[0])]] = 3.0, 10.0
    e1 = next(w[i] for i in range(1, len(w)) if abs(vecs[:, i] @ s2) > 1e-8 and w[i] > w[0] + 1e-9)
    assert abs(df["shift_s1"].iloc[-500:].mean() - w[0]) < 1e-3 * abs(w[0])
    assert abs(df["shift_s2"].iloc[-500:].mean() - e1) < 1e-2 * abs(e1), (df["shift_s2"].iloc[-500:].mean(), e1, w[:5])
    u, v = sim.replicas[0][1].v, sim.replicas[0][0].v
    R.GramSchmidt(2).orthogonalize(sim.replicas[0])
    assert abs(u.dot(v)) < 1e-9 * u.norm(2) * v.norm(2)


def test_report_to_file_on_the_cpu(built, cpu_device, tmp_path):
    """ReportToFile (reportingstrategy.jl:302-434): chunks of `chunk_size` reported steps go to an Arrow file, the in-memory
    report is emptied after every chunk, an existing file gets a numbered sibling, the metadata travel with the file, and
    load_df returns the same table a ReportDFAndInfo run produces."""
    import rimu_b200 as R
    oh, ph = cpu_device("real1d_6")
    kw = dict(start_at=ph.address, style=R.IsStochasticInteger(), time_step=0.01, last_step=250, target_walkers=200, random_seed=3)
    ref = R.solve(R.ProjectorMonteCarloProblem(ph, **kw)).dataframe()
    fn = tmp_path / "out.arrow"
    rs = R.ReportToFile(filename=fn, chunk_size=100, save_if=True)
    sim = R.solve(R.ProjectorMonteCarloProblem(ph, reporting_strategy=rs, metadata={"note": "abc"}, **kw))
    assert sim.success and rs.chunks_written == 3  # 100 + 100 + the last 50 at finalisation
    assert all(len(v) == 0 for v in sim.report.values())
    df = R.load_df(rs.filename)
    assert list(df.columns) == list(ref.columns) and len(df) == 250
    for col in ref.columns:
        assert np.array_equal(np.asarray(df[col]), np.asarray(ref[col])), col
    assert df.attrs["note"] == "abc" and "success" in df.attrs and df.attrs["num_replicas"] == "1"  # (metadata as of the first chunk)
    assert sim.dataframe().equals(df)
    # a second run with the same name does not overwrite (reportingstrategy.jl:362-381)
    rs2 = R.ReportToFile(filename=fn, chunk_size=1000, reporting_interval=10, save_if=True)
    sim2 = R.solve(R.ProjectorMonteCarloProblem(ph, reporting_strategy=rs2, **kw))
    assert rs2.filename.endswith("out-1.arrow") and len(R.load_df(rs2.filename)) == 25 and sim2.success
    # save_if = false: nothing is written
    rs3 = R.ReportToFile(filename=tmp_path / "none.arrow", save_if=False)
    R.solve(R.ProjectorMonteCarloProblem(ph, reporting_strategy=rs3, **kw))
    assert not (tmp_path / "none.arrow").exists()
    with pytest.raises(ValueError):
        R.ReportToFile(compress="gzip")


def test_device_batches_assemble_the_same_report_on_the_cpu(built, cpu_device, monkeypatch):
    """`solve(...; device_steps=K)` hands K steps to one `advance` call (rimu_advance on the device) and replays the report rows
    from the per-step statistics it returns.  Here `advance` is a host loop over the oracle-backed step with the host's own
    shift strategies, so what is under test is the driver's bookkeeping: same columns and rows as the step-by-step loop --
    incl. reporting_interval, *AfterTargetWalkers shift_mode, projections, the abort rules and solve!'s continuation."""
    import rimu_b200 as R
    from rimu_b200 import _lib, dictvectors, fciqmc
    oh, ph = cpu_device("real1d_6")
    calls = []

    def fake_advance(wm, v, pv, ham, sp, sid, *, target_walkers=0.0, zeta=0.0, xi=0.0, nsteps=1, max_length=0, boost=1.0, projectors=()):
        strat = {_lib.SHIFT_DONT_UPDATE: lambda: R.DontUpdate(target_walkers), _lib.SHIFT_LOG_UPDATE: lambda: R.LogUpdate(zeta),
                 _lib.SHIFT_LOG_UPDATE_AFTER_TARGET: lambda: R.LogUpdateAfterTargetWalkers(target_walkers, zeta),
                 _lib.SHIFT_DOUBLE_LOG_UPDATE: lambda: R.DoubleLogUpdate(target_walkers, zeta, xi),
                 _lib.SHIFT_DOUBLE_LOG_UPDATE_AFTER_TARGET: lambda: R.DoubleLogUpdateAfterTargetWalkers(target_walkers, zeta, xi)}[sid]()
        calls.append(nsteps)
        stats, shifts, dots = [], [], []
        is_int = v.style.val_type == _lib.VAL_I64
        for _ in range(nsteps):
            fciqmc.apply_operator(wm, pv, v, R.FirstOrderTransitionOperator(ham, sp.shift, sp.time_step))
            v, pv = pv, v
            s = wm.last_stats
            tnorm = float(s.inorm1) if is_int else s.norm1
            proceed = True
            if s.len > 0:
                _, proceed = strat.update(sp, tnorm)
            stats.append(s); shifts.append(sp.shift); dots.append([fr.dot(v) for fr in projectors])
            if s.len == 0 or (max_length and s.len > max_length) or not proceed:
                break
        out = (v, pv, stats, shifts, len(stats))
        return out + (np.array(dots).reshape(len(stats), len(projectors)),) if projectors else out

    monkeypatch.setattr(fciqmc, "advance", fake_advance)
    monkeypatch.setattr(FakeDVec, "handle", property(lambda self: 1), raising=False)  # "a device vector" for _batch_size
    monkeypatch.setattr(dictvectors, "FrozenDVec", FakeDVec)  # FakeDVec.freeze() returns a FakeDVec

    def run(ds, **kw):
        del calls[:]
        prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, random_seed=11, device_steps=ds, **kw)
        sim = R.solve(prob)
        return sim, sim.dataframe(), list(calls)

    cases = [dict(style=R.IsDynamicSemistochastic(), time_step=0.002, last_step=230, target_walkers=300, reporting_interval=3),
             dict(style=R.IsStochasticInteger(), time_step=0.005, last_step=150,
                  shift_strategy=R.DoubleLogUpdateAfterTargetWalkers(target_walkers=200)),
             dict(style=R.IsStochasticInteger(), time_step=0.01, last_step=400, target_walkers=5000, max_length=40),          # aborts
             dict(shift=20.0, shift_strategy=R.DontUpdate(target_walkers=150), time_step=0.01, last_step=2000)]                   # stops
    for kw in cases:
        sa, a, ca = run(1, **kw)
        sb, b, cb = run(64, **kw)
        assert ca == [] and cb and max(cb) <= 64
        assert (sa.success, sa.aborted, sa.message, sa.step) == (sb.success, sb.aborted, sb.message, sb.step)
        assert list(a.columns) == list(b.columns) and len(a) == len(b)
        for col in a.columns:
            assert list(a[col]) == list(b[col]), (kw, col)
    assert sum(cb) >= sb.step  # the DontUpdate run ended inside a batch
    # projections inside the batch; a strategy that needs the host every step switches batching off
    ref = R.GPUDVec([(ph.address, 1.0)], style=R.IsDeterministic())
    kw = dict(style=R.IsDynamicSemistochastic(), time_step=0.002, last_step=100, target_walkers=200)
    sa, a, _ = run(1, post_step_strategy=(R.ProjectedEnergy(ph, ref), R.Projector(ones=ref)), **kw)
    sb, b, cb = run(32, post_step_strategy=(R.ProjectedEnergy(ph, ref), R.Projector(ones=ref)), **kw)
    assert cb == [32, 32, 32, 4] and list(a.columns) == list(b.columns) and {"vproj", "hproj", "ones"} <= set(b.columns)
    for col in a.columns:
        assert list(a[col]) == list(b[col]), col
    sc, c, cc = run(32, post_step_strategy=(R.Timer(),), **kw)
    assert cc == [] and "time" in c.columns
    # solve! continues in batches when last_step is raised
    sim, _, _ = run(50, **kw)
    del calls[:]
    R.solve_(sim, last_step=130)
    assert sim.step == 130 and calls == [30] and len(sim.dataframe()) == 130
