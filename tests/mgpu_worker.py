"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun).  Runs FCIQMC steps with the NCCL spawn
exchange and checks the union of the ranks' vectors against the single-rank CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def energy_mode(R, rank, world):
    """BASELINE config 5's correctness check at the world size this worker runs with: FCIQMC on Transcorrelated1D (M=12,
    3 up 2 down, three-body term, non-Hermitian) through the full driver -- ProjectorMonteCarloProblem, DoubleLogUpdate,
    ProjectedEnergy over the AdjointUnknown sweep -- with the vector hash-partitioned over the ranks.  Shift and projected
    energy must agree with the oracle's exact diagonalisation within blocking-analysis error bars
    (test/lomc.jl:518-541, test/mpi_runtests.jl:140-155)."""
    import json
    from tests.cases import oracle_ham, product_ham
    oh = oracle_ham("tc_12")
    e_exact = oh.exact_energy(hermitian=False)
    R.reset_contexts()
    R.init_distributed(oh.W, records_per_peer=1 << 14)
    ph = product_ham("tc_12")
    ref = R.GPUDVec([(ph.address, 1.0)], style=R.IsDeterministic())
    prob = R.ProjectorMonteCarloProblem(ph, start_at=ph.address, style=R.IsDynamicSemistochastic(), time_step=0.002, last_step=5000,
                                        target_walkers=20_000, random_seed=11, max_length=10 ** 6,
                                        post_step_strategy=(R.ProjectedEnergy(ph, ref),))
    sim = R.solve(prob)
    assert sim.success, sim.message
    df = sim.dataframe()
    se = R.shift_estimator(df, skip=1500)
    pe = R.projected_energy(df, skip=1500)
    tol_bias = 0.01 * abs(e_exact)
    ok = abs(se.mean - e_exact) < 5 * se.err + tol_bias and abs(pe.f - e_exact) < 5 * pe.sigma_f + tol_bias
    if rank == 0:
        print("mgpu energy " + json.dumps({"model": "Transcorrelated1D M=12 3up2down 3-body", "world": world, "exact": e_exact,
                                           "shift": se.mean, "shift_err": se.err, "projected": pe.f, "projected_err": pe.sigma_f,
                                           "walkers": 20000, "steps": 5000, "within_5_sigma_plus_1pct": bool(ok)}), flush=True)
    assert ok, (se.mean, se.err, pe.f, pe.sigma_f, e_exact)


def main():
    import torch
    import torch.distributed as dist
    import rimu_b200 as R
    from oracle import oracle as orc
    from tests.cases import oracle_ham, product_ham

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    if os.environ.get("RIMU_MGPU_MODE") == "energy":
        energy_mode(R, rank, world)
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    method = os.environ.get("RIMU_B200_METHOD", "partition")
    direct = os.environ.get("RIMU_B200_P2P", "1") != "0"
    cases = [("real1d_10", "int", 0), ("rs_bose_3d_w2", "int", 0), ("mom1d_bose", "semi", 0), ("rs_f2c_4x4", "int", 0)]
    if method == "partition" and direct:  # initiator lanes travel in the sub-stream index of the direct exchange
        cases += [("real1d_10", "int", 1), ("mom1d_bose", "semi", 3)]
    for name, style_name, irule in cases:
        oh = oracle_ham(name)
        R.reset_contexts()
        ctx = R.init_distributed(oh.W, records_per_peer=1 << 12)  # small on purpose: exercises the grow-and-repeat path
        ph = product_ham(name)
        seed, dtau = 99, 0.001 if name == "rs_bose_3d_w2" else 0.01
        if style_name == "int":
            style, pop, dtype, ostyle, kw = R.IsStochasticInteger(), 20000, np.int64, orc.STYLE_INTEGER, {}
        else:
            style, pop, dtype, ostyle, kw = R.IsDynamicSemistochastic(), 5000.5, np.float64, orc.STYLE_SEMISTOCHASTIC, dict(compress_threshold=1.0)
        rule = {0: None, 1: R.Initiator(1.0), 2: R.SimpleInitiator(1.0), 3: R.CoherentInitiator(1.0)}[irule]
        v = R.GPUDVec([(ph.address, pop)], style=style, initiator=rule)  # only the owner rank keeps the entry
        wm = R.working_memory(v, seed=seed)
        ok, ov = np.array([oh.start_key], dtype=np.uint64), np.array([pop], dtype=dtype)
        shift = oh.diagonal_element(oh.start_key)
        for step in range(5):
            T = R.FirstOrderTransitionOperator(ph, shift, dtau)
            out = v.similar()
            R.apply_operator(wm, out, v, T)
            v = out
            s = wm.last_stats
            pp = orc.make_params(ostyle, shift=shift, dtau=dtau, key=orc.step_key(seed, step), initiator_rule=irule,
                                 initiator_threshold=1.0, **kw)
            ok, ov, st = oh.step(pp, ok, ov)
            lk, lv = v.download()
            parts = [None] * world
            dist.all_gather_object(parts, (lk, lv))
            gk = np.concatenate([p[0] for p in parts]).reshape(-1, oh.W)
            gv = np.concatenate([p[1] for p in parts])
            order = np.lexsort(tuple(gk[:, j] for j in range(oh.W)))
            gk, gv = gk[order], gv[order]
            assert np.array_equal(gk, ok), (name, step, "keys", len(gk), len(ok))
            # every rank holds exactly the keys it owns
            for k in lk[:50]:
                assert orc.addr_owner(k, world) == rank
            assert s.spawn_attempts == st.spawn_attempts and s.len == st.len_after, (name, step)
            if style_name == "int":
                assert np.array_equal(gv, ov), (name, step, "values")
                assert (s.ispawns, s.ideaths, s.iclones, s.izombies, s.inorm1) == (st.ispawns, st.ideaths, st.iclones, st.izombies, st.inorm1)
            else:
                assert np.allclose(gv, ov, rtol=1e-10, atol=0), (name, step)
                assert abs(s.norm1 - st.norm1) <= 1e-9 * st.norm1
                # feed the (gathered) GPU values back so that last-bit differences cannot flip branches
                ok, ov = gk, gv
            assert all(len(p[1]) > 0 for p in parts) or step < 4, "every rank should own part of the vector"
            # walkernumber_and_length is global
            wn, _ = R.walkernumber_and_length(v)
            assert abs(wn - float(np.abs(ov).sum())) <= 1e-9 * float(np.abs(ov).sum())
        if rank == 0:
            import ctypes as C
            p2p = C.c_int()
            R._lib.check(R._lib.lib().rimu_comm_p2p(ctx.handle, C.byref(p2p)))
            print(f"mgpu ok: {name} {style_name} initiator={irule} method={method} world={world} len={len(ov)} sent={s.sent_records} p2p={p2p.value}", flush=True)
    dist.barrier()
    # leave without running destructors in arbitrary order against the peers' teardown (NCCL communicators, IPC
    # mappings): every assertion has been evaluated by now
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
