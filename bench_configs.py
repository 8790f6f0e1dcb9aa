#!/usr/bin/env python
"""bench_configs.py -- throughput of the FCIQMC step / matrix-free H*v on ALL five BASELINE.json configs
(bench.py is the contract line on config 2; this script is the per-config evidence table kept in profiles/).

  python bench_configs.py --configs 1,2,4,5 --steps 20          one GPU   (6, 7: the reference's own benchmark workloads)
  torchrun --nproc-per-node N bench_configs.py --configs 4,5    hash-partitioned over N GPUs (weak scaling:
                                                                 --walkers is PER GPU)
Every line: model, style, walkers, determinants, ms/step (CUDA events on the library's stream, max over ranks),
spawn attempts/s, phase split, algorithmic HBM bytes/step and the fraction of the measured HBM peak.
The population is grown by the sampler itself from `starting_address => 10` under DoubleLogUpdate.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def fermi(m, modes):
    return tuple(1 if (i + 1) in modes else 0 for i in range(m))


def make_config(R, cid, walkers):
    """-> dict(name, ham, style, dtau, walkers (per GPU), words)"""
    if cid == 1:
        a = R.near_uniform(R.BoseFS, 10, 10)
        return dict(name="config1 HubbardReal1D BoseFS{10,10} u=6 IsStochasticInteger", ham=lambda: R.HubbardReal1D(a, u=6.0, t=1.0),
                    addr=a, style=R.IsStochasticInteger(), dtau=1e-3, walkers=walkers or 1e4, words=1)
    if cid == 2:
        a = R.BoseFS(tuple(20 if i == 9 else 0 for i in range(20)))
        return dict(name="config2 HubbardMom1D BoseFS{20,20} u=6 IsDynamicSemistochastic", ham=lambda: R.HubbardMom1D(a, u=6.0, t=1.0),
                    addr=a, style=R.IsDynamicSemistochastic(), dtau=1e-4, walkers=walkers or 1e7, words=1)
    if cid == 3:
        a = R.FermiFS2C(fermi(16, range(1, 9)), fermi(16, range(5, 13)))
        return dict(name="config3 HubbardRealSpace 4x4 FermiFS2C 8+8 IsDeterministic H*v",
                    ham=lambda: R.HubbardRealSpace(a, geometry=R.PeriodicBoundaries(4, 4), t=(1.0, 1.0), u=((0.0, 1.0), (1.0, 0.0))),
                    addr=a, style=R.IsDeterministic(), dtau=0.0, walkers=walkers or 0, words=1)
    if cid == 4:
        a = R.near_uniform(R.BoseFS, 64, 64)
        return dict(name="config4 HubbardRealSpace 4x4x4 BoseFS{64,64} u=1 IsDynamicSemistochastic (2-word addresses)",
                    ham=lambda: R.HubbardRealSpace(a, geometry=R.PeriodicBoundaries(4, 4, 4), t=1.0, u=1.0),
                    addr=a, style=R.IsDynamicSemistochastic(), dtau=1e-3, walkers=walkers or 1.25e8, words=2)
    if cid == 5:
        a = R.FermiFS2C(fermi(32, (15, 16, 17)), fermi(32, (15, 16, 17)))
        return dict(name="config5 Transcorrelated1D FermiFS2C M=32 3up3down t=1 v=1 cutoff=1 3-body IsDynamicSemistochastic",
                    ham=lambda: R.Transcorrelated1D(a, t=1.0, v=1.0, cutoff=1, three_body_term=True),
                    addr=a, style=R.IsDynamicSemistochastic(), dtau=1e-4, walkers=walkers or 1.25e7, words=1)
    # the reference's own FCIQMC benchmark workloads (benchmark/benchmarks.jl:56-73), sizes as defined there
    if cid == 6:
        a = R.BoseFS(tuple(10 if i == 9 else 0 for i in range(20)))
        return dict(name="ref-bench (10,20) HubbardMom1D BoseFS{10,20} u=1 IsDynamicSemistochastic initiator=true (benchmark/benchmarks.jl:56-64)",
                    ham=lambda: R.HubbardMom1D(a, u=1.0, t=1.0), addr=a, style=R.IsDynamicSemistochastic(), dtau=1e-4,
                    walkers=walkers or 4e4, words=1, initiator=R.Initiator(1.0))
    if cid == 7:
        a = R.near_uniform(R.BoseFS, 50, 50)
        return dict(name="ref-bench (50,50) HubbardReal1D BoseFS{50,50} u=6 IsDynamicSemistochastic (benchmark/benchmarks.jl:66-73)",
                    ham=lambda: R.HubbardReal1D(a, u=6.0, t=1.0), addr=a, style=R.IsDynamicSemistochastic(), dtau=1e-4,
                    walkers=walkers or 5e4, words=2)
    # north_star's headline wording: "2D Hubbard at 1e9 walkers" -- no such stochastic config is listed in BASELINE.json
    # (SURVEY 8g), so it is reported on config 3's model run stochastically: 4x4 Fermi-Hubbard at half filling, 1.25e8
    # walkers per GPU (= 1e9 on 8 GPUs; the sector has 1.66e8 determinants, so every determinant carries several walkers)
    if cid == 8:
        a = R.FermiFS2C(fermi(16, range(1, 9)), fermi(16, range(5, 13)))
        return dict(name="headline 2D Hubbard: HubbardRealSpace 4x4 FermiFS2C 8+8 u=1 IsDynamicSemistochastic",
                    ham=lambda: R.HubbardRealSpace(a, geometry=R.PeriodicBoundaries(4, 4), t=(1.0, 1.0), u=((0.0, 1.0), (1.0, 0.0))),
                    addr=a, style=R.IsDynamicSemistochastic(), dtau=1e-3, walkers=walkers or 1.25e8, words=1)
    raise ValueError(cid)


def run_stochastic(R, cfg, args, world, rank, dist, torch, peak):
    per_gpu = float(cfg["walkers"])
    target = per_gpu * world
    W = cfg["words"]
    ctx = R.init_distributed(W, records_per_peer=int(per_gpu * 1.5) + 4096, table_slots=1 << 22)
    H = cfg["ham"]()
    style = cfg["style"]
    is_int = style.val_type == R._lib.VAL_I64
    v = R.GPUDVec([(cfg["addr"], 10 if is_int else 10.0)], style=style, capacity=int(per_gpu * 1.6) + 4096,
                  initiator=cfg.get("initiator"))
    pv = v.similar()
    R._lib.check(R._lib.lib().rimu_vec_reserve(pv.handle, int(per_gpu * 1.6) + 4096))
    wm = R.working_memory(v, seed=args.seed)
    sp = R.ShiftParameters(R.diagonal_element(H, cfg["addr"]), 10.0, cfg["dtau"])
    strat = R.DoubleLogUpdate(target_walkers=target)

    def one():
        nonlocal v, pv
        R.apply_operator(wm, pv, v, R.FirstOrderTransitionOperator(H, sp.shift, sp.time_step))
        v, pv = pv, v
        s = wm.last_stats
        strat.update(sp, float(s.inorm1) if is_int else s.norm1)
        return s

    from bench import Settled
    t0, nsteps, settled = time.time(), 0, Settled(target, args.equil)
    while True:
        s = one()
        nsteps += 1
        tn = float(s.inorm1) if is_int else s.norm1
        if settled.update(tn, s.len) or nsteps >= args.max_growth or time.time() - t0 > args.growth_seconds:
            break
    for _ in range(3):
        one()
    acc = dict(att=0, dep=0, spawn=0.0, exch=0.0, merge=0.0, total=0.0, P=0, U=0, lb=0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # --device-steps K (one GPU): the timed steps go through rimu_advance in batches of K -- shift update and abort rules on the
    # device, no host round trip between steps (what `solve` does by default for runs without post-step strategies).  The
    # trajectory is the same; only the per-phase event times are not recorded.
    batched = args.device_steps > 1 and world == 1
    e0.record(stream)
    if batched:
        from rimu_b200 import _lib
        left = args.steps
        while left > 0:
            k = min(args.device_steps, left)
            P = len(v)
            v, pv, stats, shifts, done = R.advance(wm, v, pv, H, sp, _lib.SHIFT_DOUBLE_LOG_UPDATE, target_walkers=target,
                                                   zeta=strat.zeta, xi=strat.xi, nsteps=k)
            assert done == k, "the run ended inside the timed region"
            for s in stats:
                acc["att"] += s.spawn_attempts; acc["dep"] += s.deposits; acc["P"] += P; acc["U"] += s.local_len; acc["lb"] += s.len_before
                P = s.local_len
            left -= k
    else:
        for _ in range(args.steps):
            P = len(v)
            s = one()
            acc["att"] += s.spawn_attempts; acc["dep"] += s.deposits; acc["P"] += P; acc["U"] += s.local_len; acc["lb"] += s.len_before
            acc["spawn"] += s.ms_spawn; acc["exch"] += s.ms_exchange; acc["merge"] += s.ms_compact
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / args.steps
    K = args.steps
    E = 8 * W + 8
    P, U = acc["P"] / K, acc["U"] / K
    A1 = max(acc["dep"] / world - acc["P"], 0) / K  # non-zero spawn records per step and rank (deposits = parents' diagonal + records)
    step_bytes = (P * E + A1 * E) + (P * (E + 8) + A1 * E + U * (E + 8))
    tn = float(s.inorm1) if is_int else s.norm1
    return {"config": cfg["name"], "n_gpus": world, "walkers_per_gpu": per_gpu, "norm": tn, "determinants_per_gpu": P,
            "attempts_per_step": acc["att"] / K, "nonzero_spawns_per_step_per_gpu": A1, "ms_per_step": ms,
            "spawn_attempts_per_s": acc["att"] / K / (ms * 1e-3),
            "phase_ms": {"spawn": acc["spawn"] / K, "exchange": acc["exch"] / K, "merge": acc["merge"] / K},
            "algorithmic_bytes_per_step_per_gpu": step_bytes, "hbm_gbs": step_bytes / (ms * 1e-3) / 1e9,
            "hbm_frac_of_measured_peak": step_bytes / (ms * 1e-3) / 1e9 / peak, "growth_steps": nsteps, "steps": K,
            "shift": sp.shift, "dtau": cfg["dtau"], "words": W, "buckets_per_gpu": int(s.buckets), "max_bucket_fill": int(s.max_bucket_fill),
            "sent_records_per_step_per_gpu": int(s.sent_records),
            "driver": f"rimu_advance, batches of {min(args.device_steps, K)} steps" if batched else "rimu_step per step + host shift update"}


def run_deterministic(R, cfg, args, world, rank, dist, torch, peak):
    """config 3: K matrix-free H*v applications (one `mul!` = one Lanczos matvec) over the complete sector in the
    dense-indexed layout (csrc/sector.cuh: gather, no records).  `--dictionary-hv` times the dictionary path instead
    (grow x <- H x / |H x| from the starting determinant until the sector is filled or --max-dim)."""
    if not args.dictionary_hv and world == 1:
        H = cfg["ham"]()
        basis = R.SectorBasis(H)
        x = basis.vector([(cfg["addr"], 1.0)])
        y = basis.zeros()
        for _ in range(8):  # fill the sector
            R.mul(y, H, x)
            y.scale_(1.0 / y.norm(2))
            x, y = y, x
        times = []
        for _ in range(args.steps):
            R.mul(y, H, x)
            times.append(y.last_mul_ms)
            x, y = y, x
        ms = float(np.median(times))
        L = 64  # off-diagonals per address (incl. Pauli-blocked ones): HubbardRealSpace.jl:316-321
        # algorithmic HBM bytes of the gather: keys + x read once, y written once (neighbour reads hit L2: consecutive
        # ranks of the last component share their neighbours' cache lines)
        bytes_ = basis.dim * (8 + 8 + 8)
        return {"config": cfg["name"] + " (dense-indexed sector)", "n_gpus": 1, "dimension": basis.dim, "ms_per_matvec": ms,
                "attempts_per_matvec": basis.dim * L, "spawn_attempts_per_s": basis.dim * L / (ms * 1e-3),
                "algorithmic_bytes_per_matvec": bytes_, "hbm_gbs": bytes_ / (ms * 1e-3) / 1e9,
                "hbm_frac_of_measured_peak": bytes_ / (ms * 1e-3) / 1e9 / peak, "steps": args.steps, "words": 1}
    W = cfg["words"]
    ctx = R.init_distributed(W, records_per_peer=1 << 22, table_slots=1 << 22)
    H = cfg["ham"]()
    x = R.GPUDVec([(cfg["addr"], 1.0)], style=cfg["style"])
    wm = R.working_memory(x)
    y = x.similar()
    last, it = 0, 0
    while True:
        R.mul(y, H, x, wm)
        y.scale_(1.0 / y.norm(2))
        x, y = y, x
        it += 1
        n = wm.last_stats.len
        if n == last or n >= args.max_dim or it > 200:
            break
        last = n
    acc_att, tot_ms, sp_ms, mg_ms = 0, 0.0, 0.0, 0.0
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        R.mul(y, H, x, wm)
        s = wm.last_stats
        acc_att += s.spawn_attempts; sp_ms += s.ms_spawn + s.ms_diag; mg_ms += s.ms_compact
        x, y = y, x
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / args.steps
    return {"config": cfg["name"], "n_gpus": world, "dimension_reached": int(wm.last_stats.len), "growth_matvecs": it,
            "attempts_per_matvec": acc_att / args.steps, "ms_per_matvec": ms, "spawn_attempts_per_s": acc_att / args.steps / (ms * 1e-3),
            "phase_ms": {"spawn": sp_ms / args.steps, "merge": mg_ms / args.steps}, "steps": args.steps, "words": W}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,4,5")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--walkers", type=float, default=0, help="per GPU; 0 = the config's default")
    ap.add_argument("--equil", type=int, default=30)
    ap.add_argument("--max-growth", type=int, default=4000)
    ap.add_argument("--growth-seconds", type=float, default=300)
    ap.add_argument("--max-dim", type=float, default=2e8)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--device-steps", type=int, default=1, help="time the steps in batches of K through rimu_advance (one GPU)")
    ap.add_argument("--dictionary-hv", action="store_true", help="config 3 through the dictionary path (records + annihilation)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import rimu_b200 as R
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("gloo")
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    for cid in [int(c) for c in args.configs.split(",")]:
        cfg = make_config(R, cid, args.walkers)
        fn = run_deterministic if cid == 3 else run_stochastic
        line = fn(R, cfg, args, world, rank, dist, torch, peak)
        if rank == 0:
            print(json.dumps(line), flush=True)
        import gc
        gc.collect()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
