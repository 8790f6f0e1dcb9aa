#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native FCIQMC step (contract: see task spec / DESIGN.md).

Metric (BASELINE.json): walker spawn attempts per second (+ annihilation HBM GB/s in `extra`).
Workload at every N: BASELINE config 2 -- HubbardMom1D, BoseFS{20,20}, IsDynamicSemistochastic,
1e7 target walkers PER GPU (weak scaling; determinant space hash-partitioned across ranks).
A "step" is one FCIQMC step (apply_operator! + shift update) on an equilibrated population that the
sampler itself produced from the starting address (SURVEY.md section 8d "steady-state surrogate").

  python bench.py --gpus 1 --steps 20 --warmup 3
  torchrun --nproc-per-node N bench.py --gpus N ...        (driver launches this form)
  python bench.py --impl reference ...                      CPU arm: oracle port on all host cores
"""
import argparse
import ctypes as C
import json
import math
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "spawn_attempts_per_s"
UNIT = "attempts/s"
M_SITES, N_PART, U_INT, T_HOP, DTAU = 20, 20, 6.0, 1.0, 1e-4
START_ONR = tuple(N_PART if i == 9 else 0 for i in range(M_SITES))  # all bosons in mode 10 (k = 0)
WORKLOAD = "config2: HubbardMom1D BoseFS{20,20} u=6 t=1 dtau=1e-4 IsDynamicSemistochastic"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers", type=float, default=1e7, help="target walkers per GPU")
    ap.add_argument("--equil", type=int, default=40, help="extra equilibration steps once within 5%% of the target")
    ap.add_argument("--cpu-walkers", type=float, default=0, help="reference arm: walkers of the CPU run (0 = --walkers, the GPU arm's per-GPU workload)")
    ap.add_argument("--long-steps", type=int, default=200, help="extra resident steps timed after the K contract steps (robust ms/step; 0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--e2e-replicas", type=int, default=3, help="independent replicas in flight in the host-buffer (e2e) measurement")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- steady-state detector
class Settled:
    """The population counts as equilibrated when, over the last `window` steps, the walker number stayed within 5 % of
    the target AND drifted by less than 1.5 % of it AND the number of stored determinants changed by less than 3 %.
    (Staying inside the band alone is not enough: the first passage through it, still growing, can last tens of steps.)"""

    def __init__(self, target, window):
        self.target, self.window, self.norms, self.lens = float(target), int(window), [], []

    def update(self, norm, length):
        self.norms.append(float(norm)); self.lens.append(float(length))
        if len(self.norms) < self.window:
            return False
        n, l = self.norms[-self.window:], self.lens[-self.window:]
        t = self.target
        return (all(abs(x - t) < 0.05 * t for x in n) and abs(n[-1] - n[0]) < 0.015 * t
                and abs(l[-1] - l[0]) < 0.03 * max(l[-1], 1.0))


# --------------------------------------------------------------------------- reference arm (CPU)
def reference_arm(args):
    """The reference (pure Julia) cannot run here: the CPU arm is the oracle port of Rimu's threaded
    PDVec path (oracle.c orc_step_threaded), all host cores, on a bounded sample of the same workload
    (same model/style, fewer walkers)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    oh = orc.OracleHam("HubbardMom1D", "bose", START_ONR, u=U_INT, t=T_HOP)
    # the SAME workload as the GPU arm: 1e7 target walkers (one GPU's share when the job has several), grown on the CPU by
    # the same DoubleLogUpdate schedule
    target = args.cpu_walkers or args.walkers
    keys = np.array([oh.start_key], dtype=np.uint64)
    vals = np.array([10.0])
    shift0 = oh.diagonal_element(oh.start_key)
    shift, pnorm, zeta = shift0, 10.0, 0.08
    xi = zeta ** 2 / 4
    step = 0

    def one(step, shift):
        p = orc.make_params(orc.STYLE_SEMISTOCHASTIC, shift=shift, dtau=DTAU, compress_threshold=1.0,
                            key=orc.step_key(args.seed, step))
        return oh.step(p, keys, vals, threads=cores)

    # DoubleLogUpdate from the first step, as ProjectorMonteCarloProblem does by default
    t_budget, settled = time.time(), Settled(target, min(args.equil, 12))  # (a CPU step at 1e7 walkers takes ~1 s: equilibrate 12 steps, not 40)
    while True:
        keys, vals, st = one(step, shift)
        step += 1
        tnorm = st.norm1
        shift -= xi / DTAU * math.log(tnorm / target) + zeta / DTAU * math.log(tnorm / pnorm)
        pnorm = tnorm
        if settled.update(tnorm, st.len_after) or step >= 3000 or time.time() - t_budget > 900:
            break
    grow_s = time.time() - t_budget
    for _ in range(args.warmup):
        keys, vals, st = one(step, shift)
        step += 1
    attempts, t0 = 0, time.time()
    for _ in range(args.steps):
        keys, vals, st = one(step, shift)
        tnorm = st.norm1
        shift -= xi / DTAU * math.log(tnorm / target) + zeta / DTAU * math.log(tnorm / pnorm)
        pnorm = tnorm
        attempts += st.spawn_attempts
        step += 1
    dt = time.time() - t0
    val = attempts / dt
    world = max(1, args.gpus)
    sample = (f"{WORKLOAD}, {target:.0e} target walkers grown on the CPU in {step - args.steps - args.warmup} steps ({grow_s:.0f} s), "
              f"{len(vals)} determinants; every timed step is one full FCIQMC step over that vector"
              + (f" (= ONE GPU's share of the {world}-GPU job: bounded sample)" if world > 1 else ""))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "walkers_per_gpu": target, "target_walkers": target * world,
                   "parallelism": f"hash-partitioned x{world}",
                   "note": "CPU port of Rimu's threaded PDVec path (the reference is Julia and cannot run here; probe: julia "
                           + ("present but Rimu.jl is not installed offline" if shutil.which("julia") else "absent") + ")"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------- parity preflight (multi-rank arm)
def parity_preflight(R, rank, world):
    """Five IsStochasticInteger steps of HubbardReal1D BoseFS{10,10} (BASELINE config 1) on a fresh context spanning all
    ranks, the union of the ranks' vectors compared bit for bit -- keys, values and every integer statistic -- with the
    single-rank CPU oracle.  The oracle is used as the checker only (it is never timed here).  Every SCALE record thereby
    carries its own correctness signal for the exchange path it measures."""
    import torch.distributed as dist
    from oracle import oracle as orc
    onr = (1,) * 10
    oh = orc.OracleHam("HubbardReal1D", "bose", onr, u=6.0, t=1.0)
    addr = R.BoseFS(onr)
    H = R.HubbardReal1D(addr, u=6.0, t=1.0)
    ctx = R.init_distributed(1, records_per_peer=1 << 12, fresh=True)
    seed, dtau, pop = 4242, 0.01, 30000
    v = R.GPUDVec([(addr, pop)], style=R.IsStochasticInteger(), ctx=ctx)
    wm = R.working_memory(v, seed=seed)
    ok, ov = np.array([oh.start_key], dtype=np.uint64), np.array([pop], dtype=np.int64)
    shift = oh.diagonal_element(oh.start_key)
    verdict = "ok"
    for step in range(5):
        out = v.similar()
        R.apply_operator(wm, out, v, R.FirstOrderTransitionOperator(H, shift, dtau))
        v = out
        s = wm.last_stats
        ok, ov, st = oh.step(orc.make_params(orc.STYLE_INTEGER, shift=shift, dtau=dtau, key=orc.step_key(seed, step)), ok, ov)
        lk, lv = v.download()
        parts = [(lk, lv)]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, (lk, lv))
        gk = np.concatenate([p[0] for p in parts]).reshape(-1)
        gv = np.concatenate([p[1] for p in parts])
        order = np.argsort(gk, kind="stable")
        same = (np.array_equal(gk[order], ok.reshape(-1)) and np.array_equal(gv[order], ov)
                and (s.spawn_attempts, s.len, s.ispawns, s.ideaths, s.iclones, s.izombies, s.inorm1)
                == (st.spawn_attempts, st.len_after, st.ispawns, st.ideaths, st.iclones, st.izombies, st.inorm1))
        if not same:
            verdict = f"MISMATCH at step {step}"
            break
    del wm, v, out
    import gc
    gc.collect()
    ctx.close(collective=True)  # every rank, same point: peer mappings are closed before their buffers are freed
    return verdict


# --------------------------------------------------------------------------- our arm (GPU)
def pinned_array(R, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    R._lib.check(R._lib.lib().rimu_host_alloc(max(n, 8), C.byref(p)))
    buf = (C.c_char * max(n, 8)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape), p


def ours(args):
    import torch
    import torch.distributed as dist
    import rimu_b200 as R
    from rimu_b200 import _lib

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    per_gpu = args.walkers
    target = per_gpu * world
    slots = 1 << max(16, int(math.ceil(math.log2(per_gpu * 3))))
    ctx = R.init_distributed(1, records_per_peer=int(per_gpu * 1.5), table_slots=slots)

    if os.environ.get("RIMU_B200_LIB") and os.environ.get("RIMU_BENCH_SKIP_PREFLIGHT"):
        preflight = "skipped (kernel-tuning build with one model compiled in; never a contract run)"
    else:
        preflight = parity_preflight(R, rank, world)  # before anything is timed
        if preflight != "ok":
            raise SystemExit(f"parity preflight failed on rank {rank}: {preflight}")

    addr = R.BoseFS(START_ONR)
    H = R.HubbardMom1D(addr, u=U_INT, t=T_HOP)
    style = R.IsDynamicSemistochastic()
    v = R.GPUDVec([(addr, 10.0)], style=style, capacity=int(per_gpu * 1.5))
    pv = v.similar()
    R._lib.check(R._lib.lib().rimu_vec_reserve(pv.handle, int(per_gpu * 1.5)))
    wm = R.working_memory(v, seed=args.seed)
    shift0 = R.diagonal_element(H, addr)
    sp = R.ShiftParameters(shift0, 10.0, DTAU)
    strat = R.DoubleLogUpdate(target_walkers=target)

    def one_step():
        nonlocal v, pv
        T = R.FirstOrderTransitionOperator(H, sp.shift, sp.time_step)
        R.apply_operator(wm, pv, v, T)
        v, pv = pv, v
        s = wm.last_stats
        strat.update(sp, s.norm1)
        return s

    # DoubleLogUpdate from the first step (the reference's default); stop once the population has
    # stayed within 5 % of the target for `equil` consecutive steps
    t0, nsteps, settled = time.time(), 0, Settled(target, args.equil)
    while True:
        s = one_step()
        nsteps += 1
        if settled.update(s.norm1, s.len) or nsteps >= 3000 or time.time() - t0 > 240:
            break
    for _ in range(args.warmup):
        s = one_step()

    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- timed region: K resident steps
    launches0 = C.c_uint64()
    _lib.check(_lib.lib().rimu_ctx_launch_count(ctx.handle, C.byref(launches0)))
    clocks = ClockSampler(local if world == 1 else ",".join(str(i) for i in range(world)))  # ONE sampler (rank 0) for all GPUs of the job
    barrier()
    if rank == 0:  # (eight nvidia-smi pollers at 20 ms each contend for the driver's locks with the ranks' own launches)
        clocks.start()
    torch.cuda.profiler.start()  # cudaProfilerStart: `ncu --profile-from-start off` captures the timed steps only
    ev0.record(stream)
    tw0 = time.time()
    acc = dict(attempts=0, deposits=0, ms_diag=0.0, ms_spawn=0.0, ms_exch=0.0, ms_compact=0.0, parents=0, len_before=0, len=0)
    for _ in range(args.steps):
        parents = len(v)
        s = one_step()
        acc["attempts"] += s.spawn_attempts; acc["deposits"] += s.deposits
        acc["ms_diag"] += s.ms_diag; acc["ms_spawn"] += s.ms_spawn; acc["ms_exch"] += s.ms_exchange; acc["ms_compact"] += s.ms_compact
        acc["parents"] += parents; acc["len_before"] += s.len_before; acc["len"] += s.len
    ev1.record(stream)
    barrier()
    torch.cuda.profiler.stop()
    wall = time.time() - tw0
    ms = ev0.elapsed_time(ev1)
    launches1 = C.c_uint64()
    _lib.check(_lib.lib().rimu_ctx_launch_count(ctx.handle, C.byref(launches1)))
    # the K contract steps last ~20 ms -- one nvidia-smi sample.  A longer run of resident steps (same loop, same events)
    # gives the clock sampler something to see and a ms/step that does not hinge on 20 steps
    long_ms = None
    if args.long_steps > 0:
        evl0, evl1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        evl0.record(stream)
        for _ in range(args.long_steps):
            one_step()
        evl1.record(stream)
        barrier()
        tl = torch.tensor([evl0.elapsed_time(evl1)], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        long_ms = float(tl[0]) / args.long_steps
    t = torch.tensor([ms], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t[0])
    value = acc["attempts"] / (ms_max * 1e-3)  # attempts are already global (all-reduced by the library)

    # ---- e2e: the same step through the C ABI with HOST buffers: every step uploads its source vector from pinned
    # host memory (rimu_vec_assign), runs rimu_step and downloads the result (rimu_vec_download), all inside the
    # timed region.  `--e2e-replicas` independent replicas (Rimu's n_replicas: independent vectors, own context =
    # own stream + working memory) are in flight on host threads so that one replica's PCIe copies overlap another's
    # kernels; steps are issued in strict round-robin order so that multi-rank collectives are ordered identically
    # on every rank.  replicas=1 is the plain serial call sequence.
    n_in = len(v)
    cap_h = int(per_gpu * 1.5)
    hk, hkp = pinned_array(R, (cap_h, 1), np.uint64)
    hv, hvp = pinned_array(R, (cap_h,), np.float64)
    m = C.c_int64()
    _lib.check(_lib.lib().rimu_vec_download(v.handle, hk.ctypes.data_as(_lib._u64p), hv.ctypes.data_as(C.c_void_p), hk.shape[0], C.byref(m)))
    e2e_steps = max(3, min(args.steps, 10))
    nrep = max(1, args.e2e_replicas)
    T = R.FirstOrderTransitionOperator(H, sp.shift, sp.time_step)
    reps, pinned = [], [hkp, hvp]
    for r in range(nrep):
        rctx = ctx if r == 0 else R.init_distributed(1, records_per_peer=int(per_gpu * 1.5), table_slots=slots, fresh=True)
        rsrc = R.GPUDVec(style=style, address_type=v.address_type, capacity=cap_h, ctx=rctx)
        rdst = pv if r == 0 else R.GPUDVec(style=style, address_type=v.address_type, capacity=cap_h, ctx=rctx)
        ok, okp = pinned_array(R, (cap_h, 1), np.uint64)
        ov, ovp = pinned_array(R, (cap_h,), np.float64)
        pinned += [okp, ovp]
        reps.append(dict(ctx=rctx, src=rsrc, dst=rdst, wm=R.working_memory(rsrc, seed=args.seed + 1000 * (r + 1)), ok=ok, ov=ov,
                         attempts=0, h2d=0, d2h=0, err=None))
    turn = {"n": 0}
    cv = threading.Condition()

    def replica_loop(r, nsteps, count):
        rep = reps[r]
        mm = C.c_int64()
        try:
            for it in range(nsteps):
                _lib.check(_lib.lib().rimu_vec_assign(rep["src"].handle, hk.ctypes.data_as(_lib._u64p), hv.ctypes.data_as(C.c_void_p), n_in))
                with cv:  # strict round-robin issue order of the steps (identical on every rank)
                    cv.wait_for(lambda: turn["n"] % nrep == r or turn.get("abort"))
                if turn.get("abort"):
                    return
                try:
                    R.apply_operator(rep["wm"], rep["dst"], rep["src"], T)
                finally:
                    with cv:
                        turn["n"] += 1
                        cv.notify_all()
                _lib.check(_lib.lib().rimu_vec_download(rep["dst"].handle, rep["ok"].ctypes.data_as(_lib._u64p),
                                                        rep["ov"].ctypes.data_as(C.c_void_p), cap_h, C.byref(mm)))
                if count:
                    rep["attempts"] += rep["wm"].last_stats.spawn_attempts
                    rep["h2d"] += n_in * 16
                    rep["d2h"] += mm.value * 16
        except Exception as e:  # surfaced after join
            rep["err"] = e
            with cv:
                turn["abort"] = True
                cv.notify_all()

    def run_all(nsteps, count):
        turn["n"] = 0
        ths = [threading.Thread(target=replica_loop, args=(r, nsteps, count)) for r in range(nrep)]
        for t_ in ths:
            t_.start()
        for t_ in ths:
            t_.join()
        for rep in reps:
            if rep["err"] is not None:
                raise rep["err"]

    def measure_e2e(n_active):
        """host-buffer steps with the first n_active replicas in flight; returns the e2e record"""
        nonlocal nrep
        all_reps, saved = reps[:], nrep
        del reps[n_active:]
        nrep = n_active
        for rep_ in reps:
            rep_["attempts"] = rep_["h2d"] = rep_["d2h"] = 0
        try:
            run_all(2, False)  # warm-up: working memory of every replica context sized, pinned pages touched
            barrier()
            ev0.record(stream)
            tw = time.time()
            run_all(e2e_steps, True)  # every thread returns only after its last download has completed (stream-synchronised)
            ev1.record(stream)
            barrier()
            wall_ = time.time() - tw
            tt = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            att = sum(rep_["attempts"] for rep_ in reps)
            h2d_, d2h_ = sum(rep_["h2d"] for rep_ in reps), sum(rep_["d2h"] for rep_ in reps)
            total = e2e_steps * n_active
            return {"value": att / (float(tt[0]) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_ // total,
                    "d2h_bytes_per_step": d2h_ // total, "ms_per_step": float(tt[0]) / total, "steps": total,
                    "replicas_in_flight": n_active, "wall_ms_per_step": 1e3 * wall_ / total}
        finally:
            reps[:] = all_reps
            nrep = saved

    def copy_ceiling():
        """what the box gives this job for host<->device copies alone: every rank moves the same two buffers (source vector
        in, result out) on two streams at the same time, nothing else.  e2e cannot be faster than this; on an 8-GPU box whose
        GPUs hang off one host memory system the AGGREGATE saturates (~100-130 GB/s here, scratch/pcie_ceiling.py), which
        is why the host-buffer number stops scaling while the resident one does not."""
        nel = max(n_in * 2, 1)  # keys + values, 8 bytes each
        h_a, h_b = torch.empty(nel, dtype=torch.float64).pin_memory(), torch.empty(nel, dtype=torch.float64).pin_memory()
        d_a, d_b = torch.empty(nel, dtype=torch.float64, device="cuda"), torch.zeros(nel, dtype=torch.float64, device="cuda")
        s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()

        def go(reps):
            barrier()
            t0_ = time.time()
            for _ in range(reps):
                with torch.cuda.stream(s_a):
                    d_a.copy_(h_a, non_blocking=True)
                with torch.cuda.stream(s_b):
                    h_b.copy_(d_b, non_blocking=True)
            barrier()
            return time.time() - t0_

        go(2)
        reps = 8
        dt_ = go(reps)
        tt = torch.tensor([dt_], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return 2 * nel * 8 * reps * world / float(tt[0]) / 1e9  # GB/s, all ranks, both directions

    e2e_pipelined = measure_e2e(nrep)            # headline: independent replicas keep both PCIe directions busy
    e2e_serial = measure_e2e(1) if nrep > 1 else e2e_pipelined  # the plain call sequence: upload, step, download
    clk = clocks.stop() if rank == 0 else None  # sampled every 20 ms across ALL timed regions (resident steps and host-buffer steps)
    ceiling = copy_ceiling()
    moved = (e2e_pipelined["h2d_bytes_per_step"] + e2e_pipelined["d2h_bytes_per_step"]) * world / (e2e_pipelined["ms_per_step"] * 1e-3) / 1e9
    e2e_pipelined["copy_gbs_all_gpus"] = moved
    e2e_pipelined["platform_copy_ceiling_gbs_all_gpus"] = ceiling
    e2e_pipelined["frac_of_copy_ceiling"] = moved / ceiling if ceiling > 0 else None

    # ---- roofline of the dominant kernel (CUDA-event durations measured live inside rimu_step, per launch averages)
    K = args.steps
    E = 16  # bytes per entry: one uint64 address word + one 8-byte value
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # per-rank figures (stats are global -> divide by world for a per-GPU kernel).  Algorithmic bytes:
    #   spawn_part_kernel: P*E parents read + A'*E records written
    #   merge_kernel     : P*(E+8) parents + cached H_aa read, A'*E records read, U*(E+8) survivors + H_aa written
    P = acc["parents"] / K
    A1 = max(acc["deposits"] / world - acc["parents"], 0) / K  # non-zero spawn records per step and rank
    U = acc["len"] / world / K
    spawn_bytes = P * E + A1 * E
    merge_bytes = P * (E + 8) + A1 * E + U * (E + 8)
    kern = {"spawn_part_kernel": (acc["ms_spawn"] / K, spawn_bytes), "merge_kernel": (acc["ms_compact"] / K, merge_bytes)}
    dom = max(kern, key=lambda k: kern[k][0])
    dms, dbytes = kern[dom]
    achieved = dbytes / (dms * 1e-3) / 1e9 if dms > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("walkers_per_gpu") == per_gpu:
            traffic = tr.get(dom)
    except Exception:
        pass
    annih_bytes = merge_bytes
    annih_ms = acc["ms_compact"] / K
    step_bytes = spawn_bytes + merge_bytes

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "walkers_per_gpu": per_gpu, "target_walkers": target, "determinants_per_gpu": P,
                   "attempts_per_step": acc["attempts"] / K, "method": "partition (bucket streams + shared-memory annihilation)",
                   "l2": "inputs larger than L2: walker vector + spawn record streams of a step exceed 126 MB",
                   "growth_steps": nsteps, "equil_steps": args.equil, "parallelism": f"hash-partitioned x{world}"},
        "e2e": e2e_pipelined,
        "e2e_serial": e2e_serial,
        "gpu_launches": int(launches1.value - launches0.value),
        "parity_preflight": preflight,
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dbytes, "ms_per_launch": dms},
        "extra": {"annihilation_gbs": annih_bytes / (annih_ms * 1e-3) / 1e9 if annih_ms > 0 else None,
                  "step_hbm_gbs": step_bytes / (ms_max / K * 1e-3) / 1e9, "step_hbm_frac_of_peak": step_bytes / (ms_max / K * 1e-3) / 1e9 / peak,
                  "phase_ms_per_step": {"spawn": acc["ms_spawn"] / K, "exchange": acc["ms_exch"] / K, "merge": acc["ms_compact"] / K},
                  "wall_ms_per_step": 1e3 * wall / K, "norm": s.norm1, "shift": sp.shift,
                  "long_run": {"steps": args.long_steps, "ms_per_step": long_ms},
                  "buckets_per_gpu": int(s.buckets), "mean_bucket_fill": (P + A1) / max(int(s.buckets), 1),
                  "max_bucket_fill": int(s.max_bucket_fill)},
    }

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample, rank 0, N=1 only
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        cores = os.cpu_count() or 1
        oh = orc.OracleHam("HubbardMom1D", "bose", START_ONR, u=U_INT, t=T_HOP)
        p = orc.make_params(orc.STYLE_SEMISTOCHASTIC, shift=sp.shift, dtau=DTAU, compress_threshold=1.0, key=orc.step_key(args.seed, 0))
        # bounded sample: whole oracle steps on the equilibrated GPU vector, repeated until >= 10 s of CPU work
        nsample = n_in
        t0, att, nst = time.time(), 0, 0
        while True:
            pk = orc.make_params(orc.STYLE_SEMISTOCHASTIC, shift=sp.shift, dtau=DTAU, compress_threshold=1.0, key=orc.step_key(args.seed, nst))
            _, _, st = oh.step(pk, hk[:nsample], hv[:nsample], threads=cores)
            att += st.spawn_attempts
            nst += 1
            dt = time.time() - t0
            if dt >= 10.0 or nst >= 50:
                break
        line["cpu_baseline"] = {"value": att / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{nst} oracle steps over all {n_in} determinants of the equilibrated GPU vector, {dt:.1f} s"}
    if rank == 0:
        print(json.dumps(line))
    for p in pinned:
        _lib.lib().rimu_host_free(p)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)
