import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rimu_b200 as R
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
per_gpu = 1e7
ctx = R.init_distributed(1, records_per_peer=int(per_gpu * 1.5) + 4096, table_slots=slots)
a = R.BoseFS(tuple(20 if i == 9 else 0 for i in range(20)))
H = R.HubbardMom1D(a, u=6.0, t=1.0)
style = R.IsDynamicSemistochastic()
v = R.GPUDVec([(a, 10.0)], style=style, capacity=int(per_gpu * 1.6) + 4096)
pv = v.similar()
R._lib.check(R._lib.lib().rimu_vec_reserve(pv.handle, int(per_gpu * 1.6) + 4096))
wm = R.working_memory(v, seed=1)
sp = R.ShiftParameters(R.diagonal_element(H, a), 10.0, 1e-4)
strat = R.DoubleLogUpdate(target_walkers=per_gpu)
t0 = time.time()
for step in range(460):
    t1 = time.time()
    R.apply_operator(wm, pv, v, R.FirstOrderTransitionOperator(H, sp.shift, sp.time_step))
    v, pv = pv, v
    s = wm.last_stats
    strat.update(sp, s.norm1)
    dt = time.time() - t1
    if dt > 0.05 or step % 40 == 0:
        print(f"step {step} dt {dt*1e3:.1f} ms norm {s.norm1:.3g} len {s.len} buckets {s.buckets} fill {s.max_bucket_fill} attempts {s.spawn_attempts} t {time.time()-t0:.1f}", flush=True)
