import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rimu_b200 as R
from oracle import oracle as orc
from tests.cases import oracle_ham, product_ham
name = sys.argv[1] if len(sys.argv) > 1 else "real1d_w2"
oh, ph = oracle_ham(name), product_ham(name)
x = R.GPUDVec([(ph.address, 1.0)], style=R.IsDeterministic())
ok, ov = np.array([oh.start_key], dtype=np.uint64), np.array([1.0])
p = orc.make_params(orc.STYLE_DETERMINISTIC, plain_h=True)
for it in range(3):
    y = x.similar()
    wm = R.working_memory(x)
    R.apply_operator(wm, y, x, ph)
    s = wm.last_stats
    ok, ov, st = oh.step(p, ok, ov)
    gk, gv = y.download_sorted()
    print("  attempts", s.spawn_attempts, st.spawn_attempts, "spawns", s.spawns, st.spawns, "exact", s.exact_steps, st.exact_steps)
    print(it, "gpu len", len(gv), "oracle len", len(ov), "buckets", s.buckets, "maxfill", s.max_bucket_fill, "deposits", s.deposits, "len_before", s.len_before)
    gs = {tuple(k): v for k, v in zip(gk.tolist(), gv)}
    os_ = {tuple(k): v for k, v in zip(ok.tolist(), ov)}
    miss = [k for k in os_ if k not in gs]
    extra = [k for k in gs if k not in os_]
    bad = [k for k in os_ if k in gs and abs(gs[k] - os_[k]) > 1e-9]
    print("  missing", len(miss), "extra", len(extra), "wrong values", len(bad), "sum gpu", sum(gs.values()), "sum orc", sum(os_.values()))
    for k in miss[:5]:
        print("   miss", [hex(w) for w in k], os_[k])
    for k in bad[:5]:
        print("   bad", [hex(w) for w in k], gs[k], os_[k])
    x = y
