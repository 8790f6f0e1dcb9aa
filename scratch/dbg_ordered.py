import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rimu_b200 as R
from tests.cases import product_ham
from tests.test_gpu_energies import _grow
ph = product_ham("mom1d_bose_20")
big = _grow(R, ph, int(sys.argv[1]) if len(sys.argv) > 1 else 300_000, R.IsDynamicSemistochastic())
keys, vals = big.download()
shift = R.diagonal_element(ph, ph.address)
runs = []
for grid in ("0", "37", "0"):
    if grid != "0":
        os.environ["RIMU_B200_MERGE_GRID"] = grid
    try:
        ctx = R.Context(1)
    finally:
        os.environ.pop("RIMU_B200_MERGE_GRID", None)
    v = R.GPUDVec(style=R.IsDynamicSemistochastic(), address_type=big.address_type, ctx=ctx)
    v.assign(keys, vals)
    wm = R.working_memory(v, seed=5, ordered=True)
    per = []
    for _ in range(3):
        out = v.similar()
        R.apply_operator(wm, out, v, R.FirstOrderTransitionOperator(ph, shift, 1e-3))
        v = out
        s = wm.last_stats
        per.append((v.download_sorted(), s.norm1, s.len, s.buckets, s.spawn_attempts))
    runs.append(per)
    del v, out, wm
    ctx.close()
for a, b, tag in ((runs[0], runs[1], "grid0 vs grid37"), (runs[0], runs[2], "grid0 vs grid0")):
    for st in range(3):
        (ka, va), na, la, ba, aa = a[st]
        (kb, vb), nb_, lb, bb, ab = b[st]
        same_k = ka.shape == kb.shape and np.array_equal(ka, kb)
        ndiff = int((va.view(np.uint64) != vb.view(np.uint64)).sum()) if same_k else -1
        print(tag, "step", st, "len", la, lb, "buckets", ba, bb, "attempts", aa, ab, "keys equal", same_k, "value bit diffs", ndiff,
              "max rel", float(np.abs(va - vb).max() / np.abs(va).max()) if same_k else None, "norm equal", na == nb_, flush=True)
