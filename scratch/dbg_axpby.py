import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rimu_b200 as R
from rimu_b200 import _lib

def ref_axpby(kx, vx, ky, vy, a, b, W):
    # z = a*x + b*y on host via structured sort
    k = np.concatenate([kx, ky]); v = np.concatenate([a * vx, b * vy])
    order = np.lexsort(tuple(k[:, j] for j in range(W)))
    k, v = k[order], v[order]
    new = np.ones(len(v), bool); new[1:] = np.any(k[1:] != k[:-1], axis=1)
    idx = np.cumsum(new) - 1
    out = np.zeros(idx[-1] + 1); np.add.at(out, idx, v)
    ku = k[new]
    nz = out != 0
    return ku[nz], out[nz]

for W in (1, 2):
    at = R.AddressType(_lib.ADDR_BOSE, (20,) if W == 1 else (60,), 20 if W == 1 else 60)
    rng = np.random.default_rng(W)
    for n in (1000, 300_000, 3_000_000, 20_000_000):
        pool = rng.integers(1, 2 ** 62, size=(int(n * 1.5), W), dtype=np.uint64)
        pool = np.unique(pool, axis=0)
        ix = rng.choice(len(pool), size=n, replace=False); iy = rng.choice(len(pool), size=n, replace=False)
        kx, ky = pool[ix], pool[iy]
        vx, vy = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        x = R.GPUDVec(style=R.IsDeterministic(), address_type=at); x.assign(kx, vx)
        y = R.GPUDVec(style=R.IsDeterministic(), address_type=at); y.assign(ky, vy)
        a, b = 0.75, -1.25
        z = x.copy()
        kc, vc = z.download_sorted()
        kx_s, vx_s = x.download_sorted()
        ok_copy = np.array_equal(kc, kx_s) and np.array_equal(vc, vx_s)
        z.axpby_(b, y, a)
        kz, vz = z.download_sorted()
        kr, vr = ref_axpby(kx, vx, ky, vy, a, b, W)
        order = np.lexsort(tuple(kr[:, j] for j in range(W)))
        ok_keys = kz.shape == kr.shape and np.array_equal(kz.reshape(-1, W), kr[order])
        ok_vals = ok_keys and np.allclose(vz, vr[order], rtol=1e-12)
        d = x.dot(y)
        # host dot
        kk = np.concatenate([kx, ky]); 
        print(f"W={W} n={n}: copy {ok_copy} axpby keys {ok_keys} vals {ok_vals} len {len(vz)} vs {len(vr)} dot {d:.6f}", flush=True)
        # second in-place add on the large result (segmented vector now)
        z2 = z.copy(); z2.add_(z, -1.0)
        print("   z - z len:", len(z2), "norm", z2.norm(2), flush=True)
