import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rimu_b200 as R
from rimu_b200 import _lib
rng = np.random.default_rng(0)
W = 1
at = R.AddressType(_lib.ADDR_BOSE, (20,), 20)
n = 80
keys = rng.integers(1, 2 ** 62, size=(n, W), dtype=np.uint64)
vals = np.arange(1, n + 1).astype(np.float64)
v = R.GPUDVec(style=R.IsDeterministic(), address_type=at)
v.assign(keys, vals)
k0, v0 = v.download()
for nb in (1, 5):
    _lib.check(_lib.lib().rimu_vec_rebucket(v.handle, nb))
    k1, v1 = v.download()
    st = np.zeros(nb, dtype=np.uint64); ln = np.zeros(nb, dtype=np.uint32)
    _lib.check(_lib.lib().rimu_vec_segments(v.handle, st.ctypes.data_as(_lib._u64p), ln.ctypes.data_as(C.POINTER(C.c_uint32))))
    print("nb", nb, "seg_start", st, "seg_len", ln)
    print(" vals", v1.astype(int).tolist())
    orig = {int(k[0]): x for k, x in zip(k0, v0)}
    print(" pairs intact", sum(1 for k, x in zip(k1, v1) if orig.get(int(k[0])) == x), "of", n)
