"""The C-ABI library loads without a GPU, exports every symbol include/rimu_b200.h declares, its
structs have the layout the ctypes mirror assumes, and compute entry points FAIL LOUDLY without a
device (there is no CPU fallback).  CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "rimu_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rimu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    from rimu_b200 import _lib
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/rimu_b200.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype in rimu.jl_b200/_lib.py"
    assert set(_lib.SYMBOLS) == set(names)


def test_struct_layouts(built):
    from rimu_b200 import _lib
    L = _lib.lib()
    assert L.rimu_sizeof_ham_desc() == C.sizeof(_lib.HamDesc)
    assert L.rimu_sizeof_step_params() == C.sizeof(_lib.StepParams)
    assert L.rimu_sizeof_step_stats() == C.sizeof(_lib.StepStats)
    assert L.rimu_sizeof_shift_params() == C.sizeof(_lib.ShiftParams)


def test_no_cpu_fallback(built):
    """Without a CUDA device every compute entry point returns RIMU_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rimu_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    st = L.rimu_ctx_create(0, 1, 1 << 12, C.byref(h))
    assert st == _lib.ERR_NO_DEVICE
    assert b"no CPU fallback" in L.rimu_last_error()
    import rimu_b200 as R
    with pytest.raises(R.RimuB200Error):
        R.GPUDVec([(R.BoseFS(1, 1, 1), 1.0)])


def test_host_side_helpers_match_oracle(built):
    """rimu_addr_hash / rimu_addr_owner / rimu_step_key / rimu_philox4x32_10 are pure host functions of
    the ABI; they must agree with the oracle's independent restatement (partitioning and RNG streams
    are what make multi-GPU integer runs bit-reproducible)."""
    from oracle import oracle as orc
    from rimu_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(0)
    for W in (1, 2):
        keys = rng.integers(0, 2 ** 63, size=(200, W), dtype=np.uint64)
        for k in keys:
            kk = np.ascontiguousarray(k)
            assert L.rimu_addr_hash(kk.ctypes.data_as(_lib._u64p), W) == orc.addr_hash(k)
            for nr in (1, 2, 3, 8):
                assert L.rimu_addr_owner(kk.ctypes.data_as(_lib._u64p), W, nr) == orc.addr_owner(k, nr)
    for seed, step in ((0, 0), (1, 2), (2 ** 63 + 5, 10 ** 9)):
        out = (C.c_uint32 * 2)()
        L.rimu_step_key(seed, step, out)
        assert (out[0], out[1]) == orc.step_key(seed, step)
    ctr, key, out = (C.c_uint32 * 4)(1, 2, 3, 4), (C.c_uint32 * 2)(5, 6), (C.c_uint32 * 4)()
    L.rimu_philox4x32_10(ctr, key, out)
    assert tuple(out) == orc.philox((1, 2, 3, 4), (5, 6))
    z4, z2 = (C.c_uint32 * 4)(0, 0, 0, 0), (C.c_uint32 * 2)(0, 0)
    L.rimu_philox4x32_10(z4, z2, out)
    assert tuple(out) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)  # Random123 known answer


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under rimu.jl_b200/ (Python or CUDA sources) may import, link or mention
    loading it, and the product library must not link the oracle's shared object."""
    import subprocess
    pkg = os.path.join(ROOT, "rimu.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            src = open(os.path.join(dirpath, f), errors="ignore").read()
            assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
    from rimu_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
