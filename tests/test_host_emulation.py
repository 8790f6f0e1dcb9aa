"""The DEVICE address arithmetic checked on the CPU: rimu.jl_b200/csrc/{common,hamiltonians}.cuh and the host half of
rimu_ham_create (csrc/ham_host.h) are compiled with g++ (tests/cuda/host_ham.cpp supplies the few CUDA intrinsics) and
compared element by element with the ONR-based oracle -- diagonal_element, num_offdiagonals and EVERY get_offdiagonal(i)
of sampled addresses for every model case, plus exhaustive checks of the two primitives everything leans on (rank/select,
float-reciprocal division).  The same comparison runs on the GPU in tests/test_gpu_parity.py::test_hamiltonian_elements;
this one needs no device, so the kernels' arithmetic is under test in the CPU suite as well.  (Floating point: g++ is run with
-ffp-contract=off, nvcc with --fmad=false; sqrt and division are IEEE-exact on both.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.cases import SPECS, oracle_ham, product_ham, sample_keys

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    so, src = os.path.join(HERE, "libhost_ham.so"), os.path.join(HERE, "host_ham.cpp")
    deps = [src] + [os.path.join(ROOT, "rimu.jl_b200", "csrc", f) for f in ("common.cuh", "hamiltonians.cuh", "ham_host.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                        src, "-o", so], check=True, capture_output=True)
    L = C.CDLL(so)
    u64p = C.POINTER(C.c_uint64)
    L.emu_ham_create.restype, L.emu_ham_create.argtypes = C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
    L.emu_ham_destroy.argtypes = [C.c_void_p]
    L.emu_ham_words.restype, L.emu_ham_words.argtypes = C.c_int, [C.c_void_p]
    L.emu_diagonal.restype, L.emu_diagonal.argtypes = C.c_double, [C.c_void_p, u64p]
    L.emu_num_offdiagonals.restype, L.emu_num_offdiagonals.argtypes = C.c_longlong, [C.c_void_p, u64p]
    L.emu_offdiagonal.restype, L.emu_offdiagonal.argtypes = C.c_double, [C.c_void_p, u64p, C.c_longlong, u64p]
    L.emu_select64.restype, L.emu_select64.argtypes = C.c_int, [C.c_uint64, C.c_int]
    L.emu_udiv_small.restype, L.emu_udiv_small.argtypes = C.c_uint, [C.c_uint, C.c_uint]
    return L


def test_select_and_small_division_exhaustive(emu):
    rng = np.random.default_rng(0)
    words = [0x1, 0x8000000000000000, 0xFFFFFFFFFFFFFFFF, 0x5555555555555555, 0xF0F0F0F00F0F0F0F] + \
        [int(x) for x in rng.integers(1, 2 ** 63, size=300, dtype=np.uint64)]
    for w in words:
        pos = [i for i in range(64) if (w >> i) & 1]
        for k in range(len(pos)):
            assert emu.emu_select64(w, k) == pos[k]
    # udiv_small is specified for x < 2^21, 0 < d < 2^10: all divisors, a dense sweep of dividends including every boundary
    for d in range(1, 1024):
        xs = np.unique(np.concatenate([np.arange(0, 4096), np.arange(d - 2 if d > 2 else 0, 2 ** 21, d)[:3000],
                                       np.arange(d - 1, 2 ** 21, d)[:3000], np.arange(2 ** 21 - 2048, 2 ** 21)]))
        for x in xs[::7] if d > 64 else xs:
            assert emu.emu_udiv_small(int(x), d) == int(x) // d, (x, d)


@pytest.mark.parametrize("name", sorted(SPECS))
def test_device_hamiltonian_code_matches_oracle_on_the_host(built, emu, name):
    oh, ph = oracle_ham(name), product_ham(name)  # ph only supplies the rimu_ham_desc the product would hand to rimu_ham_create
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    assert emu.emu_ham_create(C.byref(ph.desc), C.byref(h), err, 512) == 0, err.value
    W = emu.emu_ham_words(h)
    assert W == oh.W
    keys = sample_keys(oh, 16, seed=2)
    u64p = C.POINTER(C.c_uint64)
    out = (C.c_uint64 * 2)()
    checked = 0
    for key in keys:
        kt = tuple(int(x) for x in key)
        kin = (C.c_uint64 * 2)(*(list(kt) + [0] * (2 - len(kt))))
        assert emu.emu_diagonal(h, C.cast(kin, u64p)) == oh.diagonal_element(kt), (name, "diagonal", kt)
        L = oh.num_offdiagonals(kt)
        assert emu.emu_num_offdiagonals(h, C.cast(kin, u64p)) == L
        idx = range(1, L + 1) if L <= 1500 else list(range(1, 700)) + list(range(L - 800, L + 1)) + list(range(700, L - 800, max(1, L // 600)))
        for i in idx:
            ok, ov = oh.get_offdiagonal(kt, i)
            v = emu.emu_offdiagonal(h, C.cast(kin, u64p), i - 1, C.cast(out, u64p))
            assert v == ov, (name, kt, i, v, ov)
            if ov != 0.0:
                assert tuple(int(out[j]) for j in range(W)) == tuple(ok), (name, kt, i)
            checked += 1
    assert checked > 0
    emu.emu_ham_destroy(h)


def test_host_validation_errors(emu):
    """rimu_ham_create's argument checks live in ham_host.h: exercised here without a device."""
    import rimu_b200 as R
    from rimu_b200 import _lib
    h, err = C.c_void_p(), C.create_string_buffer(512)
    d = _lib.HamDesc()
    d.model, d.addr_kind, d.num_modes, d.num_components = _lib.TRANSCORRELATED_1D, _lib.ADDR_BOSE, 4, 1
    d.num_particles[0] = 4
    assert emu.emu_ham_create(C.byref(d), C.byref(h), err, 512) == _lib.ERR_INVALID and b"not implemented" in err.value
    d.model, d.num_modes = _lib.HUBBARD_MOM_1D, 2
    assert emu.emu_ham_create(C.byref(d), C.byref(h), err, 512) == _lib.ERR_INVALID and b"at least 3 modes" in err.value
    d.model, d.num_modes, d.ndim = _lib.HUBBARD_REAL_SPACE, 6, 2
    d.dims[0], d.dims[1] = 2, 2
    assert emu.emu_ham_create(C.byref(d), C.byref(h), err, 512) == _lib.ERR_INVALID and b"correct number of sites" in err.value
    d.model, d.boundary_condition = _lib.EXTENDED_HUBBARD_REAL_1D, 7
    assert emu.emu_ham_create(C.byref(d), C.byref(h), err, 512) == _lib.ERR_INVALID and b"boundary" in err.value


class _EmuStep(C.Structure):
    _fields_ = [("style", C.c_int), ("plain_h", C.c_int), ("shift", C.c_double), ("dtau", C.c_double), ("boost", C.c_double),
                ("proj_thr", C.c_double), ("rel_thr", C.c_double), ("abs_thr", C.c_double), ("k0", C.c_uint32), ("k1", C.c_uint32),
                ("init_rule", C.c_int), ("init_thr", C.c_double)]


STYLES = {  # name -> (oracle style, integer values?, proj_threshold, plain_h)
    "deterministic_H": (0, False, 0.0, True),
    "deterministic_T": (0, False, 0.0, False),
    "integer": (1, True, 0.0, False),
    "semistochastic": (2, False, 0.0, False),
    "with_threshold": (3, False, 1.0, False),
}


@pytest.mark.parametrize("style_name", sorted(STYLES))
@pytest.mark.parametrize("name", ["ext_mom1d", "mom1d_ep", "mom1d_ep_f2c", "real1d_10", "real1d_w2", "ext1d_twisted", "real1d_ep", "mom1d_bose", "mom1d_f2c", "rs_bose_2d_hw",
                                  "rs_bose_3d_w2", "rs_f2c_4x4", "tc_7"])
def test_device_step_arithmetic_matches_oracle_on_the_host(built, emu, name, style_name):
    """Everything ONE parent deposits in a step -- the diagonal death/cloning value and every spawn attempt (Philox draw,
    off-diagonal pick, exact/stochastic decision, projection, initiator lane) -- computed with the kernels' own functions
    (csrc/step_math.cuh compiled for the host) must reproduce the oracle's step on that single parent: bit-exact for integer
    walkers, 1e-13 for Float64 (same summation order), including the statistics the step reports."""
    import math
    from oracle import oracle as orc
    style, is_int, proj_thr, plain = STYLES[style_name]
    oh, ph = oracle_ham(name), product_ham(name)
    h, err = C.c_void_p(), C.create_string_buffer(512)
    assert emu.emu_ham_create(C.byref(ph.desc), C.byref(h), err, 512) == 0, err.value
    emu.emu_parent_deposits.restype = C.c_longlong
    W = oh.W
    u64p = C.POINTER(C.c_uint64)
    shift = oh.diagonal_element(oh.start_key) + 1.5
    values = [1, 3, -2, 40, 1300] if is_int else [0.3, 1.0, -2.7, 57.4, 1300.5]
    cap = 200000
    ck, cv, cl = (C.c_uint64 * (cap * W))(), (C.c_double * cap)(), (C.c_int * cap)()
    for step, key in enumerate(sample_keys(oh, 6, seed=4)):
        kt = tuple(int(x) for x in key)
        for rule in (0, 1, 3):
            if rule and plain:
                continue
            for val in values:
                k0, k1 = orc.step_key(11, step)
                pp = orc.make_params(style, shift=shift, dtau=0.01, plain_h=plain, proj_threshold=proj_thr, key=(k0, k1),
                                     initiator_rule=rule, initiator_threshold=1.0)
                ko, vo, st = oh.step(pp, np.array([kt], dtype=np.uint64), np.array([val], dtype=np.int64 if is_int else np.float64))
                want = {tuple(int(t) for t in k): float(v) for k, v in zip(ko.reshape(len(vo), W), vo)}
                es = _EmuStep(style, int(plain), shift, 0.01, 1.0, proj_thr, 1.0, math.inf, k0, k1, rule, 1.0)
                kin = (C.c_uint64 * 2)(*(list(kt) + [0] * (2 - len(kt))))
                dv, dl, att, ex, ssum = C.c_double(), C.c_int(), C.c_longlong(), C.c_int(), C.c_double()
                n = emu.emu_parent_deposits(h, C.byref(es), C.cast(kin, u64p), C.c_double(float(val)), int(is_int), cap,
                                            C.cast(ck, u64p), cv, cl, C.byref(dv), C.byref(dl), C.byref(att), C.byref(ex), C.byref(ssum))
                assert 0 <= n <= cap
                assert att.value == st.spawn_attempts and ex.value == st.exact_steps, (name, style_name, val)
                assert math.isclose(ssum.value, float(st.ispawns) if is_int else st.spawns, rel_tol=1e-12, abs_tol=1e-300)
                # lanes -> (safe + initiator, unsafe, initiator non-zero) per address, then from_initiator_value
                acc = {}
                def add(k, v, lane):
                    a = acc.setdefault(k, [0.0, 0.0, False])
                    if lane == 1:
                        a[1] += v
                    else:
                        a[0] += v
                        a[2] |= lane == 2 and v != 0
                if dv.value != 0.0:
                    add(kt, dv.value, dl.value)
                for r in range(n):
                    add(tuple(int(ck[r * W + j]) for j in range(W)), cv[r], cl[r])
                got = {}
                for k, (a, u, fi) in acc.items():
                    v = a if rule in (0, 2) else (a + u if (fi or (rule == 3 and abs(u) > 1.0)) else a)
                    if rule == 0:
                        v = a + u  # without a rule everything is in lane 0 anyway
                    if v != 0.0:
                        got[k] = v
                assert got.keys() == want.keys(), (name, style_name, rule, val, len(got), len(want))
                for k, v in want.items():
                    assert (got[k] == v) if is_int else math.isclose(got[k], v, rel_tol=1e-13, abs_tol=1e-300), (name, style_name, rule, val, k, got[k], v)
    emu.emu_ham_destroy(h)


def _random_onr(rng, kind, comps, M):
    """random occupation numbers with the particle numbers of `comps` (edge shapes included)"""
    out = []
    kinds = ["bose" if letter == "b" else "fermi" for letter in kind[5:]] if kind.startswith("comp:") else [kind] * len(comps)
    for kind, comp in zip(kinds, comps):
        N = sum(comp)
        if kind == "bose":
            mode = rng.integers(0, 4)
            if mode == 0:  # everything in one mode (first, last or random)
                onr = [0] * M
                onr[[0, M - 1, int(rng.integers(0, M))][int(rng.integers(0, 3))]] = N
            else:
                cuts = np.sort(rng.integers(0, N + 1, size=M - 1))
                onr = list(np.diff(np.concatenate([[0], cuts, [N]])))
        else:
            onr = [0] * M
            for m in rng.choice(M, size=N, replace=False):
                onr[int(m)] = 1
        out.append(tuple(int(x) for x in onr))
    return out


@pytest.mark.parametrize("name", ["real1d_10", "real1d_w2", "real1d_ep_w2", "ext1d_hw", "mom1d_bose_20", "mom1d_odd", "mom1d_f2c",
                                  "rs_bose_2d_hw", "rs_bose_3d_w2", "rs_fermi_hw", "rs_f2c_trap", "tc_8_cut2", "tc_32",
                                  "rs_comp_bf", "rs_comp_bb_trap", "rs_comp_ffb", "rs_comp_ff_wide", "rs_comp_bf_w2"])
def test_device_hamiltonian_code_on_random_addresses(built, emu, name):
    """Addresses the BFS walk from the starting address rarely visits: all particles in the first / last mode, random
    fillings, both address widths.  diagonal, count and a spread of off-diagonals, device code (host build) vs oracle."""
    oh, ph = oracle_ham(name), product_ham(name)
    h, err = C.c_void_p(), C.create_string_buffer(512)
    assert emu.emu_ham_create(C.byref(ph.desc), C.byref(h), err, 512) == 0, err.value
    W = oh.W
    u64p = C.POINTER(C.c_uint64)
    out = (C.c_uint64 * 2)()
    rng = np.random.default_rng(12345)
    kind = SPECS[name][1]
    comps = oh.start_onr
    for _ in range(60):
        onrs = _random_onr(rng, kind if kind == "bose" or kind.startswith("comp:") else "fermi", comps, oh.M)
        key = oh.pack(onrs[0] if len(onrs) == 1 else tuple(onrs))
        kt = tuple(int(x) for x in key) if isinstance(key, (tuple, list)) else tuple(int(x) for x in np.asarray(key, dtype=np.uint64).ravel())
        kin = (C.c_uint64 * 2)(*(list(kt) + [0] * (2 - len(kt))))
        assert emu.emu_diagonal(h, C.cast(kin, u64p)) == oh.diagonal_element(kt), (name, onrs)
        L = oh.num_offdiagonals(kt)
        assert emu.emu_num_offdiagonals(h, C.cast(kin, u64p)) == L, (name, onrs)
        idx = range(1, L + 1) if L <= 200 else sorted(set(int(x) for x in rng.integers(1, L + 1, size=200)) | {1, L})
        for i in idx:
            ok, ov = oh.get_offdiagonal(kt, i)
            v = emu.emu_offdiagonal(h, C.cast(kin, u64p), i - 1, C.cast(out, u64p))
            assert v == ov, (name, onrs, i, v, ov)
            if ov != 0.0:
                assert tuple(int(out[j]) for j in range(W)) == (tuple(ok) if isinstance(ok, (tuple, list)) else (int(ok),)), (name, onrs, i)
    emu.emu_ham_destroy(h)
