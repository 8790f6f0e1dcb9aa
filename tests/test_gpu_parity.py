"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Integer/index work must be bit-exact; Float64 within 1e-12 relative (north_star tolerance)."""
import math
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.cases import SPECS, oracle_ham, product_ham, sample_keys

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def sort_kv(keys, vals):
    keys = np.asarray(keys, dtype=np.uint64).reshape(len(vals), -1)
    order = np.lexsort(tuple(keys[:, j] for j in range(keys.shape[1])))
    return keys[order], np.asarray(vals)[order]


@pytest.mark.parametrize("name", sorted(SPECS))
def test_hamiltonian_elements(built, name):
    """diagonal_element / num_offdiagonals / get_offdiagonal(i) for every i, element-wise."""
    oh, ph = oracle_ham(name), product_ham(name)
    keys = sample_keys(oh, 24, seed=1)
    d_gpu = ph.diagonal_elements(keys)
    n_gpu = ph.nums_offdiagonals(keys)
    for j, key in enumerate(keys):
        kt = tuple(int(x) for x in key)
        assert d_gpu[j] == oh.diagonal_element(kt), (name, "diagonal", kt)
        L = oh.num_offdiagonals(kt)
        assert n_gpu[j] == L
        cap = min(L, 3000)
        ko, vo = ph.offdiagonals_of_key(key, 1, cap)
        idx = range(1, cap + 1) if L <= 3000 else None
        for i in range(1, cap + 1):
            ok, ov = oh.get_offdiagonal(kt, i)
            assert vo[i - 1] == ov, (name, kt, i, vo[i - 1], ov)
            if ov != 0.0:
                assert tuple(int(x) for x in ko[i - 1]) == ok, (name, kt, i)
        if L > 3000:  # tail block (second three-body segment) spot check
            first = L - 1999
            ko, vo = ph.offdiagonals_of_key(key, first, 2000)
            for i in range(0, 2000, 7):
                ok, ov = oh.get_offdiagonal(kt, first + i)
                assert vo[i] == ov
                if ov != 0.0:
                    assert tuple(int(x) for x in ko[i]) == ok


@pytest.mark.parametrize("name", ["ext_mom1d", "mom1d_ep", "mom1d_ep_f2c", "real1d_ep", "ext1d", "ext1d_hw", "real1d_6", "real1d_w2", "mom1d_bose", "mom1d_f2c", "rs_bose_2d", "rs_bose_3d_w2",
                                  "rs_fermi", "rs_f2c_4x4", "rs_f2c_trap", "tc_7", "tc_8_cut2",
                                  "rs_comp_bb", "rs_comp_bf", "rs_comp_bb_trap", "rs_comp_ffb", "rs_comp_ff_wide", "rs_comp_bf_w2"])
def test_deterministic_hv(built, name):
    """mul!(y, H, x) three times from the starting address: keys exact, values 1e-12 relative."""
    import rimu_b200 as R
    oh, ph = oracle_ham(name), product_ham(name)
    x = R.GPUDVec([(ph.address, 1.0)], style=R.IsDeterministic())
    ok, ov = np.array([oh.start_key], dtype=np.uint64), np.array([1.0])
    p = orc.make_params(orc.STYLE_DETERMINISTIC, plain_h=True)
    for it in range(3):
        y = x.similar()
        R.mul(y, ph, x)
        ok, ov, st = oh.step(p, ok, ov)
        gk, gv = y.download_sorted()
        scale = np.abs(ov).max()
        if not np.array_equal(gk, ok):
            # a sum of terms that cancel (the symmetric potential of HubbardMom1DEP with fermion signs) is an exact zero -- no
            # entry -- in one summation order and a rounding residue in another: such keys may differ, but only with values
            # below the Float64 tolerance of the product
            got = {tuple(k): v for k, v in zip(gk.tolist(), gv.tolist())}
            want = {tuple(k): v for k, v in zip(ok.tolist(), ov.tolist())}
            for k in set(got) ^ set(want):
                assert abs(got.get(k, 0.0) - want.get(k, 0.0)) <= RTOL * scale, (name, it, k)
            common = sorted(set(got) & set(want))
            gv, ov2 = np.array([got[k] for k in common]), np.array([want[k] for k in common])
            assert np.all(np.abs(gv - ov2) <= RTOL * np.maximum(np.abs(ov2), scale)), (name, it)
            # continue from the oracle's vector so that both sides see the same input
            y = R.GPUDVec(style=R.IsDeterministic(), address_type=x.address_type, ctx=x.ctx).upload(ok, ov)
        else:
            assert np.all(np.abs(gv - ov) <= RTOL * np.maximum(np.abs(ov), scale)), (name, it, np.abs(gv - ov).max())
        x = y
        if len(ok) > 20000:
            break


@pytest.mark.parametrize("name", ["ext_mom1d", "ext_mom1d_20", "mom1d_ep", "mom1d_ep_f2c", "real1d_ep", "ext1d", "ext1d_hw", "real1d_6", "real1d_10", "real1d_w2", "mom1d_bose", "mom1d_f2c", "rs_bose_2d",
                                  "rs_bose_3d_w2", "rs_f2c_4x4", "tc_7", "rs_comp_bf", "rs_comp_ffb", "rs_comp_bbb", "rs_comp_ff_wide", "rs_comp_bf_w2"])
def test_integer_walkers_bit_exact(built, name):
    """IsStochasticInteger FCIQMC steps: same Philox streams => identical vectors and statistics."""
    import rimu_b200 as R
    oh, ph = oracle_ham(name), product_ham(name)
    seed, dtau = 20240917, 0.002 if name.startswith("tc") else 0.01
    v = R.GPUDVec([(ph.address, 500)], style=R.IsStochasticInteger())
    wm = R.working_memory(v, seed=seed)
    ok, ov = np.array([oh.start_key], dtype=np.uint64), np.array([500], dtype=np.int64)
    shift = oh.diagonal_element(oh.start_key)
    for step in range(6):
        T = R.FirstOrderTransitionOperator(ph, shift + 2.0, dtau)
        out = v.similar()
        names, vals, _, _ = R.apply_operator(wm, out, v, T)
        v = out
        pp = orc.make_params(orc.STYLE_INTEGER, shift=shift + 2.0, dtau=dtau, key=orc.step_key(seed, step))
        ok, ov, st = oh.step(pp, ok, ov)
        gk, gv = v.download_sorted()
        assert np.array_equal(gk, ok), (name, step)
        assert np.array_equal(gv, ov), (name, step)
        s = wm.last_stats
        assert (s.spawn_attempts, s.ispawns, s.ideaths, s.iclones, s.izombies) == \
            (st.spawn_attempts, st.ispawns, st.ideaths, st.iclones, st.izombies)
        assert s.len == st.len_after and s.inorm1 == st.inorm1
        assert names == ("spawn_attempts", "spawns", "deaths", "clones", "zombies")


@pytest.mark.parametrize("name", ["real1d_6", "mom1d_bose", "rs_f2c_4x4", "tc_7", "rs_comp_fb", "rs_comp_bf_w2"])
def test_semistochastic_step(built, name):
    """IsDynamicSemistochastic: exact/inexact branch, late ThresholdCompression.  Random decisions are
    keyed on addresses, so results match the oracle except for Float64 summation order."""
    import rimu_b200 as R
    oh, ph = oracle_ham(name), product_ham(name)
    seed, dtau = 77, 0.002 if name.startswith("tc") else 0.01
    style = R.IsDynamicSemistochastic()
    v = R.GPUDVec([(ph.address, 40.0)], style=style)
    wm = R.working_memory(v, seed=seed)
    ok, ov = np.array([oh.start_key], dtype=np.uint64), np.array([40.0])
    shift = oh.diagonal_element(oh.start_key)
    for step in range(5):
        T = R.FirstOrderTransitionOperator(ph, shift + 1.0, dtau)
        out = v.similar()
        names, vals, _, _ = R.apply_operator(wm, out, v, T)
        v = out
        pp = orc.make_params(orc.STYLE_SEMISTOCHASTIC, shift=shift + 1.0, dtau=dtau, compress_threshold=1.0,
                             key=orc.step_key(seed, step))
        ok, ov, st = oh.step(pp, ok, ov)
        gk, gv = v.download_sorted()
        assert np.array_equal(gk, ok), (name, step)
        assert np.allclose(gv, ov, rtol=1e-10, atol=0), (name, step)
        s = wm.last_stats
        assert (s.exact_steps, s.inexact_steps, s.spawn_attempts, s.len_before, s.len) == \
            (st.exact_steps, st.inexact_steps, st.spawn_attempts, st.len_before, st.len_after)
        assert math.isclose(s.spawns, st.spawns, rel_tol=1e-10)
        assert math.isclose(s.norm1, st.norm1, rel_tol=1e-10)
        assert names[-1] == "len_before"
        # feed the GPU result back so that last-bit differences cannot accumulate into branch flips
        ok, ov = gk, gv


def test_with_threshold_style(built):
    import rimu_b200 as R
    name = "real1d_6"
    oh, ph = oracle_ham(name), product_ham(name)
    style = R.IsStochasticWithThreshold(1.0)
    v = R.GPUDVec([(ph.address, 25.0)], style=style)
    wm = R.working_memory(v, seed=5)
    ok, ov = np.array([oh.start_key], dtype=np.uint64), np.array([25.0])
    for step in range(4):
        T = R.FirstOrderTransitionOperator(ph, 1.0, 0.01)
        out = v.similar()
        R.apply_operator(wm, out, v, T)
        v = out
        pp = orc.make_params(orc.STYLE_WITH_THRESHOLD, shift=1.0, dtau=0.01, proj_threshold=1.0, key=orc.step_key(5, step))
        ok, ov, st = oh.step(pp, ok, ov)
        gk, gv = v.download_sorted()
        assert np.array_equal(gk, ok)
        assert np.allclose(gv, ov, rtol=1e-10, atol=0)
        ok, ov = gk, gv


@pytest.mark.parametrize("method", [0, 1, 2], ids=["hash", "sort", "partition"])
@pytest.mark.parametrize("W,dtype", [(1, np.float64), (1, np.int64), (2, np.float64), (2, np.int64)])
def test_annihilate_given_spawn_list(built, W, dtype, method):
    """Annihilation of a given spawn list with each method (HBM hash table, radix sort + segmented reduce,
    bucket partition + shared-memory merge): bit-exact for Int64, 1e-12 for Float64; includes heavy
    duplication, exact cancellation, empty input."""
    if method == 1 and W == 2:
        pytest.skip("the sort method covers one-word addresses")
    import rimu_b200 as R
    rng = np.random.default_rng(3)
    at = R.AddressType(R._lib.ADDR_BOSE, (20,) if W == 1 else (60,), 20 if W == 1 else 60)
    style = R.IsStochasticInteger() if dtype == np.int64 else R.IsDeterministic()
    for n, distinct in [(0, 1), (1, 1), (1000, 10), (200_000, 5000), (300_000, 300_000)]:
        pool = rng.integers(0, 2 ** 62, size=(max(distinct, 1), W), dtype=np.uint64)
        keys = pool[rng.integers(0, max(distinct, 1), size=n)]
        vals = rng.integers(-3, 4, size=n).astype(np.int64) if dtype == np.int64 else rng.uniform(-1, 1, size=n)
        if n >= 1000:  # force exact cancellations
            keys[1], vals[1] = keys[0], -vals[0]
        v = R.GPUDVec(style=style, address_type=at)
        R._lib.check(R._lib.lib().rimu_annihilate(v.handle, np.ascontiguousarray(keys).ctypes.data_as(R._lib._u64p),
                                                  np.ascontiguousarray(vals).ctypes.data_as(R._lib._vp), n, method))
        gk, gv = v.download_sorted()
        ok, ov = orc.annihilate(W, keys, vals)
        assert np.array_equal(gk, ok)
        if dtype == np.int64:
            assert np.array_equal(gv, ov)
        else:
            # Float64 sums in a different order: the error bound scales with the magnitude that was summed per address
            ka, mag = orc.annihilate(W, keys, np.abs(vals))
            scale = dict(zip(map(tuple, ka.tolist()), mag))
            bound = np.array([scale[tuple(k)] for k in ok.tolist()]) if len(ok) else np.zeros(0)
            assert np.all(np.abs(gv - ov) <= RTOL * bound + 1e-300)


def test_vector_interface(built):
    """Subset of test/DictVectors.jl test_dvec_interface that lies on the step path."""
    import rimu_b200 as R
    a, b, c = R.BoseFS(3, 2, 1), R.BoseFS(2, 3, 1), R.BoseFS(1, 1, 4)
    v = R.GPUDVec([(a, 1.5), (b, -2.0), (a, 0.5), (c, 0.0)], style=R.IsDeterministic())
    assert len(v) == 2 and v[a] == 2.0 and v[b] == -2.0 and v[c] == 0.0
    v[c] = 4.0
    assert len(v) == 3 and v[c] == 4.0
    v[c] = 0.0
    assert len(v) == 2
    assert v.norm(1) == 4.0 and math.isclose(v.norm(2), math.sqrt(8.0)) and v.norm(math.inf) == 2.0
    w = R.GPUDVec([(a, 1.0), (c, 3.0)], style=R.IsDeterministic())
    assert v.dot(w) == 2.0 and w.dot(v) == 2.0
    z = v + w
    assert z.to_dict() == {a: 3.0, b: -2.0, c: 3.0}
    z = v - v
    assert len(z) == 0
    z = 2.0 * v
    assert z[a] == 4.0
    v.add_(w, -2.0)
    assert v.to_dict() == {a: 0.0 + 2.0 - 2.0, b: -2.0, c: -6.0} or v.to_dict() == {b: -2.0, c: -6.0}
    assert len(v.zerovector()) == 0
    vi = R.GPUDVec([(a, 3), (b, -2)], style=R.IsStochasticInteger())
    assert vi.walkernumber() == 5.0 and R.walkernumber_and_length(vi) == (5.0, 2)
    vf = R.GPUDVec(style=R.IsDeterministic(), address_type=a.address_type).copy_from(vi)
    assert vf.to_dict() == {a: 3.0, b: -2.0}
    # iteration / reductions / convenience arithmetic (test/DictVectors.jl:124-150,183-249)
    u = R.GPUDVec([(a, 3.0), (b, -4.0)], style=R.IsDeterministic())
    assert sorted(u.values()) == [-4.0, 3.0] and set(u.keys()) == {a, b} and dict(iter(u)) == {a: 3.0, b: -4.0}
    assert a in u and c not in u and u.get(c, 7) == 7
    assert u.sum() == -1.0 and u.sum(abs) == 7.0 and u.mapreduce(lambda x: x * x, max) == 16.0
    assert u.all(lambda x: x != 0) and u.any(lambda x: x < 0) and not u.any(lambda x: x > 3)
    assert u.all(lambda k: k.num_particles == 6, over="keys")
    assert math.isclose(u.normalize().norm(2), 1.0) and u.norm(2) == 5.0 and math.isclose(u.normalize(1).norm(1), 1.0)
    assert (-u).to_dict() == {a: -3.0, b: 4.0} and (u / 2).to_dict() == {a: 1.5, b: -2.0} and (u * 2) == (2 * u)
    assert u == u.copy() and u != (u + u) and "2 entries" in repr(u)


def test_error_behaviour(built):
    import rimu_b200 as R
    a = R.BoseFS(1, 1, 1)
    H = R.HubbardReal1D(a)
    v = R.GPUDVec([(a, 1.0)], style=R.IsDynamicSemistochastic())
    with pytest.raises(ValueError):
        R.mul(v.similar(), H, v)  # mul! with non-deterministic working memory (pdvec.jl:814-819)
    with pytest.raises(ValueError):
        R.apply_operator(R.working_memory(v), v, v, H)  # aliasing
    with pytest.raises(R.RimuB200Error):
        vi = R.GPUDVec([(a, 1)], style=R.IsStochasticInteger())
        R.apply_operator(R.working_memory(v), v.similar(), vi, H)  # style/eltype mismatch
    with pytest.raises(IndexError):
        R.get_offdiagonal(H, a, 7)


def test_table_growth_retry(built):
    """A deliberately tiny working table overflows, the step reports it and the retry gives the same result."""
    import rimu_b200 as R
    name = "real1d_10"
    oh, ph = oracle_ham(name), product_ham(name)
    x = R.GPUDVec([(ph.address, 1.0)], style=R.IsDeterministic())
    for _ in range(5):
        y = x.similar()
        R.mul(y, ph, x)
        x = y
    ref = x.download_sorted()
    y = x.similar()
    wm = R.working_memory(x)
    R.apply_operator(wm, y, x, ph, table_slots=1024)  # forces internal regrow of the active table
    z = x.similar()
    R.apply_operator(wm, z, x, ph)
    k1, v1 = y.download_sorted()
    k2, v2 = z.download_sorted()
    assert np.array_equal(k1, k2) and np.allclose(v1, v2, rtol=1e-12)
    assert np.array_equal(x.download_sorted()[0], ref[0])  # source untouched


@pytest.mark.parametrize("W", [1, 2])
def test_rebucket_preserves_vector(built, W):
    """Re-segmentation (partition.cuh) is a permutation of the (key, value) pairs, every bucket's entries are
    contiguous and sit in the bucket the HOST hash function assigns (guards the nvcc hash miscompile that
    tests/cuda/t_hash_miscompile.cu reproduces)."""
    import ctypes as C
    import rimu_b200 as R
    from rimu_b200 import _lib
    rng = np.random.default_rng(W)
    at = R.AddressType(_lib.ADDR_BOSE, (20,) if W == 1 else (60,), 20 if W == 1 else 60)
    for n in (1, 80, 5000, 300_000):
        keys = rng.integers(1, 2 ** 62, size=(n, W), dtype=np.uint64)
        vals = rng.uniform(1, 2, size=n)
        v = R.GPUDVec(style=R.IsDeterministic(), address_type=at)
        v.assign(keys, vals)
        k0, v0 = v.download_sorted()
        for nb in (1, 5, 64, 1000):
            _lib.check(_lib.lib().rimu_vec_rebucket(v.handle, nb))
            k1, v1 = v.download_sorted()
            assert np.array_equal(k0, k1) and np.array_equal(v0, v1), (W, n, nb)
            ku, _ = v.download()
            st, ln = np.zeros(nb, dtype=np.uint64), np.zeros(nb, dtype=np.uint32)
            _lib.check(_lib.lib().rimu_vec_segments(v.handle, st.ctypes.data_as(_lib._u64p), ln.ctypes.data_as(C.POINTER(C.c_uint32))))
            assert int(ln.sum()) == n
            bucket_of_pos = np.full(n, -1, dtype=np.int64)
            for b in range(nb):
                bucket_of_pos[int(st[b]):int(st[b]) + int(ln[b])] = b
            assert (bucket_of_pos >= 0).all()
            for pos in rng.integers(0, n, size=min(n, 200)):
                kk = np.ascontiguousarray(ku[pos])
                h = _lib.lib().rimu_addr_hash(kk.ctypes.data_as(_lib._u64p), W)
                assert ((h >> 32) * nb) >> 32 == bucket_of_pos[pos], (W, n, nb, pos)


@pytest.mark.parametrize("name,style_name", [("real1d_10", "int"), ("rs_bose_2d", "int"), ("mom1d_bose", "semi"),
                                             ("real1d_w2", "int"), ("tc_7", "semi")])
def test_heavy_parents(built, name, style_name):
    """A determinant with far more walkers than HEAVY_T (1024) attempts goes through the heavy-parent queue
    (tiles of attempts, per-off-diagonal pre-summation in shared memory); results stay bit-exact for
    integer walkers and within 1e-10 for the semistochastic style (exact columns with L > 1024 for tc_7)."""
    import rimu_b200 as R
    oh, ph = oracle_ham(name), product_ham(name)
    seed, dtau = 4242, 0.002 if name.startswith("tc") else 0.005
    shift = oh.diagonal_element(oh.start_key)
    if style_name == "int":
        style, pop, dtype, ostyle, kw = R.IsStochasticInteger(), 150_000, np.int64, orc.STYLE_INTEGER, {}
    else:
        style, pop, dtype, ostyle, kw = R.IsDynamicSemistochastic(), 30_000.5, np.float64, orc.STYLE_SEMISTOCHASTIC, dict(compress_threshold=1.0)
    v = R.GPUDVec([(ph.address, pop)], style=style)
    wm = R.working_memory(v, seed=seed)
    ok, ov = np.array([oh.start_key], dtype=np.uint64), np.array([pop], dtype=dtype)
    for step in range(3):
        T = R.FirstOrderTransitionOperator(ph, shift, dtau)
        out = v.similar()
        R.apply_operator(wm, out, v, T)
        v = out
        pp = orc.make_params(ostyle, shift=shift, dtau=dtau, key=orc.step_key(seed, step), **kw)
        ok, ov, st = oh.step(pp, ok, ov)
        gk, gv = v.download_sorted()
        assert np.array_equal(gk, ok), (name, step)
        s = wm.last_stats
        assert s.spawn_attempts == st.spawn_attempts
        if style_name == "int":
            assert np.array_equal(gv, ov), (name, step)
            assert (s.ispawns, s.ideaths, s.iclones, s.izombies) == (st.ispawns, st.ideaths, st.iclones, st.izombies)
        else:
            assert np.allclose(gv, ov, rtol=1e-10, atol=0), (name, step)
            assert math.isclose(s.spawns, st.spawns, rel_tol=1e-10)
            ok, ov = gk, gv


@pytest.mark.parametrize("W", [1, 2])
def test_axpby_in_place_beyond_table_size(built, W):
    """y <- a*x + b*y in place with more records than the working table holds: the table must be regrown BEFORE
    anything is drained into the (aliased) destination.  Regression: a partial drain used to overwrite y and the
    retry then summed a corrupted input."""
    import rimu_b200 as R
    from rimu_b200 import _lib
    at = R.AddressType(_lib.ADDR_BOSE, (20,) if W == 1 else (60,), 20 if W == 1 else 60)
    ctx = R.Context(W, table_slots=1 << 16)  # own small context so that the overflow path is certain
    rng = np.random.default_rng(10 + W)
    n = 400_000
    pool = np.unique(rng.integers(1, 2 ** 62, size=(int(n * 1.5), W), dtype=np.uint64), axis=0)
    kx, ky = pool[rng.choice(len(pool), n, replace=False)], pool[rng.choice(len(pool), n, replace=False)]
    vx, vy = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    x = R.GPUDVec(style=R.IsDeterministic(), address_type=at, ctx=ctx)
    y = R.GPUDVec(style=R.IsDeterministic(), address_type=at, ctx=ctx)
    x.assign(kx, vx)
    y.assign(ky, vy)
    a, b = 0.75, -1.25
    y.axpby_(a, x, b)  # y = a*x + b*y
    gk, gv = y.download_sorted()
    k = np.concatenate([kx, ky])
    v = np.concatenate([a * vx, b * vy])
    order = np.lexsort(tuple(k[:, j] for j in range(W)))
    k, v = k[order], v[order]
    first = np.ones(len(v), dtype=bool)
    first[1:] = np.any(k[1:] != k[:-1], axis=1)
    ref = np.zeros(int(first.sum()))
    np.add.at(ref, np.cumsum(first) - 1, v)
    rk = k[first]
    rk, ref = sort_kv(rk[ref != 0], ref[ref != 0])
    assert np.array_equal(gk.reshape(-1, W), rk)
    assert np.allclose(gv, ref, rtol=1e-12, atol=0)
    assert math.isclose(x.dot(y), float(np.dot(*_aligned(kx, vx, rk, ref, W))), rel_tol=1e-9)
    del x, y
    ctx.close()


def _aligned(kx, vx, kr, vr, W):
    """values of x on the keys of r (0 where absent), for a host-side dot product"""
    d = {tuple(int(t) for t in key): val for key, val in zip(kr, vr)}
    return vx, np.array([d.get(tuple(int(t) for t in key), 0.0) for key in kx])


@pytest.mark.parametrize("name", ["mom1d_bose_20", "rs_bose_3d_w2"])
def test_frozen_dot_matches_dense_dot(built, name):
    """dot(::FrozenDVec, v) (pdvec.jl:773-779) through rimu_vec_dot_sparse: per-key lookups in the bucket segment of a
    segmented vector, full scan of an unsegmented one; Float64 and Int64 vectors; keys that are absent contribute 0."""
    import rimu_b200 as R
    ph = product_ham(name)
    x = R.GPUDVec([(ph.address, 1.0)], style=R.IsDeterministic())
    wm = R.working_memory(x)
    for _ in range(3 if name == "rs_bose_3d_w2" else 6):
        y = x.similar()
        R.mul(y, ph, x, wm)
        y.scale_(1.0 / y.norm(np.inf))
        x = y
    assert len(x) > 2000
    keys, vals = x.download_sorted()
    rng = np.random.default_rng(3)
    pick = np.sort(rng.choice(len(vals), size=500, replace=False))
    fk = np.concatenate([keys.reshape(len(vals), -1)[pick], np.full((3, keys.reshape(len(vals), -1).shape[1]), 5, dtype=np.uint64)])  # + 3 absent keys
    fv = np.concatenate([rng.uniform(-1, 1, 500), np.ones(3)])
    fr = R.FrozenDVec(fk, fv, x.address_type)
    ref = float(np.dot(fv[:500], vals[pick]))
    assert math.isclose(fr.dot(x), ref, rel_tol=1e-12)          # segmented (result of a partitioned step)
    assert math.isclose(x.dot(fr), ref, rel_tol=1e-12)
    u = R.GPUDVec(style=R.IsDeterministic(), address_type=x.address_type)
    u.assign(keys, vals)                                          # unsegmented
    assert math.isclose(fr.dot(u), ref, rel_tol=1e-12)
    iv = R.GPUDVec(style=R.IsStochasticInteger(), address_type=x.address_type)
    ivals = np.round(vals * 1000).astype(np.int64)
    nz = ivals != 0
    iv.assign(keys.reshape(len(vals), -1)[nz], ivals[nz])
    assert math.isclose(fr.dot(iv), float(np.dot(fv[:500], ivals[pick])), rel_tol=1e-12)
    # freeze() of a device vector and the dense dot agree
    small = R.GPUDVec(style=R.IsDeterministic(), address_type=x.address_type)
    small.assign(fk[:500], fv[:500])
    assert math.isclose(small.freeze().dot(x), small.dot(x), rel_tol=1e-12)


def test_vectors_may_outlive_their_context(built):
    """Host garbage collectors finalise vectors and contexts in arbitrary order (Julia finalizers, Python __del__):
    destroying a context first must neither crash nor leave a stale CUDA error behind; using the orphaned vector
    reports an error instead of touching freed memory."""
    import rimu_b200 as R
    a = R.BoseFS((1, 1, 1))
    H = R.HubbardReal1D(a)
    ctx = R.Context(1)
    v = R.GPUDVec([(a, 2.0)], style=R.IsDeterministic(), ctx=ctx)
    w = v.similar()
    R.mul(w, H, v)
    assert len(w) > 1
    ctx.close()                      # context gone, two vectors still alive
    with pytest.raises(R.RimuB200Error):
        v.norm(2)
    del v, w                         # the last one releases the context's remains
    # the default context is unaffected and sees no stale error
    x = R.GPUDVec([(a, 1.0)], style=R.IsDeterministic())
    y = x.similar()
    R.mul(y, H, x)
    assert math.isclose(y.norm(1), sum(abs(val) for val in y.download()[1]))


# --------------------------------------------------------------------------- ordered (order-deterministic) Float64 steps
@pytest.mark.parametrize("name", ["mom1d_bose", "rs_bose_2d", "rs_f2c_4x4", "real1d_w2"])
def test_ordered_semistochastic_steps_match_the_oracle_tightly(built, name):
    """rimu_step_params.ordered: every address is summed in sorted (address, value) order.  The oracle sums in its own
    (hash-map insertion) order, so the comparison is still a Float64 one -- but at 1e-12 relative, without feeding the device
    values back into the oracle between steps (the plain mode needs that, because its summation order varies)."""
    import rimu_b200 as R
    oh, ph = oracle_ham(name), product_ham(name)
    seed, dtau = 31, 0.01
    style = R.IsDynamicSemistochastic()
    v = R.GPUDVec([(ph.address, 40.0)], style=style)
    wm = R.working_memory(v, seed=seed, ordered=True)
    ok, ov = np.array([oh.start_key], dtype=np.uint64).reshape(1, -1), np.array([40.0])
    shift = oh.diagonal_element(oh.start_key)
    for step in range(6):
        out = v.similar()
        R.apply_operator(wm, out, v, R.FirstOrderTransitionOperator(ph, shift + 1.0, dtau))
        v = out
        ok, ov, st = oh.step(orc.make_params(orc.STYLE_SEMISTOCHASTIC, shift=shift + 1.0, dtau=dtau, compress_threshold=1.0,
                                             key=orc.step_key(seed, step)), ok, ov)
        gk, gv = v.download_sorted()
        assert np.array_equal(gk, ok), (name, step)
        assert np.allclose(gv, ov, rtol=1e-12, atol=0), (name, step, np.abs(gv - ov).max())
        assert math.isclose(wm.last_stats.norm1, st.norm1, rel_tol=1e-12)


def test_ordered_steps_are_bit_reproducible(built):
    """The same four steps on a 3e5-walker vector, run on fresh contexts whose merge grids differ (so that buckets meet
    different CTAs, the spawn kernels' appends interleave differently and the entries of a segment are written in a different
    order): with ordered summation the vectors AND the walker numbers are identical bit for bit.  (The walker number is reduced
    over the SORTED positions of a bucket: a sum over "the items a thread happens to own" is not reproducible -- found with
    compute-sanitizer, whose timing made it visible.)"""
    import rimu_b200 as R
    from tests.test_gpu_energies import _grow
    ph = product_ham("mom1d_bose_20")
    big = _grow(R, ph, 300_000, R.IsDynamicSemistochastic())
    keys, vals = big.download()
    shift = R.diagonal_element(ph, ph.address)
    runs = []
    for grid in ("0", "37", "5", "0"):
        if grid != "0":
            os.environ["RIMU_B200_MERGE_GRID"] = grid
        try:
            ctx = R.Context(1)
        finally:
            os.environ.pop("RIMU_B200_MERGE_GRID", None)
        v = R.GPUDVec(style=R.IsDynamicSemistochastic(), address_type=big.address_type, ctx=ctx)
        v.assign(keys, vals)
        wm = R.working_memory(v, seed=5, ordered=True)
        norms = []
        for _ in range(4):
            out = v.similar()
            R.apply_operator(wm, out, v, R.FirstOrderTransitionOperator(ph, shift, 1e-3))
            v = out
            norms.append(wm.last_stats.norm1)
        runs.append(v.download_sorted() + (norms,))
        del v, out, wm
        ctx.close()
    k0, v0, n0 = runs[0]
    for k1, v1, n1 in runs[1:]:
        assert np.array_equal(k0, k1) and np.array_equal(v0.view(np.uint64), v1.view(np.uint64))
        assert n0 == n1
