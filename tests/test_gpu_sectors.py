"""Dense-indexed deterministic H*v over a complete sector (BASELINE config 3; csrc/sector.cuh): the device code against
(i) the oracle's dense-indexed rows -- plain-loop combinadic ranking + ONR off-diagonals, (ii) the dictionary path's `mul!`
(two independent device implementations), (iii) exact diagonalisation through the Lanczos driver.  The full-size checks
(4 up 4 down against scipy's eigsh on the oracle's matrix; 8 up 8 down = 165 636 900 determinants against sampled oracle rows
and the dictionary path) run with RIMU_B200_SLOW_TESTS=1 and print the line kept in profiles/."""
import json
import math
import os

import numpy as np
import pytest

from tests.cases import oracle_ham, product_ham
from tests.test_gpu_energies import exact_energy

pytestmark = pytest.mark.gpu

CASES = ["real1d_6", "rs_bose_2d", "rs_fermi", "rs_fermi_hw", "rs_f2c_4x4", "rs_f2c_trap", "mom1d_bose", "mom1d_f2c", "ext1d_twisted", "real1d_ep", "ext_mom1d", "mom1d_ep", "mom1d_ep_f2c"]


@pytest.mark.parametrize("name", CASES)
def test_rank_and_unrank_match_the_oracle(built, name):
    import rimu_b200 as R
    oh, ph = oracle_ham(name), product_ham(name)
    basis = R.SectorBasis(ph)
    assert basis.dim == oh.sector_dim()
    idx = np.unique(np.random.default_rng(1).integers(0, basis.dim, size=min(basis.dim, 400)))
    keys = basis.keys()
    assert np.array_equal(keys[idx], np.array([oh.sector_unrank(i)[0] for i in idx], dtype=np.uint64))
    assert np.array_equal(basis.rank(keys), np.arange(basis.dim))
    assert len(np.unique(keys)) == basis.dim
    assert basis.rank([oh.start_key if np.ndim(oh.start_key) == 0 else oh.start_key[0]])[0] == oh.sector_rank(oh.start_key)


@pytest.mark.parametrize("name", CASES)
def test_dense_hv_matches_oracle_rows_and_dictionary_path(built, name):
    import rimu_b200 as R
    oh, ph = oracle_ham(name), product_ham(name)
    basis = R.SectorBasis(ph)
    rng = np.random.default_rng(3)
    xh = rng.normal(size=basis.dim)
    xh[rng.random(basis.dim) < 0.2] = 0.0  # some exact zeros: the dictionary vector drops them
    x = basis.zeros().set(np.arange(basis.dim), xh)
    y = basis.zeros()
    R.mul(y, ph, x)
    yh = y.get()
    rows = np.arange(basis.dim) if basis.dim <= 3000 else np.unique(rng.integers(0, basis.dim, size=1500))
    want = oh.sector_rows(rows, xh)
    scale = np.abs(want).max()
    assert np.allclose(yh[rows], want, rtol=1e-12, atol=1e-12 * scale), np.abs(yh[rows] - want).max()
    # the dictionary path on the same vector (independent device implementation: spawn records + annihilation)
    xd = x.to_dvec()
    assert len(xd) == np.count_nonzero(xh)
    yd = xd.similar()
    R.mul(yd, ph, xd)
    back = basis.zeros().from_dvec(yd).get()
    assert np.allclose(back, yh, rtol=1e-11, atol=1e-12 * scale), np.abs(back - yh).max()
    # symmetry of the gather: <x|Hy> = <Hx|y>
    z = basis.zeros().set(np.arange(basis.dim), rng.normal(size=basis.dim))
    hz = basis.zeros()
    R.mul(hz, ph, z)
    assert math.isclose(x.dot(hz), y.dot(z), rel_tol=1e-10, abs_tol=1e-10)


@pytest.mark.parametrize("name", ["real1d_6", "rs_f2c_4x4", "mom1d_bose", "rs_fermi", "rs_f2c_3up3dn"])
def test_dense_lanczos_matches_exact_diagonalization(built, name):
    import rimu_b200 as R
    oh, ph = oracle_ham(name), product_ham(name)
    basis = R.SectorBasis(ph)
    start = basis.vector([(ph.address, 1.0)])
    vals, vecs, info = R.eigsolve_lanczos(ph, start, krylovdim=80, tol=1e-10, maxiter=30, full_reorth=True)
    assert info["converged"], info
    assert math.isclose(vals[0], exact_energy(oh), rel_tol=1e-9, abs_tol=1e-9), (vals[0], exact_energy(oh))


def test_non_hermitian_models_are_rejected(built):
    import rimu_b200 as R
    with pytest.raises((ValueError, R.RimuB200Error)):
        R.SectorBasis(product_ham("tc_7"))


@pytest.mark.skipif(not os.environ.get("RIMU_B200_SLOW_TESTS"), reason="full-size config-3 checks: set RIMU_B200_SLOW_TESTS=1")
def test_config3_full_size(built):
    import time
    import scipy.sparse.linalg as spla
    import rimu_b200 as R
    from oracle import oracle as orc

    def fermi(m, modes):
        return tuple(1 if (i + 1) in modes else 0 for i in range(m))

    out = {}
    # ---- 4 up 4 down (dim 3 312 400): E0 of the device Lanczos against scipy's eigsh on the oracle's sparse matrix
    up, dn = fermi(16, (1, 3, 6, 8)), fermi(16, (9, 11, 14, 16))
    oh = orc.OracleHam("HubbardRealSpace", "fermi2c", (up, dn), t=(1.0, 1.0), u=((0.0, 4.0), (4.0, 0.0)), dims=(4, 4))
    addr = R.FermiFS2C(up, dn)
    ph = R.HubbardRealSpace(addr, geometry=R.PeriodicBoundaries(4, 4), t=(1.0, 1.0), u=((0.0, 4.0), (4.0, 0.0)))
    basis = R.SectorBasis(ph)
    assert basis.dim == 1820 ** 2
    t0 = time.time()
    keys = basis.keys().reshape(-1, 1)
    Hs = oh.sparse_matrix(keys)  # rows/columns in the device's rank order
    t_build = time.time() - t0
    v0 = np.zeros(basis.dim)
    v0[oh.sector_rank(oh.start_key)] = 1.0
    e_ref = float(spla.eigsh(Hs, k=1, which="SA", v0=v0, tol=1e-12)[0][0])
    vals, vecs, info = R.eigsolve_lanczos(ph, basis.vector([(addr, 1.0)]), krylovdim=120, tol=1e-10, maxiter=20, full_reorth=True)
    assert info["converged"], info
    assert abs(vals[0] - e_ref) <= 1e-9 * abs(e_ref), (vals[0], e_ref)
    out["4up4dn"] = {"dim": basis.dim, "E0_device_lanczos": vals[0], "E0_eigsh_oracle_matrix": e_ref, "matvecs": info["matvecs"],
                     "oracle_matrix_build_s": t_build}
    del basis, vecs, Hs
    # ---- 8 up 8 down (config 3 itself, dim 165 636 900)
    up, dn = fermi(16, range(1, 9)), fermi(16, range(5, 13))
    oh = orc.OracleHam("HubbardRealSpace", "fermi2c", (up, dn), t=(1.0, 1.0), u=((0.0, 1.0), (1.0, 0.0)), dims=(4, 4))
    addr = R.FermiFS2C(up, dn)
    ph = R.HubbardRealSpace(addr, geometry=R.PeriodicBoundaries(4, 4), t=(1.0, 1.0), u=((0.0, 1.0), (1.0, 0.0)))
    basis = R.SectorBasis(ph)
    assert basis.dim == 12870 ** 2
    x = basis.vector([(addr, 1.0)])
    y = basis.zeros()
    for _ in range(12):  # power iterations fill the sector
        R.mul(y, ph, x)
        y.scale_(1.0 / y.norm(2))
        x, y = y, x
    times = []
    for _ in range(5):
        R.mul(y, ph, x)
        times.append(y.last_mul_ms)
    xh = x.get()
    rows = np.unique(np.random.default_rng(5).integers(0, basis.dim, size=20000))
    want = oh.sector_rows(rows, xh)
    got = y.gather(rows)
    scale = np.abs(want).max()
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12 * scale), np.abs(got - want).max()
    rq_dense = x.dot(y) / x.dot(x)
    # the dictionary path (table-method fallback at this size) on the same vector
    xd = x.to_dvec()
    yd = xd.similar()
    t0 = time.time()
    R.mul(yd, ph, xd)
    ms_dict = 1e3 * (time.time() - t0)
    rq_dict = xd.dot(yd) / xd.dot(xd)
    assert math.isclose(rq_dense, rq_dict, rel_tol=1e-11), (rq_dense, rq_dict)
    out["8up8dn"] = {"dim": basis.dim, "nonzeros": len(xd), "dense_ms_per_matvec": float(np.median(times)), "dictionary_ms_per_matvec": ms_dict,
                     "rayleigh_dense": rq_dense, "rayleigh_dictionary": rq_dict, "oracle_rows_checked": int(len(rows)),
                     "max_abs_row_error": float(np.abs(got - want).max())}
    print("config3 full size " + json.dumps(out))
