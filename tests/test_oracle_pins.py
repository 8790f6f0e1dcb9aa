"""Pins the CPU oracle (oracle/) against the reference's OWN known answers -- golden energies,
doctest vectors and hand-computed values that Rimu.jl v0.14.0 carries in its tests and docstrings
(SURVEY.md Appendix B).  The reference is pure Julia and cannot run here, so these fixtures are
what makes the oracle trustworthy; every case cites the reference file:line it was taken from.

CPU only (no GPU, no product code)."""
import math

import numpy as np
import pytest

from oracle import oracle as orc


def bose(model, onr, **kw):
    return orc.OracleHam(model, "bose", tuple(onr), **kw)


def fermi_onr(m, modes):
    return tuple(1 if (i + 1) in modes else 0 for i in range(m))


# ------------------------------------------------------------------ addresses
def test_bose_bit_layout():
    """bitstring.jl:464-472 / fockaddress.jl:80-81: mode 1 in the lowest bits, n ones then a 0."""
    h = bose("HubbardReal1D", (1, 0, 2))
    assert h.pack((1, 0, 2)) == (0b11001,)
    assert h.unpack((0b11001,)) == (1, 0, 2)
    h = bose("HubbardReal1D", (3, 2, 1))
    assert h.pack((3, 2, 1)) == (0b10110111,)
    # multi-word: BoseFS{40,40} needs 79 bits
    onr = tuple([1] * 40)
    h = bose("HubbardReal1D", onr)
    assert h.W == 2
    assert h.unpack(h.pack(onr)) == onr
    k = h.pack(onr)  # bits 0,2,4,...,78 set: word 0 = 0x5555..., word 1 = bits 64..78
    assert k == (0x5555555555555555, 0x5555)


def test_multichunk_bose_roundtrip_and_interaction():
    """test/BitStringAddresses.jl:152-231: multi-chunk BoseFS with known interaction sums
    (66*65, 8*7+2, 136*135)."""
    # BoseFS{66,64}: all 66 bosons in one mode -> sum n(n-1) = 66*65
    for mode in (0, 31, 63):
        onr = [0] * 62
        onr[mode % 62] = 66
        h = bose("HubbardReal1D", onr, u=2.0)
        assert h.W == 2
        key = h.pack(onr)
        assert h.unpack(key) == tuple(onr)
        assert h.diagonal_element(key) == 2.0 * 66 * 65 / 2
    onr = [0] * 60
    onr[3], onr[40], onr[59] = 8, 2, 1
    h = bose("HubbardReal1D", onr, u=1.0)
    assert h.diagonal_element(h.pack(onr)) == (8 * 7 + 2) / 2


def test_fermi_bit_layout_and_sign():
    """bitstring.jl:713-723 (bit m-1 <-> mode m); fockaddress.jl:197-206: excitation of
    FermiFS(1,1,0,0,1,1,1,1) creating (3,4), destroying (2,5) gives (1,0,1,1,0,1,1,1), -1."""
    h = orc.OracleHam("HubbardRealSpace", "fermi", (1, 1, 0, 0, 1, 1, 1, 1), dims=(8,))
    assert h.pack((1, 1, 0, 0, 1, 1, 1, 1)) == (0b11110011,)
    # the same sign rule drives the hops: a hop 2->3 passes no particle, 1->8 (periodic) passes 5 others
    key = h.pack((1, 1, 0, 0, 1, 1, 1, 1))
    offs = h.offdiagonals(key)
    # particle 2 (mode 2), +e1 -> mode 3: value -t * (+1)
    assert offs[2] == (h.pack((1, 0, 1, 0, 1, 1, 1, 1)), -1.0)


# ------------------------------------------------------------------ HubbardReal1D
def test_hopnextneighbour_doctest():
    """bosefs.jl:279-286: hopnextneighbour(BoseFS(1,0,1), 3) -> (BoseFS(2,0,0), 1.414...),
    hopnextneighbour(BoseFS(1,0,1), 4) -> (BoseFS(1,1,0), 1.0)."""
    h = bose("HubbardReal1D", (1, 0, 1), u=1.0, t=1.0)
    k = h.pack((1, 0, 1))
    k3, v3 = h.get_offdiagonal(k, 3)
    k4, v4 = h.get_offdiagonal(k, 4)
    assert h.unpack(k3) == (2, 0, 0) and v3 == -1.4142135623730951
    assert h.unpack(k4) == (1, 1, 0) and v4 == -1.0


def test_offdiagonals_doctest_real1d():
    """Interfaces/hamiltonians.jl:316-331: offdiagonals(HubbardReal1D(BoseFS(3,2,1)), addr)."""
    h = bose("HubbardReal1D", (3, 2, 1), u=1.0, t=1.0)
    got = [(h.unpack(k), v) for k, v in h.offdiagonals(h.pack((3, 2, 1)))]
    want = [((2, 3, 1), -3.0), ((2, 2, 2), -2.449489742783178), ((3, 1, 2), -2.0),
            ((4, 1, 1), -2.8284271247461903), ((4, 2, 0), -2.0), ((3, 3, 0), -1.7320508075688772)]
    assert got == want


def test_bose_hubbard_interaction_doctest():
    """bosefs.jl:393-398: bose_hubbard_interaction (2,1,1,0) -> 2, (3,0,1,0) -> 6."""
    h = bose("HubbardReal1D", (2, 1, 1, 0), u=2.0)
    assert h.diagonal_element(h.pack((2, 1, 1, 0))) == 2.0  # u * 2 / 2
    assert h.diagonal_element(h.pack((3, 0, 1, 0))) == 6.0


def test_energy_real1d_5():
    """test/lomc.jl:544-546: E0 of HubbardReal1D(BoseFS{5,5}(1,1,1,1,1)) u=t=1."""
    h = bose("HubbardReal1D", (1,) * 5, u=1.0, t=1.0)
    assert math.isclose(h.exact_eigenvalues()[0], -8.280991746582686, rel_tol=1e-12)


def test_energy_bhm_example():
    """test/KrylovKit.jl:7-15, scripts/BHM-example.jl:158: E0 = -4.0215 (atol 1e-4) for
    HubbardReal1D(near_uniform(BoseFS{6,6}); u=6, t=1)."""
    h = bose("HubbardReal1D", (1,) * 6, u=6.0, t=1.0)
    assert abs(h.exact_eigenvalues()[0] - (-4.0215)) < 1e-4


def test_energy_real1d_7():
    """test/mpi_runtests.jl:158-160."""
    h = bose("HubbardReal1D", (1,) * 7, u=6.0, t=1.0)
    assert math.isclose(h.exact_eigenvalues(max_dim=5000)[0], -4.628524493494574, rel_tol=1e-11)


def test_spectrum_real1d_3():
    """ExactDiagonalization/exact_diagonalization_problem.jl:80-84: full spectrum of
    HubbardReal1D(BoseFS(1,1,1))."""
    h = bose("HubbardReal1D", (1, 1, 1), u=1.0, t=1.0)
    want = [-5.09593, -1.51882, -1.51882, 1.55611, 1.6093, 1.6093, 4.0, 4.53982, 4.90952, 4.90952]
    assert np.allclose(h.exact_eigenvalues(), want, atol=6e-6)


# ------------------------------------------------------------------ HubbardMom1D
def test_mom1d_doctests():
    """Interfaces/hamiltonians.jl:229-280: diagonal_element 8.666666666666664, num_offdiagonals 10,
    get_offdiagonal(., 3) -> (BoseFS(2,1,3), 1.0) for HubbardMom1D(BoseFS(3,2,1))."""
    h = bose("HubbardMom1D", (3, 2, 1), u=1.0, t=1.0)
    k = h.pack((3, 2, 1))
    assert h.diagonal_element(k) == 8.666666666666664
    assert h.num_offdiagonals(k) == 10
    k3, v3 = h.get_offdiagonal(k, 3)
    assert h.unpack(k3) == (2, 1, 3) and math.isclose(v3, 1.0, rel_tol=1e-15)


def test_momentum_transfer_diagonal_doctest():
    """HubbardMom1D.jl:153-161: momentum_transfer_diagonal(HubbardMom1D(BoseFS{6,5}(1,2,3,0,0)), map) = 5.2
    (u/2M * onproduct); the full diagonal adds the kinetic part."""
    onr = (1, 2, 3, 0, 0)
    h = bose("HubbardMom1D", onr, u=1.0, t=1.0)
    _, kes = orc.mom1d_grid(5, 1.0)
    kin = sum(kes[i] * onr[i] for i in range(5))
    assert math.isclose(h.diagonal_element(h.pack(onr)) - kin, 5.2, rel_tol=1e-13)


def test_energy_mom1d_golden():
    """test/lomc.jl:642-643 (u=4) and :678-679 (u=1): BoseFS{10,10} with all ten bosons in mode 5."""
    onr = tuple(10 if i == 4 else 0 for i in range(10))
    for u, e0 in ((4.0, -9.251592973178997), (1.0, -16.36048582876015)):
        h = bose("HubbardMom1D", onr, u=u, t=1.0)
        basis = h.bfs_basis()
        import scipy.sparse.linalg as sla
        ev = sla.eigsh(h.sparse_matrix(basis), k=1, which="SA", tol=1e-12)[0][0]
        assert math.isclose(ev, e0, rel_tol=1e-10), (u, ev)


def test_energy_mom1d_small():
    """test/ExactDiagonalization.jl:183-189: E0 of HubbardMom1D(BoseFS(1,2,3)) in its BFS sector;
    DictVectors/pdvec.jl:128-141: lowest four eigenvalues of HubbardMom1D(BoseFS(0,0,5,0,0); u=6)."""
    h = bose("HubbardMom1D", (1, 2, 3), u=1.0, t=1.0)
    assert math.isclose(h.exact_eigenvalues()[0], -3.045633163020568, rel_tol=1e-12)
    h = bose("HubbardMom1D", (0, 0, 5, 0, 0), u=6.0, t=1.0)
    want = [-3.4311156892322234, 1.1821748602612363, 3.7377753753082823, 6.996390417443125]
    assert np.allclose(h.exact_eigenvalues()[:4], want, rtol=1e-12)


def test_mom1d_fermi2c_hv_doctest():
    """DictVectors/pdvec.jl:54-103: H*v for HubbardMom1D(FermiFS2C((1,1,0,0),(0,0,1,1)); t=4/pi^2, u=4)
    on the unit vector: 7 entries, diagonal 4.0 and six entries of magnitude 1; dot(dest,pv)=10 after
    adding; three-argument dot = 44."""
    a = ((1, 1, 0, 0), (0, 0, 1, 1))
    h = orc.OracleHam("HubbardMom1D", "fermi2c", a, u=4.0, t=4 / math.pi ** 2)
    p = orc.make_params(orc.STYLE_DETERMINISTIC, plain_h=True)
    keys, vals, st = h.step(p, [h.start_key], [1.0])
    assert len(vals) == 7
    d = {tuple(k): v for k, v in zip(keys.tolist(), vals)}
    assert math.isclose(d[h.start_key], 4.0, rel_tol=1e-14)
    others = sorted(abs(v) for k, v in d.items() if k != h.start_key)
    assert np.allclose(others, 1.0, rtol=1e-14)
    # dest = pv-like vector of all ones on the 7 keys... the doctest's dot(dest, op, pv) = 44 with
    # dest = op*pv + ... ; we pin the simpler identities <v|H|v> = 4 and |Hv|^2 = 16 + 6 = 22
    assert math.isclose(float(np.dot(vals, vals)), 22.0, rel_tol=1e-14)


def test_mom1d_fermi_equals_realspace():
    """test/Hamiltonians.jl:1066-1080: HubbardMom1D(FermiFS2C) and HubbardRealSpace have the same
    ground-state energy."""
    M = 6
    a = (fermi_onr(M, (1, 2, 3)), fermi_onr(M, (2, 3)))
    hm = orc.OracleHam("HubbardMom1D", "fermi2c", a, u=2.0, t=1.0)
    hr = orc.OracleHam("HubbardRealSpace", "fermi2c", a, u=((0.0, 2.0), (2.0, 0.0)), t=(1.0, 1.0), dims=(M,))
    em = min(hm.exact_eigenvalues(start_key=hm.pack(k))[0] for k in [a])
    # the real-space sector is the whole Fock space; momentum space splits into M total-momentum sectors
    er = hr.exact_eigenvalues()[0]
    sect = []
    seen = set()
    basis_r = hr.bfs_basis()
    for key in basis_r:
        kt = tuple(int(x) for x in key)
        if kt in seen:
            continue
        b = hm.bfs_basis(start_key=kt)
        for kk in b:
            seen.add(tuple(int(x) for x in kk))
        sect.append(hm.exact_eigenvalues(start_key=kt)[0])
        if len(sect) >= M:
            break
    assert math.isclose(min(sect), er, rel_tol=1e-10)
    assert em >= er - 1e-9


# ------------------------------------------------------------------ HubbardRealSpace
def test_real1d_equals_realspace_1d():
    """test/Hamiltonians.jl:322-327: HubbardReal1D == HubbardRealSpace in one dimension (exact)."""
    onr = (1, 2, 0, 1, 1)
    h1 = bose("HubbardReal1D", onr, u=2.0, t=1.5)
    h2 = bose("HubbardRealSpace", onr, u=2.0, t=1.5, dims=(5,))
    e1, e2 = h1.exact_eigenvalues(), h2.exact_eigenvalues()
    assert np.allclose(e1, e2, rtol=0, atol=1e-12)


def test_realspace_fermion_energies():
    """test/Hamiltonians.jl:362-471: 2x2 t=2 -> -8 (N=1..3), 0 (N=4); 4x4: -4, -6, -8 for N=1,2,3;
    -10 for five fermions from FermiFS((1,0,1,0,1,0,1,0,1,0,0,...)) (sector-lowest from that start
    vector); 1-D three fermions t=3.5 -> -14 on 6 sites?  (see below)."""
    for n, e in ((1, -8.0), (2, -8.0), (3, -8.0), (4, 0.0)):
        h = orc.OracleHam("HubbardRealSpace", "fermi", fermi_onr(4, range(1, n + 1)), t=2.0, dims=(2, 2))
        assert math.isclose(h.exact_energy(), e, abs_tol=1e-9), (n, h.exact_energy())
    for n, e in ((1, -4.0), (2, -6.0), (3, -8.0)):
        h = orc.OracleHam("HubbardRealSpace", "fermi", fermi_onr(16, range(1, n + 1)), t=1.0, dims=(4, 4))
        assert math.isclose(h.exact_energy(), e, rel_tol=1e-3), (n, h.exact_energy())
    h = orc.OracleHam("HubbardRealSpace", "fermi", fermi_onr(16, (1, 3, 5, 7, 9)), t=1.0, dims=(4, 4))
    assert math.isclose(h.exact_energy(max_dim=10000), -10.0, rel_tol=1e-3)
    # 1-D, three fermions, t = 3.5 -> -14 (test/Hamiltonians.jl:362-367)
    h = orc.OracleHam("HubbardRealSpace", "fermi", (1, 1, 1, 0, 0, 0), t=3.5, dims=(6,))
    assert math.isclose(h.exact_energy(), -14.0, rel_tol=1e-4)
    # non-interacting two-component chain: -3 + -6 (:369-375); interactions move it the right way (:377-389)
    a = ((1, 1, 1, 1, 0, 0), (1, 1, 0, 0, 0, 0))
    e = {}
    for uu in (0.0, 1.0, -1.0):
        h = orc.OracleHam("HubbardRealSpace", "fermi2c", a, t=(1.0, 2.0), u=((0.0, uu), (uu, 0.0)), dims=(6,))
        e[uu] = h.exact_energy()
    assert math.isclose(e[0.0], -9.0, rel_tol=1e-4) and e[1.0] > -9 and e[-1.0] < -9
    # 3x3 two-component, u = 0: -16 (:437-445)
    a = (fermi_onr(9, (1, 2, 3)), fermi_onr(9, (1, 2)))
    h = orc.OracleHam("HubbardRealSpace", "fermi2c", a, t=(1.0, 2.0), u=((0.0, 0.0), (0.0, 0.0)), dims=(3, 3))
    assert math.isclose(h.exact_energy(max_dim=10000), -16.0, rel_tol=1e-3)


def test_realspace_offdiagonal_counts():
    """test/Hamiltonians.jl:295-321: FermiFS{3,12} on periodic / hard-wall 3x4 and 4x3 grids has 12
    off-diagonals; the number of non-zero ones is 6 / 8 (periodic 3x4 / 4x3?) and 3 / 4 (hard wall)."""
    onr = fermi_onr(12, (1, 2, 3))
    counts = {}
    for dims in ((3, 4), (4, 3)):
        for fold in ((True, True), (False, False)):
            h = orc.OracleHam("HubbardRealSpace", "fermi", onr, t=1.0, dims=dims, fold=fold)
            offs = h.offdiagonals(h.start_key)
            assert len(offs) == 12
            counts[(dims, fold[0])] = sum(1 for _, v in offs if v != 0.0)
    assert counts[((3, 4), True)] == 6 and counts[((4, 3), True)] == 8
    assert counts[((3, 4), False)] == 3 and counts[((4, 3), False)] == 4


# ------------------------------------------------------------------ Transcorrelated1D
def compare_to_bethe(g, nf, m):
    """test/Hamiltonians.jl:1155-1186 restated: same start addresses, t = m^2/2, v = t*2/m*g,
    energy = lowest eigenvalue of Matrix(ham) (BFS-connected sector of the start address)."""
    c = -(-m // 2)  # cld(m, 2)
    if nf == 2:
        f1 = f2 = fermi_onr(m, (c,))
        exact = {10: 5.2187287509452015, -10: -25.640329369393125}[g]
    elif nf == 3:
        f1, f2 = fermi_onr(m, (c, c + 1)), fermi_onr(m, (c,))
        exact = {-10: -15.151863462651115}[g]
    else:
        f1 = f2 = fermi_onr(m, (c - 1, c, c + 1))
        exact = {10: 148.90448481827905, -10: -43.819879567678}[g]
    t = m ** 2 / 2
    v = t * 2 / m * g
    h = orc.OracleHam("Transcorrelated1D", "fermi2c", (f1, f2), t=t, v=v, cutoff=1, three_body_term=True)
    ev = h.exact_eigenvalues(hermitian=False, max_dim=20000)
    return abs(ev[0].real - exact)


def test_transcorrelated_vs_bethe():
    """test/Hamiltonians.jl:1188-1197: the six Bethe-ansatz assertions of the reference."""
    assert compare_to_bethe(10, 2, 7) < 0.03
    assert compare_to_bethe(-10, 2, 7) <= 0.02
    assert compare_to_bethe(-10, 3, 7) <= 0.06
    assert compare_to_bethe(10, 6, 7) < 1.5
    assert compare_to_bethe(-10, 6, 7) < 0.4
    assert compare_to_bethe(-10, 3, 6) < compare_to_bethe(-10, 3, 7)


def test_transcorrelated_high_cutoff_equals_hubbard_continuum():
    """test/Hamiltonians.jl:1199-1248: with a cutoff beyond the grid the correlation factor vanishes
    and Transcorrelated1D reduces to HubbardMom1D with the continuum dispersion and u = v
    (off-diagonals out of range are dropped instead of folded, so compare diagonal elements and the
    in-range matrix elements)."""
    M = 6
    a = (fermi_onr(M, (3, 4)), fermi_onr(M, (3,)))
    tc = orc.OracleHam("Transcorrelated1D", "fermi2c", a, t=1.0, v=1.5, cutoff=100, three_body_term=True)
    _, _, ws, us = orc.tc_tables(M, 1.0, 100)
    assert np.all(us == 0.0)
    k = tc.start_key
    # diagonal: kinetic + N1 N2 (v/M + 2 v^2 W(0)/t)
    ks, kes, _, _ = orc.tc_tables(M, 1.0, 100)
    kin = kes[2] + kes[3] + kes[2]
    assert math.isclose(tc.diagonal_element(k), kin + 2 * (1.5 / M + 2 * 1.5 ** 2 * ws[0] / 1.0), rel_tol=1e-13)
    # three-body entries all vanish
    L = tc.num_offdiagonals(k)
    n_mom = 2 * 1 * (M - 1)
    assert all(tc.get_offdiagonal(k, i)[1] == 0.0 for i in range(n_mom + 1, L + 1))


# ------------------------------------------------------------------ stochastic styles
def test_diagonal_step_known_outcomes():
    """test/StochasticStyles.jl:128-153 (deterministic branch of diagonal_step!): with
    T = FirstOrderTransitionOperator(H, shift, dtau) on BoseFS(2,0,1) (H_aa = 1 for u=1):
      T(H,10,1)   x 2.5  -> value 25.0, clones 22.5
      T(H,-0.5,0.5) x 1.0 -> value 0.25, deaths 0.75
      T(H,-10,0.5) x 1.0  -> value -4.5, deaths 1, zombies 4.5."""
    h = bose("HubbardReal1D", (2, 0, 1), u=1.0, t=1.0)
    k = h.start_key
    assert h.diagonal_element(k) == 1.0
    for shift, dtau, val, want_val, clones, deaths, zombies in (
            (10.0, 1.0, 2.5, 25.0, 22.5, 0.0, 0.0),
            (-0.5, 0.5, 1.0, 0.25, 0.0, 0.75, 0.0),
            (-10.0, 0.5, 1.0, -4.5, 0.0, 1.0, 4.5)):
        # IsStochasticWithThreshold(0) == exact diagonal, stochastic spawns; use dtau-scaled spawns but
        # only look at the diagonal statistics, which do not depend on the random stream
        p = orc.make_params(orc.STYLE_WITH_THRESHOLD, shift=shift, dtau=dtau, proj_threshold=0.0, key=(1, 2))
        keys, vals, st = h.step(p, [k], [val])
        assert (st.clones, st.deaths, st.zombies) == (clones, deaths, zombies)


def test_diagonal_step_integer_outcomes():
    """test/StochasticStyles.jl:93-126 (integer branch): T(H,10,0.5) on BoseFS(2,0,1) x 1 -> 4 or 5 clones;
    T(H,-1,0.125) x 2 -> 0 or 1 deaths; T(H,-10,0.5) x 1 -> 1 death and 4 or 5 zombies."""
    h = bose("HubbardReal1D", (2, 0, 1), u=1.0, t=1.0)
    k = h.start_key
    seen = set()
    for rep in range(20):
        key = orc.step_key(7, rep)
        _, _, st = h.step(orc.make_params(orc.STYLE_INTEGER, shift=10.0, dtau=0.5, key=key), [k], np.array([1]))
        assert st.iclones in (4, 5) and st.ideaths == 0 and st.izombies == 0
        seen.add(st.iclones)
        _, _, st = h.step(orc.make_params(orc.STYLE_INTEGER, shift=-1.0, dtau=0.125, key=key), [k], np.array([2]))
        assert st.ideaths in (0, 1) and st.iclones == 0
        _, _, st = h.step(orc.make_params(orc.STYLE_INTEGER, shift=-10.0, dtau=0.5, key=key), [k], np.array([1]))
        assert st.ideaths == 1 and st.izombies in (4, 5)
    assert seen == {4, 5}


def test_spawning_expectation_equality():
    """test/StochasticStyles.jl:187-238: all spawning strategies have the same expectation value --
    the mean over many stochastic steps equals the deterministic step."""
    h = bose("HubbardReal1D", (1, 1, 1, 1), u=2.0, t=1.0)
    k = h.start_key
    pd = orc.make_params(orc.STYLE_DETERMINISTIC, shift=0.5, dtau=0.05)
    kd, vd, _ = h.step(pd, [k], [7.0])
    exact = {tuple(kk): v for kk, v in zip(kd.tolist(), vd)}
    nrep = 4000
    for style, kw, dtype in ((orc.STYLE_INTEGER, {}, np.int64), (orc.STYLE_WITH_THRESHOLD, dict(proj_threshold=1.0), np.float64),
                             (orc.STYLE_SEMISTOCHASTIC, dict(compress_threshold=1.0), np.float64)):
        acc = {}
        for rep in range(nrep):
            p = orc.make_params(style, shift=0.5, dtau=0.05, key=orc.step_key(99, rep), **kw)
            ks, vs, _ = h.step(p, [k], np.array([7], dtype=dtype))
            for kk, v in zip(ks.tolist(), vs):
                acc[tuple(kk)] = acc.get(tuple(kk), 0.0) + float(v)
        for kk, v in exact.items():
            m = acc.get(kk, 0.0) / nrep
            assert abs(m - v) < 5 * max(abs(v), 0.3) / math.sqrt(nrep) + 0.02, (style, kk, m, v)


def test_annihilation_semantics():
    """pdworkingmemory.jl:25-29, pdvec.jl:740-744: deposits to one address are summed, exact zeros
    are deleted; order of the result is unspecified (we sort)."""
    keys = np.array([[5], [3], [5], [9], [3], [9]], dtype=np.uint64)
    vals = np.array([2, -1, 3, 4, 1, -4], dtype=np.int64)
    ko, vo = orc.annihilate(1, keys, vals)
    assert ko.ravel().tolist() == [5] and vo.tolist() == [5]
    ko, vo = orc.annihilate(1, keys, vals.astype(np.float64) * 0.5)
    assert ko.ravel().tolist() == [5] and vo.tolist() == [2.5]


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors): the RNG both sides share."""
    assert orc.philox((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert orc.philox((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert orc.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


# --------------------------------------------------------------------------- initiator rules
def test_initiator_rules_hand_computed():
    """to_initiator_value / from_initiator_value (DictVectors/initiators.jl:132-236) on a two-site Bose-Hubbard
    model where every number can be done by hand: H = HubbardReal1D(BoseFS(1,1); u=1, t=1), operator = H, threshold 1.
    Off-diagonals: (1,1) -> (0,2) twice and (2,0) twice with -sqrt(2) each (periodic, M=2); (2,0) -> (1,1) twice with -sqrt(2);
    diagonal: H(2,0)(2,0) = 1, H(1,1)(1,1) = 0."""
    oh = orc.OracleHam("HubbardReal1D", "bose", (1, 1), u=1.0, t=1.0)
    k11, k20, k02 = (oh.pack((1, 1))[0], oh.pack((2, 0))[0], oh.pack((0, 2))[0])
    r2 = math.sqrt(2.0)

    def run(rule, vec):
        p = orc.make_params(orc.STYLE_DETERMINISTIC, plain_h=True, initiator_rule=rule, initiator_threshold=1.0)
        keys = np.array([k for k, _ in vec], dtype=np.uint64)
        vals = np.array([v for _, v in vec], dtype=np.float64)
        ko, vo, st = oh.step(p, keys, vals)
        return {int(k): float(v) for k, v in zip(ko.ravel(), vo)}, st

    # case 1: (1,1) => 2 is an initiator, (2,0) => 0.5 is not.
    #   (0,2): safe -4 sqrt2;  (2,0): safe -4 sqrt2 + 0.5 (diagonal of a non-initiator is safe);  (1,1): unsafe -sqrt2 only
    vec = [(k11, 2.0), (k20, 0.5)]
    base = {k02: 2 * (-r2 * 2.0), k20: 2 * (-r2 * 2.0) + 0.5}
    for rule, extra in ((orc.NON_INITIATOR, {k11: 2 * (-r2 * 0.5)}), (orc.INITIATOR, {}), (orc.SIMPLE_INITIATOR, {}),
                        (orc.COHERENT_INITIATOR, {k11: 2 * (-r2 * 0.5)})):  # |unsafe| = 1.41 > threshold: coherent spawns count
        got, st = run(rule, vec)
        want = {**base, **extra}
        assert got.keys() == want.keys(), (rule, got)
        for k in want:
            assert math.isclose(got[k], want[k], rel_tol=1e-15), (rule, k, got[k], want[k])
        assert st.len_before == 3  # the all-unsafe entry is still an entry before from_initiator_value

    # case 2: (2,0) => 3 is an initiator (diagonal deposit 3 goes to the initiator lane), (1,1) => 0.5 is not.
    #   (2,0): initiator 3, unsafe -sqrt2;  (0,2): unsafe -sqrt2 only;  (1,1): safe -6 sqrt2
    vec = [(k20, 3.0), (k11, 0.5)]
    unsafe = 2 * (-r2 * 0.5)
    for rule, want in ((orc.NON_INITIATOR, {k20: 3.0 + unsafe, k02: unsafe, k11: 2 * (-r2 * 3.0)}),
                       (orc.INITIATOR, {k20: 3.0 + unsafe, k11: 2 * (-r2 * 3.0)}),          # unsafe counts where an initiator lives
                       (orc.SIMPLE_INITIATOR, {k20: 3.0, k11: 2 * (-r2 * 3.0)}),           # non-initiators never spawn
                       (orc.COHERENT_INITIATOR, {k20: 3.0 + unsafe, k02: unsafe, k11: 2 * (-r2 * 3.0)})):  # |unsafe| = 1.41 > 1
        got, st = run(rule, vec)
        assert got.keys() == want.keys(), (rule, got)
        for k in want:
            assert math.isclose(got[k], want[k], rel_tol=1e-15), (rule, k, got[k], want[k])


def test_initiator_rule_zero_is_the_plain_step():
    """NonInitiator must reproduce the rule-free step bit for bit (integer walkers, several steps)."""
    oh = orc.OracleHam("HubbardReal1D", "bose", (1, 1, 1, 1, 1, 1), u=6.0, t=1.0)
    k, v = np.array([oh.start_key], dtype=np.uint64), np.array([50], dtype=np.int64)
    k2, v2 = k.copy(), v.copy()
    for step in range(4):
        key = orc.step_key(3, step)
        k, v, s1 = oh.step(orc.make_params(orc.STYLE_INTEGER, shift=1.0, dtau=0.02, key=key), k, v)
        k2, v2, s2 = oh.step(orc.make_params(orc.STYLE_INTEGER, shift=1.0, dtau=0.02, key=key, initiator_rule=0), k2, v2)
        assert np.array_equal(k, k2) and np.array_equal(v, v2)
    # and with a rule the population is never larger (spawns from non-initiators onto empty sites are dropped)
    k3, v3 = np.array([oh.start_key], dtype=np.uint64), np.array([50], dtype=np.int64)
    for step in range(4):
        k3, v3, s3 = oh.step(orc.make_params(orc.STYLE_INTEGER, shift=1.0, dtau=0.02, key=orc.step_key(3, step),
                                             initiator_rule=orc.INITIATOR, initiator_threshold=1.0), k3, v3)
    assert len(v3) <= len(v)


# --------------------------------------------------------------------------- 1-D models sharing HubbardReal1D's generator
def test_hubbard_real1d_ep_pins():
    """HubbardReal1DEP (Hamiltonians/HubbardReal1DEP.jl): the reference pins it three ways (test/Hamiltonians.jl):
    :392-396 same energy as HubbardRealSpace with a trap; :1039-1043 shift_lattice; :1045-1054 a single particle in a
    wide harmonic trap has the oscillator spectrum n + 1/2 (atol 0.005) and sits at zero potential."""
    assert orc.ep_lattice(3) == [0, 1, -1] and orc.ep_lattice(4) == [0, 1, -2, -1]  # js[1] == 0, circular
    for M in (3, 4):
        js = orc.ep_lattice(M)
        k = M // 2
        assert (js[-k:] + js[:-k] if k else js) == list(range(-(M // 2), -(M // 2) + M))  # shift_lattice_inv(js) == is
    a = (1, 2, 3, 4)
    h1 = orc.OracleHam("HubbardReal1DEP", "bose", a, u=2.0, t=3.0, v_ho=4.0)
    h2 = orc.OracleHam("HubbardRealSpace", "bose", a, u=2.0, t=3.0, dims=(4,), trap=((4.0,),))
    assert math.isclose(h1.exact_energy(), h2.exact_energy(), rel_tol=1e-12)
    m, l0 = 100, 10
    t, v_ho = 0.5 * l0 ** 2, 0.5 / l0 ** 2
    h = orc.OracleHam("HubbardReal1DEP", "bose", tuple(1 if i == 0 else 0 for i in range(m)), t=t, v_ho=v_ho)
    assert h.diagonal_element(h.start_key) == 0.0
    ev = h.exact_eigenvalues() + 2 * t
    assert np.allclose(ev[:3], [0.5, 1.5, 2.5], atol=0.005)


def test_extended_hubbard_real1d_pins():
    """ExtendedHubbardReal1D (Hamiltonians/ExtendedHubbardReal1D.jl): hand-computed diagonal; v = 0 is HubbardReal1D;
    boundary conditions as asserted in test/Hamiltonians.jl:1580-1603 (twisted: boundary hop changes sign, diagonal
    unchanged; hard wall: boundary hop is 0)."""
    a = (1, 0, 2, 1)
    h = orc.OracleHam("ExtendedHubbardReal1D", "bose", a, u=1.0, v=2.0, t=3.0)
    # sum n(n-1) = 2; neighbours: n3 n4 = 2, ring closure n4 n1 = 1 -> 3
    assert h.diagonal_element(h.start_key) == 1.0 * 2 / 2 + 2.0 * 3
    hw = orc.OracleHam("ExtendedHubbardReal1D", "bose", a, u=1.0, v=2.0, t=3.0, boundary_condition="hard_wall")
    assert hw.diagonal_element(hw.start_key) == 1.0 * 2 / 2 + 2.0 * 2
    plain = orc.OracleHam("HubbardReal1D", "bose", a, u=1.0, t=3.0)
    h0 = orc.OracleHam("ExtendedHubbardReal1D", "bose", a, u=1.0, v=0.0, t=3.0)
    basis = plain.bfs_basis()
    assert np.array_equal(h0.bfs_basis(), basis)
    assert (abs(plain.sparse_matrix(basis) - h0.sparse_matrix(basis))).max() == 0.0
    tw = orc.OracleHam("ExtendedHubbardReal1D", "bose", a, u=1.0, v=2.0, t=3.0, boundary_condition="twisted")
    key = h.start_key
    k_per, me = h.get_offdiagonal(key, 2)       # chosen = 2: first occupied mode (site 1) hops LEFT across the boundary
    k_tw, me_tw = tw.get_offdiagonal(key, 2)
    k_hw, me_hw = hw.get_offdiagonal(key, 2)
    assert me != 0.0 and me_tw == -me and k_tw == k_per and me_hw == 0.0
    assert tw.diagonal_element(key) == h.diagonal_element(key)
    assert h.get_offdiagonal(key, 1) == tw.get_offdiagonal(key, 1) == hw.get_offdiagonal(key, 1)  # interior hop
    # Hermitian for all three boundary conditions
    for ham in (h, tw, hw):
        b = ham.bfs_basis()
        mat = ham.sparse_matrix(b)
        assert abs(mat - mat.T).max() < 1e-14


# --------------------------------------------------------------------------- the committed golden fixture
def _golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_values.json")) as f:
        return json.load(f)


def _golden_ham(e):
    p = dict(e["params"])
    if "dims" in p:
        p["dims"] = tuple(p["dims"])
    return orc.OracleHam(e["model"], e["kind"], tuple(e["onr"]), **p)


def _close(got, e, want):
    if "abs_tol" in e:
        return abs(got - want) <= e["abs_tol"]
    return math.isclose(got, want, rel_tol=e["rel_tol"])


@pytest.mark.parametrize("entry", _golden()["energies"], ids=lambda e: e["id"])
def test_golden_fixture_energy(entry):
    """tests/golden/reference_values.json: values transcribed from the reference's tests (source = file:line there)."""
    h = _golden_ham(entry)
    basis = h.bfs_basis(max_dim=200_000)
    if len(basis) <= 5000:  # dense, with the overlap criterion (a sparse solve leaks into other symmetry sectors by rounding)
        got = h.exact_energy(max_dim=200_000)
    else:  # Krylov solve from the single start determinant, as the reference's exact_energy helper does
        import scipy.sparse.linalg as sla
        v0 = np.zeros(len(basis))
        v0[0] = 1.0
        got = float(sla.eigsh(h.sparse_matrix(basis), k=1, which="SA", v0=v0, tol=1e-12)[0][0])
    assert _close(got, entry, entry["value"]), (entry["id"], got, entry["value"], entry["source"])


def test_golden_fixture_spectra_and_elements():
    g = _golden()
    for e in g["spectra"]:
        ev = _golden_ham(e).exact_eigenvalues()[:len(e["values"])]
        for got, want in zip(ev, e["values"]):
            assert _close(float(got), e, want), (e["id"], got, want)
    for e in g["matrix_elements"]:
        h = _golden_ham(e)
        k = h.pack(tuple(e["onr"]))
        assert h.diagonal_element(k) == e["diagonal_element"] and h.num_offdiagonals(k) == e["num_offdiagonals"]
        k2, v2 = h.get_offdiagonal(k, e["get_offdiagonal"]["chosen"])
        assert h.unpack(k2) == tuple(e["get_offdiagonal"]["onr"]) and math.isclose(v2, e["get_offdiagonal"]["value"], rel_tol=1e-15)


# --------------------------------------------------------------------------- momentum-space siblings
def test_hubbard_mom1d_ep_relations_of_the_reference():
    """test/Hamiltonians.jl:1082-1106: (i) exact_energy(HubbardReal1DEP) == exact_energy(HubbardMom1DEP) for the same
    parameters; (ii) without a potential the fermionic HubbardMom1DEP matrix IS the HubbardMom1D matrix; (iii) two fermions of
    opposite spin and two bosons have the same ground-state energy in the trap."""
    real = orc.OracleHam("HubbardReal1DEP", "bose", (1, 1, 1, 1, 1), u=1.2, t=2.0, v_ho=2.0)
    mom = orc.OracleHam("HubbardMom1DEP", "bose", (0, 0, 5, 0, 0), u=1.2, t=2.0, v_ho=2.0)
    assert math.isclose(real.exact_energy(), mom.exact_energy(), rel_tol=1e-10)
    c = ((0, 1, 0, 1, 0, 0), (0, 0, 1, 0, 0, 0))
    plain = orc.OracleHam("HubbardMom1D", "fermi2c", c, u=2.0)
    ep0 = orc.OracleHam("HubbardMom1DEP", "fermi2c", c, u=2.0, v_ho=0.0)
    # the EP model lists (N1 + N2)(M - 1) more off-diagonals, all with value 0 when v_ho = 0: compare through the full sector
    b = np.array([plain.sector_unrank(i) for i in range(plain.sector_dim())]).reshape(-1, 1)
    assert np.array_equal(plain.sparse_matrix(b).toarray(), ep0.sparse_matrix(b).toarray())
    for disp in ("continuum", "hubbard"):
        bose = orc.OracleHam("HubbardMom1DEP", "bose", (0, 0, 2, 0, 0), v_ho=1.5, dispersion=disp)
        fermi = orc.OracleHam("HubbardMom1DEP", "fermi2c", ((0, 0, 1, 0, 0), (0, 0, 1, 0, 0)), v_ho=1.5, dispersion=disp)
        assert math.isclose(bose.exact_energy(), fermi.exact_energy(), rel_tol=1e-10, abs_tol=1e-12)


def test_extended_hubbard_mom1d_spectrum_is_contained_in_the_real_space_one():
    """test/Hamiltonians.jl:1627-1639 (boundary_condition = 0, bosons): every eigenvalue of ExtendedHubbardMom1D (one momentum
    sector) is an eigenvalue of ExtendedHubbardReal1D, to 8 digits."""
    hm = orc.OracleHam("ExtendedHubbardMom1D", "bose", (0, 0, 3, 0, 0, 0))
    hr = orc.OracleHam("ExtendedHubbardReal1D", "bose", (0, 0, 3, 0, 0, 0))
    em = np.round(hm.exact_eigenvalues(), 8)
    er = np.round(hr.exact_eigenvalues(), 8)
    assert len(em) >= 8 and all(np.any(np.abs(er - x) < 2e-8) for x in em)
    # the v = 0 limit is HubbardMom1D
    a = (0, 1, 2, 0, 1, 0)
    h0 = orc.OracleHam("ExtendedHubbardMom1D", "bose", a, u=1.5, v=0.0, t=1.0)
    h1 = orc.OracleHam("HubbardMom1D", "bose", a, u=1.5, t=1.0)
    bs = h1.bfs_basis()
    assert np.allclose(h0.sparse_matrix(bs).toarray(), h1.sparse_matrix(bs).toarray(), rtol=1e-14, atol=1e-14)


# ------------------------------------------------------------------ general CompositeFS on HubbardRealSpace
def _comp(letters, onr, **kw):
    return orc.OracleHam("HubbardRealSpace", "comp:" + letters, onr, **kw)


def test_composite_two_component_relations_of_the_reference():
    """test/Hamiltonians.jl:328-360: a single particle is a boson and a fermion at once, so BoseFS x BoseFS,
    BoseFS x FermiFS and the two swapped orders (t, u permuted accordingly) have the same exact energy."""
    three, one = (1, 1, 1, 0, 0, 0), (1, 0, 0, 0, 0, 0)
    e2 = _comp("bb", (three, one), t=(1.0, 4.0), u=((2.0, 3.0), (3.0, 0.0)), dims=(6,)).exact_energy()
    e3 = _comp("bf", (three, one), t=(1.0, 4.0), u=((2.0, 3.0), (3.0, 0.0)), dims=(6,)).exact_energy()
    e4 = _comp("bb", (one, three), t=(4.0, 1.0), u=((0.0, 3.0), (3.0, 2.0)), dims=(6,)).exact_energy()
    e5 = _comp("fb", (one, three), t=(4.0, 1.0), u=((0.0, 3.0), (3.0, 2.0)), dims=(6,)).exact_energy()
    assert math.isclose(e2, e3, rel_tol=1e-4) and math.isclose(e3, e4, rel_tol=1e-4) and math.isclose(e4, e5, rel_tol=1e-4)
    # :396-410 the same with a trap: component order does not matter
    e3t = _comp("bb", (three, one), trap=((1.0,), (4.0,)), u=((2.0, 3.0), (3.0, 0.0)), dims=(6,)).exact_energy()
    e4t = _comp("bb", (one, three), trap=((4.0,), (1.0,)), u=((0.0, 3.0), (3.0, 2.0)), dims=(6,)).exact_energy()
    assert math.isclose(e3t, e4t, rel_tol=1e-4)


def test_composite_fermions_equal_the_fermifs2c_path():
    """CompositeFS(FermiFS, FermiFS) in the general packed layout is the same operator as the FermiFS2C path that is pinned by
    test/Hamiltonians.jl:362-389 (-3 + -6 without interaction): element by element over the whole sector."""
    a = ((1, 1, 1, 1, 0, 0), (1, 1, 0, 0, 0, 0))
    for uu in (0.0, 1.0, -1.0):
        kw = dict(t=(1.0, 2.0), u=((0.0, uu), (uu, 0.0)), dims=(6,))
        hg, hf = _comp("ff", a, **kw), orc.OracleHam("HubbardRealSpace", "fermi2c", a, **kw)
        assert hg.W == hf.W == 1  # 12 bits either way, and the same bit layout: the keys coincide
        basis = hf.bfs_basis(None, 100000)
        for key in basis:
            kt = tuple(int(x) for x in np.atleast_1d(key))
            assert hg.diagonal_element(kt) == hf.diagonal_element(kt)
            assert hg.offdiagonals(kt) == hf.offdiagonals(kt)
    assert math.isclose(_comp("ff", a, t=(1.0, 2.0), u=((0.0, 0.0), (0.0, 0.0)), dims=(6,)).exact_energy(), -9.0, rel_tol=1e-4)


def test_composite_interaction_and_potential_by_hand():
    """HubbardRealSpace.jl:18-106 by hand for three components (fermion, fermion, boson) on a 2x2 grid:
    interaction = u_33/2 sum n(n-1) + sum_{s<t} u_st sum_i n_s(i) n_t(i); potential = sum_c sum_i v_c . x_i^2 n_c(i)."""
    onr = ((1, 1, 0, 0), (1, 0, 0, 1), (0, 0, 2, 1))
    u = ((0.0, 1.0, 2.0), (1.0, 0.0, 3.0), (2.0, 3.0, 1.5))
    h = _comp("ffb", onr, t=(1.0, 2.0, 0.5), u=u, dims=(2, 2), fold=(True, False))
    # self: 1.5 * (2*1)/2 = 1.5; f1.f2 = 1 (site 1) -> 1.0; f1.b = 0; f2.b = 1 (site 4) -> 3.0
    assert h.diagonal_element(h.start_key) == 1.5 + 1.0 + 3.0
    assert h.num_offdiagonals(h.start_key) == (2 + 2 + 2) * 4
    # hops are listed component by component (:383-391); every hop changes exactly one component
    for i, (k, v) in enumerate(h.offdiagonals(h.start_key)):
        if v == 0.0:
            continue
        new = h.unpack(k)
        changed = [c for c in range(3) if new[c] != onr[c]]
        assert changed == [i // 8]
        c = i // 8
        if c < 2:
            assert abs(v) == (1.0, 2.0)[c]  # fermions: -t * (+-1)
        else:  # bosons: -t sqrt(n_src (n_dst + 1)) (fockaddress.jl:559-567)
            (src,) = [j for j in range(4) if new[2][j] == onr[2][j] - 1]
            (dst,) = [j for j in range(4) if new[2][j] == onr[2][j] + 1]
            assert v == -0.5 * math.sqrt(onr[2][src] * (onr[2][dst] + 1))
    ht = _comp("ffb", onr, u=u, dims=(2, 2), trap=((1.0, 0.0), (0.0, 2.0), (0.5, 0.25)))
    # positions x = (-1, 0) per dimension, column-major sites: site1 (-1,-1), site2 (0,-1), site3 (-1,0), site4 (0,0)
    pot = (1.0 * 1 + 0.0) + (0.0 + 0.0) + (2.0 * 1 + 0.0) + 2 * (0.5 * 1 + 0.0) + 1 * 0.0
    assert math.isclose(ht.diagonal_element(ht.start_key), 5.5 + pot, rel_tol=1e-15)


def test_composite_pack_roundtrip_two_words():
    onr = (tuple(2 if i % 3 == 0 else 1 for i in range(27)), fermi_onr(27, (1, 5, 9, 14, 27)))
    h = _comp("bf", onr, dims=(3, 3, 3))
    assert h.W == 2 and h.unpack(h.start_key) == onr
    # component 0 occupies the low N + M - 1 bits in the BoseFS layout, component 1 the next M bits
    x = h.start_key[0] | (h.start_key[1] << 64)
    nb = sum(onr[0]) + 27 - 1
    assert bin(x & ((1 << nb) - 1)).count("1") == sum(onr[0]) and (x >> nb) == sum(1 << m for m in range(27) if onr[1][m])
