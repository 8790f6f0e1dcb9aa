"""ctypes front-end of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  The product (rimu.jl_b200) never does.

The oracle restates Rimu.jl v0.14.0's algorithm for the FCIQMC-step path in plain C on
occupation-number representations; see the header of oracle.c for the reference
file:line list.  Parity status: the reference is pure Julia and cannot run in this
image, so the oracle is pinned against the reference's own known answers
(tests/test_oracle_pins.py).

Table construction below (momentum grids, transcorrelated W(k)/u(k), trap potential)
restates the reference constructors:
  Hamiltonians/HubbardMom1D.jl:49-65, Transcorrelated1D.jl:72-89,113-179,
  HubbardRealSpace.jl:214-227.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
MAXM = 128

BOSE, FERMI, FERMI2C, COMPOSITE = 0, 1, 2, 3
MAXC = 4
HUBBARD_REAL_1D, HUBBARD_MOM_1D, HUBBARD_REAL_SPACE, TRANSCORRELATED_1D = 0, 1, 2, 3
HUBBARD_REAL_1D_EP, EXTENDED_HUBBARD_REAL_1D = 4, 5
EXTENDED_HUBBARD_MOM_1D, HUBBARD_MOM_1D_EP = 6, 7
BC_PERIODIC, BC_HARD_WALL, BC_TWISTED = 0, 1, 2
STYLE_DETERMINISTIC, STYLE_INTEGER, STYLE_SEMISTOCHASTIC, STYLE_WITH_THRESHOLD = 0, 1, 2, 3


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc if missing or stale."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Ham(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("addr_kind", C.c_int32), ("M", C.c_int32), ("ncomp", C.c_int32),
        ("N", C.c_int32 * 2),
        ("ndim", C.c_int32), ("dims", C.c_int32 * 3), ("fold", C.c_int32 * 3),
        ("cutoff", C.c_int32), ("three_body", C.c_int32), ("has_pot", C.c_int32), ("words", C.c_int32),
        ("u", C.c_double), ("t", C.c_double), ("v", C.c_double),
        ("tc", C.c_double * 2), ("umat", C.c_double * 4),
        ("ks", C.c_double * MAXM), ("kes", C.c_double * MAXM), ("ws", C.c_double * MAXM), ("us", C.c_double * MAXM),
        ("pot", C.c_double * (2 * MAXM)),
        ("ckind", C.c_int32 * MAXC), ("Nc", C.c_int32 * MAXC),
        ("tcs", C.c_double * MAXC), ("umats", C.c_double * (MAXC * MAXC)),
    ]


class StepParams(C.Structure):
    _fields_ = [
        ("style", C.c_int32), ("plain_h", C.c_int32),
        ("shift", C.c_double), ("dtau", C.c_double), ("boost", C.c_double),
        ("proj_threshold", C.c_double), ("rel_threshold", C.c_double), ("abs_threshold", C.c_double),
        ("compress_threshold", C.c_double),
        ("key", C.c_uint32 * 2),
        ("initiator_rule", C.c_int32), ("pad_", C.c_int32),
        ("initiator_threshold", C.c_double),
    ]


# InitiatorRule ids (DictVectors/initiators.jl:132-236)
NON_INITIATOR, INITIATOR, SIMPLE_INITIATOR, COHERENT_INITIATOR = 0, 1, 2, 3


class StepStats(C.Structure):
    _fields_ = [
        ("exact_steps", C.c_int64), ("inexact_steps", C.c_int64), ("spawn_attempts", C.c_int64),
        ("len_before", C.c_int64), ("len_after", C.c_int64),
        ("spawns", C.c_double), ("deaths", C.c_double), ("clones", C.c_double), ("zombies", C.c_double),
        ("norm1", C.c_double),
        ("ispawns", C.c_int64), ("ideaths", C.c_int64), ("iclones", C.c_int64), ("izombies", C.c_int64),
        ("inorm1", C.c_int64),
    ]

    def asdict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u64p, i64p, f64p, i32p, u32p = (C.POINTER(t) for t in (C.c_uint64, C.c_int64, C.c_double, C.c_int32, C.c_uint32))
        hp = C.POINTER(_Ham)
        L.orc_pack_onr.argtypes = [hp, i32p, u64p]
        L.orc_unpack_onr.argtypes = [hp, u64p, i32p]
        L.orc_diagonal.argtypes = [hp, u64p]; L.orc_diagonal.restype = C.c_double
        L.orc_num_offdiagonals.argtypes = [hp, u64p]; L.orc_num_offdiagonals.restype = C.c_long
        L.orc_offdiagonal.argtypes = [hp, u64p, C.c_long, u64p]; L.orc_offdiagonal.restype = C.c_double
        L.orc_philox.argtypes = [u32p, u32p, u32p]
        L.orc_addr_hash.argtypes = [u64p, C.c_int]; L.orc_addr_hash.restype = C.c_uint64
        L.orc_addr_owner.argtypes = [u64p, C.c_int, C.c_int]; L.orc_addr_owner.restype = C.c_int
        L.orc_step.argtypes = [hp, C.POINTER(StepParams), C.c_long, u64p, C.c_void_p, u64p, C.c_void_p, C.c_long, C.POINTER(StepStats)]
        L.orc_step.restype = C.c_long
        L.orc_step_threaded.argtypes = L.orc_step.argtypes + [C.c_int]
        L.orc_step_threaded.restype = C.c_long
        L.orc_annihilate.argtypes = [C.c_int, C.c_int, C.c_long, u64p, C.c_void_p, u64p, C.c_void_p, C.c_long]
        L.orc_annihilate.restype = C.c_long
        L.orc_bfs_basis.argtypes = [hp, u64p, C.c_long, u64p]; L.orc_bfs_basis.restype = C.c_long
        L.orc_coo_matrix.argtypes = [hp, C.c_long, u64p, C.c_long, i64p, i64p, f64p]; L.orc_coo_matrix.restype = C.c_long
        assert L.orc_sizeof_ham() == C.sizeof(_Ham)
        assert L.orc_sizeof_params() == C.sizeof(StepParams)
        assert L.orc_sizeof_stats() == C.sizeof(StepStats)
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# --------------------------------------------------------------------------- table builders
def mom1d_grid(M, t, dispersion="hubbard"):
    """HubbardMom1D.jl:55-63."""
    step = 2 * math.pi / M
    start = -math.pi * (1 + 1 / M) + step if M % 2 else -math.pi + step
    ks = np.array([start + i * step for i in range(M)])
    kes = -2 * t * np.cos(ks) if dispersion == "hubbard" else t * ks ** 2
    return ks, kes


def tc_w_function(n, nc):
    """Transcorrelated1D.jl:164-178."""
    prefactor = -1 / (8 * math.pi ** 2)
    n = abs(n)
    if n == 0:
        x = math.pi ** 2 / 6 - sum(1 / (q * q) for q in range(1, nc))
    elif 2 * nc > n > 0:
        x = 1 / n * sum(1 / q for q in range(nc, n + nc))
    else:
        x = 1 / n * sum(1 / q for q in range(nc, n + nc)) - 0.5 * sum(1 / (q * (n - q)) for q in range(nc, n - nc + 1))
    return prefactor * x


def tc_tables(M, t, cutoff):
    """Transcorrelated1D.jl:72-78,113-116,135-141."""
    i_to_n = lambda i: i - M // 2 - (M % 2)
    ks = np.array([i_to_n(i) * 2 * math.pi / M for i in range(1, M + 1)])
    kes = t * ks ** 2
    ws = np.array([tc_w_function(n, cutoff) for n in range(M)])
    us = np.array([(-1 / (2 * (n * 2 * math.pi / M))) if abs(n) >= cutoff else 0.0 for n in range(1, M + 1)])
    return ks, kes, ws, us


def trap_potential(dims, v):
    """HubbardRealSpace.jl:214-227: pot[c, site] = sum_d v[c,d] * x_d^2, x from -fld(M,2)."""
    dims = tuple(dims)
    v = np.atleast_2d(np.asarray(v, dtype=float))
    nsite = int(np.prod(dims))
    pot = np.zeros((v.shape[0], nsite))
    for site in range(nsite):
        idx, x2 = site, []
        for d in dims:
            x = idx % d - d // 2
            idx //= d
            x2.append(x * x)
        for c in range(v.shape[0]):
            pot[c, site] = sum(v[c, k] * x2[k] for k in range(len(dims)))
    return pot


def ep_lattice(M):
    """shift_lattice(range(-fld(M,2); length=M)) (HubbardReal1DEP.jl:9,55-56): positions with js[0] = 0."""
    is_ = list(range(-(M // 2), -(M // 2) + M))
    k = -(-M // 2)  # cld(M, 2)
    return is_[-k:] + is_[:-k]  # circshift(is, k)


def momentum_space_harmonic_potential(M, v):
    """HubbardMom1DEP.jl:14-31: 1/M * real(fft(v * j^2)) over the shifted lattice, symmetrised."""
    js = ep_lattice(M)
    mom = np.fft.fft(np.array([v * j * j for j in js], dtype=float))
    for i in range(1, M // 2 + 1):
        mom[M - i] = mom[i]
    return (1 / M) * np.real(mom)


class OracleHam:
    """A Hamiltonian of the oracle.

    model: 'HubbardReal1D' | 'HubbardMom1D' | 'HubbardRealSpace' | 'Transcorrelated1D'
    kind:  'bose' | 'fermi' | 'fermi2c' | 'comp:<letters>' -- a general CompositeFS, one letter per component:
           'b' BoseFS, 'f' FermiFS (e.g. 'comp:bf'); HubbardRealSpace only
    onr:   occupation numbers of the starting address (tuple, or tuple of two tuples)
    """

    MODELS = {"HubbardReal1D": HUBBARD_REAL_1D, "HubbardMom1D": HUBBARD_MOM_1D,
              "HubbardRealSpace": HUBBARD_REAL_SPACE, "Transcorrelated1D": TRANSCORRELATED_1D,
              "HubbardReal1DEP": HUBBARD_REAL_1D_EP, "ExtendedHubbardReal1D": EXTENDED_HUBBARD_REAL_1D,
              "ExtendedHubbardMom1D": EXTENDED_HUBBARD_MOM_1D, "HubbardMom1DEP": HUBBARD_MOM_1D_EP}
    KINDS = {"bose": BOSE, "fermi": FERMI, "fermi2c": FERMI2C}

    def __init__(self, model, kind, onr, u=1.0, t=1.0, v=1.0, dims=None, fold=None, trap=None,
                 cutoff=1, three_body_term=True, dispersion="hubbard", v_ho=1.0, boundary_condition="periodic"):
        h = _Ham()
        composite = kind.startswith("comp:")
        h.model, h.addr_kind = self.MODELS[model], COMPOSITE if composite else self.KINDS[kind]
        comps = [tuple(onr)] if kind in ("bose", "fermi") else [tuple(c) for c in onr]
        M = len(comps[0])
        assert all(len(c) == M for c in comps) and M <= MAXM
        h.M, h.ncomp = M, len(comps)
        if composite:
            letters = kind[5:]
            assert model == "HubbardRealSpace" and len(letters) == len(comps) and 2 <= len(comps) <= MAXC
            bits = 1  # one spare bit (empty-slot sentinel)
            for c, (letter, comp) in enumerate(zip(letters, comps)):
                h.ckind[c], h.Nc[c] = {"b": BOSE, "f": FERMI}[letter], sum(comp)
                bits += sum(comp) + M - 1 if letter == "b" else M
        else:
            for c, comp in enumerate(comps):
                h.N[c] = sum(comp)
            bits = (h.N[0] + M - 1) + 1 if kind == "bose" else M * len(comps)  # bosons keep one spare bit (empty-slot sentinel)
        h.words = (bits + 63) // 64
        assert h.words <= 2
        self.model, self.kind, self.M, self.W = model, kind, M, h.words
        ncomp = len(comps)
        if model == "HubbardRealSpace":
            dims = (M,) if dims is None else tuple(dims)
            fold = (True,) * len(dims) if fold is None else tuple(fold)
            assert int(np.prod(dims)) == M
            h.ndim = len(dims)
            for d in range(len(dims)):
                h.dims[d], h.fold[d] = dims[d], int(fold[d])
            tt = np.ones(ncomp) * np.asarray(t, dtype=float)
            uu = np.ones((ncomp, ncomp)) * np.asarray(u, dtype=float)
            for c in range(ncomp):
                if composite:
                    h.tcs[c] = tt[c]
                else:
                    h.tc[c] = tt[c]
                for c2 in range(ncomp):
                    if composite:
                        h.umats[c + ncomp * c2] = uu[c, c2]
                    else:
                        h.umat[c + 2 * c2] = uu[c, c2]
            if trap is not None and np.any(np.asarray(trap) != 0):
                pot = trap_potential(dims, np.asarray(trap, dtype=float).reshape(ncomp, len(dims)))
                h.has_pot = 1
                for c in range(ncomp):
                    for i in range(M):
                        h.pot[c * M + i] = pot[c, i]
        elif model == "HubbardMom1D":
            h.u, h.t = float(u), float(t)
            ks, kes = mom1d_grid(M, float(t), dispersion)
            for i in range(M):
                h.ks[i], h.kes[i] = ks[i], kes[i]
        elif model == "ExtendedHubbardMom1D":  # ExtendedHubbardMom1D.jl:37-57 (boundary_condition = 0)
            h.u, h.t, h.v = float(u), float(t), float(v)
            ks, kes = mom1d_grid(M, float(t), dispersion)
            step = 2 * math.pi / M
            for i in range(M):
                h.ks[i], h.kes[i] = ks[i], kes[i]
                h.ws[i] = math.cos(i * 2 * math.pi / M)   # cos(q * 2pi / M) of get_offdiagonal (:99-102), q = 0 .. M-1
                h.us[i] = math.cos(i * step)              # cos((mode_j - mode_i) * step) of the diagonal (excitations.jl:152)
        elif model == "HubbardMom1DEP":  # HubbardMom1DEP.jl:75-96
            h.u, h.t = float(u), float(t)
            ks, kes = mom1d_grid(M, float(t), dispersion)
            ep = momentum_space_harmonic_potential(M, float(v_ho))
            h.has_pot = 1
            for i in range(M):
                h.ks[i], h.kes[i], h.pot[i] = ks[i], kes[i], ep[i]
        elif model == "Transcorrelated1D":
            h.t, h.v, h.cutoff, h.three_body = float(t), float(v), int(cutoff), int(three_body_term)
            ks, kes, ws, us = tc_tables(M, float(t), int(cutoff))
            for i in range(M):
                h.ks[i], h.kes[i], h.ws[i], h.us[i] = ks[i], kes[i], ws[i], us[i]
        elif model == "HubbardReal1DEP":  # HubbardReal1DEP.jl:52-59: eps_i = v_ho * j_i^2, j = shift_lattice(-M//2 .. )
            h.u, h.t = float(u), float(t)
            js = ep_lattice(M)
            h.has_pot = 1
            for i in range(M):
                h.pot[i] = float(v_ho) * js[i] ** 2
        elif model == "ExtendedHubbardReal1D":  # ExtendedHubbardReal1D.jl:54-66
            h.u, h.t, h.v = float(u), float(t), float(v)
            h.fold[0] = {"periodic": BC_PERIODIC, "hard_wall": BC_HARD_WALL, "twisted": BC_TWISTED}[boundary_condition]
        else:
            h.u, h.t = float(u), float(t)
        self.h = h
        self.start_onr = comps
        self.start_key = self.pack(onr)

    # -- codec
    def pack(self, onr):
        comps = [tuple(onr)] if self.kind in ("bose", "fermi") else [tuple(c) for c in onr]
        flat = np.array([x for c in comps for x in c], dtype=np.int32)
        out = np.zeros(self.W, dtype=np.uint64)
        lib().orc_pack_onr(C.byref(self.h), _p(flat, C.c_int32), _p(out, C.c_uint64))
        return tuple(int(x) for x in out)

    def unpack(self, key):
        k = np.array(key, dtype=np.uint64).reshape(self.W)
        out = np.zeros(self.h.ncomp * self.M, dtype=np.int32)
        lib().orc_unpack_onr(C.byref(self.h), _p(k, C.c_uint64), _p(out, C.c_int32))
        if self.kind not in ("bose", "fermi"):
            return tuple(tuple(int(x) for x in out[c * self.M:(c + 1) * self.M]) for c in range(self.h.ncomp))
        return tuple(int(x) for x in out)

    def _key(self, key):
        return np.ascontiguousarray(np.array(key, dtype=np.uint64).reshape(self.W))

    # -- Hamiltonian interface (1-based `chosen` as in the reference)
    def diagonal_element(self, key):
        return lib().orc_diagonal(C.byref(self.h), _p(self._key(key), C.c_uint64))

    def num_offdiagonals(self, key):
        return lib().orc_num_offdiagonals(C.byref(self.h), _p(self._key(key), C.c_uint64))

    def get_offdiagonal(self, key, chosen):
        out = np.zeros(self.W, dtype=np.uint64)
        v = lib().orc_offdiagonal(C.byref(self.h), _p(self._key(key), C.c_uint64), chosen, _p(out, C.c_uint64))
        return tuple(int(x) for x in out), v

    def offdiagonals(self, key):
        return [self.get_offdiagonal(key, i) for i in range(1, self.num_offdiagonals(key) + 1)]

    # -- step
    def step(self, params: StepParams, keys, vals, threads=0):
        """apply_operator!: returns (keys_out[n,W] ascending, vals_out[n], StepStats)."""
        is_int = params.style == STYLE_INTEGER
        keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64).reshape(-1, self.W))
        vals = np.ascontiguousarray(np.asarray(vals, dtype=np.int64 if is_int else np.float64))
        n = keys.shape[0]
        cap = max(1024, 4 * n)
        st = StepStats()
        while True:
            ko = np.zeros((cap, self.W), dtype=np.uint64)
            vo = np.zeros(cap, dtype=vals.dtype)
            args = [C.byref(self.h), C.byref(params), n, _p(keys, C.c_uint64), vals.ctypes.data_as(C.c_void_p),
                    _p(ko, C.c_uint64), vo.ctypes.data_as(C.c_void_p), cap, C.byref(st)]
            r = lib().orc_step_threaded(*args, threads) if threads > 0 else lib().orc_step(*args)
            if r >= 0:
                return ko[:r].copy(), vo[:r].copy(), st
            cap = -r + 16

    def bfs_basis(self, start_key=None, max_dim=2_000_000):
        start = self._key(self.start_key if start_key is None else start_key)
        basis = np.zeros((max_dim, self.W), dtype=np.uint64)
        dim = lib().orc_bfs_basis(C.byref(self.h), _p(start, C.c_uint64), max_dim, _p(basis, C.c_uint64))
        if dim < 0:
            raise RuntimeError("basis larger than max_dim")
        return basis[:dim].copy()

    def sparse_matrix(self, basis):
        import scipy.sparse as sp
        basis = np.ascontiguousarray(basis, dtype=np.uint64)
        dim = basis.shape[0]
        cap = max(1024, dim * 8)
        while True:
            rows, cols = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
            vals = np.zeros(cap, dtype=np.float64)
            nnz = lib().orc_coo_matrix(C.byref(self.h), dim, _p(basis, C.c_uint64), cap,
                                       _p(rows, C.c_int64), _p(cols, C.c_int64), _p(vals, C.c_double))
            if nnz >= 0:
                break
            cap = -nnz
        return sp.coo_matrix((vals[:nnz], (rows[:nnz], cols[:nnz])), shape=(dim, dim)).tocsr()

    # -- dense-indexed sector (config 3): combinadic rank / unrank and sampled rows of y = H x
    def sector_dim(self):
        lib().orc_sector_dim.restype = C.c_long
        return int(lib().orc_sector_dim(C.byref(self.h)))

    def sector_rank(self, key):
        lib().orc_sector_rank.restype = C.c_long
        k = self._key(key)
        return int(lib().orc_sector_rank(C.byref(self.h), _p(k, C.c_uint64)))

    def sector_unrank(self, idx):
        k = np.zeros(self.W, dtype=np.uint64)
        lib().orc_sector_unrank(C.byref(self.h), C.c_long(int(idx)), _p(k, C.c_uint64))
        return k

    def sector_rows(self, idx, x):
        """y[idx] of y = H x for a dense sector vector x (host array of sector_dim doubles)"""
        idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int64))
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        out = np.zeros(len(idx), dtype=np.float64)
        lib().orc_sector_rows(C.byref(self.h), C.c_long(len(idx)), _p(idx, C.c_int64), _p(x, C.c_double), _p(out, C.c_double))
        return out

    def exact_eigenvalues(self, start_key=None, max_dim=200_000, hermitian=True):
        """All eigenvalues of H in the BFS-connected sector of the start address (dense)."""
        basis = self.bfs_basis(start_key, max_dim)
        H = self.sparse_matrix(basis).toarray()
        if hermitian:
            return np.linalg.eigvalsh(H)
        ev = np.linalg.eigvals(H)
        return ev[np.argsort(ev.real)]

    def exact_energy(self, start_key=None, max_dim=200_000, hermitian=True, overlap_tol=1e-10):
        """What the reference's `exact_energy` test helper returns (test/Hamiltonians.jl:13-17):
        the lowest eigenvalue whose eigenvector overlaps the single start determinant."""
        basis = self.bfs_basis(start_key, max_dim)
        H = self.sparse_matrix(basis).toarray()
        if hermitian:
            w, v = np.linalg.eigh(H)
            for i in range(len(w)):
                # degenerate blocks: test the projection of e_0 onto the eigenspace
                sel = np.abs(w - w[i]) < 1e-9
                if np.linalg.norm(v[0, sel]) > overlap_tol:
                    return w[i]
            raise RuntimeError("no overlap")
        w = np.linalg.eigvals(H)
        return np.sort(w.real)[0]


def make_params(style, shift=0.0, dtau=0.01, boost=1.0, plain_h=False, proj_threshold=0.0,
                rel_threshold=1.0, abs_threshold=math.inf, compress_threshold=0.0, key=(0, 0),
                initiator_rule=0, initiator_threshold=1.0) -> StepParams:
    p = StepParams()
    p.style, p.plain_h = style, int(plain_h)
    p.shift, p.dtau, p.boost = shift, dtau, boost
    p.proj_threshold, p.rel_threshold, p.abs_threshold = proj_threshold, rel_threshold, abs_threshold
    p.compress_threshold = compress_threshold
    p.key[0], p.key[1] = key
    p.initiator_rule, p.initiator_threshold = int(initiator_rule), float(initiator_threshold)
    return p


def annihilate(W, keys, vals):
    """Sum a spawn list by key, drop exact zeros, ascending key order."""
    vals = np.ascontiguousarray(vals)
    is_int = vals.dtype == np.int64
    keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64).reshape(-1, W))
    n = keys.shape[0]
    ko = np.zeros((max(n, 1), W), dtype=np.uint64)
    vo = np.zeros(max(n, 1), dtype=vals.dtype)
    r = lib().orc_annihilate(W, int(is_int), n, _p(keys, C.c_uint64), vals.ctypes.data_as(C.c_void_p),
                             _p(ko, C.c_uint64), vo.ctypes.data_as(C.c_void_p), max(n, 1))
    assert r >= 0
    return ko[:r].copy(), vo[:r].copy()


def philox(ctr, key):
    c = np.array(ctr, dtype=np.uint32); k = np.array(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(o, C.c_uint32))
    return tuple(int(x) for x in o)


def addr_hash(key):
    k = np.atleast_1d(np.array(key, dtype=np.uint64))
    return int(lib().orc_addr_hash(_p(k, C.c_uint64), len(k)))


def addr_owner(key, nranks):
    k = np.atleast_1d(np.array(key, dtype=np.uint64))
    return int(lib().orc_addr_owner(_p(k, C.c_uint64), len(k), nranks))


def step_key(seed: int, step: int):
    """Per-step Philox key = splitmix64(seed ^ splitmix64(step)); same derivation as the host
    driver in rimu.jl_b200 (restated here so the oracle does not import the product)."""
    def splitmix64(x):
        x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)
    k = splitmix64((seed & 0xFFFFFFFFFFFFFFFF) ^ splitmix64(step & 0xFFFFFFFFFFFFFFFF))
    return (k & 0xFFFFFFFF, k >> 32)
