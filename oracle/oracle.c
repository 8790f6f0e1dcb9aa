/*
 * oracle.c -- CPU restatement of the Rimu.jl FCIQMC step.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path in rimu.jl_b200/csrc.  It is NOT part
 * of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product path never calls into it.
 *
 * The reference (RimuQMC/Rimu.jl v0.14.0) is pure Julia and Julia is not installed in this
 * image, so the reference cannot be executed; this restatement is pinned against the
 * reference's own known answers instead (tests/test_oracle_pins.py: golden energies,
 * doctest vectors and hand-computed step statistics listed in SURVEY.md Appendix B).
 *
 * Design: everything is written in occupation-number representation (ONR), the way the
 * reference's own slow test helpers are (test/excitation_tests.jl:32-69), NOT with the
 * bit tricks the CUDA kernels use -- so oracle and kernels are independent derivations
 * of the same semantics and only meet at the packed-key interchange format.
 *
 * Indices `chosen` are 1-based exactly as in the reference.
 *
 * Reference files followed (path:line under the Rimu.jl source tree):
 *   BitStringAddresses/bitstring.jl:464-545,713-792   bit layouts of BoseFS / FermiFS
 *   BitStringAddresses/fockaddress.jl:258-275,559-567  OccupiedModeMap, excitation value
 *   Hamiltonians/HubbardReal1D.jl:51-62, bosefs.jl:297-345
 *   Hamiltonians/HubbardMom1D.jl:49-65,131-205, excitations.jl:26-137,199-238
 *   Hamiltonians/HubbardRealSpace.jl:18-106,279-391, geometry.jl:127-139,161-181,232-235
 *   Hamiltonians/Transcorrelated1D.jl:72-89,113-387
 *   fciqmc.jl:78-112  FirstOrderTransitionOperator
 *   StochasticStyles/spawning.jl:9-93,152-243,358-385, styles.jl:11-25,76-105,175-214,
 *   StochasticStyles/compression.jl:8-42
 *   Interfaces/dictvectors.jl:112-140, DictVectors/pdworkingmemory.jl:21-31,191-309
 *   DictVectors/initiators.jl:22-45,132-236  InitiatorValue lanes, to_/from_initiator_value
 *   Hamiltonians/HubbardReal1DEP.jl:9,47-92, ExtendedHubbardReal1D.jl:30-135, bosefs.jl:355-369
 *
 * RNG: the reference draws from Julia's task-local Xoshiro256++; that stream cannot be matched ("parity unpinned" for the
 * random stream, by design).  Oracle and kernels share a Philox4x32-10 function of (step key, address hash, attempt
 * index, stream id), pinned by Random123 known answers, so THEIR parity is bit-exact; parity with the reference for
 * stochastic runs is distributional (blocking analysis, expectation-equality tests).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXM 128

#define ORC_MAXC 4
enum { ORC_BOSE = 0, ORC_FERMI = 1, ORC_FERMI2C = 2,
       ORC_COMPOSITE = 3 }; /* CompositeFS of 2..ORC_MAXC BoseFS / FermiFS components (multicomponent.jl:10-34), HubbardRealSpace */
enum { ORC_HUBBARD_REAL_1D = 0, ORC_HUBBARD_MOM_1D = 1, ORC_HUBBARD_REAL_SPACE = 2,
       ORC_TRANSCORRELATED_1D = 3, ORC_HUBBARD_REAL_1D_EP = 4, ORC_EXTENDED_HUBBARD_REAL_1D = 5,
       ORC_EXTENDED_HUBBARD_MOM_1D = 6, /* ExtendedHubbardMom1D.jl:37-117 (bosons): ws[q] = cos(q*2pi/M), us[d] = cos(d*(2pi/M)) */
       ORC_HUBBARD_MOM_1D_EP = 7 };     /* HubbardMom1DEP.jl:68-257: pot[k] = ep[k+1], the momentum-space harmonic potential */
/* ExtendedHubbardReal1D boundary_condition (ExtendedHubbardReal1D.jl:12-22), stored in orc_ham.fold[0] */
enum { ORC_BC_PERIODIC = 0, ORC_BC_HARD_WALL = 1, ORC_BC_TWISTED = 2 };
enum { ORC_STYLE_DETERMINISTIC = 0, ORC_STYLE_INTEGER = 1, ORC_STYLE_SEMISTOCHASTIC = 2,
       ORC_STYLE_WITH_THRESHOLD = 3 };

typedef struct {
    int32_t model, addr_kind, M, ncomp;
    int32_t N[2];
    int32_t ndim, dims[3], fold[3];
    int32_t cutoff, three_body, has_pot, words;
    double u, t, v;
    double tc[2], umat[4]; /* umat[i + 2*j] = u[i,j], column major like Julia */
    double ks[ORC_MAXM], kes[ORC_MAXM], ws[ORC_MAXM], us[ORC_MAXM];
    double pot[2 * ORC_MAXM]; /* pot[c*M + site] */
    /* ORC_COMPOSITE: kind (ORC_BOSE / ORC_FERMI) and particle number of every component, t[c], u[i + ncomp*j] */
    int32_t ckind[ORC_MAXC], Nc[ORC_MAXC];
    double tcs[ORC_MAXC], umats[ORC_MAXC * ORC_MAXC];
} orc_ham;

typedef struct { int n[ORC_MAXC][ORC_MAXM]; } orc_onr;
static void onr_copy(const orc_ham *h, orc_onr *dst, const orc_onr *src) { memcpy(dst, src, sizeof(int) * ORC_MAXM * h->ncomp); }
typedef struct { int len; int occ[ORC_MAXM]; int mode[ORC_MAXM]; } orc_map;

/* ------------------------------------------------------------------ helpers */
static int mod1i(int x, int y) { int r = (x - 1) % y; if (r < 0) r += y; return r + 1; }
/* fldmod1 for x >= 1 (Julia Base.fldmod1) */
static void fldmod1i(long x, long y, long *f, long *m) { *f = (x - 1) / y + 1; *m = (x - 1) % y + 1; }

/* OccupiedModeMap (fockaddress.jl:258-275): occupied modes in ascending mode order */
static void build_map(const int *n, int M, orc_map *mp) {
    mp->len = 0;
    for (int i = 0; i < M; i++)
        if (n[i] > 0) { mp->occ[mp->len] = n[i]; mp->mode[mp->len] = i + 1; mp->len++; }
}

/* bosonic excitation on an ONR (test/excitation_tests.jl:32-49 ==
 * fockaddress.jl:559-567 bose_excitation_value): destructions first, each in reverse
 * tuple order.  n is modified only when the move is legal.  Returns sqrt(value). */
static double bose_excite(int *n, int M, const int *cre, const int *des, int k) {
    int tmp[ORC_MAXM];
    memcpy(tmp, n, sizeof(int) * M);
    long value = 1;
    for (int j = k - 1; j >= 0; j--) { int d = des[j] - 1; value *= tmp[d]; tmp[d]--; }
    for (int j = k - 1; j >= 0; j--) { int c = cre[j] - 1; tmp[c]++; value *= tmp[c]; }
    if (value == 0) return 0.0;
    memcpy(n, tmp, sizeof(int) * M);
    return sqrt((double)value);
}

/* fermionic excitation on an ONR (test/excitation_tests.jl:50-69 ==
 * bitstring.jl:773-792 fermi_excitation). */
static double fermi_excite(int *n, int M, const int *cre, const int *des, int k) {
    int tmp[ORC_MAXM];
    memcpy(tmp, n, sizeof(int) * M);
    int num = 0;
    for (int j = k - 1; j >= 0; j--) {
        int d = des[j] - 1;
        for (int q = 0; q < d; q++) num += tmp[q];
        if (tmp[d] == 0) return 0.0;
        tmp[d]--;
    }
    for (int j = k - 1; j >= 0; j--) {
        int c = cre[j] - 1;
        for (int q = 0; q < c; q++) num += tmp[q];
        if (tmp[c] != 0) return 0.0;
        tmp[c]++;
    }
    memcpy(n, tmp, sizeof(int) * M);
    return (num % 2 == 0) ? 1.0 : -1.0;
}

static double excite(int kind_is_bose, int *n, int M, const int *cre, const int *des, int k) {
    return kind_is_bose ? bose_excite(n, M, cre, des, k) : fermi_excite(n, M, cre, des, k);
}

/* ------------------------------------------------------------------ key codec
 * Device interchange layout: W little-endian uint64 words, word 0 least significant.
 * BoseFS (bitstring.jl:464-472): mode 1 in the lowest bits, n ones then a 0 separator.
 * FermiFS (bitstring.jl:713-723): bit m-1 <-> mode m.
 * Two fermion components: component c occupies bits [c*M, (c+1)*M).
 * General CompositeFS: the components' bit strings side by side from the low bits (BoseFS: N_c + M - 1 bits, FermiFS: M). */
static void setbit(uint64_t *w, int pos) { w[pos >> 6] |= (uint64_t)1 << (pos & 63); }
static int getbit(const uint64_t *w, int pos) { return (int)((w[pos >> 6] >> (pos & 63)) & 1); }

void orc_pack(const orc_ham *h, const orc_onr *o, uint64_t *w) {
    for (int j = 0; j < h->words; j++) w[j] = 0;
    if (h->addr_kind == ORC_COMPOSITE) {
        int base = 0;
        for (int c = 0; c < h->ncomp; c++) {
            if (h->ckind[c] == ORC_BOSE) {
                int pos = base;
                for (int m = 0; m < h->M; m++) {
                    for (int q = 0; q < o->n[c][m]; q++) setbit(w, pos++);
                    pos++;
                }
                base += h->Nc[c] + h->M - 1;
            } else {
                for (int m = 0; m < h->M; m++) if (o->n[c][m]) setbit(w, base + m);
                base += h->M;
            }
        }
    } else if (h->addr_kind == ORC_BOSE) {
        int pos = 0;
        for (int m = 0; m < h->M; m++) {
            for (int q = 0; q < o->n[0][m]; q++) setbit(w, pos++);
            pos++;
        }
    } else {
        for (int c = 0; c < h->ncomp; c++)
            for (int m = 0; m < h->M; m++)
                if (o->n[c][m]) setbit(w, c * h->M + m);
    }
}

void orc_unpack(const orc_ham *h, const uint64_t *w, orc_onr *o) {
    memset(o, 0, sizeof(int) * ORC_MAXM * h->ncomp);
    if (h->addr_kind == ORC_COMPOSITE) {
        int base = 0;
        for (int c = 0; c < h->ncomp; c++) {
            if (h->ckind[c] == ORC_BOSE) {
                int B = h->Nc[c] + h->M - 1, mode = 0;
                for (int pos = 0; pos < B; pos++) {
                    if (getbit(w, base + pos)) o->n[c][mode]++;
                    else mode++;
                }
                base += B;
            } else {
                for (int m = 0; m < h->M; m++) o->n[c][m] = getbit(w, base + m);
                base += h->M;
            }
        }
    } else if (h->addr_kind == ORC_BOSE) {
        int B = h->N[0] + h->M - 1, mode = 0;
        for (int pos = 0; pos < B; pos++) {
            if (getbit(w, pos)) o->n[0][mode]++;
            else mode++;
        }
    } else {
        for (int c = 0; c < h->ncomp; c++)
            for (int m = 0; m < h->M; m++) o->n[c][m] = getbit(w, c * h->M + m);
    }
}

/* flat int32 ONR <-> key, for python */
void orc_pack_onr(const orc_ham *h, const int32_t *onr, uint64_t *w) {
    orc_onr o; memset(&o, 0, sizeof(o));
    for (int c = 0; c < h->ncomp; c++) for (int m = 0; m < h->M; m++) o.n[c][m] = onr[c * h->M + m];
    orc_pack(h, &o, w);
}
void orc_unpack_onr(const orc_ham *h, const uint64_t *w, int32_t *onr) {
    orc_onr o; orc_unpack(h, w, &o);
    for (int c = 0; c < h->ncomp; c++) for (int m = 0; m < h->M; m++) onr[c * h->M + m] = o.n[c][m];
}

/* ------------------------------------------------------------------ geometry
 * geometry.jl:127-139,161-175,232-235: column-major site index; direction k<=D is +e_k,
 * k>D is -e_{k-D}; folded if periodic else 0 when leaving the grid. */
static int neighbor_site(const orc_ham *h, int mode, int chosen) {
    int D = h->ndim, idx = mode - 1, x[3];
    for (int d = 0; d < D; d++) { x[d] = idx % h->dims[d] + 1; idx /= h->dims[d]; }
    if (chosen <= D) x[chosen - 1] += 1; else x[chosen - D - 1] -= 1;
    for (int d = 0; d < D; d++) {
        if (h->fold[d]) x[d] = mod1i(x[d], h->dims[d]);
        else if (x[d] < 1 || x[d] > h->dims[d]) return 0;
    }
    int lin = 0, stride = 1;
    for (int d = 0; d < D; d++) { lin += (x[d] - 1) * stride; stride *= h->dims[d]; }
    return lin + 1;
}

/* ------------------------------------------------------------------ Transcorrelated1D helpers
 * Transcorrelated1D.jl:113-116,135-149,164-179,196-224 */
static double tc_n_to_k(int n, int M) { return n * 2.0 * M_PI / M; }
static double tc_corr(const orc_ham *h, int n) {
    int a = n < 0 ? -n : n;
    if (a == 0) return 0.0;
    return (n > 0 ? 1.0 : -1.0) * h->us[a - 1];
}
static double tc_w(const orc_ham *h, int n) { return h->ws[(n < 0 ? -n : n)]; }
static double tc_t_function(const orc_ham *h, int p, int q, int k) {
    int M = h->M;
    double k_pi = tc_n_to_k(k, M), pmq_pi = tc_n_to_k(p - q, M), cor_k = tc_corr(h, k);
    return h->v / M + 2 * h->v / M * (cor_k * k_pi - cor_k * pmq_pi) + 2 * h->v * h->v / h->t * tc_w(h, k);
}
static double tc_q_function(const orc_ham *h, int k, int l) {
    int M = h->M;
    return -(h->v * h->v) / (h->t * ((double)M * M)) * tc_corr(h, k) * tc_corr(h, l);
}

/* ------------------------------------------------------------------ diagonal elements */
static long bose_interaction(const int *n, int M) { /* bosefs.jl:400-428 */
    long r = 0;
    for (int i = 0; i < M; i++) r += (long)n[i] * (n[i] - 1);
    return r;
}

static double tc_three_body_diag(const orc_ham *h, const orc_map *m1, const orc_map *m2) {
    /* Transcorrelated1D.jl:236-246 transcorrelated_diagonal */
    double value = 0.0;
    for (int p = 0; p < m1->len; p++)
        for (int q = 0; q < p; q++) {
            int k = m1->mode[p] - m1->mode[q];
            double qkk = tc_q_function(h, -k, k);
            value += 2 * qkk * m2->len;
        }
    return value;
}

double orc_diagonal_onr(const orc_ham *h, const orc_onr *o) {
    int M = h->M;
    switch (h->model) {
    case ORC_HUBBARD_REAL_1D: /* HubbardReal1D.jl:55-57 */
        return h->u * (double)bose_interaction(o->n[0], M) / 2;
    case ORC_HUBBARD_REAL_1D_EP: { /* HubbardReal1DEP.jl:82-87: sum over occupied modes of u n (n-1) / 2 + ep[mode] n */
        double s = 0.0; int first = 1;
        for (int m = 0; m < M; m++) {
            int n = o->n[0][m];
            if (!n) continue;
            double term = h->u * n * (n - 1) / 2 + h->pot[m] * n;
            s = first ? term : s + term; first = 0;
        }
        return s;
    }
    case ORC_EXTENDED_HUBBARD_REAL_1D: { /* ExtendedHubbardReal1D.jl:101-126 */
        orc_map mp; build_map(o->n[0], M, &mp);
        long ext = 0, reg = 0; int pmode = 0, pocc = 0;
        for (int i = 0; i < mp.len; i++) {
            if (pmode == mp.mode[i] - 1) ext += (long)pocc * mp.occ[i]; /* prev starts as the zero index (mode 0, occ 0) */
            reg += (long)mp.occ[i] * (mp.occ[i] - 1);
            pmode = mp.mode[i]; pocc = mp.occ[i];
        }
        if (h->fold[0] != ORC_BC_HARD_WALL && mp.len > 0) {
            long last = mp.mode[mp.len - 1] == M ? mp.occ[mp.len - 1] : 0;
            long firstn = mp.mode[0] == 1 ? mp.occ[0] : 0;
            ext += last * firstn;
        }
        return h->u * (double)reg / 2 + h->v * (double)ext;
    }
    case ORC_HUBBARD_MOM_1D: case ORC_EXTENDED_HUBBARD_MOM_1D: case ORC_HUBBARD_MOM_1D_EP: {
        /* HubbardMom1D.jl:163-181, excitations.jl:126-160; ExtendedHubbardMom1D.jl:85-89 + excitations.jl:145-156;
         * HubbardMom1DEP.jl:151-166 + excitations.jl:274-279 */
        orc_map ma; build_map(o->n[0], M, &ma);
        if (h->addr_kind == ORC_BOSE) {
            double ke = 0.0;
            for (int i = 0; i < ma.len; i++) ke += h->kes[ma.mode[i] - 1] * ma.occ[i];
            long onproduct = 0;
            for (int i = 0; i < ma.len; i++) {
                onproduct += (long)ma.occ[i] * (ma.occ[i] - 1);
                for (int j = 0; j < i; j++) onproduct += 4L * ma.occ[i] * ma.occ[j];
            }
            double value = ke + h->u / (2 * M) * (double)onproduct;
            if (h->model == ORC_EXTENDED_HUBBARD_MOM_1D) { /* + (v / M) * extended_momentum_transfer_diagonal(map, 2pi / M) */
                double ext = 0.0;
                for (int i = 0; i < ma.len; i++) {
                    ext += (double)((long)ma.occ[i] * (ma.occ[i] - 1));
                    for (int j = 0; j < i; j++) {
                        int d = ma.mode[i] - ma.mode[j]; /* cos((mode_j - mode_i) * step) = cos(|d| * step): us[|d|] */
                        ext += (double)(2L * ma.occ[i] * ma.occ[j]) * (1 + h->us[d]);
                    }
                }
                value += (h->v / M) * ext;
            } else if (h->model == ORC_HUBBARD_MOM_1D_EP) {
                long n = 0;
                for (int i = 0; i < ma.len; i++) n += ma.occ[i];
                value += (double)n * h->pot[0];
            }
            return value;
        } else {
            orc_map mb; build_map(o->n[1], M, &mb);
            double ka = 0.0, kb = 0.0;
            for (int i = 0; i < ma.len; i++) ka += h->kes[ma.mode[i] - 1] * ma.occ[i];
            for (int i = 0; i < mb.len; i++) kb += h->kes[mb.mode[i] - 1] * mb.occ[i];
            double value = ka + kb + h->u / (2 * M) * (double)(2 * ma.len * mb.len);
            if (h->model == ORC_HUBBARD_MOM_1D_EP) value = value + (double)ma.len * h->pot[0] + (double)mb.len * h->pot[0];
            return value;
        }
    }
    case ORC_HUBBARD_REAL_SPACE: { /* HubbardRealSpace.jl:18-75,90-106,279-293 */
        double interaction = 0.0;
        int C = h->ncomp;
        if (h->addr_kind == ORC_COMPOSITE) {
            /* local_interaction(::CompositeFS, u) = _interactions(components, u) (:18-75), recursive:
             *   _interactions((a, as...), m) = self(a, m[1,1]) + _interaction_col(a, as, m[2:N,1]) + _interactions(as, m[2:N,2:N])
             *   _interaction_col(a, (b, bs...), (u, us...)) = u * dot(occupied_modes(a), occupied_modes(b)) + _interaction_col(a, bs, us)
             * self = u * sum n(n-1) / 2 for a BoseFS, 0 for a FermiFS; both recursions bottom out in 0 */
            double rest = 0.0;
            for (int i = C - 1; i >= 0; i--) {
                double row = 0.0;
                for (int j = C - 1; j > i; j--) {
                    long dot = 0;
                    for (int m = 0; m < M; m++) dot += (long)o->n[i][m] * o->n[j][m];
                    row = h->umats[j + C * i] * (double)dot + row;
                }
                double self = h->ckind[i] == ORC_BOSE ? h->umats[i + C * i] * (double)bose_interaction(o->n[i], M) / 2 : 0.0;
                rest = (self + row) + rest;
            }
            interaction = rest;
            int allz = 1;
            for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) if (h->umats[i + C * j] != 0.0) allz = 0;
            if (allz) interaction = 0.0; /* u_mat = iszero(u) ? nothing : ... (:205) */
            double potc = 0.0;
            if (h->has_pot)
                for (int c = 0; c < C; c++) {
                    double pe = 0.0;
                    for (int i = 0; i < M; i++) if (o->n[c][i]) pe += o->n[c][i] * h->pot[c * M + i];
                    potc += pe;
                }
            return interaction + potc;
        }
        if (C == 1) {
            if (h->addr_kind == ORC_BOSE) interaction = h->umat[0] * (double)bose_interaction(o->n[0], M) / 2;
        } else {
            /* _interactions: (self_1 + u[2,1]*cross) + (self_2 + 0) + 0.0 */
            double self1 = 0.0, self2 = 0.0; /* fermion components do not self-interact */
            long cross = 0;
            for (int i = 0; i < M; i++) cross += (long)o->n[0][i] * o->n[1][i];
            interaction = (self1 + h->umat[1] * (double)cross) + (self2 + 0.0);
        }
        int allzero = 1;
        for (int i = 0; i < C * C; i++) if (h->umat[i] != 0.0) allzero = 0;
        if (allzero) interaction = 0.0;
        double pot = 0.0;
        if (h->has_pot)
            for (int c = 0; c < C; c++) {
                double pe = 0.0;
                for (int i = 0; i < M; i++) if (o->n[c][i]) pe += o->n[c][i] * h->pot[c * M + i];
                pot += pe;
            }
        return interaction + pot;
    }
    case ORC_TRANSCORRELATED_1D: { /* Transcorrelated1D.jl:228-267 */
        orc_map m1, m2; build_map(o->n[0], M, &m1); build_map(o->n[1], M, &m2);
        double k1 = 0.0, k2 = 0.0;
        for (int i = 0; i < m1.len; i++) k1 += h->kes[m1.mode[i] - 1] * m1.occ[i];
        for (int i = 0; i < m2.len; i++) k2 += h->kes[m2.mode[i] - 1] * m2.occ[i];
        double mtd = (double)(2 * m1.len * m2.len) * (h->v / M + 2 * h->v * h->v / h->t * tc_w(h, 0)) / 2;
        double value = k1 + k2 + mtd;
        if (h->three_body) value += tc_three_body_diag(h, &m1, &m2) + tc_three_body_diag(h, &m2, &m1);
        return value;
    }
    }
    return NAN;
}

/* ------------------------------------------------------------------ num_offdiagonals */
long orc_num_offdiagonals_onr(const orc_ham *h, const orc_onr *o) {
    int M = h->M;
    orc_map ma, mb; build_map(o->n[0], M, &ma);
    if (h->ncomp == 2) build_map(o->n[1], M, &mb); else mb.len = 0;
    switch (h->model) {
    case ORC_HUBBARD_REAL_1D: case ORC_HUBBARD_REAL_1D_EP: case ORC_EXTENDED_HUBBARD_REAL_1D:
        return 2L * ma.len; /* HubbardReal1D.jl:51-53, HubbardReal1DEP.jl:78-80, ExtendedHubbardReal1D.jl:88-90 */
    case ORC_HUBBARD_MOM_1D: case ORC_EXTENDED_HUBBARD_MOM_1D: case ORC_HUBBARD_MOM_1D_EP: {
        /* HubbardMom1D.jl:131-144; ExtendedHubbardMom1D.jl:75-83; HubbardMom1DEP.jl:178-186,223-235 */
        long ep = h->model == ORC_HUBBARD_MOM_1D_EP ? (long)(ma.len + mb.len) * (M - 1) : 0;
        if (h->addr_kind == ORC_BOSE) {
            long s = ma.len, d = 0;
            for (int i = 0; i < ma.len; i++) d += ma.occ[i] > 1;
            return s * (s - 1) * (M - 2) + d * (M - 1) + ep;
        } else if (h->addr_kind == ORC_FERMI2C)
            return (long)h->N[0] * h->N[1] * (M - 1) + ep;
        return 0;
    }
    case ORC_HUBBARD_REAL_SPACE: /* HubbardRealSpace.jl:309-315,371-374 */
        if (h->addr_kind == ORC_COMPOSITE) {
            long s = 0;
            for (int c = 0; c < h->ncomp; c++) { orc_map mc; build_map(o->n[c], M, &mc); s += mc.len; }
            return s * 2 * h->ndim;
        }
        return (long)(ma.len + mb.len) * 2 * h->ndim;
    case ORC_TRANSCORRELATED_1D: { /* Transcorrelated1D.jl:277-297 */
        long N1 = ma.len, N2 = mb.len;
        long n_mom = N1 * N2 * (M - 1);
        long n1 = h->three_body ? N1 * (N1 - 1) * N2 * M * M : 0;
        long n2 = h->three_body ? N2 * (N2 - 1) * N1 * M * M : 0;
        return n_mom + n1 + n2;
    }
    }
    return -1;
}

/* ------------------------------------------------------------------ off-diagonal elements */

/* excitations.jl:82-119: two-component momentum transfer.  a,b: ONRs (modified when value != 0).
 * Returns val_a*val_b; writes p,q,-k into prm. */
static double mom_transfer_2c(int M, int *na, int *nb, const orc_map *ma, const orc_map *mb,
                              long chosen, int fold, int prm[3]) {
    long src_a, rem, dst_a, src_b;
    fldmod1i(chosen, (long)(M - 1) * mb->len, &src_a, &rem);
    fldmod1i(rem, mb->len, &dst_a, &src_b);
    int src_a_mode = ma->mode[src_a - 1], src_b_mode = mb->mode[src_b - 1];
    if (dst_a >= src_a_mode) dst_a += 1;
    int mom_change = (int)dst_a - src_a_mode;
    int dst_b = src_b_mode - mom_change;
    prm[0] = src_a_mode; prm[1] = src_b_mode; prm[2] = -mom_change;
    int da = (int)dst_a;
    if (fold) { da = mod1i(da, M); dst_b = mod1i(dst_b, M); }
    else if (!(0 < da && da <= M) || !(0 < dst_b && dst_b <= M)) return 0.0;
    int ta[ORC_MAXM], tb[ORC_MAXM];
    memcpy(ta, na, sizeof(int) * M); memcpy(tb, nb, sizeof(int) * M);
    int c1[1] = {da}, d1[1] = {src_a_mode}, c2[1] = {dst_b}, d2[1] = {src_b_mode};
    double va = fermi_excite(ta, M, c1, d1, 1);
    double vb = fermi_excite(tb, M, c2, d2, 1);
    /* the reference returns both (possibly unchanged) addresses and val_a*val_b; callers
     * only use the new address when the product is non-zero */
    if (va * vb != 0.0) { memcpy(na, ta, sizeof(int) * M); memcpy(nb, tb, sizeof(int) * M); }
    return va * vb;
}

/* excitations.jl:199-238 for fermions: (p,q,s,p_k,q_l), first index fastest */
static double tc_three_body(int M, int *na, int *nb, const orc_map *ma, const orc_map *mb,
                            long i, int *k_out, int *l_out) {
    long N1 = ma->len, N2 = mb->len;
    long idx = i - 1;
    long p = idx % N1 + 1; idx /= N1;
    long q = idx % (N1 - 1) + 1; idx /= (N1 - 1);
    long s = idx % N2 + 1; idx /= N2;
    long p_k = idx % M + 1; idx /= M;
    long q_l = idx % M + 1;
    if (q >= p) q += 1;
    int pm = ma->mode[p - 1], qm = ma->mode[q - 1], sm = mb->mode[s - 1];
    int k = pm - (int)p_k, l = (int)q_l - qm, s_kl = sm + k - l;
    *k_out = k; *l_out = l;
    if (k == 0 || l == 0) return 0.0;
    if (pm == q_l && qm == p_k) return 0.0;
    if (s_kl > M || s_kl < 1) return 0.0;
    int ta[ORC_MAXM], tb[ORC_MAXM];
    memcpy(ta, na, sizeof(int) * M); memcpy(tb, nb, sizeof(int) * M);
    int cre[2] = {(int)p_k, (int)q_l}, des[2] = {qm, pm};
    double v1 = fermi_excite(ta, M, cre, des, 2);
    int c2[1] = {s_kl}, d2[1] = {sm};
    double v2 = fermi_excite(tb, M, c2, d2, 1);
    if (v1 * v2 != 0.0) { memcpy(na, ta, sizeof(int) * M); memcpy(nb, tb, sizeof(int) * M); }
    return v1 * v2;
}

/* get_offdiagonal(h, address, chosen) -> value, new address written to `out`.
 * When the value is 0 the returned address equals the input (as the reference does). */
double orc_offdiagonal_onr(const orc_ham *h, const orc_onr *in, long chosen, orc_onr *out) {
    int M = h->M;
    onr_copy(h, out, in);
    orc_map ma, mb; build_map(in->n[0], M, &ma);
    if (h->ncomp == 2) build_map(in->n[1], M, &mb); else mb.len = 0;
    switch (h->model) {
    case ORC_HUBBARD_REAL_1D: case ORC_HUBBARD_REAL_1D_EP: case ORC_EXTENDED_HUBBARD_REAL_1D: {
        /* bosefs.jl:270-274,347-369; HubbardReal1D.jl:59-62, HubbardReal1DEP.jl:89-92, ExtendedHubbardReal1D.jl:128-135 */
        int site = (int)((chosen + 1) >> 1);
        int src = ma.mode[site - 1];
        int dir = (chosen & 1) ? 1 : -1;
        int dst = mod1i(src + dir, M);
        int cre[1] = {dst}, des[1] = {src};
        double val = bose_excite(out->n[0], M, cre, des, 1);
        if (h->model == ORC_EXTENDED_HUBBARD_REAL_1D) {
            int on_boundary = (src == 1 && dir == -1) || (src == M && dir == 1);
            if (on_boundary && h->fold[0] == ORC_BC_TWISTED) val = -val;
            else if (on_boundary && h->fold[0] == ORC_BC_HARD_WALL) val = 0.0;
        }
        return -h->t * val;
    }
    case ORC_HUBBARD_MOM_1D: case ORC_EXTENDED_HUBBARD_MOM_1D: case ORC_HUBBARD_MOM_1D_EP: {
        if (h->model == ORC_HUBBARD_MOM_1D_EP) { /* the external-potential block after the momentum-transfer block */
            long n_mom;
            if (h->addr_kind == ORC_BOSE) {
                long s = ma.len, d = 0;
                for (int i = 0; i < ma.len; i++) d += ma.occ[i] > 1;
                n_mom = s * (s - 1) * (M - 2) + d * (M - 1);
            } else n_mom = (long)ma.len * mb.len * (M - 1);
            if (chosen > n_mom) { /* momentum_external_potential_excitation, excitations.jl:257-267 */
                long i = chosen - n_mom;
                int comp = 0;
                if (h->addr_kind != ORC_BOSE && i > (long)ma.len * (M - 1)) { i -= (long)ma.len * (M - 1); comp = 1; }
                const orc_map *mp = comp ? &mb : &ma;
                long p, q;
                fldmod1i(i, M - 1, &p, &q);
                int pmode = mp->mode[p - 1];
                if (q >= pmode) q += 1;            /* leave out the diagonal term */
                int k = pmode - (int)q;            /* change in momentum */
                int km = ((k % M) + M) % M;
                double factor = h->pot[km];
                int cre[1] = {(int)q}, des[1] = {pmode};
                double val = excite(h->addr_kind == ORC_BOSE, out->n[comp], M, cre, des, 1);
                return val * factor;
            }
        }
        if (h->addr_kind == ORC_BOSE) { /* excitations.jl:26-80; HubbardMom1D.jl:182-188 */
            long singlies = ma.len;
            long dbl = chosen - singlies * (singlies - 1) * (M - 2);
            int src[2], dst[2];
            long mom_change;
            if (dbl > 0) {
                long d;
                fldmod1i(dbl, M - 1, &d, &mom_change);
                int idx = 0;
                for (int i = 0; i < ma.len; i++) {
                    d -= ma.occ[i] >= 2;
                    if (d == 0) { idx = i; break; }
                }
                src[0] = src[1] = ma.mode[idx];
            } else {
                long pair, fst, snd, f_hole, s_hole;
                fldmod1i(chosen, M - 2, &pair, &mom_change);
                fldmod1i(pair, singlies - 1, &fst, &snd);
                if (snd < fst) { f_hole = snd; s_hole = fst; }
                else { f_hole = fst; s_hole = snd + 1; }
                src[0] = ma.mode[f_hole - 1]; src[1] = ma.mode[s_hole - 1];
                if (mom_change >= src[1] - src[0]) mom_change += 1;
            }
            dst[0] = mod1i(src[0] + (int)mom_change, M);
            dst[1] = mod1i(src[1] - (int)mom_change, M);
            double val = bose_excite(out->n[0], M, dst, src, 2);
            if (h->model == ORC_EXTENDED_HUBBARD_MOM_1D) /* ExtendedHubbardMom1D.jl:99-102: q = -mom_change, cos even */
                return h->u * val / (2 * M) + h->v * h->ws[mom_change] * val / M;
            return h->u / (2 * M) * val;
        } else { /* HubbardMom1D.jl:189-199 */
            int prm[3];
            double val = mom_transfer_2c(M, out->n[0], out->n[1], &ma, &mb, chosen, 1, prm);
            return h->u / M * val;
        }
    }
    case ORC_HUBBARD_REAL_SPACE: { /* HubbardRealSpace.jl:316-338,383-391 */
        int nb = 2 * h->ndim, comp = 0;
        long c = chosen;
        if (h->addr_kind == ORC_COMPOSITE) { /* _getindex over the components' hop lists (:383-391) */
            orc_map mc;
            for (comp = 0; ; comp++) {
                build_map(in->n[comp], M, &mc);
                if (c <= (long)mc.len * nb || comp == h->ncomp - 1) break;
                c -= (long)mc.len * nb;
            }
            long particle, neigh;
            fldmod1i(c, nb, &particle, &neigh);
            int src = mc.mode[particle - 1];
            int dstsite = neighbor_site(h, src, (int)neigh);
            if (dstsite == 0) return 0.0;
            int cre[1] = {dstsite}, des[1] = {src};
            double val = excite(h->ckind[comp] == ORC_BOSE, out->n[comp], M, cre, des, 1);
            return -h->tcs[comp] * val;
        }
        if (c > (long)ma.len * nb) { c -= (long)ma.len * nb; comp = 1; }
        const orc_map *mp = comp ? &mb : &ma;
        long particle, neigh;
        fldmod1i(c, nb, &particle, &neigh);
        int src = mp->mode[particle - 1];
        int dstsite = neighbor_site(h, src, (int)neigh);
        if (dstsite == 0) return 0.0;
        int cre[1] = {dstsite}, des[1] = {src};
        double val = excite(h->addr_kind == ORC_BOSE, out->n[comp], M, cre, des, 1);
        return -h->tc[comp] * val;
    }
    case ORC_TRANSCORRELATED_1D: { /* Transcorrelated1D.jl:299-387 */
        long N1 = ma.len, N2 = mb.len;
        long n_mom = N1 * N2 * (M - 1);
        long n1 = h->three_body ? N1 * (N1 - 1) * N2 * M * M : 0;
        long n2 = h->three_body ? N2 * (N2 - 1) * N1 * M * M : 0;
        if (chosen <= n_mom) {
            int prm[3];
            double value = mom_transfer_2c(M, out->n[0], out->n[1], &ma, &mb, chosen, 0, prm);
            if (value != 0.0) value *= tc_t_function(h, prm[0], prm[1], prm[2]);
            return value;
        } else if (chosen <= n_mom + n1) {
            int k, l;
            double value = tc_three_body(M, out->n[0], out->n[1], &ma, &mb, chosen - n_mom, &k, &l);
            value *= tc_q_function(h, k, l);
            if (value == 0.0) onr_copy(h, out, in);
            return value;
        } else if (chosen <= n_mom + n1 + n2) {
            int k, l;
            double value = tc_three_body(M, out->n[1], out->n[0], &mb, &ma, chosen - n_mom - n1, &k, &l);
            value *= tc_q_function(h, k, l);
            if (value == 0.0) onr_copy(h, out, in);
            return value;
        }
        return NAN;
    }
    }
    return NAN;
}

/* key-level wrappers (python entry points) */
double orc_diagonal(const orc_ham *h, const uint64_t *key) {
    orc_onr o; orc_unpack(h, key, &o); return orc_diagonal_onr(h, &o);
}
long orc_num_offdiagonals(const orc_ham *h, const uint64_t *key) {
    orc_onr o; orc_unpack(h, key, &o); return orc_num_offdiagonals_onr(h, &o);
}
double orc_offdiagonal(const orc_ham *h, const uint64_t *key, long chosen, uint64_t *key_out) {
    orc_onr o, n; orc_unpack(h, key, &o);
    double v = orc_offdiagonal_onr(h, &o, chosen, &n);
    orc_pack(h, &n, key_out);
    return v;
}

/* ------------------------------------------------------------------ Philox4x32-10 + hashing
 * The RNG is NOT the reference's (Julia Xoshiro256++, stream parity impossible by design,
 * SURVEY.md section 8c).  Both oracle and kernels use Philox4x32-10 keyed on
 * (step key, address hash, attempt index), so integer-walker steps are bit-reproducible
 * between CPU and GPU irrespective of thread/GPU geometry. */
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void orc_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox4x32_10(ctr, key, out); }

static inline uint64_t fmix64(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    return h;
}
static inline uint64_t addr_hash(const uint64_t *w, int W) {
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (int j = 0; j < W; j++) h = fmix64(h ^ w[j]);
    return h;
}
uint64_t orc_addr_hash(const uint64_t *w, int W) { return addr_hash(w, W); }
/* owner rank of an address: fastrange of the high 32 hash bits (reference: fastrange_hash,
 * pdvec.jl:6-9 + communicators.jl:77-81; placement does not influence results) */
static inline int addr_owner(uint64_t h, int nranks) { return (int)(((h >> 32) * (uint64_t)nranks) >> 32); }
int orc_addr_owner(const uint64_t *w, int W, int nranks) { return addr_owner(addr_hash(w, W), nranks); }

enum { STREAM_SPAWN = 0, STREAM_DIAG = 1, STREAM_COMPRESS = 2 };
static inline double u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)a << 21) ^ ((uint64_t)b >> 11)) * (1.0 / 9007199254740992.0);
}
static inline void rng_draw(uint64_t h, uint64_t k, int stream, const uint32_t key[2], uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)h, (uint32_t)(h >> 32), (uint32_t)k,
                       ((uint32_t)stream << 28) | (uint32_t)((k >> 32) & 0x0fffffffu)};
    philox4x32_10(ctr, key, out);
}

/* ------------------------------------------------------------------ accumulation map (the "working memory") */
typedef union { double f; int64_t i; } orc_val;
/* vals = safe lane (the only one without an initiator rule); vals_u / vals_i = unsafe / initiator lanes of
 * InitiatorValue (DictVectors/initiators.jl:22-45) */
typedef struct { uint64_t *keys; orc_val *vals, *vals_u, *vals_i; uint8_t *used; size_t cap, count; int W, is_int; } orc_mapv;
enum { LANE_SAFE = 0, LANE_UNSAFE = 1, LANE_INIT = 2 };

static void mapv_init(orc_mapv *m, int W, int is_int, size_t cap) {
    size_t c = 64; while (c < cap) c <<= 1;
    m->cap = c; m->count = 0; m->W = W; m->is_int = is_int;
    m->keys = (uint64_t *)malloc(sizeof(uint64_t) * W * c);
    m->vals = (orc_val *)calloc(c, sizeof(orc_val));
    m->vals_u = (orc_val *)calloc(c, sizeof(orc_val));
    m->vals_i = (orc_val *)calloc(c, sizeof(orc_val));
    m->used = (uint8_t *)calloc(c, 1);
}
static void mapv_free(orc_mapv *m) { free(m->keys); free(m->vals); free(m->vals_u); free(m->vals_i); free(m->used); }
static void mapv_add_lane(orc_mapv *m, const uint64_t *key, orc_val v, int lane);
static void mapv_add(orc_mapv *m, const uint64_t *key, orc_val v) { mapv_add_lane(m, key, v, LANE_SAFE); }
static void mapv_grow(orc_mapv *m) {
    orc_mapv n; mapv_init(&n, m->W, m->is_int, m->cap * 2);
    for (size_t s = 0; s < m->cap; s++) if (m->used[s]) {
        mapv_add_lane(&n, m->keys + s * m->W, m->vals[s], LANE_SAFE);
        mapv_add_lane(&n, m->keys + s * m->W, m->vals_u[s], LANE_UNSAFE);
        mapv_add_lane(&n, m->keys + s * m->W, m->vals_i[s], LANE_INIT);
    }
    mapv_free(m); *m = n;
}
static void mapv_add_lane(orc_mapv *m, const uint64_t *key, orc_val v, int lane) {
    if ((m->count + 1) * 2 > m->cap) mapv_grow(m);
    size_t mask = m->cap - 1, s = (size_t)addr_hash(key, m->W) & mask;
    for (;;) {
        if (!m->used[s]) {
            m->used[s] = 1; memcpy(m->keys + s * m->W, key, sizeof(uint64_t) * m->W);
            m->vals[s].i = 0; m->vals_u[s].i = 0; m->vals_i[s].i = 0; m->count++;
        }
        if (memcmp(m->keys + s * m->W, key, sizeof(uint64_t) * m->W) == 0) {
            orc_val *dst = lane == LANE_SAFE ? &m->vals[s] : lane == LANE_UNSAFE ? &m->vals_u[s] : &m->vals_i[s];
            if (m->is_int) dst->i += v.i; else dst->f += v.f;
            return;
        }
        s = (s + 1) & mask;
    }
}
static int mapv_get(const orc_mapv *m, const uint64_t *key, orc_val *v) {
    size_t mask = m->cap - 1, s = (size_t)addr_hash(key, m->W) & mask;
    for (;;) {
        if (!m->used[s]) return 0;
        if (memcmp(m->keys + s * m->W, key, sizeof(uint64_t) * m->W) == 0) { *v = m->vals[s]; return 1; }
        s = (s + 1) & mask;
    }
}

/* deposit sink: one map (serial DVec semantics) or T row-maps chosen by target segment
 * (PDWorkingMemory column, pdworkingmemory.jl:21-31) */
typedef struct { orc_mapv *maps; int T; } orc_sink;
static inline void sink_add(orc_sink *s, const uint64_t *key, orc_val v, int lane) {
    int row = s->T > 1 ? addr_owner(addr_hash(key, s->maps[0].W) << 32, s->T) : 0;
    mapv_add_lane(&s->maps[row], key, v, lane);
}

/* ------------------------------------------------------------------ step */
typedef struct {
    int32_t style;          /* ORC_STYLE_* */
    int32_t plain_h;        /* 1: operator is H itself (mul!, pdvec.jl:810-822); 0: T = 1 + dtau (S - H) */
    double shift, dtau, boost;
    double proj_threshold;  /* on-the-fly projection threshold of the spawning strategy */
    double rel_threshold, abs_threshold; /* DynamicSemistochastic (spawning.jl:358-378) */
    double compress_threshold;           /* ThresholdCompression, 0 = NoCompression */
    uint32_t key[2];        /* Philox key for this step */
    int32_t initiator_rule; /* 0 NonInitiator, 1 Initiator, 2 SimpleInitiator, 3 CoherentInitiator (initiators.jl:132-236) */
    int32_t pad_;
    double initiator_threshold;
} orc_step_params;

/* to_initiator_value (initiators.jl:142-158): which lane of the child's InitiatorValue a deposit goes to */
static inline int deposit_lane(const orc_step_params *p, int diagonal, double parent_val) {
    if (p->initiator_rule == 0) return LANE_SAFE;
    int is_initiator = fabs(parent_val) > p->initiator_threshold;
    if (diagonal) return is_initiator ? LANE_INIT : LANE_SAFE;
    return is_initiator ? LANE_SAFE : LANE_UNSAFE;
}
/* from_initiator_value (initiators.jl:136-138, 177-179, 201-207) */
static inline double from_initiator_f(const orc_step_params *p, double safe, double unsafe, double init) {
    switch (p ? p->initiator_rule : 0) {
    case 1: return safe + init + (init != 0.0 ? 1.0 : 0.0) * unsafe;
    case 2: return safe + init;
    case 3: return (init != 0.0 || fabs(unsafe) > p->initiator_threshold) ? init + safe + unsafe : init + safe;
    default: return safe;
    }
}
static inline int64_t from_initiator_i(const orc_step_params *p, int64_t safe, int64_t unsafe, int64_t init) {
    switch (p ? p->initiator_rule : 0) {
    case 1: return safe + init + (init != 0 ? 1 : 0) * unsafe;
    case 2: return safe + init;
    case 3: return (init != 0 || fabs((double)unsafe) > p->initiator_threshold) ? init + safe + unsafe : init + safe;
    default: return safe;
    }
}

typedef struct {
    int64_t exact_steps, inexact_steps, spawn_attempts, len_before, len_after;
    double spawns, deaths, clones, zombies, norm1;     /* float styles */
    int64_t ispawns, ideaths, iclones, izombies, inorm1; /* integer style */
} orc_step_stats;

static inline double sgn(double x) { return (x > 0) - (x < 0); }

/* projected_deposit! (spawning.jl:9-45): returns deposited value (as double; exact for ints
 * below 2^53) */
static double projected_deposit(orc_sink *w, int is_int, const uint64_t *key, double val,
                                double threshold, double r, int lane) {
    if (is_int) {
        int64_t nv = (int64_t)sgn(val) * (int64_t)floor(fabs(val) + r);
        if (nv != 0) { orc_val v; v.i = nv; sink_add(w, key, v, lane); }
        return (double)nv;
    }
    double a = fabs(val);
    if (a < threshold) {
        if (r < a / threshold) val = sgn(val) * threshold; else val = 0.0;
    }
    if (val != 0.0) { orc_val v; v.f = val; sink_add(w, key, v, lane); }
    return val;
}

static void clones_deaths_zombies(double res, double val, double *clones, double *deaths, double *zombies) {
    /* spawning.jl:79-93 */
    *clones = *deaths = *zombies = 0;
    if (res > val) *clones = fabs(res - val);
    else if (sgn(res) != sgn(val)) { *deaths = fabs(val); *zombies = fabs(res); }
    else *deaths = fabs(res - val);
}

/* apply_column! for one parent (styles.jl:21-25,101-105,210-214 + fciqmc.jl:93-112) */
static void apply_column(const orc_ham *h, const orc_step_params *p, orc_sink *w,
                         const uint64_t *key, orc_val pval, orc_step_stats *st) {
    int is_int = p->style == ORC_STYLE_INTEGER, W = h->words;
    double val = is_int ? (double)pval.i : pval.f;
    uint64_t hsh = addr_hash(key, W);
    uint32_t rnd[4];
    orc_onr o; orc_unpack(h, key, &o);

    /* diagonal_step! */
    double hd = orc_diagonal_onr(h, &o);
    double d = p->plain_h ? hd : 1 - p->dtau * (hd - p->shift);
    rng_draw(hsh, 0, STREAM_DIAG, p->key, rnd);
    double res = projected_deposit(w, is_int, key, d * val, is_int ? 0.0 : p->proj_threshold, u53(rnd[1], rnd[2]), deposit_lane(p, 1, val));
    double cl, de, zo; clones_deaths_zombies(res, val, &cl, &de, &zo);
    if (is_int) { st->iclones += (int64_t)cl; st->ideaths += (int64_t)de; st->izombies += (int64_t)zo; }
    else { st->clones += cl; st->deaths += de; st->zombies += zo; }

    /* spawn! */
    long L = orc_num_offdiagonals_onr(h, &o);
    if (L <= 0) return;
    int exact;
    if (p->style == ORC_STYLE_DETERMINISTIC) exact = 1;
    else if (p->style == ORC_STYLE_SEMISTOCHASTIC) {
        double thresh = fmin(p->abs_threshold, (double)L);
        double amount = p->boost * fabs(val) * p->rel_threshold;
        exact = amount >= thresh;
    } else exact = 0;
    double spawns = 0;
    orc_onr child; uint64_t ckey[2];
    if (exact) { /* spawn!(Exact) spawning.jl:174-182 */
        for (long i = 1; i <= L; i++) {
            double m = orc_offdiagonal_onr(h, &o, i, &child);
            if (!p->plain_h) m = -m * p->dtau;
            double r = 0.0;
            if (p->proj_threshold > 0) { rng_draw(hsh, (uint64_t)(i - 1), STREAM_SPAWN, p->key, rnd); r = u53(rnd[1], rnd[2]); }
            orc_pack(h, &child, ckey);
            spawns += fabs(projected_deposit(w, 0, ckey, val * m, p->proj_threshold, r, deposit_lane(p, 0, val)));
        }
        st->exact_steps += 1; st->spawn_attempts += L;
    } else { /* spawn!(WithReplacement) spawning.jl:232-243 + random_offdiagonal hamiltonians.jl:361-370 */
        int64_t n = (int64_t)floor(fabs(val) * p->boost); if (n < 1) n = 1;
        double magnitude = val / (double)n;
        for (int64_t k = 0; k < n; k++) {
            rng_draw(hsh, (uint64_t)k, STREAM_SPAWN, p->key, rnd);
            long i = (long)(((uint64_t)rnd[0] * (uint64_t)L) >> 32) + 1;
            double m = orc_offdiagonal_onr(h, &o, i, &child);
            if (!p->plain_h) m = -m * p->dtau;
            /* spawning.jl:240: mat_elem * magnitude / prob with prob = 1 / L (hamiltonians.jl:366); written as the
             * multiplication by L -- one rounding instead of two, within the 1e-12 Float64 tolerance of the reference's value */
            double nv = m * magnitude * (double)L;
            orc_pack(h, &child, ckey);
            spawns += fabs(projected_deposit(w, is_int, ckey, nv, is_int ? 0.0 : p->proj_threshold, u53(rnd[1], rnd[2]), deposit_lane(p, 0, val)));
        }
        st->inexact_steps += 1; st->spawn_attempts += n;
    }
    if (is_int) st->ispawns += (int64_t)spawns; else st->spawns += spawns;
}

typedef struct { uint64_t k[2]; orc_val v; } orc_rec;
static int rec_cmp1(const void *a, const void *b) {
    const orc_rec *x = (const orc_rec *)a, *y = (const orc_rec *)b;
    if (x->k[1] != y->k[1]) return x->k[1] < y->k[1] ? -1 : 1;
    if (x->k[0] != y->k[0]) return x->k[0] < y->k[0] ? -1 : 1;
    return 0;
}

/* move_and_compress! for one map (pdworkingmemory.jl:262-273, compression.jl:18-26): drops exact
 * zeros, applies ThresholdCompression, collects survivors (unsorted) and partial statistics. */
static long map_collect(const orc_mapv *w, const orc_step_params *p, orc_rec *recs, orc_step_stats *st) {
    int W = w->W, is_int = w->is_int;
    long n = 0, len_before = 0;
    uint32_t rnd[4];
    for (size_t s = 0; s < w->cap; s++) {
        if (!w->used[s]) continue;
        orc_val v = w->vals[s];
        /* entries whose InitiatorValue is entirely zero are deleted on deposit (pdworkingmemory.jl:25-29); the others
         * are counted by len_before and converted with from_initiator_value (pdworkingmemory.jl:268-270) */
        if (is_int ? (v.i == 0 && w->vals_u[s].i == 0 && w->vals_i[s].i == 0) : (v.f == 0.0 && w->vals_u[s].f == 0.0 && w->vals_i[s].f == 0.0)) continue;
        len_before++;
        if (is_int) v.i = from_initiator_i(p, v.i, w->vals_u[s].i, w->vals_i[s].i);
        else v.f = from_initiator_f(p, v.f, w->vals_u[s].f, w->vals_i[s].f);
        if (is_int ? v.i == 0 : v.f == 0.0) continue; /* setindex! of a zero deletes */
        if (!is_int && p && p->compress_threshold > 0) {
            double prob = fabs(v.f) / p->compress_threshold;
            if (prob < 1) {
                rng_draw(addr_hash(w->keys + s * W, W), 0, STREAM_COMPRESS, p->key, rnd);
                v.f = (prob > u53(rnd[1], rnd[2])) ? p->compress_threshold * sgn(v.f) : 0.0;
            }
            if (v.f == 0.0) continue;
        }
        recs[n].k[0] = w->keys[s * W]; recs[n].k[1] = W > 1 ? w->keys[s * W + 1] : 0; recs[n].v = v;
        n++;
    }
    if (st) {
        st->len_before += len_before; st->len_after += n;
        for (long i = 0; i < n; i++) {
            if (is_int) st->inorm1 += recs[i].v.i < 0 ? -recs[i].v.i : recs[i].v.i;
            else st->norm1 += fabs(recs[i].v.f);
        }
    }
    return n;
}
static long recs_write(const orc_rec *recs, long n, int W, int is_int, uint64_t *keys_out, void *vals_out, long cap) {
    if (n > cap) return -n;
    for (long i = 0; i < n; i++) {
        for (int j = 0; j < W; j++) keys_out[i * W + j] = recs[i].k[j];
        if (is_int) ((int64_t *)vals_out)[i] = recs[i].v.i; else ((double *)vals_out)[i] = recs[i].v.f;
    }
    return n;
}
/* serial export in ascending key order (canonical form for parity comparisons) */
static long map_export(orc_mapv *w, const orc_step_params *p, uint64_t *keys_out, void *vals_out,
                       long cap, orc_step_stats *st) {
    orc_rec *recs = (orc_rec *)malloc(sizeof(orc_rec) * (w->count + 1));
    long n = map_collect(w, p, recs, st);
    qsort(recs, n, sizeof(orc_rec), rec_cmp1);
    long r = recs_write(recs, n, w->W, w->is_int, keys_out, vals_out, cap);
    free(recs);
    return r;
}

/* apply_operator! (Interfaces/dictvectors.jl:112-140): serial semantics.
 * keys: n*W words; vals: n x (double | int64).  If rank filter nranks>1 is given, only
 * children owned by `rank` are kept (used to check the partitioned multi-GPU path). */
long orc_step(const orc_ham *h, const orc_step_params *p, long n, const uint64_t *keys, const void *vals,
              uint64_t *keys_out, void *vals_out, long cap_out, orc_step_stats *st) {
    int is_int = p->style == ORC_STYLE_INTEGER, W = h->words;
    memset(st, 0, sizeof(*st));
    orc_mapv w; mapv_init(&w, W, is_int, (size_t)n * 4 + 64);
    orc_sink sink = {&w, 1};
    for (long i = 0; i < n; i++) {
        orc_val v;
        if (is_int) v.i = ((const int64_t *)vals)[i]; else v.f = ((const double *)vals)[i];
        apply_column(h, p, &sink, keys + i * W, v, st);
    }
    long r = map_export(&w, p, keys_out, vals_out, cap_out, st);
    mapv_free(&w);
    return r;
}

/* Threaded restatement structured like the reference's PDVec path
 * (pdworkingmemory.jl:191-309): the vector lives in T hash segments; thread t spawns segment t into
 * its private column of T row-maps (perform_spawns!), row r is merged over columns
 * (collect_local!), then compressed into segment r of the target (move_and_compress!).  Output is
 * the concatenation of the segments (unsorted, as in the reference).  Used as cpu_baseline. */
long orc_step_threaded(const orc_ham *h, const orc_step_params *p, long n, const uint64_t *keys,
                       const void *vals, uint64_t *keys_out, void *vals_out, long cap_out,
                       orc_step_stats *st, int T) {
    int is_int = p->style == ORC_STYLE_INTEGER, W = h->words;
    memset(st, 0, sizeof(*st));
    if (T < 1) T = 1;
    orc_mapv *grid = (orc_mapv *)malloc(sizeof(orc_mapv) * T * T); /* grid[col*T + row] */
    orc_step_stats *sts = (orc_step_stats *)calloc(T, sizeof(orc_step_stats));
    for (int i = 0; i < T * T; i++) mapv_init(&grid[i], W, is_int, (size_t)(n * 2 / (T * T)) + 64);
    /* segment membership of the source vector (in the reference the PDVec is already stored that way) */
    int32_t *seg = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    long *seg_start = (long *)calloc(T + 1, sizeof(long));
    long *order = (long *)malloc(sizeof(long) * (n + 1));
#pragma omp parallel for num_threads(T) schedule(static)
    for (long i = 0; i < n; i++) seg[i] = addr_owner(addr_hash(keys + i * W, W) << 32, T);
    for (long i = 0; i < n; i++) seg_start[seg[i] + 1]++;
    for (int t = 0; t < T; t++) seg_start[t + 1] += seg_start[t];
    {
        long *fill = (long *)malloc(sizeof(long) * T);
        for (int t = 0; t < T; t++) fill[t] = seg_start[t];
        for (long i = 0; i < n; i++) order[fill[seg[i]]++] = i;
        free(fill);
    }
    /* perform_spawns! */
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; t++) {
        orc_sink sink = {&grid[t * T], T};
        for (long q = seg_start[t]; q < seg_start[t + 1]; q++) {
            long i = order[q];
            orc_val v;
            if (is_int) v.i = ((const int64_t *)vals)[i]; else v.f = ((const double *)vals)[i];
            apply_column(h, p, &sink, keys + i * W, v, &sts[t]);
        }
    }
    /* collect_local! + move_and_compress! per row */
    orc_rec **rows = (orc_rec **)calloc(T, sizeof(orc_rec *));
    long *rown = (long *)calloc(T, sizeof(long));
    orc_step_stats *cst = (orc_step_stats *)calloc(T, sizeof(orc_step_stats));
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int r = 0; r < T; r++) {
        for (int c = 1; c < T; c++) {
            orc_mapv *src = &grid[c * T + r];
            for (size_t s = 0; s < src->cap; s++)
                if (src->used[s]) { /* (the unsafe / initiator lanes are all zero unless an initiator rule is on) */
                    mapv_add_lane(&grid[r], src->keys + s * W, src->vals[s], LANE_SAFE);
                    if (src->vals_u[s].i != 0) mapv_add_lane(&grid[r], src->keys + s * W, src->vals_u[s], LANE_UNSAFE);
                    if (src->vals_i[s].i != 0) mapv_add_lane(&grid[r], src->keys + s * W, src->vals_i[s], LANE_INIT);
                }
        }
        rows[r] = (orc_rec *)malloc(sizeof(orc_rec) * (grid[r].count + 1));
        rown[r] = map_collect(&grid[r], p, rows[r], &cst[r]);
    }
    long total = 0;
    for (int t = 0; t < T; t++) {
        st->exact_steps += sts[t].exact_steps; st->inexact_steps += sts[t].inexact_steps;
        st->spawn_attempts += sts[t].spawn_attempts;
        st->spawns += sts[t].spawns; st->deaths += sts[t].deaths; st->clones += sts[t].clones; st->zombies += sts[t].zombies;
        st->ispawns += sts[t].ispawns; st->ideaths += sts[t].ideaths; st->iclones += sts[t].iclones; st->izombies += sts[t].izombies;
        st->len_before += cst[t].len_before; st->len_after += cst[t].len_after;
        st->norm1 += cst[t].norm1; st->inorm1 += cst[t].inorm1;
        total += rown[t];
    }
    long ret = total;
    if (total > cap_out) ret = -total;
    else {
        long off = 0;
        for (int r = 0; r < T; r++) {
            recs_write(rows[r], rown[r], W, is_int, keys_out + off * W,
                       is_int ? (void *)((int64_t *)vals_out + off) : (void *)((double *)vals_out + off), rown[r]);
            off += rown[r];
        }
    }
    for (int r = 0; r < T; r++) free(rows[r]);
    free(rows); free(rown); free(cst); free(seg); free(seg_start); free(order);
    for (int i = 0; i < T * T; i++) mapv_free(&grid[i]);
    free(grid); free(sts);
    return ret;
}

/* annihilation of a given spawn list: sum by key, drop exact zeros, ascending key order */
long orc_annihilate(int W, int is_int, long n, const uint64_t *keys, const void *vals,
                    uint64_t *keys_out, void *vals_out, long cap_out) {
    orc_mapv w; mapv_init(&w, W, is_int, (size_t)n * 2 + 64);
    for (long i = 0; i < n; i++) {
        orc_val v;
        if (is_int) v.i = ((const int64_t *)vals)[i]; else v.f = ((const double *)vals)[i];
        mapv_add(&w, keys + i * W, v);
    }
    long r = map_export(&w, NULL, keys_out, vals_out, cap_out, NULL);
    mapv_free(&w);
    return r;
}

/* ------------------------------------------------------------------ exact diagonalisation support
 * basis_breadth_first_search.jl:216-399: BFS over non-zero off-diagonals from a start address;
 * emits COO triplets of H restricted to the connected sector. */
long orc_bfs_basis(const orc_ham *h, const uint64_t *start, long max_dim, uint64_t *basis) {
    int W = h->words;
    orc_mapv seen; mapv_init(&seen, W, 1, 1024);
    long dim = 0, head = 0;
    orc_val v; v.i = 1;
    memcpy(basis, start, sizeof(uint64_t) * W); dim = 1; mapv_add(&seen, start, v);
    while (head < dim) {
        orc_onr o, c; uint64_t ck[2];
        orc_unpack(h, basis + head * W, &o);
        long L = orc_num_offdiagonals_onr(h, &o);
        for (long i = 1; i <= L; i++) {
            double m = orc_offdiagonal_onr(h, &o, i, &c);
            if (m == 0.0) continue;
            orc_pack(h, &c, ck);
            orc_val tmp;
            if (!mapv_get(&seen, ck, &tmp)) {
                if (dim >= max_dim) { mapv_free(&seen); return -1; }
                v.i = dim + 1; mapv_add(&seen, ck, v);
                memcpy(basis + dim * W, ck, sizeof(uint64_t) * W); dim++;
            }
        }
        head++;
    }
    mapv_free(&seen);
    return dim;
}

/* COO of H on a given basis; entries whose child is outside the basis are dropped.
 * Returns nnz (or -needed if cap too small). */
long orc_coo_matrix(const orc_ham *h, long dim, const uint64_t *basis, long cap, int64_t *rows, int64_t *cols, double *vals) {
    int W = h->words;
    orc_mapv idx; mapv_init(&idx, W, 1, (size_t)dim * 2);
    for (long i = 0; i < dim; i++) { orc_val v; v.i = i; mapv_add(&idx, basis + i * W, v); }
    long nnz = 0;
    for (long j = 0; j < dim; j++) {
        orc_onr o, c; uint64_t ck[2];
        orc_unpack(h, basis + j * W, &o);
        double d = orc_diagonal_onr(h, &o);
        if (nnz < cap) { rows[nnz] = j; cols[nnz] = j; vals[nnz] = d; } nnz++;
        long L = orc_num_offdiagonals_onr(h, &o);
        for (long i = 1; i <= L; i++) {
            double m = orc_offdiagonal_onr(h, &o, i, &c);
            if (m == 0.0) continue;
            orc_pack(h, &c, ck);
            orc_val r;
            if (!mapv_get(&idx, ck, &r)) continue;
            if (nnz < cap) { rows[nnz] = r.i; cols[nnz] = j; vals[nnz] = m; } nnz++;
        }
    }
    mapv_free(&idx);
    return nnz <= cap ? nnz : -nnz;
}

/* ------------------------------------------------------------------ dense-indexed rows of y = H x over a complete sector
 * (SURVEY 8c lesson 3: the independent oracle for config 3 at sizes where no matrix can be built).  The addresses of the
 * sector -- every bit string with the right number of set bits per component -- are numbered by their combinadic rank
 * sum_j C(p_j, j) (set bits p_1 < p_2 < ..., j = 1, 2, ...); two components: rank(comp 0) * dim(comp 1) + rank(comp 1).
 * Everything here is plain loops over a Pascal triangle, independent of the byte tables the device code uses; the
 * off-diagonals come from the ONR code above. */
static uint64_t pascal_[65][66];
static int pascal_ready_ = 0;
static void pascal_init(void) {
    if (pascal_ready_) return;
    for (int n = 0; n <= 64; n++) {
        pascal_[n][0] = 1;
        for (int k = 1; k <= 65; k++) pascal_[n][k] = 0;
        for (int k = 1; k <= n; k++) {
            uint64_t a = pascal_[n - 1][k - 1], b = k <= n - 1 ? pascal_[n - 1][k] : 0;
            pascal_[n][k] = (a > ((uint64_t)1 << 62) || b > ((uint64_t)1 << 62)) ? ~(uint64_t)0 : a + b;
        }
    }
    pascal_ready_ = 1;
}
static void sector_shape(const orc_ham *h, int *ncomp, int bits[2], int ones[2], int shift[2]) {
    if (h->addr_kind == ORC_BOSE) { *ncomp = 1; bits[0] = h->N[0] + h->M - 1; ones[0] = h->N[0]; shift[0] = 0; }
    else if (h->addr_kind == ORC_FERMI) { *ncomp = 1; bits[0] = h->M; ones[0] = h->N[0]; shift[0] = 0; }
    else { *ncomp = 2; for (int c = 0; c < 2; c++) { bits[c] = h->M; ones[c] = h->N[c]; shift[c] = c * h->M; } }
}
long orc_sector_dim(const orc_ham *h) {
    pascal_init();
    int nc, bits[2], ones[2], shift[2];
    sector_shape(h, &nc, bits, ones, shift);
    uint64_t d = pascal_[bits[0]][ones[0]];
    if (nc == 2) d *= pascal_[bits[1]][ones[1]];
    return (long)d;
}
long orc_sector_rank(const orc_ham *h, const uint64_t *key) {
    pascal_init();
    int nc, bits[2], ones[2], shift[2];
    sector_shape(h, &nc, bits, ones, shift);
    uint64_t r = 0;
    for (int c = 0; c < nc; c++) {
        uint64_t rc = 0;
        int j = 0;
        for (int p = 0; p < bits[c]; p++)
            if (getbit(key, shift[c] + p)) { j++; rc += pascal_[p][j]; }
        r = c == 0 ? rc : r * pascal_[bits[1]][ones[1]] + rc;
    }
    return (long)r;
}
void orc_sector_unrank(const orc_ham *h, long idx, uint64_t *key) {
    pascal_init();
    int nc, bits[2], ones[2], shift[2];
    sector_shape(h, &nc, bits, ones, shift);
    uint64_t r[2] = {(uint64_t)idx, 0};
    if (nc == 2) { uint64_t d1 = pascal_[bits[1]][ones[1]]; r[0] = (uint64_t)idx / d1; r[1] = (uint64_t)idx % d1; }
    key[0] = 0; if (h->words > 1) key[1] = 0;
    for (int c = 0; c < nc; c++) {
        uint64_t rem = r[c];
        int p = bits[c] - 1;
        for (int j = ones[c]; j >= 1; j--) {
            while (pascal_[p][j] > rem) p--;
            rem -= pascal_[p][j];
            setbit(key, shift[c] + p);
            p--;
        }
    }
}
/* y[i] = H_ii x[i] + sum_k H_{c_k, i} x[rank(c_k)] for the n sampled rows idx[] (H real symmetric) */
void orc_sector_rows(const orc_ham *h, long n, const int64_t *idx, const double *x, double *y_out) {
    pascal_init();
#pragma omp parallel for schedule(dynamic, 64)
    for (long r = 0; r < n; r++) {
        uint64_t key[2] = {0, 0}, ckey[2];
        orc_sector_unrank(h, (long)idx[r], key);
        double acc = orc_diagonal(h, key) * x[idx[r]];
        long L = orc_num_offdiagonals(h, key);
        for (long k = 1; k <= L; k++) {
            ckey[0] = ckey[1] = 0;
            double m = orc_offdiagonal(h, key, k, ckey);
            if (m != 0.0) acc += m * x[orc_sector_rank(h, ckey)];
        }
        y_out[r] = acc;
    }
}

int orc_sizeof_ham(void) { return (int)sizeof(orc_ham); }
int orc_sizeof_params(void) { return (int)sizeof(orc_step_params); }
int orc_sizeof_stats(void) { return (int)sizeof(orc_step_stats); }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
