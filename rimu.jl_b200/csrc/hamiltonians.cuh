// hamiltonians.cuh -- device-side Fock-address arithmetic and the four lattice Hamiltonians.
//
// Works directly on the bit-packed addresses (no ONR expansion): occupied-mode lookups are
// select/ctz/popc chains, boson moves are single-bit delete/insert, fermion signs are popcounts.
// Semantics follow (path:line under Rimu.jl src/):
//   BitStringAddresses/bitstring.jl:464-545 (BoseFS bits), :713-792 (FermiFS bits, sign rule)
//   BitStringAddresses/fockaddress.jl:559-567 (bosonic excitation value)
//   Hamiltonians/HubbardReal1D.jl:51-62 + BitStringAddresses/bosefs.jl:270-345 (hop enumeration)
//   Hamiltonians/HubbardMom1D.jl:131-205 + excitations.jl:26-137 (momentum transfer)
//   Hamiltonians/HubbardRealSpace.jl:18-106,279-391 + geometry.jl:161-175,232-235
//   Hamiltonians/Transcorrelated1D.jl:113-387 + excitations.jl:199-238
// Off-diagonal index `i` is 0-based here (= reference's chosen-1); enumeration order is the
// reference's.  Floating-point expressions are written in the reference's evaluation order and the
// library is compiled with --fmad=false so that values are bit-identical to an IEEE CPU evaluation.
#pragma once
#include "common.cuh"

enum HamKind {
    HK_REAL1D_BOSE = 0,
    HK_MOM1D_BOSE = 1,
    HK_MOM1D_F2C = 2,
    HK_RS_BOSE = 3,
    HK_RS_FERMI = 4,
    HK_RS_F2C = 5,
    HK_TC_F2C = 6,
    HK_RS_COMP = 7,   // HubbardRealSpace over a general CompositeFS (2..4 BoseFS / FermiFS components, one or two words)
    // "plain" kinds: the code of their base kind with HamDev::variant == 0 known at COMPILE time -- HubbardReal1D and HubbardMom1D
    // themselves (BASELINE configs 1 and 2).  The variants' branches (EP / Extended models) cost the hot kernels of the plain
    // models 1.6 % (gpurun_out/r2_sweep_var0.log); ham_build_host picks the plain kind whenever variant == 0.
    HK_REAL1D_BOSE_PLAIN = 8,
    HK_MOM1D_BOSE_PLAIN = 9,
    HK_COUNT = 10
};
template <int HKX> struct HkBase {
    static constexpr int value = HKX == HK_REAL1D_BOSE_PLAIN ? (int)HK_REAL1D_BOSE : HKX == HK_MOM1D_BOSE_PLAIN ? (int)HK_MOM1D_BOSE : HKX;
    static constexpr bool plain = value != HKX;
};
#define HAM_MAX_COMP 4

struct HamDev {
    int hk, M, N0, N1, ndim, nnb, cutoff, three_body, has_pot, umat_zero;
    int variant, bc; // HK_REAL1D_BOSE: 0 HubbardReal1D, 1 HubbardReal1DEP (pot = eps_i), 2 ExtendedHubbardReal1D (v, bc = RIMU_BC_*)
                     // HK_MOM1D_BOSE / HK_MOM1D_F2C: 0 HubbardMom1D, 1 ExtendedHubbardMom1D (v; ws, us = cosine tables),
                     //                               2 HubbardMom1DEP (pot = ep, momentum-space harmonic potential)
    double v_m;      // v / M
    double u, t, v, tc0, tc1, u00, u10;
    double u_2m, u_m; // u / (2M), u / M: the same IEEE divisions the reference evaluates per element, done once on the host
    const double *kes, *ws, *us, *pot; // device tables
    const unsigned char *nbr;          // HubbardRealSpace: nbr[(site-1)*nnb + dir] = neighbour site (1-based) or 0
    // HK_RS_COMP: component c is a BoseFS (cbose[c]) or FermiFS bit string of cbits[c] bits at bit offset coff[c]
    int ncomp, cbose[HAM_MAX_COMP], coff[HAM_MAX_COMP], cbits[HAM_MAX_COMP];
    double tcs[HAM_MAX_COMP], umat[HAM_MAX_COMP * HAM_MAX_COMP]; // t[c]; u[i + ncomp * j]
};

// the model variant of a kind (see HamDev::variant).  RIMU_TUNE_VARIANT0 (kernel-tuning builds only) compiles the plain models
// alone, to measure what the variants' branches cost the benchmark configuration.
#ifdef RIMU_TUNE_VARIANT0
#define HVARIANT(h) 0
#else
#define HVARIANT(h) (HkBase<HKX>::plain ? 0 : (h).variant) /* HKX: the kind the enclosing function was instantiated for */
#endif
#if defined(__CUDACC__) || defined(RIMU_HOST_EMULATION)
// exact x / d for x < 2^21, 0 < d < 2^10 (off-diagonal index decoding): (x + 0.5) / d is at least 0.5/d away from
// an integer, far more than the float rounding error at these magnitudes, so truncation gives floor(x / d)
DEV unsigned udiv_small(unsigned x, unsigned d) { return (unsigned)__float2uint_rz(((float)x + 0.5f) * __frcp_rn((float)d)); }

// ---------------------------------------------------------------- boson primitives
template <class B> DEV int bose_mode_offset(B x, int m) { return m == 1 ? 0 : select_((B)~x, m - 2) + 1; }
template <class B> DEV B delete_bit(B x, int p) { return (x & lowmask<B>(p)) | ((x >> (p + 1)) << p); }
template <class B> DEV B insert_one(B x, int p) { return (x & lowmask<B>(p)) | (((B)1) << p) | ((x >> p) << (p + 1)); }
// a_m: returns occupation before (0 = illegal, x untouched)
template <class B> DEV int bose_destroy(B &x, int m) {
    int off = bose_mode_offset(x, m);
    int n = cto_((B)(x >> off));
    if (n) x = delete_bit(x, off);
    return n;
}
// a^dagger_m: returns occupation after
template <class B> DEV int bose_create(B &x, int m) {
    int off = bose_mode_offset(x, m);
    int n = cto_((B)(x >> off));
    x = insert_one(x, off);
    return n + 1;
}
template <class B> DEV int bose_num_occupied(B x) { return popc_((B)(x & ~(x << 1))); }
template <class B> DEV int bose_num_doubly(B x) { return popc_((B)(x & ~(x << 1) & (x >> 1))); }
// k-th (0-based) occupied mode in ascending order -> (mode 1-based, occupation, bit offset of its first particle).
// Occupied modes start where a 1 follows a 0 (or sits at bit 0); the k-th such bit is found with a rank/select
// step instead of walking the modes (OccupiedModeMap lookups, fockaddress.jl:258-275, in O(1)).
template <class B> DEV void bose_kth_occupied(B x, int k, int &mode, int &occ, int &off) {
    const B starts = x & ~(x << 1);
    off = select_(starts, k);
    mode = off - popc_((B)(x & lowmask<B>(off))) + 1;
    occ = cto_((B)(x >> off));
}
template <class B> DEV void bose_kth_occupied(B x, int k, int &mode, int &occ) {
    int off;
    bose_kth_occupied(x, k, mode, occ, off);
}
// d-th (0-based) mode with occupation >= 2
template <class B> DEV void bose_kth_doubly(B x, int d, int &mode, int &occ, int &off) {
    const B starts = x & ~(x << 1) & (x >> 1);
    off = select_(starts, d);
    mode = off - popc_((B)(x & lowmask<B>(off))) + 1;
    occ = cto_((B)(x >> off));
}
// Occupied modes of a bosonic bit string in ascending order: `for (BoseModes<B> it(x); it.next(md, n);)` yields the 0-based
// mode and its occupation.  The string is bit-reversed ONCE, so that every scan is a count-leading-zeros (one FLO per word)
// instead of a bit reversal + FLO per scan: these all run on the quarter-rate XU pipe, which the diagonal elements of new
// determinants kept busy (profiles/r2_merge_ncu.md).
DEV u64 brev_(u64 x) { return __brevll(x); }
DEV u128 brev_(u128 x) { return ((u128)__brevll((u64)x) << 64) | (u128)__brevll((u64)(x >> 64)); }
DEV int clz_(u64 x) { return __clzll((long long)x); } // 64 for x == 0
DEV int clz_(u128 x) { const u64 hi = (u64)(x >> 64); return hi ? __clzll((long long)hi) : 64 + __clzll((long long)(u64)x); }
template <class B> struct BoseModes {
    B yb;   // remaining bits, reversed: the next bit of the string is the most significant one
    int md; // zeros passed so far = index of the mode the next particle belongs to
    DEV explicit BoseModes(B x) : yb(brev_(x)), md(0) {}
    DEV bool next(int &mode, int &n) {
        if (yb == 0) return false;
        const int z = clz_(yb); yb <<= z; md += z;
        n = clz_((B)~yb); yb <<= n; // (a valid address has a spare zero bit, so n < bit width)
        mode = md;
        return true;
    }
};
template <class B> DEV long long bose_interaction(B x) { // sum n(n-1)
    int r = 0, md, n;
    for (BoseModes<B> it(x); it.next(md, n);) r += n * (n - 1);
    return r;
}

// ---------------------------------------------------------------- fermion primitives (one component in a u64)
DEV bool fermi_destroy(u64 &f, int m, int &cnt) {
    u64 bit = 1ull << (m - 1);
    if (!(f & bit)) return false;
    cnt += __popcll(f & (bit - 1));
    f ^= bit;
    return true;
}
DEV bool fermi_create(u64 &f, int m, int &cnt) {
    u64 bit = 1ull << (m - 1);
    if (f & bit) return false;
    cnt += __popcll(f & (bit - 1));
    f ^= bit;
    return true;
}
DEV double parity_sign(int cnt) { return (cnt & 1) ? -1.0 : 1.0; }
DEV double kes_sum(const double *kes, u64 f) { // dot(kes, OccupiedModeMap) ascending modes
    double s = 0.0;
    while (f) { int b = __ffsll((long long)f) - 1; f &= f - 1; s += kes[b] * 1; }
    return s;
}

// ---------------------------------------------------------------- Transcorrelated1D scalar functions
DEV double tc_n_to_k(int n, int M) { return n * 2.0 * 3.14159265358979323846 / M; }
DEV double tc_corr(const HamDev &h, int n) {
    int a = n < 0 ? -n : n;
    if (a == 0) return 0.0;
    return (n > 0 ? 1.0 : -1.0) * h.us[a - 1];
}
DEV double tc_w(const HamDev &h, int n) { return h.ws[n < 0 ? -n : n]; }
DEV double tc_t_function(const HamDev &h, int p, int q, int k) {
    int M = h.M;
    double k_pi = tc_n_to_k(k, M), pmq_pi = tc_n_to_k(p - q, M), cor_k = tc_corr(h, k);
    return h.v / M + 2 * h.v / M * (cor_k * k_pi - cor_k * pmq_pi) + 2 * h.v * h.v / h.t * tc_w(h, k);
}
DEV double tc_q_function(const HamDev &h, int k, int l) {
    int M = h.M;
    return -(h.v * h.v) / (h.t * ((double)M * M)) * tc_corr(h, k) * tc_corr(h, l);
}
DEV double tc_three_body_diag(const HamDev &h, u64 fa, int nb) {
    double value = 0.0;
    // p over occupied modes ascending, q over those below p (ascending) -- same order as the reference loop
    u64 fp = fa;
    while (fp) {
        int pm = __ffsll((long long)fp); fp &= fp - 1; // 1-based mode
        u64 fq = fa & ((1ull << (pm - 1)) - 1);
        while (fq) {
            int qm = __ffsll((long long)fq); fq &= fq - 1;
            int k = pm - qm;
            double qkk = tc_q_function(h, -k, k);
            value += 2 * qkk * nb;
        }
    }
    return value;
}

// two-component momentum transfer (excitations.jl:82-119). i is 0-based.
DEV double mom_transfer_2c(int M, u64 &fa, u64 &fb, int Nb, long long i64_, bool fold, int &p, int &q, int &mk) {
    const unsigned i = (unsigned)i64_, nb = (unsigned)Nb; // indices are < 2^31 (checked at rimu_ham_create)
    const unsigned per_a = (unsigned)(M - 1) * nb;
    const unsigned qa = udiv_small(i, per_a); // i < N1*N2*(M-1) <= 2^15, per_a < 2^10
    int src_a = (int)qa;
    const unsigned rem = i - qa * per_a;
    const unsigned qb = udiv_small(rem, nb);
    int dst_a = (int)qb + 1; // 1..M-1
    int src_b = (int)(rem - qb * nb);
    int src_a_mode = select_(fa, src_a) + 1, src_b_mode = select_(fb, src_b) + 1;
    if (dst_a >= src_a_mode) dst_a += 1;
    int mom = dst_a - src_a_mode;
    int dst_b = src_b_mode - mom;
    p = src_a_mode; q = src_b_mode; mk = -mom;
    if (fold) {
        if (dst_b < 1) dst_b += M; else if (dst_b > M) dst_b -= M;
    } else if (dst_b < 1 || dst_b > M) return 0.0;
    u64 ta = fa, tb = fb;
    int ca = 0, cb = 0;
    fermi_destroy(ta, src_a_mode, ca);
    if (!fermi_create(ta, dst_a, ca)) return 0.0;
    fermi_destroy(tb, src_b_mode, cb);
    if (!fermi_create(tb, dst_b, cb)) return 0.0;
    fa = ta; fb = tb;
    return parity_sign(ca) * parity_sign(cb);
}

// three-body term (excitations.jl:199-238), fermions. i 0-based; (p,q,s,p_k,q_l) first index fastest.
DEV double tc_three_body(int M, u64 &fa, u64 &fb, int N1, int N2, long long i64_, int &k, int &l) {
    unsigned i = (unsigned)i64_; // indices are < 2^31 (checked at rimu_ham_create)
    int p = (int)(i % (unsigned)N1); i /= (unsigned)N1;
    int q = (int)(i % (unsigned)(N1 - 1)); i /= (unsigned)(N1 - 1);
    int s = (int)(i % (unsigned)N2); i /= (unsigned)N2;
    int p_k = (int)(i % (unsigned)M) + 1; i /= (unsigned)M;
    int q_l = (int)(i % (unsigned)M) + 1;
    if (q >= p) q += 1;
    int pm = select_(fa, p) + 1, qm = select_(fa, q) + 1, sm = select_(fb, s) + 1;
    k = pm - p_k; l = q_l - qm;
    int s_kl = sm + k - l;
    if (k == 0 || l == 0) return 0.0;
    if (pm == q_l && qm == p_k) return 0.0;
    if (s_kl > M || s_kl < 1) return 0.0;
    u64 ta = fa, tb = fb;
    int ca = 0, cb = 0;
    // destructions (q,p) applied last-first: p then q; creations (p_k,q_l): q_l then p_k
    if (!fermi_destroy(ta, pm, ca)) return 0.0;
    if (!fermi_destroy(ta, qm, ca)) return 0.0;
    if (!fermi_create(ta, q_l, ca)) return 0.0;
    if (!fermi_create(ta, p_k, ca)) return 0.0;
    if (!fermi_destroy(tb, sm, cb)) return 0.0;
    if (!fermi_create(tb, s_kl, cb)) return 0.0;
    fa = ta; fb = tb;
    return parity_sign(ca) * parity_sign(cb);
}

// ---------------------------------------------------------------- the Hamiltonian interface
// ---------------------------------------------------------------- general CompositeFS (HK_RS_COMP)
template <class B> DEV B comp_get(const HamDev &h, B x, int c) { return (B)(x >> h.coff[c]) & lowmask<B>(h.cbits[c]); }
template <class B> DEV B comp_put(const HamDev &h, B x, int c, B xc) {
    return (x & ~(lowmask<B>(h.cbits[c]) << h.coff[c])) | (xc << h.coff[c]);
}
template <class B> DEV int comp_num_occupied(const HamDev &h, B xc, int c) { return h.cbose[c] ? bose_num_occupied(xc) : popc_(xc); }
// occupation of 0-based mode md in component c
template <class B> DEV int comp_occupation(const HamDev &h, B xc, int c, int md) {
    if (!h.cbose[c]) return (int)((u64)(xc >> md) & 1ull);
    const int off = bose_mode_offset(xc, md + 1);
    return cto_((B)(xc >> off));
}
// dot(occupied_modes(a), occupied_modes(b)) = sum_m n_a(m) n_b(m) (fockaddress.jl:692-719); an exact integer, any order
template <class B> DEV long long comp_overlap(const HamDev &h, B x, int a, int b) {
    const B xa = comp_get(h, x, a), xb = comp_get(h, x, b);
    if (!h.cbose[a] && !h.cbose[b]) return popc_((B)(xa & xb));
    if (!h.cbose[a]) { // iterate the fermions, look the boson occupation up
        long long r = 0; u64 f = (u64)xa;
        while (f) { int md = __ffsll((long long)f) - 1; f &= f - 1; r += comp_occupation(h, xb, b, md); }
        return r;
    }
    long long r = 0; int md, n;
    for (BoseModes<B> it(xa); it.next(md, n);) r += (long long)n * comp_occupation(h, xb, b, md);
    return r;
}

template <int HKX, class B> DEV double ham_diagonal(const HamDev &h, B x) {
    constexpr int HK = HkBase<HKX>::value; // (plain kinds run their base kind's code with variant == 0 folded in)
    const int M = h.M;
    if constexpr (HK == HK_REAL1D_BOSE) {
        if (HVARIANT(h) == 0) return h.u * (double)bose_interaction(x) / 2;
        if (HVARIANT(h) == 1) { // HubbardReal1DEP.jl:82-87: sum over occupied modes (ascending) of u n (n-1) / 2 + eps[mode] n
            double s = 0.0; bool first = true;
            int md, n;
            for (BoseModes<B> it(x); it.next(md, n);) {
                const double term = h.u * n * (n - 1) / 2 + h.pot[md] * n;
                s = first ? term : s + term; first = false;
            }
            return s;
        }
        // ExtendedHubbardReal1D.jl:101-126: u sum n(n-1) / 2 + v sum n_j n_j+1 (ring unless hard wall)
        long long ext = 0, reg = 0;
        int md, n, pmode = -1, pocc = 0, first_mode = -1, first_occ = 0;
        for (BoseModes<B> it(x); it.next(md, n);) { // md = 0-based mode of this occupied block
            if (pmode == md - 1) ext += (long long)pocc * n;
            reg += (long long)n * (n - 1);
            if (first_mode < 0) { first_mode = md; first_occ = n; }
            pmode = md; pocc = n;
        }
        if (h.bc != 1 /* hard wall */ && pmode >= 0) ext += (long long)(pmode == M - 1 ? pocc : 0) * (first_mode == 0 ? first_occ : 0);
        return h.u * (double)reg / 2 + h.v * (double)ext;
    } else if constexpr (HK == HK_MOM1D_BOSE) {
        double ke = 0.0;
        int lin = 0, md, n;
        for (BoseModes<B> it(x); it.next(md, n);) {
            ke += h.kes[md] * n;
            lin += n * (n - 1);
        }
        // sum n(n-1) + 4 sum_{i>j} n_i n_j with sum n = N and sum n^2 = lin + N
        const int ntot = h.N0;
        const long long onproduct = (long long)lin + 2LL * ((long long)ntot * ntot - lin - ntot);
        double value = ke + h.u_2m * (double)onproduct;
        if (HVARIANT(h) == 1) { // + (v / M) * extended_momentum_transfer_diagonal(map, 2pi / M)  (excitations.jl:145-156)
            double ext = 0.0;
            int mi, ni;
            for (BoseModes<B> it(x); it.next(mi, ni);) {
                ext += (double)(ni * (ni - 1));
                int mj, nj;
                for (BoseModes<B> jt(x); jt.next(mj, nj) && mj < mi;) ext += (double)(2 * ni * nj) * (1 + h.us[mi - mj]);
            }
            value += h.v_m * ext;
        } else if (HVARIANT(h) == 2) value += (double)ntot * h.pot[0]; // momentum_external_potential_diagonal (excitations.jl:274-279)
        return value;
    } else if constexpr (HK == HK_MOM1D_F2C) {
        u64 mask = (1ull << M) - 1, fa = (u64)x & mask, fb = ((u64)x >> M) & mask;
        double ka = kes_sum(h.kes, fa), kb = kes_sum(h.kes, fb);
        double value = ka + kb + h.u_2m * (double)(2 * __popcll(fa) * __popcll(fb));
        if (HVARIANT(h) == 2) value = value + (double)__popcll(fa) * h.pot[0] + (double)__popcll(fb) * h.pot[0];
        return value;
    } else if constexpr (HK == HK_RS_BOSE) {
        double interaction = h.umat_zero ? 0.0 : h.u00 * (double)bose_interaction(x) / 2;
        double pot = 0.0;
        if (h.has_pot) {
            int md, n; double pe = 0.0;
            for (BoseModes<B> it(x); it.next(md, n);) pe += n * h.pot[md];
            pot += pe;
        }
        return interaction + pot;
    } else if constexpr (HK == HK_RS_FERMI) {
        double pot = 0.0;
        if (h.has_pot) {
            u64 f = (u64)x; double pe = 0.0;
            while (f) { int b = __ffsll((long long)f) - 1; f &= f - 1; pe += 1 * h.pot[b]; }
            pot += pe;
        }
        return 0.0 + pot;
    } else if constexpr (HK == HK_RS_F2C) {
        u64 mask = (1ull << M) - 1, fa = (u64)x & mask, fb = ((u64)x >> M) & mask;
        double interaction = h.umat_zero ? 0.0 : (0.0 + h.u10 * (double)__popcll(fa & fb)) + (0.0 + 0.0);
        double pot = 0.0;
        if (h.has_pot) {
            double pe = 0.0; u64 f = fa;
            while (f) { int b = __ffsll((long long)f) - 1; f &= f - 1; pe += 1 * h.pot[b]; }
            pot += pe;
            pe = 0.0; f = fb;
            while (f) { int b = __ffsll((long long)f) - 1; f &= f - 1; pe += 1 * h.pot[M + b]; }
            pot += pe;
        }
        return interaction + pot;
    } else if constexpr (HK == HK_RS_COMP) {
        // local_interaction(::CompositeFS, u) = _interactions(components, u) (HubbardRealSpace.jl:18-75):
        // (self_1 + row_1) + ((self_2 + row_2) + (... + 0.0)), row_i = u[i+1,i] n_i.n_{i+1} + (u[i+2,i] n_i.n_{i+2} + (... + 0))
        const int C = h.ncomp;
        double interaction = 0.0;
        if (!h.umat_zero) {
            for (int i = C - 1; i >= 0; i--) {
                const B xi = comp_get(h, x, i);
                double row = 0.0;
                for (int j = C - 1; j > i; j--) row = h.umat[j + C * i] * (double)comp_overlap(h, x, i, j) + row;
                const double self = h.cbose[i] ? h.umat[i + C * i] * (double)bose_interaction(xi) / 2 : 0.0;
                interaction = (self + row) + interaction;
            }
        }
        double pot = 0.0;
        if (h.has_pot) { // external_potential(::CompositeFS, pot::Matrix) (:99-106): per component, occupied modes ascending
            for (int c = 0; c < C; c++) {
                const B xc = comp_get(h, x, c);
                double pe = 0.0;
                if (h.cbose[c]) { int md, n; for (BoseModes<B> it(xc); it.next(md, n);) pe += n * h.pot[c * M + md]; }
                else { u64 f = (u64)xc; while (f) { int b = __ffsll((long long)f) - 1; f &= f - 1; pe += 1 * h.pot[c * M + b]; } }
                pot += pe;
            }
        }
        return interaction + pot;
    } else { // HK_TC_F2C
        u64 mask = (1ull << M) - 1, fa = (u64)x & mask, fb = ((u64)x >> M) & mask;
        int n1 = __popcll(fa), n2 = __popcll(fb);
        double k1 = kes_sum(h.kes, fa), k2 = kes_sum(h.kes, fb);
        double mtd = (double)(2 * n1 * n2) * (h.v / M + 2 * h.v * h.v / h.t * tc_w(h, 0)) / 2;
        double value = k1 + k2 + mtd;
        if (h.three_body) value += tc_three_body_diag(h, fa, n2) + tc_three_body_diag(h, fb, n1);
        return value;
    }
}

template <int HKX, class B> DEV long long ham_num_offdiagonals(const HamDev &h, B x) {
    constexpr int HK = HkBase<HKX>::value; // (plain kinds run their base kind's code with variant == 0 folded in)
    const int M = h.M;
    if constexpr (HK == HK_REAL1D_BOSE) {
        return 2LL * bose_num_occupied(x);
    } else if constexpr (HK == HK_MOM1D_BOSE) {
        long long s = bose_num_occupied(x), d = bose_num_doubly(x);
        return s * (s - 1) * (M - 2) + d * (M - 1) + (HVARIANT(h) == 2 ? s * (M - 1) : 0);
    } else if constexpr (HK == HK_MOM1D_F2C) {
        return (long long)h.N0 * h.N1 * (M - 1) + (HVARIANT(h) == 2 ? (long long)(h.N0 + h.N1) * (M - 1) : 0);
    } else if constexpr (HK == HK_RS_BOSE) {
        return (long long)bose_num_occupied(x) * h.nnb;
    } else if constexpr (HK == HK_RS_FERMI) {
        return (long long)__popcll((u64)x) * h.nnb;
    } else if constexpr (HK == HK_RS_F2C) {
        return (long long)__popcll((u64)x) * h.nnb; // both components
    } else if constexpr (HK == HK_RS_COMP) {
        long long s = 0;
        for (int c = 0; c < h.ncomp; c++) s += comp_num_occupied(h, comp_get(h, x, c), c);
        return s * h.nnb;
    } else {
        long long N1 = h.N0, N2 = h.N1;
        long long n = N1 * N2 * (M - 1);
        if (h.three_body) n += N1 * (N1 - 1) * N2 * M * M + N2 * (N2 - 1) * N1 * M * M;
        return n;
    }
}

// returns H_{out,x} for the i-th (0-based) off-diagonal; out == x whenever the value is 0
template <int HKX, class B> DEV double ham_offdiagonal(const HamDev &h, B x, long long i, B &out) {
    constexpr int HK = HkBase<HKX>::value; // (plain kinds run their base kind's code with variant == 0 folded in)
    const int M = h.M;
    out = x;
    if constexpr (HK == HK_REAL1D_BOSE) {
        int mode, ns, off;
        bose_kth_occupied(x, (int)(i >> 1), mode, ns, off);
        int dst = (i & 1) ? (mode == 1 ? M : mode - 1) : (mode == M ? 1 : mode + 1); // chosen odd <=> i even: hop right
        B y = delete_bit(x, off);
        int nd = bose_create(y, dst);
        out = y;
        double val = sqrt((double)(ns * nd));
        if (HVARIANT(h) == 2) { // hopnextneighbour(b, i, boundary_condition) bosefs.jl:355-369
            const bool on_boundary = (i & 1) ? mode == 1 : mode == M;
            if (on_boundary && h.bc == 2) val = -val;
            else if (on_boundary && h.bc == 1) { val = 0.0; out = x; }
        }
        return -h.t * val;
    } else if constexpr (HK == HK_MOM1D_BOSE) {
        const int s = bose_num_occupied(x);
        const int ndiff = s * (s - 1) * (M - 2);
        const int ii = (int)i;
        if (HVARIANT(h) == 2) { // HubbardMom1DEP: the external-potential block follows the momentum-transfer block
            const int nmom = ndiff + bose_num_doubly(x) * (M - 1);
            if (ii >= nmom) { // momentum_external_potential_excitation (excitations.jl:257-267): a^dagger_q a_p, q != p
                const unsigned e = (unsigned)(ii - nmom), mm1 = (unsigned)(M - 1);
                const unsigned p = udiv_small(e, mm1);
                int q = (int)(e - p * mm1) + 1, pmode, np_, poff;
                bose_kth_occupied(x, (int)p, pmode, np_, poff);
                if (q >= pmode) q += 1;
                int k = pmode - q;
                if (k < 0) k += M;
                B y = delete_bit(x, poff);
                const int nq = bose_create(y, q);
                out = y;
                return sqrt((double)(np_ * nq)) * h.pot[k];
            }
        }
        int src0, src1, off0, off1, n0, n1, mom;
        if (ii >= ndiff) { // both particles from one mode with n >= 2
            const unsigned dbl = (unsigned)(ii - ndiff), mm1 = (unsigned)(M - 1);
            const int d = (int)udiv_small(dbl, mm1);
            mom = (int)(dbl - (unsigned)d * mm1) + 1;
            bose_kth_doubly(x, d, src0, n0, off0);
            src1 = src0; off1 = off0; n1 = n0; n0 = n0 - 1; // a_src1 first: n, then a_src0 on the same mode: n - 1
        } else {
            const unsigned mm2 = (unsigned)(M - 2), sm1 = (unsigned)(s - 1);
            const unsigned pair = udiv_small((unsigned)ii, mm2);
            mom = (int)((unsigned)ii - pair * mm2) + 1;
            const unsigned fq = udiv_small(pair, sm1);
            int fst = (int)fq + 1, snd = (int)(pair - fq * sm1) + 1; // 1-based as in fldmod1
            int f_hole, s_hole;
            if (snd < fst) { f_hole = snd; s_hole = fst; } else { f_hole = fst; s_hole = snd + 1; }
            bose_kth_occupied(x, f_hole - 1, src0, n0, off0);
            bose_kth_occupied(x, s_hole - 1, src1, n1, off1);
            if (mom >= src1 - src0) mom += 1;
        }
        int dst0 = src0 + mom, dst1 = src1 - mom;
        if (dst0 > M) dst0 -= M;
        if (dst1 < 1) dst1 += M;
        // excitation(add, dst, src): destroy src[1], src[0]; create dst[1], dst[0].  off0 <= off1, so removing the
        // particle at off1 first leaves off0 valid.
        B y = delete_bit(delete_bit(x, off1), off0);
        int value = n1 * n0;
        value *= bose_create(y, dst1);
        value *= bose_create(y, dst0);
        out = y;
        if (HVARIANT(h) == 1) { // ExtendedHubbardMom1D.jl:99-102: u * onproduct / 2M + v * cos(q * 2pi / M) * onproduct / M, q = -mom
            const double op = sqrt((double)value);
            return h.u * op / (2 * M) + h.v * h.ws[mom] * op / M;
        }
        return h.u_2m * sqrt((double)value);
    } else if constexpr (HK == HK_MOM1D_F2C) {
        u64 mask = (1ull << M) - 1, fa = (u64)x & mask, fb = ((u64)x >> M) & mask;
        if (HVARIANT(h) == 2) { // HubbardMom1DEP.jl:223-257: then N1 (M-1) one-body moves of component a, N2 (M-1) of component b
            const long long nmom = (long long)h.N0 * h.N1 * (M - 1);
            if (i >= nmom) {
                unsigned e = (unsigned)(i - nmom);
                const unsigned mm1 = (unsigned)(M - 1);
                const int comp = e >= (unsigned)h.N0 * mm1 ? 1 : 0;
                if (comp) e -= (unsigned)h.N0 * mm1;
                u64 f = comp ? fb : fa;
                const unsigned p = udiv_small(e, mm1);
                int q = (int)(e - p * mm1) + 1;
                const int pmode = select_(f, (int)p) + 1;
                if (q >= pmode) q += 1;
                int k = pmode - q;
                if (k < 0) k += M;
                int cnt = 0;
                fermi_destroy(f, pmode, cnt);
                if (!fermi_create(f, q, cnt)) return 0.0;
                if (comp) fb = f; else fa = f;
                out = (B)(fa | (fb << M));
                return parity_sign(cnt) * h.pot[k];
            }
        }
        int p, q, mk;
        double val = mom_transfer_2c(M, fa, fb, h.N1, i, true, p, q, mk);
        if (val != 0.0) out = (B)(fa | (fb << M));
        return h.u_m * val;
    } else if constexpr (HK == HK_RS_BOSE) {
        const unsigned ii = (unsigned)i, nnb = (unsigned)h.nnb;
        const unsigned pq = udiv_small(ii, nnb);
        int particle = (int)pq, neigh = (int)(ii - pq * nnb);
        int mode, ns, off;
        bose_kth_occupied(x, particle, mode, ns, off);
        int dst = h.nbr[(mode - 1) * h.nnb + neigh];
        if (dst == 0) return 0.0;
        B y = delete_bit(x, off);
        int nd = bose_create(y, dst);
        out = y;
        return -h.tc0 * sqrt((double)(ns * nd));
    } else if constexpr (HK == HK_RS_FERMI || HK == HK_RS_F2C) {
        u64 mask = (M >= 64) ? ~0ull : ((1ull << M) - 1);
        u64 fa = (u64)x & mask, fb = (HK == HK_RS_F2C) ? (((u64)x >> M) & mask) : 0;
        long long na = (long long)__popcll(fa) * h.nnb;
        int comp = 0;
        if (HK == HK_RS_F2C && i >= na) { comp = 1; i -= na; }
        u64 f = comp ? fb : fa;
        const unsigned ii = (unsigned)i, nnb = (unsigned)h.nnb;
        const unsigned pq = udiv_small(ii, nnb);
        int particle = (int)pq, neigh = (int)(ii - pq * nnb);
        int mode = select_(f, particle) + 1;
        int dst = h.nbr[(mode - 1) * h.nnb + neigh];
        if (dst == 0) return 0.0;
        int cnt = 0;
        fermi_destroy(f, mode, cnt);
        if (!fermi_create(f, dst, cnt)) return 0.0;
        if (comp) fb = f; else fa = f;
        out = (HK == HK_RS_F2C) ? (B)(fa | (fb << M)) : (B)fa;
        return -(comp ? h.tc1 : h.tc0) * parity_sign(cnt);
    } else if constexpr (HK == HK_RS_COMP) {
        // HubbardRealSpaceOffdiagonals (HubbardRealSpace.jl:340-391): the components' hop lists one after the other
        int c = 0;
        B xc = comp_get(h, x, 0);
        for (;;) {
            const long long nc = (long long)comp_num_occupied(h, xc, c) * h.nnb;
            if (i < nc || c == h.ncomp - 1) break;
            i -= nc; c++;
            xc = comp_get(h, x, c);
        }
        const unsigned ii = (unsigned)i, nnb = (unsigned)h.nnb;
        const unsigned pq = udiv_small(ii, nnb);
        const int particle = (int)pq, neigh = (int)(ii - pq * nnb);
        double value;
        if (h.cbose[c]) {
            int mode, ns, off;
            bose_kth_occupied(xc, particle, mode, ns, off);
            const int dst = h.nbr[(mode - 1) * h.nnb + neigh];
            if (dst == 0) return 0.0;
            B y = delete_bit(xc, off);
            const int nd = bose_create(y, dst);
            xc = y;
            value = sqrt((double)(ns * nd));
        } else {
            u64 f = (u64)xc;
            const int mode = select_(f, particle) + 1;
            const int dst = h.nbr[(mode - 1) * h.nnb + neigh];
            if (dst == 0) return 0.0;
            int cnt = 0;
            fermi_destroy(f, mode, cnt);
            if (!fermi_create(f, dst, cnt)) return 0.0;
            xc = (B)f;
            value = parity_sign(cnt);
        }
        out = comp_put(h, x, c, xc);
        return -h.tcs[c] * value;
    } else { // HK_TC_F2C
        u64 mask = (1ull << M) - 1, fa = (u64)x & mask, fb = ((u64)x >> M) & mask;
        long long N1 = h.N0, N2 = h.N1;
        long long n_mom = N1 * N2 * (M - 1);
        long long n1 = h.three_body ? N1 * (N1 - 1) * N2 * M * M : 0;
        double value;
        if (i < n_mom) {
            int p, q, mk;
            value = mom_transfer_2c(M, fa, fb, (int)N2, i, false, p, q, mk);
            if (value != 0.0) value *= tc_t_function(h, p, q, mk);
        } else if (i < n_mom + n1) {
            int k, l;
            value = tc_three_body(M, fa, fb, (int)N1, (int)N2, i - n_mom, k, l);
            value *= tc_q_function(h, k, l);
        } else {
            int k, l;
            value = tc_three_body(M, fb, fa, (int)N2, (int)N1, i - n_mom - n1, k, l);
            value *= tc_q_function(h, k, l);
        }
        if (value != 0.0) out = (B)(fa | (fb << M));
        return value;
    }
}
#endif // __CUDACC__
