// host_ham.cpp -- the DEVICE address arithmetic (rimu.jl_b200/csrc/common.cuh + hamiltonians.cuh), compiled for the host.
//
// The bit-trick implementations the CUDA kernels use (rank/select occupied-mode lookups, single-bit boson moves, popcount
// fermion signs, float-reciprocal index decoding) are plain C++ apart from a handful of intrinsics; this file supplies
// those intrinsics and exports diagonal_element / num_offdiagonals / get_offdiagonal so that tests/test_host_emulation.py
// can compare them with the ONR-based oracle for every model WITHOUT a GPU.  Build: g++ -std=c++17 -O2 -ffp-contract=off
// (no FMA contraction, like nvcc --fmad=false).  Test infrastructure only.
#include <cmath>
#include <cstdint>
#include <cstring>

#define RIMU_HOST_EMULATION 1
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __popc(unsigned int x) { return __builtin_popcount(x); }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned long long __brevll(unsigned long long x) {
    unsigned long long r = 0;
    for (int i = 0; i < 64; i++) { r = (r << 1) | (x & 1ull); x >>= 1; }
    return r;
}
static inline unsigned int __float2uint_rz(float x) { return (unsigned int)x; }   // truncation toward zero
static inline float __frcp_rn(float x) { return 1.0f / x; }                      // IEEE round-to-nearest division

#include "../../rimu.jl_b200/csrc/ham_host.h" // brings hamiltonians.cuh; the same host code rimu_ham_create runs
#include "../../rimu.jl_b200/csrc/step_math.cuh" // the per-deposit arithmetic the spawn and merge kernels call

template <class F> static auto dispatch(int hk, int W, F &&f) {
    switch (hk) {
    case HK_REAL1D_BOSE: return W == 1 ? f(std::integral_constant<int, HK_REAL1D_BOSE>(), u64()) : f(std::integral_constant<int, HK_REAL1D_BOSE>(), u128());
    case HK_MOM1D_BOSE: return W == 1 ? f(std::integral_constant<int, HK_MOM1D_BOSE>(), u64()) : f(std::integral_constant<int, HK_MOM1D_BOSE>(), u128());
    case HK_REAL1D_BOSE_PLAIN: return W == 1 ? f(std::integral_constant<int, HK_REAL1D_BOSE_PLAIN>(), u64()) : f(std::integral_constant<int, HK_REAL1D_BOSE_PLAIN>(), u128());
    case HK_MOM1D_BOSE_PLAIN: return W == 1 ? f(std::integral_constant<int, HK_MOM1D_BOSE_PLAIN>(), u64()) : f(std::integral_constant<int, HK_MOM1D_BOSE_PLAIN>(), u128());
    case HK_MOM1D_F2C: return f(std::integral_constant<int, HK_MOM1D_F2C>(), u64());
    case HK_RS_BOSE: return W == 1 ? f(std::integral_constant<int, HK_RS_BOSE>(), u64()) : f(std::integral_constant<int, HK_RS_BOSE>(), u128());
    case HK_RS_FERMI: return f(std::integral_constant<int, HK_RS_FERMI>(), u64());
    case HK_RS_F2C: return f(std::integral_constant<int, HK_RS_F2C>(), u64());
    case HK_RS_COMP: return W == 1 ? f(std::integral_constant<int, HK_RS_COMP>(), u64()) : f(std::integral_constant<int, HK_RS_COMP>(), u128());
    default: return f(std::integral_constant<int, HK_TC_F2C>(), u64());
    }
}
template <class B> static B load(const uint64_t *k) { return (B)k[0]; }
template <> u128 load<u128>(const uint64_t *k) { return ((u128)k[1] << 64) | (u128)k[0]; }
static void store(uint64_t *k, u64 x) { k[0] = x; }
static void store(uint64_t *k, u128 x) { k[0] = (u64)x; k[1] = (u64)(x >> 64); }

struct EmuHam { HamHostImage img; };

extern "C" {
// rimu_ham_create without the upload: validation, scalars and tables exactly as the product builds them
int emu_ham_create(const rimu_ham_desc *d, EmuHam **out, char *err, int errlen) {
    EmuHam *h = new EmuHam();
    int rc = ham_build_host(d, &h->img);
    if (rc) { snprintf(err, errlen, "%s", h->img.error.c_str()); delete h; return rc; }
    ham_set_tables(&h->img.dev, h->img.tables.data(), h->img.nbr.empty() ? nullptr : h->img.nbr.data());
    *out = h;
    return 0;
}
void emu_ham_destroy(EmuHam *h) { delete h; }
int emu_ham_words(const EmuHam *h) { return h->img.W; }
int emu_ham_kind(const EmuHam *h) { return h->img.hk; }
double emu_diagonal(const EmuHam *eh, const uint64_t *key) {
    const HamDev *h = &eh->img.dev;
    return dispatch(h->hk, eh->img.W, [&](auto hk, auto b) { typedef decltype(b) B; return ham_diagonal<decltype(hk)::value, B>(*h, load<B>(key)); });
}
long long emu_num_offdiagonals(const EmuHam *eh, const uint64_t *key) {
    const HamDev *h = &eh->img.dev;
    return dispatch(h->hk, eh->img.W, [&](auto hk, auto b) { typedef decltype(b) B; return ham_num_offdiagonals<decltype(hk)::value, B>(*h, load<B>(key)); });
}
// i is 0-based (= the reference's chosen - 1)
double emu_offdiagonal(const EmuHam *eh, const uint64_t *key, long long i, uint64_t *key_out) {
    const HamDev *h = &eh->img.dev;
    return dispatch(h->hk, eh->img.W, [&](auto hk, auto b) {
        typedef decltype(b) B;
        B out;
        double v = ham_offdiagonal<decltype(hk)::value, B>(*h, load<B>(key), i, out);
        store(key_out, out);
        return v;
    });
}
// All deposits one parent makes in a step, computed with the DEVICE functions (step_math.cuh) in the order the kernels use
// them: attempts_for -> for every attempt spawn_attempt (Philox draw, off-diagonal, projection) + deposit_lane; plus the
// diagonal deposit (the few lines the merge kernel applies to a parent, restated here around project_value / rng_draw).
// Values are returned as doubles (exact for Int64 walkers below 2^53).  Returns the number of non-zero spawn records.
struct EmuStep { int style, plain_h; double shift, dtau, boost, proj_thr, rel_thr, abs_thr; uint32_t k0, k1; int init_rule; double init_thr; };
long long emu_parent_deposits(const EmuHam *eh, const EmuStep *es, const uint64_t *key, double val, int is_int, long long cap,
                              uint64_t *child_keys, double *child_vals, int *child_lanes,
                              double *diag_val, int *diag_lane, long long *attempts, int *exact_out, double *spawned_sum) {
    const HamDev *h = &eh->img.dev;
    StepDev p;
    memset(&p, 0, sizeof(p));
    p.style = es->style; p.plain_h = es->plain_h; p.shift = es->shift; p.dtau = es->dtau; p.boost = es->boost;
    p.proj_thr = es->proj_thr; p.rel_thr = es->rel_thr; p.abs_thr = es->abs_thr; p.k0 = es->k0; p.k1 = es->k1;
    p.nranks = 1; p.init_rule = es->init_rule; p.init_thr = es->init_thr;
    const int W = eh->img.W;
    return dispatch(h->hk, W, [&](auto hk, auto b) -> long long {
        typedef decltype(b) B;
        constexpr int HK = decltype(hk)::value;
        constexpr int WW = sizeof(B) / 8;
        const B k = load<B>(key);
        const u64 hkey = hash_bits(k);
        // diagonal step (partition.cuh merge_kernel staging; spawning.jl:73-77, fciqmc.jl:93-96)
        const double hd = ham_diagonal<HK, B>(*h, k);
        const double d = p.plain_h ? hd : 1 - p.dtau * (hd - p.shift);
        double rr = 0.0;
        const double thr = is_int ? 0.0 : p.proj_thr;
        if (is_int || thr > 0.0) { u32 rnd[4]; rng_draw(hkey, 0, STREAM_DIAG, p.k0, p.k1, rnd); rr = u53(rnd[1], rnd[2]); }
        *diag_val = is_int ? (double)project_value<i64>(d * val, thr, rr) : project_value<double>(d * val, thr, rr);
        *diag_lane = (int)deposit_lane(p, true, val);
        // spawning
        const long long L = ham_num_offdiagonals<HK, B>(*h, k);
        u64 nat = 0;
        const bool exact = attempts_for(p, val, L, nat);
        *attempts = (long long)nat; *exact_out = exact;
        long long nrec = 0;
        double ssum = 0.0;
        for (u64 a = 0; a < nat; a++) {
            B child; long long ci; double sp;
            double nv;
            if (is_int) nv = (double)spawn_attempt<HK, WW, i64>(*h, p, k, hkey, val, L, nat, exact, a, child, ci, sp);
            else nv = spawn_attempt<HK, WW, double>(*h, p, k, hkey, val, L, nat, exact, a, child, ci, sp);
            ssum += sp;
            if (nv != 0.0) {
                if (nrec < cap) { store(child_keys + nrec * W, child); child_vals[nrec] = nv; child_lanes[nrec] = (int)deposit_lane(p, false, val); }
                nrec++;
            }
        }
        *spawned_sum = ssum;
        return nrec;
    });
}
// select_ / udiv_small are the two primitives everything else leans on: exported for exhaustive checks
int emu_select64(uint64_t v, int k) { return select_((u64)v, k); }
unsigned emu_udiv_small(unsigned x, unsigned d) { return udiv_small(x, d); }
}
