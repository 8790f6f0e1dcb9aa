// step_math.cuh -- the arithmetic of ONE deposit of the FCIQMC step, free of any CUDA-only construct: sign, projection
// (projected_deposit!, spawning.jl:9-45), the exact/stochastic decision and attempt count (spawning.jl:232-243,364-378),
// the initiator lane of a deposit (initiators.jl:142-158) and one spawn attempt.  The kernels (kernels.cuh, partition.cuh)
// call these; tests/cuda/host_ham.cpp compiles the same code for the host so that the CPU suite checks it against the
// oracle without a GPU.
#pragma once
#include "hamiltonians.cuh"
#include <type_traits>

// Device-resident controller of a BATCH of steps (rimu_advance): the shift update of every step runs on the GPU, so the host
// enqueues step k+1 before step k has finished -- nobody waits for the walker number to cross PCIe.  Null in rimu_step.
struct StepCtl {
    double shift, pnorm;   // DefaultShiftParameters (shiftstrategy.jl:32-38) after the last finished step
    int shift_mode;        // ... its shift_mode flag (the *AfterTargetWalkers strategies)
    int stop;              // 0 running; 1 the run ended after steps_done steps (dead population, max_length, strategy);
                           // 2 step number steps_done ran out of working memory / vector capacity: every later kernel is a no-op
    long long steps_done;
    unsigned long long n;  // entries of the current source vector
};

struct StepDev {
    int style, plain_h;
    double shift, dtau, boost, proj_thr, rel_thr, abs_thr, compress_thr;
    u32 k0, k1;
    int rank, nranks;
    int init_rule;      // RIMU_INITIATOR_*: 0 = no initiator lanes
    double init_thr;
    int ordered;        // order-deterministic Float64 summation (audit mode)
    const StepCtl *ctl; // batch mode: shift, source length and the stop flag come from device memory
};

#if defined(__CUDACC__) || defined(RIMU_HOST_EMULATION)
DEV double sgn_(double x) { return (double)((x > 0) - (x < 0)); }

// projected_deposit! (spawning.jl:9-45): returns the value actually deposited (0 = nothing)
template <class VT> DEV VT project_value(double val, double threshold, double r);
template <> DEV i64 project_value<i64>(double val, double, double r) {
    return (i64)sgn_(val) * (i64)floor(fabs(val) + r);
}
template <> DEV double project_value<double>(double val, double threshold, double r) {
    double a = fabs(val);
    if (a < threshold) val = (r < a / threshold) ? sgn_(val) * threshold : 0.0;
    return val;
}

// spawn!(::DynamicSemistochastic) decision (spawning.jl:364-378) + attempt count (spawning.jl:234)
DEV bool attempts_for(const StepDev &p, double val, long long L, u64 &n) {
    if (L <= 0) { n = 0; return false; }
    bool exact;
    if (p.style == 0) exact = true;
    else if (p.style == 2) {
        double thresh = fmin(p.abs_thr, (double)L);
        exact = p.boost * fabs(val) * p.rel_thr >= thresh;
    } else exact = false;
    if (exact) n = (u64)L;
    else {
        double f = floor(fabs(val) * p.boost);
        n = f < 1.0 ? 1ull : (u64)f;
    }
    return exact;
}

// lane of a deposit (to_initiator_value, DictVectors/initiators.jl:142-158): diagonal deposits of an initiator
// (|parent value| > threshold) are "initiator" (2), of anybody else "safe" (0); spawns of an initiator are "safe",
// of a non-initiator "unsafe" (1).  Without a rule everything is lane 0.
enum { LANE_SAFE = 0, LANE_UNSAFE = 1, LANE_INIT = 2 };
DEV u32 deposit_lane(const StepDev &p, bool diagonal, double parent_val) {
    if (p.init_rule == 0) return LANE_SAFE;
    const bool is_initiator = fabs(parent_val) > p.init_thr;
    if (diagonal) return is_initiator ? LANE_INIT : LANE_SAFE;
    return is_initiator ? LANE_SAFE : LANE_UNSAFE;
}

// one spawn attempt k of a parent (spawning.jl:174-182 Exact, :232-243 WithReplacement).
// Returns the value to deposit (0 = nothing); ci = off-diagonal index used, child = its address.
template <int HK, int W, class VT>
DEV VT spawn_attempt(const HamDev &h, const StepDev &p, typename BitsT<W>::type key, u64 hkey, double val,
                     long long L, u64 nat, bool exact, u64 k, typename BitsT<W>::type &child, long long &ci,
                     double &spawned) {
    typedef typename BitsT<W>::type B;
    constexpr bool is_int = std::is_integral<VT>::value;
    if (exact) {
        ci = (long long)k;
        double m = ham_offdiagonal<HK, B>(h, key, ci, child);
        if (!p.plain_h) m = -m * p.dtau;
        double r = 0.0;
        if (p.proj_thr > 0.0) {
            u32 rnd[4];
            rng_draw(hkey, k, STREAM_SPAWN, p.k0, p.k1, rnd);
            r = u53(rnd[1], rnd[2]);
        }
        double nv = project_value<double>(val * m, p.proj_thr, r);
        spawned = fabs(nv);
        if constexpr (is_int) return (VT)0; else return nv;
    }
    u32 rnd[4];
    rng_draw(hkey, k, STREAM_SPAWN, p.k0, p.k1, rnd);
    ci = (long long)(((u64)rnd[0] * (u64)L) >> 32);
    double m = ham_offdiagonal<HK, B>(h, key, ci, child);
    if (!p.plain_h) m = -m * p.dtau;
    // spawning.jl:237-241: magnitude = val / n; value = matrix element * magnitude / prob with prob = 1 / L.  The division by
    // 1 / L is written as the multiplication by L (one rounding instead of two; the oracle does the same -- north_star asks
    // 1e-12 for Float64, not the reference's last bit), and val / 1 is skipped.
    const double magnitude = nat == 1 ? val : val / (double)nat;
    const double nv0 = m * magnitude * (double)L;
    VT nv = project_value<VT>(nv0, is_int ? 0.0 : p.proj_thr, u53(rnd[1], rnd[2]));
    spawned = fabs((double)nv);
    return nv;
}

#endif
