// partition.cuh -- the partitioned FCIQMC step: bucket streams in HBM + shared-memory annihilation.
//
// Why (profiles/r1_baseline_ncu_summary.md): random read-modify-writes into one big HBM hash table are
// latency-bound (10-20 % of DRAM throughput) and move 4x the algorithmic bytes.  Here every HBM access is
// a stream:
//
//   walker vector : dense SoA (keys[n][W], vals[n]) that is SEGMENTED by bucket: the entries of bucket b
//                   are contiguous at [seg_start[b], seg_start[b]+seg_len[b]).  bucket(addr) is a fastrange
//                   of the address hash, so a determinant always lives in the same bucket.
//   record streams: rec[bucket][sub-stream][rcap] of 16-byte (W=1) / 32-byte (W=2) {address, value} records.  One
//                   sub-stream per (source rank, initiator lane): a sender owns the positions it writes, and the lane
//                   of a deposit (DictVectors/initiators.jl:22-45) is the stream index, never part of the record.
//   K1 spawn      : one CTA per chunk of 256 parents (the next chunk's parents are loaded meanwhile); attempt counts are
//                   scanned inside the CTA and the attempts are spread over its threads (one thread per spawn attempt).
//                   Every non-zero spawn is appended to the record stream of the CHILD's bucket: one counter atomic and
//                   ONE 16-byte store (two for W=2).  Children owned by another GPU are bucketed the same way into local
//                   staging streams.  Parents with more than HEAVY_T attempts go to a queue.
//   K2 heavy      : queue items are cut into tiles of HEAVY_TILE attempts, one CTA per tile; stochastic
//                   spawns of one parent are pre-summed per off-diagonal index in shared memory, so a
//                   determinant with 10^6 walkers emits at most L records per tile.
//   push (multi-GPU): one warp per (destination rank, lane, bucket) run copies the staged records into sub-stream
//                   [bucket][this rank, lane] of the OWNER's streams (peer stores over NVLink) and writes its fill.
//   K3 merge      : one CTA per bucket stages the bucket's parents (applying the diagonal step,
//                   spawning.jl:73-93) and the records of all its sub-streams in shared memory (metadata two buckets
//                   ahead, next bucket bulk-prefetched into L2), annihilates them there (open addressing on item
//                   indices claimed with 32-bit shared CAS; 64-bit shared adds only for genuine duplicates),
//                   collapses the initiator lanes (from_initiator_value), applies ThresholdCompression
//                   (compression.jl:18-26) as a dense pass and appends the survivors to the target vector, which is
//                   thereby segmented again.  walkernumber/length and the step statistics are reduced in the same
//                   pass (pdvec.jl:896-902).
//
// Algorithmic HBM bytes per step with P parents, A' records, U survivors, E = 8W+8:
//   K1: P*E read + A'*E written;  K3: P*(E+8) + A'*E read, U*(E+8) written (+8: cached diagonal elements).
#pragma once
#include "kernels.cuh"

#ifndef PART_NT
#define PART_NT 256        // threads of a merge CTA
#endif
#ifndef PART_CAP
#define PART_CAP 2048      // items (parents + records) a bucket may hold
#endif
#ifndef PART_MINB
#define PART_MINB 4        // merge CTAs resident per SM (register cap)
#endif
#ifndef SPAWN_NT
#define SPAWN_NT 256       // threads (= parents per chunk) of a spawn CTA
#endif
#ifndef SPAWN_MINB
#define SPAWN_MINB 4       // spawn CTAs resident per SM (register cap)
#endif
#ifndef PART_BALANCED_TAIL
#define PART_BALANCED_TAIL 1
#endif
#ifndef PART_ASYNC_PARENTS
#define PART_ASYNC_PARENTS 0 // (1 after validation on the GPU) merge: parents are staged with cp.async like the records (all loads of a bucket in flight at once)
#endif
// kernel-tuning builds only (-DRIMU_TUNE_FIXED_STEP): the step parameters of BASELINE config 2 (IsDynamicSemistochastic with late
// compression, one rank, no initiator rule, FirstOrderTransitionOperator) as compile-time constants -- measures what the runtime
// branches on them cost.  Shipping builds copy the parameter struct as it is.
#ifdef RIMU_TUNE_FIXED_STEP
#define RIMU_TUNE_FIX_STEP(p, p_in) StepDev p = p_in; p.style = 2; p.plain_h = 0; p.nranks = 1; p.rank = 0; p.init_rule = 0; p.ordered = 0; p.proj_thr = 0.0; p.ctl = nullptr;
#else
#define RIMU_TUNE_FIX_STEP(p, p_in) const StepDev p = step_view<FS>(p_in);
#endif
// FS ("fixed step parameters"): the kernels of the default production run -- IsDynamicSemistochastic with late compression on
// one GPU through rimu_step, no initiator rule, not ordered -- are instantiated a second time with those parameters as
// compile-time constants: the branches on style / plain_h / nranks / init_rule / proj_threshold fold away (spawn kernel
// 60 -> 54 registers).  rimu_step picks the instantiation when the call's parameters match (fixed_step_ok); results are identical.
template <bool FS> DEV StepDev step_view(const StepDev &in) {
    StepDev p = in;
    if constexpr (FS) { p.style = 2; p.plain_h = 0; p.nranks = 1; p.rank = 0; p.init_rule = 0; p.ordered = 0; p.proj_thr = 0.0; p.ctl = nullptr; }
    return p;
}
static inline bool fixed_step_ok(const StepDev &p, bool is_int) {
#ifdef RIMU_NO_FIXED_STEP
    (void)p; (void)is_int; return false;
#else
    return !is_int && p.style == 2 && !p.plain_h && p.nranks == 1 && !p.init_rule && !p.ordered && p.proj_thr == 0.0 && !p.ctl;
#endif
}
#define HEAVY_T 1024       // parents with more attempts than this are queued for K2
#define HEAVY_TILE 8192    // attempts per K2 work item
#define ACC_MAX 4096       // K2 pre-sums per off-diagonal index when L <= ACC_MAX

template <int W> struct PartCap { static constexpr int value = PART_CAP; };
#ifdef PART_CAP_W2 // tuning: a smaller bucket for two-word addresses (their items are 34 bytes of shared memory each)
template <> struct PartCap<2> { static constexpr int value = PART_CAP_W2; };
#endif

struct PartDev {       // bucket record streams (working memory of the partitioned step)
    u32 nb;            // buckets per rank (identical on all ranks of a step)
    u32 rcap;          // records per sub-stream
    u32 nsrc;          // sub-streams = destination ranks (1, or all ranks in direct mode) x lanes.  The SAME number of
                       //   sub-streams feeds every bucket of the merge: (source rank, lane)
    u32 nlane;         // 1, or 3 with an initiator rule: sub-stream (rank, lane) holds the safe / unsafe / initiator deposits
                       //   (DictVectors/initiators.jl:22-45) -- the lane travels in the stream index, not in the record
    u32 me;            // this rank's first sub-stream (rank * nlane; 0 when there is one rank)
    u64 *rec;          // [nsrc][nb][rcap] records, array of structures: W=1 {key, value} (16 B), W=2 {k0, k1, value, pad} (32 B).
                       //   Sub-stream (d, lane) holds what THIS rank produced for bucket b of rank d.
    u32 *rcnt;         // [nsrc][nb] fill of every sub-stream
    // direct mode (multi-GPU, CUDA IPC peer memory): nothing is pushed.  The owner's merge kernel PULLS sub-stream
    // (me, lane) of every rank's streams over NVLink while it annihilates -- the transfer overlaps the shared-memory work
    // of the other resident CTAs, there is no exchange pass and no receive pass.
    int direct;
    u32 ppc;           // parents per spawn chunk (<= SPAWN_NT; set per launch): small vectors are cut into smaller chunks so that
                       //   the spawn kernel's latency chain runs on every SM instead of a handful of CTAs
    const u64 *peer_rec[RIMU_MAX_RANKS];   // rec of every rank ([this rank] = rec)
    const u32 *peer_rcnt[RIMU_MAX_RANKS];  // rcnt of every rank
};
template <int W> struct RecWords { static constexpr int value = W == 1 ? 2 : 4; };
struct SegSrc {        // a segmented vector, read side (diag: cached diagonal elements or null)
    const u64 *keys; const u64 *vals; const u64 *seg_start; const u32 *seg_len; const double *diag;
};
struct SegDst {        // a segmented vector, write side
    u64 *keys; u64 *vals; u64 *seg_start; u32 *seg_len; u64 cap; double *diag;
};
struct HeavyItem { i64 parent; u64 nattempts; u64 tile_base; u64 exact; };
struct HeavyDev { HeavyItem *items; u64 *packed; /* (count << 32) | tiles */ u64 cap; };

#ifdef __CUDACC__
// local bucket of an address hash.  The owner rank is the integer part of x*R/2^32 (addr_owner); the
// bucket is a fastrange of the fractional part, so it is uniform within every rank's share.
DEV u32 bucket_of(u64 h, int nranks, u32 nb) {
    u32 x = (u32)(h >> 32);
    u32 y = nranks > 1 ? x * (u32)nranks : x;
    return __umulhi(y, nb);
}

// bulk L2 prefetch of [p, p + bytes): one instruction per contiguous range (cp.async.bulk.prefetch, 16-byte granules;
// the range is widened to 16-byte alignment, which stays inside the 256-byte granules cudaMalloc hands out)
DEV void l2_prefetch(const void *p, u64 bytes) {
    const u64 a = (u64)p & ~15ull;
    const u32 sz = (u32)((((u64)p + bytes + 15ull) & ~15ull) - a);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(sz) : "memory");
}

template <int W> DEV void store_rec(u64 *p, typename BitsT<W>::type key, u64 vbits);
template <> DEV void store_rec<1>(u64 *p, u64 key, u64 vbits) { *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(key, vbits); }
template <> DEV void store_rec<2>(u64 *p, u128 key, u64 vbits) {
    reinterpret_cast<ulonglong2 *>(p)[0] = make_ulonglong2((u64)key, (u64)(key >> 64));
    reinterpret_cast<ulonglong2 *>(p)[1] = make_ulonglong2(vbits, 0ull);
}
template <int W> DEV void load_rec(const u64 *p, typename BitsT<W>::type &key, u64 &vbits);
template <> DEV void load_rec<1>(const u64 *p, u64 &key, u64 &vbits) {
    const ulonglong2 r = *reinterpret_cast<const ulonglong2 *>(p); key = r.x; vbits = r.y;
}
template <> DEV void load_rec<2>(const u64 *p, u128 &key, u64 &vbits) {
    const ulonglong2 a = reinterpret_cast<const ulonglong2 *>(p)[0]; key = ((u128)a.y << 64) | (u128)a.x;
    vbits = p[2];
}

// append to sub-stream `slot` of the LOCAL bucket of hash h
template <int W, class VT>
DEV void append_record(const PartDev &pt, StatsDev *st, typename BitsT<W>::type key, u64 h, int nranks, VT v, u32 slot) {
    const u32 b = bucket_of(h, nranks, pt.nb);
    const u64 run = (u64)slot * pt.nb + b;
    const u32 pos = atomicAdd(&pt.rcnt[run], 1u);
    if (pos < pt.rcap) {
        union { VT v; u64 b; } cv; cv.v = v;
        store_rec<W>(pt.rec + (run * pt.rcap + pos) * RecWords<W>::value, key, cv.b);
    } else st->overflow_table = 1;
}

// CTA-collective routing of at most one spawn record per thread.  Every thread of the CTA must call this.
//  * one rank, or direct mode: the record is appended to sub-stream (owner rank, lane) of the child's bucket in LOCAL memory
//    -- one counter atomic and one 16/32-byte store, whoever owns the child.  Records for other GPUs are never copied: their
//    owner's merge kernel reads them in place over NVLink.
//  * staged mode (no peer access): records for other ranks are staged per peer -- ranks within the CTA come from a
//    shared-memory counter, ONE global atomic per peer and CTA reserves a contiguous run in that peer's exchange segment
//    (the reference packs per-rank buffers serially, communicators.jl:421-444); NCCL send/recv moves the segments.
// Returns true when the record leaves this rank (the sent_records statistic).
struct RouteSmem { u32 cnt[RIMU_MAX_RANKS]; u64 base[RIMU_MAX_RANKS]; };
template <int W, class VT>
DEV bool route_record(const PartDev &pt, const ExchangeDev &x, const StepDev &p, StatsDev *st, RouteSmem &rs,
                      bool has, typename BitsT<W>::type key, VT v, u32 lane) {
    u64 h = 0;
    int owner = p.rank;
    if (has) {
        h = hash_bits(key);
        if (p.nranks > 1) owner = addr_owner(h, p.nranks);
    }
    if (p.nranks > 1 && !pt.direct) { // uniform over the grid
        if (threadIdx.x < RIMU_MAX_RANKS) rs.cnt[threadIdx.x] = 0;
        __syncthreads();
        const bool remote = has && owner != p.rank;
        u32 lrank = 0;
        if (remote) lrank = atomicAdd(&rs.cnt[owner], 1u);
        __syncthreads();
        if ((int)threadIdx.x < p.nranks && rs.cnt[threadIdx.x])
            rs.base[threadIdx.x] = atomicAdd(&x.counts[threadIdx.x], (u64)rs.cnt[threadIdx.x]);
        __syncthreads();
        if (remote) {
            const u64 idx = rs.base[owner] + lrank;
            if (idx < x.cap) {
                union { VT v; u64 b; } cv; cv.v = v;
                store_key<W>(x.keys + ((u64)owner * x.cap + idx) * W, key);
                x.vals[(u64)owner * x.cap + idx] = cv.b;
            } else st->overflow_xchg = 1;
        }
        if (has && owner == p.rank) append_record<W, VT>(pt, st, key, h, p.nranks, v, lane);
        return remote;
    }
    if (has) append_record<W, VT>(pt, st, key, h, p.nranks, v, (u32)owner * pt.nlane + lane);
    return has && owner != p.rank;
}

// ---------------------------------------------------------------- K1: spawning, CTA-local work distribution
template <int HK, int W, class VT, bool FS = false>
__global__ void __launch_bounds__(SPAWN_NT, SPAWN_MINB)
spawn_part_kernel(const HamDev h, const StepDev p_in, const u64 *__restrict__ keys, const VT *__restrict__ vals, i64 n,
                  PartDev pt, ExchangeDev xch, HeavyDev hv, StatsDev *st) {
    typedef typename BitsT<W>::type B;
    RIMU_TUNE_FIX_STEP(p, p_in)
    if (p.ctl) { // batch of steps: the source is the previous step's result, its length is known on the device only
        if (p.ctl->stop) return;
        n = (i64)p.ctl->n;
    }
    __shared__ u32 s_off[SPAWN_NT + 1];
    __shared__ u64 s_keys[SPAWN_NT * W];
    __shared__ VT s_vals[SPAWN_NT];
    __shared__ u32 s_L[SPAWN_NT];   // off-diagonal count | exact flag in bit 31
    __shared__ u32 s_warp[SPAWN_NT / 32];
    __shared__ RouteSmem s_route;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double spawns = 0.0;
    i64 exact_steps = 0, inexact_steps = 0, attempts = 0, nsent = 0;
    const i64 PPC = (i64)pt.ppc; // parents per chunk (SPAWN_NT unless the vector is small)
    const bool has_parent = tid < (int)PPC;
    const i64 nchunks = (n + PPC - 1) / PPC;
    // the parent of the NEXT chunk is loaded while the current chunk is processed (its HBM latency was the largest
    // single stall of this kernel: 21 % of the samples in profiles/r1_partition_ncu_summary.md)
    B nkey = 0; VT npv = (VT)0;
    { const i64 j0 = (i64)blockIdx.x * PPC + tid; if (has_parent && j0 < n) { nkey = load_key<W>(keys + j0 * W); npv = vals[j0]; } }
    for (i64 chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const i64 j = chunk * PPC + tid;
        const B key = nkey; const VT pv = npv;
        { const i64 jn = j + (i64)gridDim.x * PPC; if (has_parent && jn < n) { nkey = load_key<W>(keys + jn * W); npv = vals[jn]; } }
        u32 cnt = 0;
        if (has_parent && j < n) {
            long long L = ham_num_offdiagonals<HK, B>(h, key);
            u64 c64;
            bool exact = attempts_for(p, (double)pv, L, c64);
            if (c64) { if (exact) exact_steps++; else inexact_steps++; }
            attempts += (i64)c64;
            if (c64 > HEAVY_T) {
                u64 ntiles = (c64 + HEAVY_TILE - 1) / HEAVY_TILE;
                u64 old = atomicAdd(hv.packed, (1ull << 32) | ntiles);
                u64 idx = old >> 32;
                if (idx < hv.cap) { HeavyItem it; it.parent = j; it.nattempts = c64; it.tile_base = old & 0xffffffffull; it.exact = exact; hv.items[idx] = it; }
                else st->overflow_table = 1;
                c64 = 0;
            }
            cnt = (u32)c64;
            s_L[tid] = (u32)L | (exact ? 0x80000000u : 0u);
            s_keys[tid * W] = (u64)key;
            if constexpr (W == 2) s_keys[tid * W + 1] = (u64)(key >> 64);
            s_vals[tid] = pv;
        }
        // CTA exclusive scan of cnt
        u32 incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        u32 base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SPAWN_NT / 32; w++) { u32 t = s_warp[w]; if (w < wid) base += t; total += t; }
        s_off[tid] = base + incl - cnt;
        if (tid == 0) s_off[SPAWN_NT] = total;
        __syncthreads();
        for (u32 a0 = 0; a0 < total; a0 += SPAWN_NT) {
            const u32 a = a0 + tid;
            B child = 0;
            VT nv = (VT)0;
            u32 rlane = 0;
            if (a < total) {
                // parent of attempt a = last lo with s_off[lo] <= a.  When every earlier parent of the chunk makes exactly one
                // attempt (the common case at ~1 walker per determinant) that is parent a itself: two loads instead of
                // the 8-step binary search
                int lo = (int)a < SPAWN_NT ? (int)a : 0;
                if (!(a < SPAWN_NT && s_off[lo] == a && s_off[lo + 1] == a + 1)) {
                    lo = 0;
                    int hi = SPAWN_NT;
#pragma unroll
                    for (int it = 0; it < 8; it++) { int mid = (lo + hi) >> 1; if (s_off[mid] <= a) lo = mid; else hi = mid; }
                }
                const u64 k = a - s_off[lo];
                B key;
                if constexpr (W == 1) key = s_keys[lo]; else key = ((u128)s_keys[lo * 2 + 1] << 64) | (u128)s_keys[lo * 2];
                const double val = (double)s_vals[lo];
                const u32 Lx = s_L[lo];
                const long long L = (long long)(Lx & 0x7fffffffu);
                const bool exact = (Lx >> 31) != 0;
                const u64 nat = s_off[lo + 1] - s_off[lo]; // light parents only: their attempt count is the scan difference
                long long ci; double sp;
                nv = spawn_attempt<HK, W, VT>(h, p, key, hash_bits(key), val, L, nat, exact, k, child, ci, sp);
                spawns += sp;
                rlane = deposit_lane(p, false, val);
            }
            nsent += route_record<W, VT>(pt, xch, p, st, s_route, nv != (VT)0, child, nv, rlane);
        }
        __syncthreads();
    }
    if (std::is_integral<VT>::value) stat_add(&st->ispawns, (i64)spawns); else stat_add(&st->spawns, spawns);
    stat_add(&st->exact_steps, exact_steps); stat_add(&st->inexact_steps, inexact_steps);
    stat_add(&st->spawn_attempts, attempts);
    if (p.nranks > 1) stat_add(&st->sent, nsent);
}

// ---------------------------------------------------------------- K2: heavy parents, one CTA per tile of attempts
template <int HK, int W, class VT>
__global__ void __launch_bounds__(SPAWN_NT)
spawn_heavy_kernel(const HamDev h, const StepDev p, const u64 *__restrict__ keys, const VT *__restrict__ vals,
                   PartDev pt, ExchangeDev xch, HeavyDev hv, StatsDev *st) {
    typedef typename BitsT<W>::type B;
    __shared__ u64 acc[ACC_MAX];
    __shared__ RouteSmem s_route;
    if (p.ctl && p.ctl->stop) return;
    const u64 packed = *hv.packed;
    const u64 total = packed & 0xffffffffull;
    u64 nitems = packed >> 32;
    if (nitems > hv.cap) nitems = hv.cap;
    double spawns = 0.0;
    i64 nsent = 0;
    for (u64 w = blockIdx.x; w < total; w += gridDim.x) {
        u64 lo = 0, hi = nitems;
        while (hi - lo > 1) { u64 mid = (lo + hi) >> 1; if (hv.items[mid].tile_base <= w) lo = mid; else hi = mid; }
        const HeavyItem it = hv.items[lo];
        const u64 a0 = (w - it.tile_base) * HEAVY_TILE;
        const u64 a1 = min(a0 + (u64)HEAVY_TILE, it.nattempts);
        const B key = load_key<W>(keys + it.parent * W);
        const double val = (double)vals[it.parent];
        const long long L = ham_num_offdiagonals<HK, B>(h, key);
        const bool exact = it.exact != 0;
        const bool agg = !exact && L <= ACC_MAX && !p.ordered; // (pre-sums use shared atomics: their order is not reproducible)
        const u64 hkey = hash_bits(key);
        const u32 rlane = deposit_lane(p, false, val); // one parent per tile: one lane
        if (agg) { for (int c = threadIdx.x; c < L; c += SPAWN_NT) acc[c] = 0ull; __syncthreads(); }
        for (u64 ab = a0; ab < a1; ab += SPAWN_NT) {
            const u64 a = ab + threadIdx.x;
            B child = 0;
            VT nv = (VT)0;
            if (a < a1) {
                long long ci; double sp;
                nv = spawn_attempt<HK, W, VT>(h, p, key, hkey, val, L, it.nattempts, exact, a, child, ci, sp);
                spawns += sp;
                if (agg && nv != (VT)0) { atomic_add_val<VT>(&acc[ci], nv); nv = (VT)0; }
            }
            if (!agg) nsent += route_record<W, VT>(pt, xch, p, st, s_route, nv != (VT)0, child, nv, rlane);
        }
        if (agg) {
            __syncthreads();
            for (int cb = 0; cb < L; cb += SPAWN_NT) {
                const int c = cb + threadIdx.x;
                B child = 0;
                union { u64 b; VT v; } cv; cv.b = 0;
                if (c < L) { cv.b = acc[c]; if (cv.v != (VT)0) ham_offdiagonal<HK, B>(h, key, c, child); }
                nsent += route_record<W, VT>(pt, xch, p, st, s_route, cv.v != (VT)0, child, cv.v, rlane);
            }
            __syncthreads();
        }
    }
    if (std::is_integral<VT>::value) stat_add(&st->ispawns, (i64)spawns); else stat_add(&st->spawns, spawns);
    if (p.nranks > 1) stat_add(&st->sent, nsent);
}

// ---------------------------------------------------------------- append a flat record list to the bucket streams
// (received exchange records, uploads, linear combinations).  Non-local keys are dropped (pdvec.jl:336-349).
template <int W, class VT>
__global__ void __launch_bounds__(RIMU_TPB)
append_records_kernel(const u64 *__restrict__ keys, const VT *__restrict__ vals, i64 n, double scale, int use_scale,
                      int rank, int nranks, PartDev pt, u32 slot, StatsDev *st) {
    typedef typename BitsT<W>::type B;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        B key = load_key<W>(keys + i * W);
        VT v = vals[i];
        if (use_scale) v = (VT)(scale * (double)v);
        if (v == (VT)0) continue;
        u64 hh = hash_bits(key);
        if (nranks > 1 && addr_owner(hh, nranks) != rank) continue;
        append_record<W, VT>(pt, st, key, hh, nranks, v, slot);
    }
}

// ---------------------------------------------------------------- K0: diagonal step as records
// Used when the source vector is not segmented for this step's bucket count (fresh uploads, a changed
// bucket count): the parents' diagonal deposits travel through the bucket streams like spawns, which
// costs one extra record write + read per parent instead of a re-segmentation pass.
template <int HK, int W, class VT>
__global__ void __launch_bounds__(RIMU_TPB)
diag_append_kernel(const HamDev h, const StepDev p, const u64 *__restrict__ keys, const VT *__restrict__ vals,
                   const double *__restrict__ diag, i64 n, PartDev pt, StatsDev *st) {
    typedef typename BitsT<W>::type B;
    constexpr bool is_int = std::is_integral<VT>::value;
    double clones = 0.0, deaths = 0.0, zombies = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const B key = load_key<W>(keys + i * W);
        const double val = (double)vals[i];
        const u64 hk = hash_bits(key);
        const double hd = diag ? diag[i] : ham_diagonal<HK, B>(h, key);
        const double d = p.plain_h ? hd : 1 - p.dtau * (hd - p.shift);
        double rr = 0.0;
        const double thr = is_int ? 0.0 : p.proj_thr;
        if (is_int || thr > 0.0) {
            u32 rnd[4];
            rng_draw(hk, 0, STREAM_DIAG, p.k0, p.k1, rnd);
            rr = u53(rnd[1], rnd[2]);
        }
        const VT v = project_value<VT>(d * val, thr, rr);
        const double rs = (double)v;
        if (rs > val) clones += fabs(rs - val);
        else if (sgn_(rs) != sgn_(val)) { deaths += fabs(val); zombies += fabs(rs); }
        else deaths += fabs(rs - val);
        if (v != (VT)0) append_record<W, VT>(pt, st, key, hk, p.nranks, v, pt.me + deposit_lane(p, true, val));
    }
    if (is_int) { stat_add(&st->iclones, (i64)clones); stat_add(&st->ideaths, (i64)deaths); stat_add(&st->izombies, (i64)zombies); }
    else { stat_add(&st->clones, clones); stat_add(&st->deaths, deaths); stat_add(&st->zombies, zombies); }
}

// ---------------------------------------------------------------- K3: per-bucket annihilation in shared memory
// MODE 0: FCIQMC step (diagonal step on the parents, compression);  MODE 1: plain sum of records (+ alpha * parents)
//
// Shared memory per CTA (CAP items): keys[CAP][W] | vals[CAP] | owner[2*CAP] (u32 item index per table slot) |
// pidx[CAP] (u16: index of the parent item that was annihilated into this owner, for its cached H_aa).
// Placement is one pass: an item claims the first free slot of its probe sequence with a 32-bit shared CAS on
// the owner table; keys are compared through the item index, so nothing has to be published after a claim.
// After placement the owner table is dead and is reused, one region per warp, for the lists of entries that still need a
// compression draw or a fresh H_aa -- every warp finishes ITS items with __syncwarp only.  CTA barriers per bucket: after
// staging, after placement, one for the survivor scan (the last warp to arrive takes the cursor atomic), one at the end.
//
// Staging: spawn records need no arithmetic, so they are copied global -> shared with cp.async (LDGSTS): every record
// load of the bucket is in flight at once, and the copies -- NVLink loads for sub-streams that live on other GPUs -- complete
// while the parents go through the diagonal step.

DEV void cp_async8(void *smem_dst, const void *gsrc) {
    const u32 s = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
DEV void cp_async16(void *smem_dst, const void *gsrc) {
    const u32 s = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> DEV void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// slot of an address in the bucket's shared-memory table.  The items of a bucket are a pseudo-random subset of the addresses
// (the bucket is a range of the fmix64 hash), so a multiplicative hash of the folded words spreads them; it costs 3 integer
// multiplies instead of the two 64-bit multiply chains of fmix64, which is kept for what needs avalanche: buckets and RNG.
template <u32 BITS> DEV u32 slot_hash(u64 k0, u64 k1) {
    u32 x = (u32)k0 ^ ((u32)(k0 >> 32) * 0x85EBCA6Bu) ^ ((u32)k1 * 0xC2B2AE35u) ^ ((u32)(k1 >> 32) * 0x27D4EB2Fu);
    return (x * 0x9E3779B1u) >> (32 - BITS);
}
template <u32 N> struct Log2 { static constexpr u32 value = 1 + Log2<N / 2>::value; };
template <> struct Log2<1> { static constexpr u32 value = 0; };

// ORD (Float64 steps, audit mode): order-deterministic summation.  The records of a bucket arrive in whatever order the
// spawn kernels' counter atomics happened to hand out, and the hash-table placement adds duplicates with shared-memory
// atomics -- both make the last bits of a Float64 sum differ from run to run.  With ORD the bucket's items are SORTED by
// (address, value bits) with a bitonic network in shared memory (the hash table is not used at all), every address is summed
// by one thread in that order, and the walker number is reduced in a fixed order as well: the step's result is a pure
// function of its inputs, bit for bit.  ~10x slower than the hash placement; selected by rimu_step_params.ordered.
template <int HK, int W, class VT, int MODE, bool INIT = false, bool ORD = false, bool FS = false>
__global__ void __launch_bounds__(PART_NT, PART_MINB)
merge_kernel(const HamDev h, const StepDev p_in, SegSrc src, double alpha, PartDev pt, SegDst dst, StatsDev *st, double *ord_partials = nullptr) {
    typedef typename BitsT<W>::type B;
    RIMU_TUNE_FIX_STEP(p, p_in)
    constexpr bool is_int = std::is_integral<VT>::value;
    constexpr int CAP = PartCap<W>::value;
    constexpr int R = CAP / PART_NT;
    constexpr int NW = PART_NT / 32;
    constexpr u32 TBITS = Log2<2 * CAP>::value;
    constexpr u32 TMASK = 2 * CAP - 1;
    constexpr u32 WLIST = 2 * CAP / NW;  // owner-table entries per warp once the table is dead (>= R * 32 items of a warp)
    constexpr u32 NIL = 0xffffffffu;
    constexpr u32 NOPARENT = 0x3fffu, OWNFLAG = 0x4000u, IFLAG = 0x8000u; // pidx: parent item index | (ORD) "sums its address" | "initiator lane is non-zero"
    static_assert(CAP < (int)NOPARENT, "item indices must fit below the flags");
    static_assert((2 * CAP & (2 * CAP - 1)) == 0, "table size must be a power of two");
    static_assert(WLIST >= (u32)R * 32, "a warp's list region must hold all of its items");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *skeys = reinterpret_cast<u64 *>(smem_raw);
    u64 *svals = skeys + CAP * W;
    u32 *owner = reinterpret_cast<u32 *>(svals + CAP);
    unsigned short *pidx = reinterpret_cast<unsigned short *>(owner + 2 * CAP);
    // initiator rule (MODE 0 only): svals accumulates safe + initiator, sunsafe the unsafe lane (CAP more u64 of shared
    // memory, allocated by the host only for such steps); whether the initiator lane is non-zero is one bit, because
    // only an address's own parent can deposit there
    constexpr bool initm = MODE == 0 && INIT; // separate instantiation: the plain step pays nothing for the lanes
    u64 *sunsafe = reinterpret_cast<u64 *>(pidx + CAP);
    if (p.ctl && p.ctl->stop) return;         // batch of steps that has ended: leave both vectors as they are
    const double shift = p.ctl ? p.ctl->shift : p.shift;
    __shared__ u32 s_warp[NW];
    __shared__ u64 s_base;
    __shared__ u32 s_arrive;
    __shared__ u32 s_nlist;
    __shared__ double s_wnorm[NW]; // ORD: per-warp walker number of the current bucket
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    u32 *wlist = owner + wid * WLIST;
    double norm1 = 0.0, clones = 0.0, deaths = 0.0, zombies = 0.0;
    i64 inorm1 = 0;
    u32 len_before = 0, len = 0, ndep = 0; // per-thread counts (32 bits are plenty; registers are the scarce resource here)
    u32 max_fill = 0;
    u32 nrec_sum = 0;
    // Bucket metadata (segment of parents, fill of every source's sub-stream) runs two buckets ahead of the merge, and
    // the next bucket's parents and local records are pulled into L2 with bulk prefetches while this one is merged, so that
    // staging sees L2 latency instead of two dependent HBM round trips (metadata, then data).
    constexpr int RW = RecWords<W>::value;
    constexpr int MAXSUB = RIMU_MAX_RANKS * 3;  // sub-streams per bucket: ranks x lanes
    __shared__ u32 s_cnt[3][MAXSUB];            // ring: sub-stream fills of this bucket, the next, the one after
    __shared__ u32 s_np[3];                     // ... and the parents' segment (length, start)
    __shared__ u64 s_p0[3];
    const u32 nsrc = pt.nsrc;
    const bool meta_thread = tid == PART_NT - 1; // loads the segment metadata; threads 0..nsrc-1 load the fills
    // Sub-stream q = (source rank q / nlane, lane q % nlane) of this rank's buckets lives in the SOURCE rank's streams as
    // its sub-stream (this rank, lane): local memory for q / nlane == this rank, peer memory (NVLink loads) otherwise.
    __shared__ const u64 *s_recbase[MAXSUB];
    __shared__ const u32 *s_cntbase[MAXSUB];
    __shared__ unsigned char s_local[MAXSUB];
    if ((u32)tid < nsrc) {
        const u32 srank = (u32)tid / pt.nlane, sub = pt.me + (u32)tid % pt.nlane;
        const u64 *rb = pt.direct ? pt.peer_rec[srank] : pt.rec;
        const u32 *cb = pt.direct ? pt.peer_rcnt[srank] : pt.rcnt;
        s_recbase[tid] = rb + (u64)sub * pt.nb * pt.rcap * RW;
        s_cntbase[tid] = cb + (u64)sub * pt.nb;
        s_local[tid] = (!pt.direct || srank * pt.nlane == pt.me) ? 1 : 0;
    }
    if (tid == 0) s_arrive = 0;
    // the kinetic-energy table of the momentum-space models is read once per occupied mode by every H_aa evaluation: keep
    // it in shared memory (29-cycle loads instead of a trip through L1/L2)
    constexpr int BK = HkBase<HK>::value;
    __shared__ double s_kes[BK == HK_MOM1D_BOSE || BK == HK_MOM1D_F2C || BK == HK_TC_F2C ? 64 : 1];
    HamDev hl = h;
    if constexpr (BK == HK_MOM1D_BOSE || BK == HK_MOM1D_F2C || BK == HK_TC_F2C) {
        if (h.kes && h.M <= 64) {
            if (tid < h.M) s_kes[tid] = h.kes[tid];
            hl.kes = s_kes;
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const u64 bq = (u64)blockIdx.x + (u64)q * gridDim.x;
        if (meta_thread) {
            const u32 npq = (bq < pt.nb && src.seg_len) ? src.seg_len[bq] : 0u;
            s_np[q] = npq;
            s_p0[q] = npq ? src.seg_start[bq] : 0ull;
        }
        if ((u32)tid < nsrc) s_cnt[q][tid] = bq < pt.nb ? s_cntbase[tid][bq] : 0u;
    }
    __syncthreads();
    u32 ring = 0;
    for (u32 b = blockIdx.x; b < pt.nb; b += gridDim.x, ring = ring == 2 ? 0 : ring + 1) {
        const u32 cur = ring, nxt = ring == 2 ? 0 : ring + 1, nn = nxt == 2 ? 0 : nxt + 1;
        const u32 np = s_np[cur];
        const u64 p0 = s_p0[cur];
        // metadata two buckets ahead: loaded now, stored to the ring at the end of this iteration (its latency is hidden)
        u32 pending_cnt = 0, pending_np = 0;
        u64 pending_p0 = 0;
        {
            const u64 b2 = (u64)b + 2ull * gridDim.x;
            if (b2 < pt.nb) {
                if (meta_thread) {
                    pending_np = src.seg_len ? src.seg_len[b2] : 0u;
                    pending_p0 = pending_np ? src.seg_start[b2] : 0ull;
                }
                if ((u32)tid < nsrc) pending_cnt = s_cntbase[tid][b2];
            }
            const u64 b1 = (u64)b + gridDim.x;
            if (b1 < pt.nb) {
                if ((u32)tid < nsrc && s_local[tid]) { // (peer memory is not cached in this GPU's L2)
                    const u32 c1 = s_cnt[nxt][tid] < pt.rcap ? s_cnt[nxt][tid] : pt.rcap;
                    if (c1) l2_prefetch(s_recbase[tid] + (b1 * pt.rcap) * RW, (u64)c1 * RW * 8);
                }
                if (meta_thread && s_np[nxt]) {
                    const u32 np1 = s_np[nxt];
                    const u64 p1 = s_p0[nxt];
                    l2_prefetch(src.keys + p1 * W, (u64)np1 * W * 8);
                    l2_prefetch(src.vals + p1, (u64)np1 * 8);
                    if (src.diag) l2_prefetch(src.diag + p1, (u64)np1 * 8);
                }
            }
        }
        u32 nrec = 0;
        bool sub_over = false;
        for (u32 q = 0; q < nsrc; q++) { const u32 cq = s_cnt[cur][q]; nrec += cq; sub_over |= cq > pt.rcap; }
        // one rank: this CTA is the only reader of the bucket's fill counters (they sit in the ring by now), so it clears
        // them for the next step -- the host then has no counter array to memset between steps
        if (!pt.direct && (u32)tid < nsrc) pt.rcnt[(u64)tid * pt.nb + b] = 0u;
        const u32 n = np + nrec;
        max_fill = max(max_fill, n);
        nrec_sum += nrec;
        if (sub_over || n > (u32)CAP) { // uniform over the CTA: the host retries with more buckets
            if (tid == 0) { st->overflow_table = 1; dst.seg_len[b] = 0; dst.seg_start[b] = 0; if constexpr (ORD) ord_partials[b] = 0.0; }
            if ((u32)tid < nsrc) s_cnt[nn][tid] = pending_cnt;
            if (meta_thread) { s_np[nn] = pending_np; s_p0[nn] = pending_p0; }
            __syncthreads();
            continue;
        }
        for (int i = tid; i < 2 * CAP; i += PART_NT) owner[i] = NIL;
        if (tid == 0) s_nlist = 0;
        const int rmax = (int)((n + PART_NT - 1) / PART_NT); // uniform: rounds of PART_NT items this bucket needs
        // Item of this thread in round r: tid + r * PART_NT in the full rounds; the items of the partial last round are dealt
        // out across ALL warps (lane-major), so that every warp owns the same number of items +-1 and the warps reach the
        // barriers together (with the plain mapping the low warps did one round more than the high ones)
        const u32 rfull = n / PART_NT;
        const u32 tail_item = rfull * PART_NT + (u32)lane * NW + (u32)wid;
#if PART_BALANCED_TAIL
        auto item_of = [&](int r) -> u32 { return (u32)r < rfull ? (u32)tid + (u32)r * PART_NT : tail_item; };
#else
        (void)rfull; (void)tail_item;
        auto item_of = [&](int r) -> u32 { return (u32)tid + (u32)r * PART_NT; };
#endif
        // ---- stage the spawn records: asynchronous copies straight into the item arrays (no lanes to sort out)
        u32 valid = 0;
        constexpr bool async_parents = PART_ASYNC_PARENTS && !initm;
#ifndef PART_RH
#define PART_RH 4
#endif
        constexpr int RH = R < PART_RH ? R : PART_RH; // (parents beyond the first 4 rounds -- more than 1024 per bucket -- read H_aa when they need it)
        [[maybe_unused]] double hdr[RH];  // async_parents: cached H_aa of this thread's parents, loaded while the copies are in flight
        if constexpr (async_parents) {
            // parents first (their own commit group): key and value go global -> shared without passing through registers;
            // the diagonal step below reads them back once this thread's copies have landed
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r >= rmax) break;
                const u32 i = item_of(r);
                if (i >= np) continue;
                const u64 gp = p0 + i;
                if constexpr (W == 1) cp_async8(skeys + i, src.keys + gp); else cp_async16(skeys + 2 * i, src.keys + 2 * gp);
                cp_async8(svals + i, src.vals + gp);
                if constexpr (MODE == 0) { if (r < RH && src.diag) hdr[r < RH ? r : 0] = src.diag[gp]; }
            }
            cp_async_commit();
        }
        if constexpr (!initm) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r >= rmax) break;
                const u32 i = item_of(r);
                if (i < np || i >= n) continue;
                u32 j = i - np, q = 0; // record j of the bucket -> (source sub-stream q, position j)
                if (nsrc > 1) for (u32 cq = s_cnt[cur][0]; j >= cq; cq = s_cnt[cur][q]) { j -= cq; q++; }
                const u64 *rp = s_recbase[q] + ((u64)b * pt.rcap + j) * RW;
                if constexpr (W == 1) { cp_async8(skeys + i, rp); cp_async8(svals + i, rp + 1); }
                else { cp_async16(skeys + 2 * i, rp); cp_async8(svals + i, rp + 2); }
                pidx[i] = (unsigned short)NOPARENT;
                valid |= 1u << r; // spawn kernels append non-zero values only
                ndep++;
            }
        }
        if constexpr (async_parents) { cp_async_commit(); cp_async_wait_group<1>(); } // this thread's parents have landed; its records may still be in flight
        // ---- stage parents (with the diagonal step); with initiator lanes the records go through registers as well
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (r >= rmax) break;
            const u32 i = item_of(r);
            if (i >= n) continue;
            if (!initm && i >= np) continue;
            B key; VT v;
            u32 ilane = LANE_SAFE;
            if (i < np) {
                union { u64 b; VT v; } cv;
                if constexpr (async_parents) {
                    if constexpr (W == 1) key = skeys[i]; else key = ((u128)skeys[i * 2 + 1] << 64) | (u128)skeys[i * 2];
                    cv.b = svals[i];
                } else {
                    key = load_key<W>(src.keys + (p0 + i) * W);
                    cv.b = src.vals[p0 + i];
                }
                const VT pv = cv.v;
                if constexpr (MODE == 0) {
                    // diagonal_step! (spawning.jl:73-77) through FirstOrderTransitionOperator (fciqmc.jl:93-96)
                    const double val = (double)pv;
                    double hd;
                    if constexpr (async_parents) hd = src.diag ? (r < RH ? hdr[r < RH ? r : 0] : src.diag[p0 + i]) : ham_diagonal<HK, B>(hl, key);
                    else hd = src.diag ? src.diag[p0 + i] : ham_diagonal<HK, B>(hl, key);
                    const double d = p.plain_h ? hd : 1 - p.dtau * (hd - shift);
                    double rr = 0.0;
                    const double thr = is_int ? 0.0 : p.proj_thr;
                    if (is_int || thr > 0.0) {
                        u32 rnd[4];
                        rng_draw(hash_bits(key), 0, STREAM_DIAG, p.k0, p.k1, rnd);
                        rr = u53(rnd[1], rnd[2]);
                    }
                    v = project_value<VT>(d * val, thr, rr);
                    ilane = deposit_lane(p, true, val);
                    const double rs = (double)v; // clones_deaths_zombies (spawning.jl:79-93)
                    if (rs > val) clones += fabs(rs - val);
                    else if (sgn_(rs) != sgn_(val)) { deaths += fabs(val); zombies += fabs(rs); }
                    else deaths += fabs(rs - val);
                } else {
                    v = (VT)(alpha * (double)pv);
                }
            } else {
                u32 j = i - np, q = 0;
                if (nsrc > 1) for (u32 cq = s_cnt[cur][0]; j >= cq; cq = s_cnt[cur][q]) { j -= cq; q++; }
                union { u64 b; VT v; } cv;
                load_rec<W>(s_recbase[q] + ((u64)b * pt.rcap + j) * RW, key, cv.b);
                v = cv.v;
                if (pt.nlane > 1) ilane = q % pt.nlane;
            }
            u32 pflag = NOPARENT;
            if (initm && v != (VT)0) {
                if (ilane == LANE_INIT) pflag |= IFLAG;
                union { u64 b; VT v; } cz; cz.v = (VT)0; sunsafe[i] = cz.b;
            }
            pidx[i] = (unsigned short)pflag;
            if (v != (VT)0) {
                if constexpr (!async_parents) { // (async: the key is where the copy put it)
                    skeys[i * W] = (u64)key;
                    if constexpr (W == 2) skeys[i * W + 1] = (u64)(key >> 64);
                }
                union { u64 b; VT v; } cv; cv.v = v;
                if (initm && ilane == LANE_UNSAFE) { sunsafe[i] = cv.b; cv.v = (VT)0; }
                svals[i] = cv.b;
                valid |= 1u << r;
                ndep++;
            } else if constexpr (ORD) svals[i] = 0ull; // (the sort recognises items without a deposit by their zero value)
        }
        if constexpr (!initm) cp_async_wait_all();
        __syncthreads();
        u32 own = 0;
        [[maybe_unused]] u32 P2 = 2; // ORD: length of the sorted index array (power of two >= n)
        if constexpr (ORD) {
            // ---- ordered annihilation: sort the item indices by (address, value bits), sum every run in that order
            u32 *idx = owner;
            while (P2 < n) P2 <<= 1; // uniform
            for (u32 s = tid; s < P2; s += PART_NT) idx[s] = (s < n && svals[s] != 0ull) ? s : NIL;
            __syncthreads();
            auto same_key = [&](u32 a, u32 b) {
                bool eq = skeys[a * W] == skeys[b * W];
                if constexpr (W == 2) eq = eq && skeys[a * W + 1] == skeys[b * W + 1];
                return eq;
            };
            auto less = [&](u32 a, u32 b) { // strict total order; NIL sorts last
                if (a == NIL) return false;
                if (b == NIL) return true;
                if constexpr (W == 2) { if (skeys[a * 2 + 1] != skeys[b * 2 + 1]) return skeys[a * 2 + 1] < skeys[b * 2 + 1]; }
                if (skeys[a * W] != skeys[b * W]) return skeys[a * W] < skeys[b * W];
                if (svals[a] != svals[b]) return svals[a] < svals[b];
                return a < b;
            };
            for (u32 k = 2; k <= P2; k <<= 1)
                for (u32 j = k >> 1; j > 0; j >>= 1) {
                    for (u32 t = tid; t < P2 / 2; t += PART_NT) {
                        const u32 i1 = ((t & ~(j - 1u)) << 1) | (t & (j - 1u)), i2 = i1 + j;
                        const bool up = (i1 & k) == 0u;
                        const u32 a = idx[i1], b2 = idx[i2];
                        if (less(b2, a) == up) { idx[i1] = b2; idx[i2] = a; }
                    }
                    __syncthreads();
                }
            for (u32 s = tid; s < P2; s += PART_NT) {
                const u32 a = idx[s];
                if (a == NIL) continue;
                if (s > 0 && same_key(idx[s - 1], a)) continue; // not the head of its run
                union { u64 b; VT v; } acc; acc.v = (VT)0;
                u32 pi = NOPARENT;
                for (u32 e = s; e < P2; e++) {
                    const u32 m = idx[e];
                    if (m == NIL || !same_key(m, a)) break;
                    union { u64 b; VT v; } cv; cv.b = svals[m];
                    acc.v += cv.v;
                    if (m < np) pi = m;
                }
                svals[a] = acc.b;
                pidx[a] = (unsigned short)(pi | OWNFLAG);
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r >= rmax) break;
                if (!((valid >> r) & 1u)) continue;
                if (pidx[item_of(r)] & OWNFLAG) own |= 1u << r;
            }
        } else {
        // ---- placement: claim a slot (CAS) or annihilate into the item that owns this address
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (r >= rmax) break;
            if (!((valid >> r) & 1u)) continue;
            const u32 i = item_of(r);
            const u64 k0 = skeys[i * W];
            u64 k1 = 0; if constexpr (W == 2) k1 = skeys[i * W + 1];
            u32 s = slot_hash<TBITS>(k0, k1);
            for (;;) {
                u32 o = owner[s];
                if (o == NIL) {
                    o = atomicCAS(&owner[s], NIL, i);
                    if (o == NIL) { own |= 1u << r; break; }
                }
                bool eq = skeys[o * W] == k0;
                if constexpr (W == 2) eq = eq && skeys[o * W + 1] == k1;
                if (eq) {
                    union { u64 b; VT v; } cv; cv.b = svals[i];
                    if (cv.v != (VT)0) atomic_add_val<VT>(&svals[o], cv.v);
                    if (initm) {
                        union { u64 b; VT v; } cu; cu.b = sunsafe[i];
                        if (cu.v != (VT)0) atomic_add_val<VT>(&sunsafe[o], cu.v);
                    }
                    if constexpr (MODE == 0) {
                        // the parent's cached H_aa (and its initiator flag) follow the address; an initiator-lane RECORD
                        // (unsegmented source) only carries the flag.  One writer per address in either case.
                        if (i < np) pidx[o] = (unsigned short)(i | (pidx[i] & IFLAG));
                        else if (initm && (pidx[i] & IFLAG)) pidx[o] = (unsigned short)(pidx[o] | IFLAG);
                    }
                    break;
                }
                s = (s + 1) & TMASK;
            }
        }
        } // !ORD
        __syncthreads(); // every sum is complete; the owner table is dead from here on (reused as per-warp lists)
        // ---- from_initiator_value (initiators.jl:136-138,177-179,201-207; pdworkingmemory.jl:268-270): collapse the lanes
        if (initm) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r >= rmax) break;
                if (!((own >> r) & 1u)) continue;
                const u32 i = item_of(r);
                union { u64 b; VT v; } ca, cu; ca.b = svals[i]; cu.b = sunsafe[i];
                const bool fi = (pidx[i] & IFLAG) != 0;
                if (ca.v == (VT)0 && cu.v == (VT)0 && !fi) continue; // all lanes zero: no entry
                len_before++;
                VT out = ca.v;
                if (p.init_rule == 1) { if (fi) out = ca.v + cu.v; }
                else if (p.init_rule == 3) { if (fi || fabs((double)cu.v) > p.init_thr) out = ca.v + cu.v; }
                ca.v = out; svals[i] = ca.b;
            }
        }
        // ---- ThresholdCompression (compression.jl:18-26) as a dense pass: the warp gathers ITS owners whose |value| is
        // below the threshold into its list so that the Philox draw runs over full warps instead of once per round for the
        // few lanes that need it.  Only this warp touches these entries from here on: __syncwarp is all it takes.
        const bool compressed = !is_int && MODE == 0 && p.compress_thr > 0.0;
        if constexpr (ORD && !is_int && MODE == 0) {
            // audit mode: every thread compresses its own entries in place -- the sorted index array (in the owner table's
            // memory, where the per-warp lists would go) is still needed for the walker number below
            if (p.compress_thr > 0.0) { // uniform
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (r >= rmax) break;
                    if (!((own >> r) & 1u)) continue;
                    const u32 i = item_of(r);
                    union { u64 b; double v; } cv; cv.b = svals[i];
                    if (cv.v == 0.0) continue;
                    len_before++;
                    if (fabs(cv.v) >= p.compress_thr) continue;
                    B key;
                    if constexpr (W == 1) key = skeys[i]; else key = ((u128)skeys[i * 2 + 1] << 64) | (u128)skeys[i * 2];
                    const double prob = p.compress_thr == 1.0 ? fabs(cv.v) : fabs(cv.v) / p.compress_thr;
                    u32 rnd[4];
                    rng_draw(hash_bits(key), 0, STREAM_COMPRESS, p.k0, p.k1, rnd);
                    cv.v = (prob > u53(rnd[1], rnd[2])) ? p.compress_thr * sgn_(cv.v) : 0.0;
                    svals[i] = cv.b;
                }
            }
        } else if constexpr (!is_int && MODE == 0) {
            if (p.compress_thr > 0.0) { // uniform
                u32 wn = 0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (r >= rmax) break;
                    const u32 i = item_of(r);
                    bool need = false;
                    if ((own >> r) & 1u) {
                        union { u64 b; double v; } cv; cv.b = svals[i];
                        if (cv.v != 0.0) { if (!initm) len_before++; need = fabs(cv.v) < p.compress_thr; }
                    }
                    const u32 bal = __ballot_sync(0xffffffffu, need);
                    if (need) wlist[wn + __popc(bal & lt_mask)] = i;
                    wn += __popc(bal);
                }
                __syncwarp();
                for (u32 j = lane; j < wn; j += 32) {
                    const u32 i = wlist[j];
                    B key;
                    if constexpr (W == 1) key = skeys[i]; else key = ((u128)skeys[i * 2 + 1] << 64) | (u128)skeys[i * 2];
                    union { u64 b; double v; } cv; cv.b = svals[i];
                    const double prob = p.compress_thr == 1.0 ? fabs(cv.v) : fabs(cv.v) / p.compress_thr; // < 1 for every listed entry (x / 1.0 == x)
                    u32 rnd[4];
                    rng_draw(hash_bits(key), 0, STREAM_COMPRESS, p.k0, p.k1, rnd);
                    cv.v = (prob > u53(rnd[1], rnd[2])) ? p.compress_thr * sgn_(cv.v) : 0.0;
                    svals[i] = cv.b;
                }
                __syncwarp();
            }
        }
        // ---- drop zeros, count survivors
        u32 keep = 0, cnt = 0;
        double bnorm = 0.0; // ORD: this thread's share of the bucket's walker number
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (r >= rmax) break;
            if (!((own >> r) & 1u)) continue;
            const u32 i = item_of(r);
            union { u64 b; VT v; } cv; cv.b = svals[i];
            const VT v = cv.v;
            if (v == (VT)0) continue; // exact zeros are deleted (pdworkingmemory.jl:25-29); so are compressed-away entries
            keep |= 1u << r; cnt++;
            if (is_int) inorm1 += (i64)(v < (VT)0 ? -v : v);
            else if (!ORD) norm1 += fabs((double)v);
        }
        if constexpr (ORD && !is_int) {
            // The bucket's walker number in a CANONICAL order.  Which thread owns which survivor depends on where the spawn
            // kernels' counter atomics happened to put the records, so a sum over "my items" would differ from run to run in its
            // last bits.  The sorted index array does not: position s holds the s-th smallest (address, value) item whatever the
            // arrival order was.  Thread t adds the run heads at positions t, t + NT, ... in increasing order, lanes and warps
            // meet in a fixed tree, buckets are summed in index order afterwards (ordered_sum_kernel).
            __syncthreads(); // every compressed value is in place
            const u32 *idx = owner;
            for (u32 s = tid; s < P2; s += PART_NT) {
                const u32 a = idx[s];
                if (a == NIL || !(pidx[a] & OWNFLAG)) continue;
                union { u64 b; double v; } cv; cv.b = svals[a];
                bnorm += fabs(cv.v); // (entries that were compressed away add an exact zero)
            }
            const double wsum = warp_sum(bnorm);
            if (lane == 0) s_wnorm[wid] = wsum;
        }
        if (!compressed && !initm) len_before += cnt;
        // ---- survivor scan: warp totals meet in shared memory; the LAST warp to arrive reserves the segment with the one
        // cursor atomic of the bucket while the others already wait at the barrier
        u32 incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
        if (lane == 31) {
            s_warp[wid] = incl;
            __threadfence_block();
            if (atomicAdd(&s_arrive, 1u) == (u32)NW - 1u) {
                __threadfence_block();
                u32 tot = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) tot += reinterpret_cast<volatile u32 *>(s_warp)[w];
                const u64 base = tot ? atomicAdd(&st->out_count, (u64)tot) : 0ull;
                s_base = base;
                dst.seg_start[b] = base;
                dst.seg_len[b] = (base + tot <= dst.cap) ? tot : 0u;
                s_arrive = 0;
            }
        }
        __syncthreads();
        if constexpr (ORD) {
            if (tid == 0) { double s = 0.0; for (int w = 0; w < NW; w++) s += s_wnorm[w]; ord_partials[b] = s; }
        }
        u32 wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) { u32 t = s_warp[w]; if (w < wid) wbase += t; total += t; }
        const u64 base = s_base;
        const bool fits = base + total <= dst.cap;
        if (fits) {
            u32 rel = wbase + incl - cnt;
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r >= rmax) break;
                const u32 i = item_of(r);
                bool fresh = false;
                if ((keep >> r) & 1u) {
                    const u64 at = base + rel;
                    dst.keys[at * W] = skeys[i * W];
                    if constexpr (W == 2) dst.keys[at * W + 1] = skeys[i * W + 1];
                    dst.vals[at] = svals[i];
                    if constexpr (MODE == 0) {
                        // H_aa of the survivor: cached with its parent, or (new determinant) evaluated below
                        const u32 pi = i < np ? i : ((u32)pidx[i] & NOPARENT);
                        if (src.diag && pi != NOPARENT) dst.diag[at] = src.diag[p0 + pi];
                        else fresh = true;
                    }
                }
                if constexpr (MODE == 0) {
                    // survivors that need a fresh H_aa are gathered CTA-wide ((item << 16) | position in the segment): the
                    // evaluation is the most expensive per-entry operation of the kernel and must run over full warps
                    const u32 bal = __ballot_sync(0xffffffffu, fresh);
                    if (bal) {
                        u32 lb = 0;
                        if (lane == 0) lb = atomicAdd(&s_nlist, (u32)__popc(bal));
                        lb = __shfl_sync(0xffffffffu, lb, 0);
                        if (fresh) owner[lb + __popc(bal & lt_mask)] = (i << 16) | rel;
                    }
                }
                if ((keep >> r) & 1u) rel++;
            }
        }
        len += cnt;
        if ((u32)tid < nsrc) s_cnt[nn][tid] = pending_cnt; // (loaded at the top of this iteration: its latency is long gone)
        if (meta_thread) { s_np[nn] = pending_np; s_p0[nn] = pending_p0; }
        if constexpr (MODE == 0) {
            // dense evaluation of H_aa for the gathered survivors (a per-lane evaluation inside the output loop would cost
            // a full divergent warp pass per new entry)
            __syncthreads();
            const u32 nl = fits ? s_nlist : 0u;
            for (u32 j = tid; j < nl; j += PART_NT) {
                const u32 e = owner[j], i = e >> 16, rel = e & 0xffffu;
                B key;
                if constexpr (W == 1) key = skeys[i]; else key = ((u128)skeys[i * 2 + 1] << 64) | (u128)skeys[i * 2];
                dst.diag[base + rel] = ham_diagonal<HK, B>(hl, key);
            }
        }
        __syncthreads(); // shared memory is reused by the next bucket
    }
    stat_add(&st->len_before, (i64)len_before);
    stat_add(&st->len, (i64)len);
    stat_add(&st->deposits, (i64)ndep);
    if (is_int) {
        stat_add(&st->inorm1, inorm1);
        if (MODE == 0) { stat_add(&st->iclones, (i64)clones); stat_add(&st->ideaths, (i64)deaths); stat_add(&st->izombies, (i64)zombies); }
    } else {
        if constexpr (!ORD) stat_add(&st->norm1, norm1); // (ORD: one partial per bucket, summed in bucket order afterwards)
        if (MODE == 0) { stat_add(&st->clones, clones); stat_add(&st->deaths, deaths); stat_add(&st->zombies, zombies); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) max_fill = max(max_fill, __shfl_xor_sync(0xffffffffu, max_fill, o));
    if (lane == 0) atomicMax(&st->max_fill, (unsigned long long)max_fill);
    if (tid == 0 && nrec_sum) atomicAdd(&st->records, (u64)nrec_sum);
}

// ---------------------------------------------------------------- controller of a batch of steps (rimu_advance)
// After the merge of a step: update_shift_parameters! (strategies_and_params/shiftstrategy.jl:77-213) and the abort rules of
// advance! (fciqmc.jl:126-181: dead population, max_length, a strategy that asks to stop), evaluated by ONE device thread so
// that the next step's kernels -- already enqueued -- find their shift, their source length and the stop flag in HBM.
struct AdvanceDev {
    int strategy;          // RIMU_SHIFT_*
    int is_int;
    double target_walkers, zeta, xi, dtau;
    long long max_length;  // 0 = no limit
    u64 dst_cap;           // capacity of the vector this step wrote
    u64 heavy_cap;
};
static __global__ void advance_ctl_kernel(StepCtl *ctl, const StatsDev *st, AdvanceDev a, double *shift_log) {
    if (ctl->stop) return;
    if (st->overflow_table || st->out_count > a.dst_cap || (st->heavy_packed >> 32) > a.heavy_cap) { ctl->stop = 2; return; }
    const double tnorm = a.is_int ? (double)st->inorm1 : st->norm1;
    const long long len = st->len;
    bool proceed = true;
    if (len > 0) {
        double shift = ctl->shift;
        const double pnorm = ctl->pnorm;
        switch (a.strategy) {
        case 0: proceed = tnorm < a.target_walkers; break;                                   // DontUpdate (pnorm untouched)
        case 1: shift -= a.zeta / a.dtau * log(tnorm / pnorm); ctl->pnorm = tnorm; break;     // LogUpdate
        case 2:                                                                               // LogUpdateAfterTargetWalkers
            if (ctl->shift_mode || tnorm > a.target_walkers) { ctl->shift_mode = 1; shift -= a.zeta / a.dtau * log(tnorm / pnorm); }
            ctl->pnorm = tnorm; break;
        case 3:                                                                               // DoubleLogUpdate
            shift -= a.xi / a.dtau * log(tnorm / a.target_walkers) + a.zeta / a.dtau * log(tnorm / pnorm);
            ctl->pnorm = tnorm; break;
        default:                                                                              // DoubleLogUpdateAfterTargetWalkers
            if (ctl->shift_mode || tnorm > a.target_walkers) {
                ctl->shift_mode = 1;
                shift -= a.xi / a.dtau * log(tnorm / a.target_walkers) + a.zeta / a.dtau * log(tnorm / pnorm);
            }
            ctl->pnorm = tnorm; break;
        }
        ctl->shift = shift;
    }
    shift_log[0] = ctl->shift; shift_log[1] = (double)ctl->shift_mode;
    ctl->n = st->out_count;
    ctl->steps_done += 1;
    if (len == 0 || (a.max_length > 0 && len > a.max_length) || !proceed) ctl->stop = 1;
}

// sum of the per-warp partial walker numbers of an ordered merge, in index order (one thread: the order IS the point)
static __global__ void ordered_sum_kernel(const double *__restrict__ partials, u32 n, double *out) {
    double s = 0.0;
    for (u32 i = 0; i < n; i++) s += partials[i];
    *out = s;
}

// ---------------------------------------------------------------- dot(::FrozenDVec, v): few keys against a big vector
// One CTA per query key: the key's bucket segment (<= a few thousand entries) is scanned, or the whole vector when it is not
// segmented.  Keys owned by another rank contribute there.
template <int W, class VT>
__global__ void __launch_bounds__(RIMU_TPB)
dot_sparse_kernel(const u64 *__restrict__ qkeys, const double *__restrict__ qvals, i64 nq, const u64 *__restrict__ keys,
                  const VT *__restrict__ vals, i64 n, const u64 *__restrict__ seg_start, const u32 *__restrict__ seg_len, u32 nb,
                  int rank, int nranks, double *__restrict__ out) {
    typedef typename BitsT<W>::type B;
    for (i64 q = blockIdx.x; q < nq; q += gridDim.x) {
        const B key = load_key<W>(qkeys + q * W);
        const u64 h = hash_bits(key);
        if (nranks > 1 && addr_owner(h, nranks) != rank) continue;
        i64 lo = 0, hi = n;
        if (nb) { const u32 b = bucket_of(h, nranks, nb); lo = (i64)seg_start[b]; hi = lo + (i64)seg_len[b]; }
        for (i64 i = lo + threadIdx.x; i < hi; i += blockDim.x)
            if (load_key<W>(keys + i * W) == key) atomicAdd(out, qvals[q] * (double)vals[i]); // at most one match per key
    }
}

// ---------------------------------------------------------------- re-segmentation of a vector for a new bucket count
template <int W>
__global__ void __launch_bounds__(RIMU_TPB)
bucket_count_kernel(const u64 *__restrict__ keys, i64 n, int nranks, u32 nb, u32 *__restrict__ counts) {
    typedef typename BitsT<W>::type B;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        B key = load_key<W>(keys + i * W);
        atomicAdd(&counts[bucket_of(hash_bits(key), nranks, nb)], 1u);
    }
}
// exclusive scan of counts -> seg_start (single CTA; nb is small compared with the vector); fill[] is zeroed
static __global__ void __launch_bounds__(1024)
bucket_scan_kernel(const u32 *__restrict__ counts, u32 nb, u64 *__restrict__ seg_start, u32 *__restrict__ seg_len, u32 *__restrict__ fill) {
    __shared__ u64 warp_tot[32];
    __shared__ u64 carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (u32 start = 0; start < nb; start += 1024) {
        u32 i = start + threadIdx.x;
        u64 v = i < nb ? counts[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u64 up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        u64 base = carry_s, tot = 0;
        for (int w = 0; w < 32; w++) { u64 t = warp_tot[w]; if (w < wid) base += t; tot += t; }
        if (i < nb) { seg_start[i] = base + incl - v; seg_len[i] = (u32)v; fill[i] = 0; }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += tot;
        __syncthreads();
    }
}
template <int W>
__global__ void __launch_bounds__(RIMU_TPB)
bucket_scatter_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ vals, const double *__restrict__ diag, i64 n, int nranks, u32 nb,
                      const u64 *__restrict__ seg_start, u32 *__restrict__ fill, u64 *__restrict__ okeys, u64 *__restrict__ ovals,
                      double *__restrict__ odiag) {
    typedef typename BitsT<W>::type B;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        B key = load_key<W>(keys + i * W);
        u32 b = bucket_of(hash_bits(key), nranks, nb);
        u64 at = seg_start[b] + atomicAdd(&fill[b], 1u);
        store_key<W>(okeys + at * W, key);
        ovals[at] = vals[i];
        if (diag) odiag[at] = diag[i]; // the cached diagonal elements move with their entries
    }
}
#endif // __CUDACC__
