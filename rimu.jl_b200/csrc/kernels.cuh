// kernels.cuh -- the FCIQMC step on one B200: diagonal/count, scan, spawn, annihilate, compress.
//
// Data layout in HBM
//   walker vector  : dense SoA, keys[n][W] (uint64 words, 16-B vector loads for W=2) + vals[n] (8 B)
//   working table  : open addressing, linear probing, one slot = {key words, value} padded to
//                    16 B (W=1) or 32 B (W=2) so that key and value share one 32-B DRAM sector.
//                    A slot is claimed with a 64-bit (W=1) or 128-bit (W=2, atom.cas.b128) CAS on
//                    the key and accumulated with a fire-and-forget RED on the value.  The table is
//                    the analogue of PDWorkingMemory (pdworkingmemory.jl:86-109): every deposit of a
//                    step lands in it, which *is* annihilation (pdworkingmemory.jl:21-31).
//   the compaction pass streams the table once, applies ThresholdCompression
//   (compression.jl:18-26), writes the new dense vector, re-empties the slots it visited and reduces
//   walkernumber/length (pdvec.jl:896-902) in the same sweep.
//
// Work decomposition for spawning (styles.jl apply_column!, spawning.jl spawn!): a determinant with
// value v makes max(floor(|v| boost),1) attempts (or L when it is spawned exactly), which is heavily
// skewed.  K1 computes per-parent attempt counts and a block-local scan, K2 scans the block totals,
// K3 runs ONE THREAD PER ATTEMPT over tiles of the global attempt index space; each tile stages its
// parents (keys, values, offsets) in shared memory and threads locate their parent by binary search.
#pragma once
#include "step_math.cuh" // StepDev + the per-deposit arithmetic (host-compilable, see tests/cuda/host_ham.cpp)
#include <type_traits>

#define RIMU_TPB 256
#define RIMU_TILE_ITEMS 4
#define RIMU_TILE (RIMU_TPB * RIMU_TILE_ITEMS)
#define RIMU_MAX_PROBE 2048
#define RIMU_MAX_RANKS 16

static const u64 EMPTY_KEY = ~0ull;


struct StatsDev {
    // block A: 16 x i64, summed over ranks
    i64 exact_steps, inexact_steps, spawn_attempts, len_before, len;
    i64 ispawns, ideaths, iclones, izombies, inorm1;
    i64 deposits; // non-zero deposits (diagonal + spawns) = table read-modify-writes
    i64 overflow_table, overflow_vec, overflow_xchg;
    u64 out_count;      // entries produced by compaction (may exceed capacity -> overflow_vec)
    u64 total_attempts; // written by the block-total scan
    // block B: 8 x double (first 5 summed over ranks by the step)
    double spawns, deaths, clones, zombies, norm1;
    double dot, norm2, norminf; // scratch for dot / norms
    // block C: local bookkeeping of the partitioned step (never reduced over ranks)
    u64 max_fill;       // fullest bucket (parents + records) seen by merge_kernel
    u64 records;        // spawn records this rank's merge read from its bucket streams (all source ranks)
    i64 sent;           // records this rank produced for buckets of other ranks (sent_records)
    u64 grow_flag;      // after the all-reduce: number of ranks whose result did not fit their target vector
    u64 heavy_packed;   // HeavyDev::packed lives here, so that one memset clears the step's statistics and the heavy-parent queue
};
#define RIMU_STATS_NI64 16 /* 'sent' was replaced by 'deposits' */
#define RIMU_STATS_NF64_STEP 5

struct TableDev {
    u64 *slots; // W=1: 2 u64 per slot; W=2: 4 u64 per slot
    u64 mask;   // active slots - 1
};

struct ExchangeDev { // per-peer send staging for multi-GPU spawn exchange
    u64 *keys;       // [nranks][cap][W]
    u64 *vals;       // [nranks][cap]
    u64 *counts;     // [nranks]
    u64 cap;
};

#ifdef __CUDACC__
template <class VT> DEV void atomic_add_val(u64 *p, VT v);
template <> DEV void atomic_add_val<double>(u64 *p, double v) { atomicAdd(reinterpret_cast<double *>(p), v); }
template <> DEV void atomic_add_val<i64>(u64 *p, i64 v) { atomicAdd(p, (u64)v); }

DEV void cas128(u64 *addr, u64 cmp_lo, u64 cmp_hi, u64 new_lo, u64 new_hi, u64 &old_lo, u64 &old_hi) {
    asm volatile(
        "{\n"
        ".reg .b128 c, n, o;\n"
        "mov.b128 c, {%2, %3};\n"
        "mov.b128 n, {%4, %5};\n"
        "atom.global.relaxed.gpu.cas.b128 o, [%6], c, n;\n"
        "mov.b128 {%0, %1}, o;\n"
        "}\n"
        : "=l"(old_lo), "=l"(old_hi)
        : "l"(cmp_lo), "l"(cmp_hi), "l"(new_lo), "l"(new_hi), "l"(addr)
        : "memory");
}

// deposit!(w, key, val): find-or-claim the slot of `key`, add `v`.  Returns false on overflow.
template <int W, class VT> DEV bool table_add(const TableDev &t, typename BitsT<W>::type key, u64 h, VT v) {
    u64 s = h & t.mask;
    if constexpr (W == 1) {
        for (int probe = 0; probe < RIMU_MAX_PROBE; probe++) {
            u64 *slot = t.slots + 2 * s;
            u64 cur = *slot;
            if (cur == EMPTY_KEY) cur = atomicCAS(slot, EMPTY_KEY, (u64)key);
            if (cur == EMPTY_KEY || cur == (u64)key) { atomic_add_val<VT>(slot + 1, v); return true; }
            s = (s + 1) & t.mask;
        }
    } else {
        const u64 klo = (u64)key, khi = (u64)(key >> 64);
        for (int probe = 0; probe < RIMU_MAX_PROBE; probe++) {
            u64 *slot = t.slots + 4 * s;
            ulonglong2 cur = *reinterpret_cast<const ulonglong2 *>(slot);
            if (cur.y == EMPTY_KEY) { // empty (valid keys never have an all-ones high word) or not yet visible
                u64 olo, ohi;
                cas128(slot, EMPTY_KEY, EMPTY_KEY, klo, khi, olo, ohi);
                if (olo == EMPTY_KEY && ohi == EMPTY_KEY) { cur.x = klo; cur.y = khi; }
                else { cur.x = olo; cur.y = ohi; }
            }
            if (cur.x == klo && cur.y == khi) { atomic_add_val<VT>(slot + 2, v); return true; }
            s = (s + 1) & t.mask;
        }
    }
    return false;
}

// read-only lookup (after all inserts of the building kernel completed)
template <int W> DEV bool table_find(const TableDev &t, typename BitsT<W>::type key, u64 h, u64 &val_bits) {
    u64 s = h & t.mask;
    for (int probe = 0; probe < RIMU_MAX_PROBE; probe++) {
        if constexpr (W == 1) {
            const u64 *slot = t.slots + 2 * s;
            u64 cur = slot[0];
            if (cur == (u64)key) { val_bits = slot[1]; return true; }
            if (cur == EMPTY_KEY) return false;
        } else {
            const u64 *slot = t.slots + 4 * s;
            ulonglong2 cur = *reinterpret_cast<const ulonglong2 *>(slot);
            if (cur.x == (u64)key && cur.y == (u64)(key >> 64)) { val_bits = slot[2]; return true; }
            if (cur.x == EMPTY_KEY && cur.y == EMPTY_KEY) return false;
        }
        s = (s + 1) & t.mask;
    }
    return false;
}

// route one deposit: local table or the owner's exchange segment
template <int W, class VT>
DEV void deposit(const TableDev &t, const ExchangeDev &x, const StepDev &p, StatsDev *st,
                 typename BitsT<W>::type key, VT v) {
    u64 h = hash_bits(key);
    if (p.nranks > 1) {
        int owner = addr_owner(h, p.nranks);
        if (owner != p.rank) {
            u64 idx = atomicAdd(&x.counts[owner], 1ull);
            if (idx < x.cap) {
                store_key<W>(x.keys + ((u64)owner * x.cap + idx) * W, key);
                union { VT v; u64 b; } cv; cv.v = v;
                x.vals[(u64)owner * x.cap + idx] = cv.b;
            } else st->overflow_xchg = 1;
            return;
        }
    }
    if (!table_add<W, VT>(t, key, h, v)) st->overflow_table = 1;
}

// block-wide reductions of a few accumulators into StatsDev with one atomic per warp
DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DEV i64 warp_sum(i64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DEV void stat_add(double *dst, double v) { v = warp_sum(v); if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(dst, v); }
DEV void stat_add(i64 *dst, i64 v) { v = warp_sum(v); if ((threadIdx.x & 31) == 0 && v != 0) atomicAdd((u64 *)dst, (u64)v); }

// statistics <-> packed doubles for the single per-step all-reduce (dir 0: pack, 1: unpack).  Two more doubles ride along:
// the records merged (block C; summed so that every rank sizes the next bucket count identically) and "my result did not fit
// my target vector" (out_count > dst_cap), so that all ranks decide to grow-and-repeat together without another collective.
#define RIMU_STATS_NPACK (RIMU_STATS_NI64 + RIMU_STATS_NF64_STEP + 2)
static __global__ void pack_stats_kernel(const StatsDev *st, double *buf, int /*dir: pack only*/, u64 dst_cap) {
    const int i = threadIdx.x;
    const i64 *ints = reinterpret_cast<const i64 *>(st);
    const double *dbl = reinterpret_cast<const double *>(reinterpret_cast<const char *>(st) + RIMU_STATS_NI64 * sizeof(i64));
    if (i < RIMU_STATS_NI64) buf[i] = (double)ints[i];
    if (i < RIMU_STATS_NF64_STEP) buf[RIMU_STATS_NI64 + i] = dbl[i];
    if (i == 31) {
        const int at = RIMU_STATS_NI64 + RIMU_STATS_NF64_STEP;
        buf[at] = (double)st->records;
        buf[at + 1] = st->out_count > dst_cap ? 1.0 : 0.0;
    }
}

#endif // __CUDACC__ (the host half of the packing is plain C++)
// the all-reduced doubles back into the statistics block (host side; mirrors pack_stats_kernel)
static inline void unpack_stats_host(StatsDev *st, const double *buf) {
    i64 *ints = reinterpret_cast<i64 *>(st);
    double *dbl = reinterpret_cast<double *>(reinterpret_cast<char *>(st) + RIMU_STATS_NI64 * sizeof(i64));
    for (int i = 0; i < RIMU_STATS_NI64; i++) ints[i] = (i64)llrint(buf[i]);
    for (int i = 0; i < RIMU_STATS_NF64_STEP; i++) dbl[i] = buf[RIMU_STATS_NI64 + i];
    const int at = RIMU_STATS_NI64 + RIMU_STATS_NF64_STEP;
    st->records = (u64)llrint(buf[at]);
    st->grow_flag = (u64)llrint(buf[at + 1]);
}
#ifdef __CUDACC__
// ---------------------------------------------------------------- K1: diagonal step + attempt counts
template <int HK, int W, class VT>
__global__ void __launch_bounds__(RIMU_TPB)
diag_count_kernel(const HamDev h, const StepDev p, const u64 *__restrict__ keys, const VT *__restrict__ vals, i64 n,
                  TableDev tab, ExchangeDev xch, u64 *__restrict__ local_off, u64 *__restrict__ block_tot, StatsDev *st) {
    typedef typename BitsT<W>::type B;
    __shared__ u64 warp_tot[RIMU_TPB / 32];
    i64 gid = (i64)blockIdx.x * RIMU_TPB + threadIdx.x;
    u64 cnt = 0;
    double clones = 0, deaths = 0, zombies = 0;
    i64 exact_steps = 0, inexact_steps = 0, ndep = 0;
    if (gid < n) {
        B key = load_key<W>(keys + gid * W);
        VT pv = vals[gid];
        double val = (double)pv;
        // diagonal_step! (spawning.jl:73-77) through FirstOrderTransitionOperator (fciqmc.jl:93-96)
        double hd = ham_diagonal<HK, B>(h, key);
        double d = p.plain_h ? hd : 1 - p.dtau * (hd - p.shift);
        double r = 0.0;
        constexpr bool is_int = std::is_integral<VT>::value;
        double thr = is_int ? 0.0 : p.proj_thr;
        if (is_int || thr > 0.0) {
            u32 rnd[4];
            rng_draw(hash_bits(key), 0, STREAM_DIAG, p.k0, p.k1, rnd);
            r = u53(rnd[1], rnd[2]);
        }
        VT res = project_value<VT>(d * val, thr, r);
        if (res != (VT)0) { deposit<W, VT>(tab, xch, p, st, key, res); ndep = 1; }
        // clones_deaths_zombies (spawning.jl:79-93)
        double rs = (double)res;
        if (rs > val) clones = fabs(rs - val);
        else if (sgn_(rs) != sgn_(val)) { deaths = fabs(val); zombies = fabs(rs); }
        else deaths = fabs(rs - val);
        long long L = ham_num_offdiagonals<HK, B>(h, key);
        bool exact = attempts_for(p, val, L, cnt);
        if (cnt) { if (exact) exact_steps = 1; else inexact_steps = 1; }
    }
    // block-local exclusive scan of cnt
    u64 incl = cnt;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    u64 base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < RIMU_TPB / 32; w++) { u64 t = warp_tot[w]; if (w < wid) base += t; tot += t; }
    if (gid < n) local_off[gid] = base + incl - cnt;
    if (threadIdx.x == 0) block_tot[blockIdx.x] = tot;
    if (std::is_integral<VT>::value) {
        stat_add(&st->iclones, (i64)clones); stat_add(&st->ideaths, (i64)deaths); stat_add(&st->izombies, (i64)zombies);
    } else {
        stat_add(&st->clones, clones); stat_add(&st->deaths, deaths); stat_add(&st->zombies, zombies);
    }
    stat_add(&st->exact_steps, exact_steps); stat_add(&st->inexact_steps, inexact_steps); stat_add(&st->deposits, ndep);
}

// ---------------------------------------------------------------- K2: exclusive scan of block totals (single CTA)
static __global__ void __launch_bounds__(1024) scan_blocks_kernel(const u64 *__restrict__ block_tot, i64 nblocks,
                                                            u64 *__restrict__ block_base, StatsDev *st) {
    __shared__ u64 warp_tot[32];
    __shared__ u64 carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (i64 start = 0; start < nblocks; start += 1024) {
        i64 i = start + threadIdx.x;
        u64 v = i < nblocks ? block_tot[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u64 up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        u64 base = carry_s, tot = 0;
        for (int w = 0; w < 32; w++) { u64 t = warp_tot[w]; if (w < wid) base += t; tot += t; }
        if (i < nblocks) block_base[i] = base + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) { block_base[nblocks] = carry_s; st->total_attempts = carry_s; st->spawn_attempts = (i64)carry_s; }
}

// global attempt offset of parent j
DEV u64 parent_offset(const u64 *__restrict__ block_base, const u64 *__restrict__ local_off, i64 j, i64 n) {
    if (j >= n) return block_base[(n + RIMU_TPB - 1) / RIMU_TPB];
    return block_base[j / RIMU_TPB] + local_off[j];
}

// ---------------------------------------------------------------- K3: spawning, one thread per attempt
template <int HK, int W, class VT>
__global__ void __launch_bounds__(RIMU_TPB)
spawn_kernel(const HamDev h, const StepDev p, const u64 *__restrict__ keys, const VT *__restrict__ vals, i64 n,
             const u64 *__restrict__ block_base, const u64 *__restrict__ local_off,
             TableDev tab, ExchangeDev xch, StatsDev *st) {
    typedef typename BitsT<W>::type B;
    constexpr bool is_int = std::is_integral<VT>::value;
    __shared__ u64 s_off[RIMU_TILE + 1];
    __shared__ u64 s_keys[RIMU_TILE * W];
    __shared__ VT s_vals[RIMU_TILE];
    __shared__ i64 s_p0;
    const u64 total = st->total_attempts;
    const i64 nblk = (n + RIMU_TPB - 1) / RIMU_TPB;
    double spawns = 0.0;
    i64 ndep = 0;
    for (u64 tile = blockIdx.x; tile * RIMU_TILE < total; tile += gridDim.x) {
        const u64 t0 = tile * RIMU_TILE;
        const u64 t1 = min(t0 + (u64)RIMU_TILE, total);
        // locate the parent containing attempt t0: last j with offset(j) <= t0  (warp 0, 32-ary search)
        if (threadIdx.x < 32) {
            int lane = threadIdx.x;
            // coarse: over block_base[0..nblk]
            i64 lo = 0, hi = nblk; // invariant: block_base[lo] <= t0 < block_base[hi] (block_base[nblk] = total > t0)
            while (hi - lo > 1) {
                i64 span = hi - lo;
                i64 step = (span + 31) / 32;
                i64 probe = lo + (i64)(lane + 1) * step;
                bool le = probe < hi ? (block_base[probe] <= t0) : false;
                unsigned m = __ballot_sync(0xffffffffu, le);
                int c = __popc(m); // probes are monotone: first c lanes true
                i64 nlo = lo + (i64)c * step, nhi = lo + (i64)(c + 1) * step;
                lo = c ? nlo : lo;
                hi = nhi < hi ? nhi : hi;
            }
            // fine: inside block lo, parents [lo*TPB, min(n,(lo+1)*TPB))
            i64 b0 = lo * RIMU_TPB, b1 = min(n, b0 + (i64)RIMU_TPB);
            u64 base = block_base[lo];
            int cnt = 0;
            for (i64 j = b0 + lane; j < b1; j += 32) cnt += (base + local_off[j] <= t0) ? 1 : 0;
            cnt = (int)warp_sum((i64)cnt);
            if (lane == 0) s_p0 = b0 + cnt - 1; // offsets are non-decreasing; cnt >= 1
        }
        __syncthreads();
        i64 pbase = s_p0;
        // windows of parents starting at pbase until the tile is covered
        for (;;) {
            for (int j = threadIdx.x; j <= RIMU_TILE; j += RIMU_TPB) s_off[j] = parent_offset(block_base, local_off, pbase + j, n);
            for (int j = threadIdx.x; j < RIMU_TILE; j += RIMU_TPB) {
                i64 g = pbase + j;
                if (g < n) {
                    B k = load_key<W>(keys + g * W);
                    s_keys[j * W] = (u64)k;
                    if constexpr (W == 2) s_keys[j * W + 1] = (u64)(k >> 64);
                    s_vals[j] = vals[g];
                }
            }
            __syncthreads();
            const u64 wlo = s_off[0], whi = s_off[RIMU_TILE];
#pragma unroll 1
            for (int it = 0; it < RIMU_TILE_ITEMS; it++) {
                u64 a = t0 + (u64)it * RIMU_TPB + threadIdx.x;
                if (a >= t1 || a < wlo || a >= whi) continue;
                // last j in [0, TILE) with s_off[j] <= a
                int lo = 0, hi = RIMU_TILE;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_off[mid] <= a) lo = mid; else hi = mid; }
                const u64 k = a - s_off[lo];
                B key;
                if constexpr (W == 1) key = s_keys[lo];
                else key = ((u128)s_keys[lo * 2 + 1] << 64) | (u128)s_keys[lo * 2];
                const VT pv = s_vals[lo];
                const double val = (double)pv;
                const long long L = ham_num_offdiagonals<HK, B>(h, key);
                u64 nat;
                const bool exact = attempts_for(p, val, L, nat);
                B child;
                if (exact) { // spawn!(Exact) spawning.jl:174-182
                    double m = ham_offdiagonal<HK, B>(h, key, (long long)k, child);
                    if (!p.plain_h) m = -m * p.dtau;
                    double r = 0.0;
                    if (p.proj_thr > 0.0) {
                        u32 rnd[4];
                        rng_draw(hash_bits(key), k, STREAM_SPAWN, p.k0, p.k1, rnd);
                        r = u53(rnd[1], rnd[2]);
                    }
                    double nv = project_value<double>(val * m, p.proj_thr, r);
                    if (nv != 0.0) {
                        if constexpr (!is_int) { deposit<W, VT>(tab, xch, p, st, child, nv); ndep++; }
                        spawns += fabs(nv);
                    }
                } else { // spawn!(WithReplacement) spawning.jl:232-243
                    u32 rnd[4];
                    rng_draw(hash_bits(key), k, STREAM_SPAWN, p.k0, p.k1, rnd);
                    long long i = (long long)(((u64)rnd[0] * (u64)L) >> 32);
                    double m = ham_offdiagonal<HK, B>(h, key, i, child);
                    if (!p.plain_h) m = -m * p.dtau;
                    double magnitude = nat == 1 ? val : val / (double)nat;
                    double nv0 = m * magnitude * (double)L; // (= m * magnitude / (1 / L) up to one rounding; see step_math.cuh)
                    VT nv = project_value<VT>(nv0, is_int ? 0.0 : p.proj_thr, u53(rnd[1], rnd[2]));
                    if (nv != (VT)0) {
                        deposit<W, VT>(tab, xch, p, st, child, nv); ndep++;
                        spawns += fabs((double)nv);
                    }
                }
            }
            __syncthreads();
            if (whi >= t1 || pbase + RIMU_TILE >= n) break;
            pbase += RIMU_TILE;
        }
    }
    if (is_int) stat_add(&st->ispawns, (i64)spawns);
    else stat_add(&st->spawns, spawns);
    stat_add(&st->deposits, ndep);
}

// ---------------------------------------------------------------- generic record insertion (exchange receive, upload, axpby)
template <int W, class VT>
__global__ void __launch_bounds__(RIMU_TPB)
insert_records_kernel(const u64 *__restrict__ keys, const VT *__restrict__ vals, i64 n, double scale, int use_scale,
                      int rank, int nranks, TableDev tab, StatsDev *st) {
    typedef typename BitsT<W>::type B;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        B key = load_key<W>(keys + i * W);
        VT v = vals[i];
        if (use_scale) v = (VT)(scale * (double)v);
        if (v == (VT)0) continue;
        u64 h = hash_bits(key);
        if (nranks > 1 && addr_owner(h, nranks) != rank) continue; // setindex! drops non-local keys, pdvec.jl:336-349
        if (!table_add<W, VT>(tab, key, h, v)) st->overflow_table = 1;
    }
}

// ---------------------------------------------------------------- K5: move_and_compress! + walkernumber_and_length
template <int W, class VT>
__global__ void __launch_bounds__(RIMU_TPB)
compact_kernel(TableDev tab, const StepDev p, u64 *__restrict__ out_keys, VT *__restrict__ out_vals, u64 out_cap, StatsDev *st) {
    typedef typename BitsT<W>::type B;
    constexpr bool is_int = std::is_integral<VT>::value;
    constexpr int SW = W == 1 ? 2 : 4;
    const u64 nslots = tab.mask + 1;
    double norm1 = 0.0;
    i64 inorm1 = 0, len_before = 0, len = 0;
    const int lane = threadIdx.x & 31;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 nround = (nslots + stride - 1) / stride;
    for (u64 rd = 0; rd < nround; rd++) {
        u64 s = rd * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        bool keep = false;
        B key = 0;
        VT v = 0;
        if (s < nslots) {
            u64 *slot = tab.slots + SW * s;
            bool used;
            if constexpr (W == 1) {
                ulonglong2 sv = *reinterpret_cast<const ulonglong2 *>(slot);
                used = sv.x != EMPTY_KEY;
                key = sv.x;
                union { u64 b; VT v; } cv; cv.b = sv.y; v = cv.v;
                if (used) *reinterpret_cast<ulonglong2 *>(slot) = make_ulonglong2(EMPTY_KEY, 0ull);
            } else {
                ulonglong2 k2 = *reinterpret_cast<const ulonglong2 *>(slot);
                used = !(k2.x == EMPTY_KEY && k2.y == EMPTY_KEY);
                if (used) {
                    key = ((u128)k2.y << 64) | (u128)k2.x;
                    union { u64 b; VT v; } cv; cv.b = slot[2]; v = cv.v;
                    *reinterpret_cast<ulonglong2 *>(slot) = make_ulonglong2(EMPTY_KEY, EMPTY_KEY);
                    slot[2] = 0ull;
                }
            }
            if (used && v != (VT)0) { // exact zeros are deleted (pdworkingmemory.jl:25-29)
                len_before++;
                if constexpr (!is_int) {
                    if (p.compress_thr > 0.0) { // ThresholdCompression (compression.jl:18-26)
                        double prob = fabs(v) / p.compress_thr;
                        if (prob < 1) {
                            u32 rnd[4];
                            rng_draw(hash_bits(key), 0, STREAM_COMPRESS, p.k0, p.k1, rnd);
                            v = (prob > u53(rnd[1], rnd[2])) ? p.compress_thr * sgn_(v) : 0.0;
                        }
                    }
                }
                keep = v != (VT)0;
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            u64 base = 0;
            if (lane == 0) base = atomicAdd(&st->out_count, (u64)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) {
                u64 idx = base + __popc(m & ((1u << lane) - 1));
                if (idx < out_cap) {
                    store_key<W>(out_keys + idx * W, key);
                    out_vals[idx] = v;
                }
                len++;
                if (is_int) inorm1 += (i64)(v < (VT)0 ? -v : v); else norm1 += fabs((double)v);
            }
        }
    }
    stat_add(&st->len_before, len_before);
    stat_add(&st->len, len);
    if (is_int) stat_add(&st->inorm1, inorm1); else stat_add(&st->norm1, norm1);
}

static __global__ void table_fill_empty_kernel(u64 *slots, u64 nslots, int W) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    if (W == 1) {
        for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += stride)
            reinterpret_cast<ulonglong2 *>(slots)[s] = make_ulonglong2(EMPTY_KEY, 0ull);
    } else {
        for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += stride) {
            reinterpret_cast<ulonglong2 *>(slots)[2 * s] = make_ulonglong2(EMPTY_KEY, EMPTY_KEY);
            reinterpret_cast<ulonglong2 *>(slots)[2 * s + 1] = make_ulonglong2(0ull, 0ull);
        }
    }
}

// ---------------------------------------------------------------- dense vector kernels
template <class VT> __global__ void norm_kernel(const VT *__restrict__ vals, i64 n, StatsDev *st) {
    double n1 = 0.0, n2 = 0.0, ninf = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double a = fabs((double)vals[i]);
        n1 += a; n2 += a * a; ninf = fmax(ninf, a);
    }
    stat_add(&st->norm1, n1);
    stat_add(&st->norm2, n2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ninf = fmax(ninf, __shfl_xor_sync(0xffffffffu, ninf, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long *>(&st->norminf), (unsigned long long)__double_as_longlong(ninf));
}

// scale! / eltype conversion.  An Int64 vector can only hold integral results: anything else raises the flag (the reference
// throws InexactError there instead of truncating), and zeros can then never appear silently.
template <class VT> __global__ void scale_kernel(VT *vals, i64 n, double alpha, i64 *inexact) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double r = alpha * (double)vals[i];
        if (std::is_integral<VT>::value && (r != rint(r) || fabs(r) >= 9007199254740992.0)) { *inexact = 1; continue; }
        vals[i] = (VT)r;
    }
}

template <class From, class To> __global__ void convert_vals_kernel(const From *__restrict__ in, To *__restrict__ out, i64 n, i64 *inexact) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const From x = in[i];
        if (std::is_integral<To>::value && !std::is_integral<From>::value && ((double)x != rint((double)x) || fabs((double)x) >= 9007199254740992.0)) *inexact = 1;
        out[i] = (To)x;
    }
}

// dot(x, y): x was inserted into the table, y is streamed and looked up
template <int W, class VT>
__global__ void lookup_dot_kernel(const u64 *__restrict__ keys, const VT *__restrict__ vals, i64 n, TableDev tab, StatsDev *st) {
    typedef typename BitsT<W>::type B;
    double acc = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        B key = load_key<W>(keys + i * W);
        u64 bits;
        if (table_find<W>(tab, key, hash_bits(key), bits)) {
            union { u64 b; VT v; } cv; cv.b = bits;
            acc += (double)cv.v * (double)vals[i];
        }
    }
    stat_add(&st->dot, acc);
}

// getindex: linear scan for one key
template <int W, class VT>
__global__ void find_key_kernel(const u64 *__restrict__ keys, const VT *__restrict__ vals, i64 n, u64 k0, u64 k1, VT *out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        bool eq = keys[i * W] == k0;
        if (W == 2) eq = eq && keys[i * W + 1] == k1;
        if (eq) *out = vals[i];
    }
}

// ---------------------------------------------------------------- element-wise Hamiltonian hooks (tests / interface parity)
template <int HK, int W>
__global__ void ham_diag_kernel(const HamDev h, const u64 *__restrict__ keys, i64 n, double *out, i64 *nod) {
    typedef typename BitsT<W>::type B;
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    B key = load_key<W>(keys + i * W);
    if (out) out[i] = ham_diagonal<HK, B>(h, key);
    if (nod) nod[i] = ham_num_offdiagonals<HK, B>(h, key);
}
template <int HK, int W>
__global__ void ham_offdiag_kernel(const HamDev h, const u64 *__restrict__ key_in, i64 first0, i64 count, u64 *keys_out, double *vals_out) {
    typedef typename BitsT<W>::type B;
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    B key = load_key<W>(key_in), child;
    vals_out[i] = ham_offdiagonal<HK, B>(h, key, first0 + i, child);
    store_key<W>(keys_out + i * W, child);
}
#endif // __CUDACC__
