// internal.cuh -- what the translation units of librimu_b200.so share: error plumbing, the NCCL function table, the
// handle structs and the host helpers of the step.  api.cu holds the C ABI and everything that does not depend on the
// Hamiltonian kind; step_hk.cu is compiled once per HamKind (-DRIMU_HK=n) so that the kernels of the seven models
// build in parallel.  Nothing here is exported: RIMU_INTERNAL symbols have hidden visibility.
#pragma once
#include "../../include/rimu_b200.h"
#include "partition.cuh"
#include "sector.cuh"
#include "ham_host.h"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#define RIMU_INTERNAL __attribute__((visibility("hidden")))

// ---------------------------------------------------------------- errors
RIMU_INTERNAL int fail(int code, const char *fmt, ...);
RIMU_INTERNAL extern size_t g_alloc_fail_bytes;
RIMU_INTERNAL std::string oom_note();
// cudaMalloc with diagnostics: RIMU_B200_TRACE_ALLOC=1 logs every allocation above 64 MiB; a failure reports the
// request and the free/total device memory and clears CUDA's "last error" so that it cannot surface at a later,
// unrelated cudaGetLastError() check
template <class T> static cudaError_t rimu_malloc(T **p, size_t bytes) {
    static const bool trace = getenv("RIMU_B200_TRACE_ALLOC") != nullptr;
    cudaError_t e = cudaMalloc((void **)p, bytes);
    if (trace && bytes >= (64u << 20)) {
        size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot);
        fprintf(stderr, "[rimu_b200] cudaMalloc %.1f MiB -> %s (free %.1f GiB of %.1f GiB)\n", bytes / 1048576.0,
                e == cudaSuccess ? "ok" : cudaGetErrorString(e), fr / 1073741824.0, tot / 1073741824.0);
    }
    if (e != cudaSuccess) { g_alloc_fail_bytes = bytes; cudaGetLastError(); }
    return e;
}
// Entry of every API call that touches the device: select the context's GPU and drop any stale, non-sticky error that an
// earlier benign failure in this thread (ours, NCCL's or the host framework's) left in CUDA's "last error" slot -- otherwise
// it would be reported by the first cudaGetLastError() after one of OUR launches.  Sticky errors survive this and are
// still caught by the next call.
static inline cudaError_t enter_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    cudaGetLastError();
    return e;
}
#define CUDA_TRY(x)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess)                                                                        \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? RIMU_ERR_NO_DEVICE : RIMU_ERR_CUDA, \
                        "%s failed: %s (%s:%d)%s", #x, cudaGetErrorString(e_), __FILE__, __LINE__,    \
                        e_ == cudaErrorMemoryAllocation ? oom_note().c_str() : "");                   \
    } while (0)
#define TRY(x)              \
    do {                    \
        int r_ = (x);       \
        if (r_ != 0) return r_; \
    } while (0)

// ---------------------------------------------------------------- NCCL (resolved lazily; same soname as torch's bundled copy)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclInt64 = 4, ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char *(*GetErrorString)(int);
};
RIMU_INTERNAL extern NcclApi g_nccl;
RIMU_INTERNAL int nccl_load();
#define NCCL_TRY(x)                                                                              \
    do {                                                                                         \
        int e_ = (x);                                                                            \
        if (e_ != 0) return fail(RIMU_ERR_NCCL, "%s failed: %s", #x, g_nccl.GetErrorString(e_)); \
        cudaGetLastError(); /* NCCL succeeded: whatever benign CUDA error it left behind is not ours */ \
    } while (0)

// ---------------------------------------------------------------- handles
struct rimu_ctx {
    int device, W, sm_count;
    cudaStream_t stream;
    u64 *table;
    u64 table_slots; // capacity (power of two)
    StatsDev *d_stats, *h_stats, *h_stats_local;
    u64 *local_off, *block_tot, *block_base;
    u64 scratch_parents;
    cudaEvent_t ev[8];
    double *d_red = nullptr; // packed statistics for the single all-reduce (RIMU_STATS_NPACK doubles)
    double *h_red = nullptr; // ... and their pinned host copy
    unsigned long long launches; // kernels launched by this context (bench bookkeeping)
    // staging for host <-> device transfers
    u64 *stage_keys; void *stage_vals; u64 stage_cap;
    // comm
    ncclComm_t comm;
    int rank, nranks;
    ExchangeDev xch;
    u64 *recv_keys, *recv_vals, recv_cap;
    u64 *d_allcounts, *h_allcounts;
    double *d_reduce;
    // partitioned step (partition.cuh): bucket record streams, heavy-parent queue, re-segmentation scratch
    int method;              // RIMU_ANNIHILATE_PARTITION (default) or RIMU_ANNIHILATE_HASH
    PartDev part;            // record streams of the FCIQMC step (direct mode: peers store into them over NVLink)
    u64 part_nb_cap;         // buckets the record streams are allocated for
    PartDev lpart;           // direct mode only: private streams for local operations (upload, axpby, annihilate), which
    u64 lpart_nb_cap;        //   run between steps while a faster peer may already be filling the step streams
    int direct;              // multi-GPU: peers can map each other's streams (CUDA IPC) -> spawned records are stored
                             //   straight into the owner's bucket sub-streams, no receive pass
    u64 last_dst_uid, last_dst_version; double last_g_len; // identity (uid, never reused) + state of the previous step's result
    u64 next_vec_uid;        // vectors are numbered in creation order: the same numbers on every rank (same call sequence)
    u32 merge_grid_cap;      // RIMU_B200_MERGE_GRID: cap on merge CTAs (tests force many buckets per CTA with it)
    int live_vecs;           // vectors created on this context and not yet destroyed
    int detached;            // rimu_comm_detach ran: the peer mappings are closed, steps are no longer possible
    int dead;                // rimu_ctx_destroy was called while vectors were alive: the struct (and stream) live on until the
                             //   last of them is destroyed (host GCs finalise vectors and contexts in arbitrary order) // global length of the previous step's result
    HeavyDev heavy;
    u32 *bucket_tmp; u64 bucket_tmp_cap; // [2][nb] counts / fill
    u64 xch_worst;           // largest per-peer record count seen by a failed exchange
    u64 xch_want;            // per-peer capacity the staging buffers get when they are first needed
    int p2p_used;            // the last exchange went peer-direct: counts are read from h_allcounts after the final sync
    void *peer_open[2][RIMU_MAX_RANKS]; // IPC-opened peer stream buffers (records, sub-stream fills)
    char *d_ipc, *h_ipc;     // all-gather scratch for the IPC handles
    alignas(16) char sort_scratch[96];   // SortScratch of sort.cu (opaque here)
    double rec_per_parent;   // running estimate: records appended per parent (sizes the bucket count)
    double *d_ord; u32 ord_cap; // per-warp partial walker numbers of an ordered merge (fixed-order reduction)
    int rcnt_clean;          // one rank: every fill counter of `part` is zero (the last merge cleared what it consumed)
    u32 ovf_nb; double ovf_expected; // the last bucket count that overflowed and the expected item count it overflowed at:
                             //   choose_buckets stays above it until the vector has shrunk (no flip-flop between a count
                             //   that overflows and the retry's larger one)
    u64 *spare_keys, *spare_vals; double *spare_diag; u64 spare_cap; // ping-pong partner of rebucket(): re-segmenting a
                             //   vector swaps it into these buffers and keeps the old ones as the next spare (no cudaMalloc,
                             //   no synchronisation when a host-fed vector is re-segmented every step)
    u64 last_max_fill;
    // rimu_advance (batches of steps without host synchronisation): one statistics block per step of a chunk, the device-resident
    // controller, the per-step shift log, and a snapshot of the chunk's first source vector (restored when a step of the chunk
    // ran out of working memory: the chunk is then repeated step by step)
    StatsDev *d_ring = nullptr, *h_ring = nullptr;
    StepCtl *d_ctl = nullptr, *h_ctl = nullptr;
    double *d_shiftlog = nullptr, *h_shiftlog = nullptr;
    u64 *snap_keys = nullptr, *snap_vals = nullptr, *snap_seg_start = nullptr; double *snap_diag = nullptr; u32 *snap_seg_len = nullptr;
    u64 snap_cap = 0, snap_nb_cap = 0;
    u32 adv_grid = 0;        // spawn grid of the current batch
    u64 *proj_keys = nullptr; double *proj_vals = nullptr; u64 proj_cap = 0; // frozen projectors of rimu_advance, resident for the call
    double *d_projlog = nullptr, *h_projlog = nullptr;                        // [RIMU_ADVANCE_CHUNK][RIMU_MAX_PROJECTORS] dots
};
// parents per spawn chunk: SPAWN_NT for vectors that fill the GPU anyway; small vectors are cut so that ~every SM gets a chunk
static inline u32 spawn_chunk_parents(i64 n, int sm_count) {
    u32 ppc = SPAWN_NT;
    while (ppc > 32 && (i64)(ppc / 2) * sm_count >= n) ppc /= 2;
    return ppc;
}
#define RIMU_ADVANCE_CHUNK 128        /* steps enqueued per host synchronisation */
#define RIMU_ADVANCE_MAX_N (1u << 21) /* batches pay off while a step is launch/latency bound; beyond this, step by step */
RIMU_INTERNAL int enter_ctx(rimu_ctx *c);
struct rimu_ham {
    rimu_ham_desc desc;
    int hk, W, device;
    HamDev dev;
    double *d_tables;
    unsigned char *d_nbr;
    u64 uid; // identifies this Hamiltonian in the vectors' diagonal-element caches
};
struct rimu_vec {
    rimu_ctx *ctx;
    int vt;
    u64 cap;
    i64 n;
    u64 *keys;
    void *vals;
    // bucket segmentation (partition.cuh); nb == 0: not segmented
    u32 nb;
    u64 seg_cap;
    u64 *seg_start;
    u32 *seg_len;
    // cache of diagonal_element(H, address) per entry, written by the partitioned step; valid for the
    // Hamiltonian with uid diag_uid (0 = invalid).  Saves re-evaluating H_aa for every parent every step.
    double *diag;
    u64 diag_cap, diag_uid;
    u64 version;             // bumped by every mutating API call (the same call sequence runs on every rank)
    u64 uid;                 // creation number within the context (a freed vector's address may be reused, its uid never is)
};

static inline u64 next_pow2(u64 x) { u64 p = 1; while (p < x) p <<= 1; return p; }
static inline int grid_for(i64 n, int sm_count, int per_sm = 8) {
    i64 g = (n + RIMU_TPB - 1) / RIMU_TPB;
    i64 cap = (i64)sm_count * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---------------------------------------------------------------- host helpers shared by api.cu and step_hk.cu
RIMU_INTERNAL int ensure_scratch(rimu_ctx *c, u64 parents);
RIMU_INTERNAL int ensure_seg(rimu_vec *v, u32 nb);
RIMU_INTERNAL int ensure_diag(rimu_vec *v);
RIMU_INTERNAL int ensure_part(rimu_ctx *c, u32 nb, u32 nlane = 1);
RIMU_INTERNAL int ensure_heavy(rimu_ctx *c, u64 parents);
RIMU_INTERNAL int rebucket(rimu_vec *v, u32 nb); // re-segment a vector for nb buckets (contents unchanged; cached H_aa follow)
RIMU_INTERNAL int exchange_spawns(rimu_ctx *c, int vt, u64 slots, i64 *sent_out, bool to_streams);
static inline u32 part_cap_items(int W) { return W == 1 ? (u32)PartCap<1>::value : (u32)PartCap<2>::value; }
static inline size_t part_smem_bytes(int W) { u32 cap = part_cap_items(W); return (size_t)cap * W * 8 + (size_t)cap * 8 + (size_t)cap * 2 * 4 + (size_t)cap * 2; }

template <int HK, int W> struct HkTag { static constexpr int hk = HK; static constexpr int w = W; };
template <class F> static int dispatch_wv(int W, int vt, F &&f) {
    if (W == 1) return vt == RIMU_VAL_F64 ? f(HkTag<0, 1>(), double()) : f(HkTag<0, 1>(), i64());
    return vt == RIMU_VAL_F64 ? f(HkTag<0, 2>(), double()) : f(HkTag<0, 2>(), i64());
}
// drain the working table into dst (no compression unless p says so)
template <int W, class VT> static int compact_into(rimu_ctx *c, rimu_vec *dst, u64 slots, const StepDev &p) {
    TableDev tab{c->table, slots - 1};
    compact_kernel<W, VT><<<grid_for((i64)slots, c->sm_count, 16), RIMU_TPB, 0, c->stream>>>(
        tab, p, dst->keys, (VT *)dst->vals, dst->cap, c->d_stats);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------- per-HamKind entry points (step_hk.cu, one object per kind)
struct HkOps {
    // one attempt of the step with a fixed bucket count / table size (rimu_step owns the retry loop)
    int (*step)(rimu_ctx *c, const rimu_ham *h, const StepDev &p, rimu_vec *src, rimu_vec *dst, bool use_part, bool is_int,
                u32 nb, u64 slots, i64 *sent);
    // element-wise hooks on device buffers: diagonal_element / num_offdiagonals, get_offdiagonal(first0 .. first0+count)
    int (*diag)(rimu_ctx *c, const rimu_ham *h, const u64 *d_keys, i64 n, double *d_out, i64 *d_nod);
    int (*offdiag)(rimu_ctx *c, const rimu_ham *h, const u64 *d_key, i64 first0, i64 count, u64 *d_keys_out, double *d_vals);
    // y = H x over a complete sector in dense (combinadic-rank) indexing: sector.cuh
    int (*sector_mul)(rimu_ctx *c, const rimu_ham *h, const SectorDev *s, const u64 *d_keys, const double *d_x, double *d_y, u64 dim);
    // batch mode (rimu_advance): enqueue the kernels of ONE partitioned step whose shift / source length / stop flag live in
    // *p.ctl; statistics go to *st; no events, no memsets, no synchronisation.  The caller has sized every buffer.
    int (*enqueue)(rimu_ctx *c, const rimu_ham *h, const StepDev &p, rimu_vec *src, rimu_vec *dst, bool is_int, u32 nb, StatsDev *st);
};
RIMU_INTERNAL const HkOps *rimu_hk_ops_0();
RIMU_INTERNAL const HkOps *rimu_hk_ops_1();
RIMU_INTERNAL const HkOps *rimu_hk_ops_2();
RIMU_INTERNAL const HkOps *rimu_hk_ops_3();
RIMU_INTERNAL const HkOps *rimu_hk_ops_4();
RIMU_INTERNAL const HkOps *rimu_hk_ops_5();
RIMU_INTERNAL const HkOps *rimu_hk_ops_6();
RIMU_INTERNAL const HkOps *rimu_hk_ops_7();
RIMU_INTERNAL const HkOps *rimu_hk_ops_8();
RIMU_INTERNAL const HkOps *rimu_hk_ops_9();

