// api.cu -- C ABI (include/rimu_b200.h) over the kernels in kernels.cuh.
// Host-side orchestration of one FCIQMC step = apply_operator! (pdworkingmemory.jl:297-309):
//   perform_spawns!  -> diag_count_kernel + scan_blocks_kernel + spawn_kernel
//   collect_local!   -> implicit (all deposits accumulate in the one working table)
//   synchronize_remote! -> count all-gather + grouped ncclSend/ncclRecv + insert_records_kernel
//   move_and_compress!  -> compact_kernel (also walkernumber_and_length)
#include "internal.cuh"

// ---------------------------------------------------------------- errors
static thread_local std::string g_err;
int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
size_t g_alloc_fail_bytes = 0;
std::string oom_note() {
    size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot);
    char b[160];
    snprintf(b, sizeof(b), " [requested %.1f MiB; device has %.1f GiB free of %.1f GiB]", g_alloc_fail_bytes / 1048576.0,
             fr / 1073741824.0, tot / 1073741824.0);
    return b;
}
extern "C" const char *rimu_last_error(void) { return g_err.c_str(); }
extern "C" int rimu_version(void) { return 100; }
extern "C" int rimu_sizeof_ham_desc(void) { return (int)sizeof(rimu_ham_desc); }
extern "C" int rimu_sizeof_step_params(void) { return (int)sizeof(rimu_step_params); }
extern "C" int rimu_sizeof_step_stats(void) { return (int)sizeof(rimu_step_stats); }
extern "C" int rimu_sizeof_shift_params(void) { return (int)sizeof(rimu_shift_params); }

// ---------------------------------------------------------------- NCCL (resolved lazily; same soname as torch's bundled copy)
NcclApi g_nccl;
int nccl_load() {
    if (g_nccl.lib) return 0;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(RIMU_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                   \
    *(void **)(&g_nccl.field) = dlsym(lib, name);                                          \
    if (!g_nccl.field) return fail(RIMU_ERR_NCCL, "libnccl lacks symbol %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather"); SYM(AllReduce, "ncclAllReduce"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = lib;
    return 0;
}
int enter_ctx(rimu_ctx *c) {
    if (!c) return fail(RIMU_ERR_INVALID, "null context");
    if (c->dead) return fail(RIMU_ERR_INVALID, "this context has been destroyed (a vector outlived it)");
    CUDA_TRY(enter_device(c->device));
    return 0;
}
// ---------------------------------------------------------------- context
static int table_fill(rimu_ctx *c, u64 slots) {
    table_fill_empty_kernel<<<grid_for((i64)slots, c->sm_count, 16), RIMU_TPB, 0, c->stream>>>(c->table, slots, c->W);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int rimu_ctx_create(int device, int words, uint64_t table_slots, rimu_ctx **out) {
    if (!out || (words != 1 && words != 2)) return fail(RIMU_ERR_INVALID, "rimu_ctx_create: words must be 1 or 2");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(RIMU_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(RIMU_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    CUDA_TRY(cudaSetDevice(device));
    rimu_ctx *c = new rimu_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device; c->W = words; c->nranks = 1; c->rank = 0;
    c->method = RIMU_ANNIHILATE_PARTITION; c->rec_per_parent = 1.5;
    if (const char *m = getenv("RIMU_B200_METHOD")) { if (!strcmp(m, "hash")) c->method = RIMU_ANNIHILATE_HASH; }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->merge_grid_cap = (u32)c->sm_count * 32;
    if (const char *g = getenv("RIMU_B200_MERGE_GRID")) { int v = atoi(g); if (v > 0) c->merge_grid_cap = (u32)v; }
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->table_slots = next_pow2(table_slots < 1024 ? 1024 : table_slots);
    CUDA_TRY(rimu_malloc(&c->table, c->table_slots * (words == 1 ? 16 : 32)));
    CUDA_TRY(rimu_malloc(&c->d_stats, sizeof(StatsDev)));
    CUDA_TRY(cudaMallocHost(&c->h_stats, sizeof(StatsDev)));
    CUDA_TRY(cudaMallocHost(&c->h_stats_local, sizeof(StatsDev)));
    CUDA_TRY(rimu_malloc(&c->d_reduce, 64 * sizeof(double)));
    for (int i = 0; i < 8; i++) CUDA_TRY(cudaEventCreate(&c->ev[i]));
    TRY(table_fill(c, c->table_slots));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out = c;
    return 0;
}

static void p2p_teardown(rimu_ctx *c);
struct SortScratch;
extern "C" int rimu_sort_scratch_bytes(void);
extern "C" int rimu_sort_annihilate_w1(cudaStream_t stream, SortScratch *s, const u64 *keys, const u64 *vals, long long n, int is_int,
                                       int key_bits, u64 *out_keys, u64 *out_vals, u64 out_cap, u64 *d_cursor);
extern "C" void rimu_sort_scratch_free(SortScratch *s);
static void ctx_free_struct(rimu_ctx *c) {
    cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}
extern "C" int rimu_ctx_destroy(rimu_ctx *c) {
    if (!c || c->dead) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    p2p_teardown(c);
    rimu_sort_scratch_free((SortScratch *)c->sort_scratch);
    cudaFree(c->d_ipc); cudaFreeHost(c->h_ipc);
    if (c->comm && g_nccl.lib) g_nccl.CommDestroy(c->comm);
    cudaFree(c->table); cudaFree(c->d_stats); cudaFreeHost(c->h_stats); cudaFreeHost(c->h_stats_local);
    cudaFree(c->local_off); cudaFree(c->block_tot); cudaFree(c->block_base);
    cudaFree(c->stage_keys); cudaFree(c->stage_vals);
    cudaFree(c->xch.keys); cudaFree(c->xch.vals); cudaFree(c->xch.counts);
    cudaFree(c->recv_keys); cudaFree(c->recv_vals); cudaFree(c->d_allcounts); cudaFreeHost(c->h_allcounts);
    cudaFree(c->d_reduce);
    cudaFree(c->part.rec); cudaFree(c->part.rcnt);
    cudaFree(c->lpart.rec); cudaFree(c->lpart.rcnt);
    cudaFree(c->heavy.items); cudaFree(c->bucket_tmp);
    cudaFree(c->spare_keys); cudaFree(c->spare_vals); cudaFree(c->spare_diag);
    for (int i = 0; i < 8; i++) cudaEventDestroy(c->ev[i]);
    cudaFree(c->d_red); cudaFreeHost(c->h_red); cudaFree(c->d_ord);
    cudaFree(c->d_ring); cudaFreeHost(c->h_ring); cudaFree(c->d_ctl); cudaFreeHost(c->h_ctl); cudaFree(c->d_shiftlog); cudaFreeHost(c->h_shiftlog);
    cudaFree(c->snap_keys); cudaFree(c->snap_vals); cudaFree(c->snap_diag); cudaFree(c->snap_seg_start); cudaFree(c->snap_seg_len);
    cudaFree(c->proj_keys); cudaFree(c->proj_vals); cudaFree(c->d_projlog); cudaFreeHost(c->h_projlog);
    cudaGetLastError(); // teardown is best effort (e.g. closing an IPC mapping whose exporter is already gone): never leave a stale error behind
    if (c->live_vecs > 0) { // vectors still point at this context: keep the struct and the stream until the last one goes
        c->dead = 1;
        return 0;
    }
    ctx_free_struct(c);
    return 0;
}
extern "C" int rimu_ctx_make_current(rimu_ctx *c) { if (!c) return fail(RIMU_ERR_INVALID, "null context"); TRY(enter_ctx(c)); return 0; }
extern "C" int rimu_ctx_synchronize(rimu_ctx *c) { CUDA_TRY(cudaStreamSynchronize(c->stream)); return 0; }
extern "C" int rimu_ctx_table_slots(rimu_ctx *c, uint64_t *out) { *out = c->table_slots; return 0; }
extern "C" int rimu_ctx_stream(rimu_ctx *c, void **s) { *s = (void *)c->stream; return 0; }
extern "C" int rimu_ctx_launch_count(rimu_ctx *c, uint64_t *out) { *out = c->launches; return 0; }
extern "C" int rimu_ctx_set_method(rimu_ctx *c, int method) {
    if (method != RIMU_ANNIHILATE_HASH && method != RIMU_ANNIHILATE_PARTITION)
        return fail(RIMU_ERR_INVALID, "step method must be RIMU_ANNIHILATE_HASH or RIMU_ANNIHILATE_PARTITION");
    c->method = method;
    return 0;
}
extern "C" int rimu_ctx_get_method(rimu_ctx *c, int *method) { *method = c->method; return 0; }
extern "C" int rimu_host_alloc(uint64_t bytes, void **out) { CUDA_TRY(cudaMallocHost(out, bytes ? bytes : 1)); return 0; }
extern "C" int rimu_host_free(void *p) { if (p) CUDA_TRY(cudaFreeHost(p)); return 0; }
extern "C" int rimu_ctx_resize_table(rimu_ctx *c, uint64_t table_slots) {
    TRY(enter_ctx(c));
    u64 slots = next_pow2(table_slots < 1024 ? 1024 : table_slots);
    if (slots == c->table_slots) return 0;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (slots > c->table_slots) { // growing: keep the old table if the new one cannot be allocated
        u64 *nt = nullptr;
        size_t fr = 0, tot = 0;
        CUDA_TRY(cudaMemGetInfo(&fr, &tot));
        const size_t need = slots * (c->W == 1 ? 16 : 32), have = c->table_slots * (c->W == 1 ? 16 : 32);
        if (need > fr + have) { g_alloc_fail_bytes = need; return fail(RIMU_ERR_CUDA, "working table of %llu slots does not fit%s", (unsigned long long)slots, oom_note().c_str()); }
        if (need > fr) { cudaFree(c->table); c->table = nullptr; c->table_slots = 0; } // (contents are scratch; only room matters)
        cudaError_t e = rimu_malloc(&nt, need);
        if (e != cudaSuccess) {
            if (!c->table) { // put a minimal table back so that the context stays usable
                CUDA_TRY(rimu_malloc(&c->table, 1024 * (c->W == 1 ? 16 : 32)));
                c->table_slots = 1024;
                TRY(table_fill(c, 1024));
            }
            return fail(RIMU_ERR_CUDA, "working table of %llu slots: %s%s", (unsigned long long)slots, cudaGetErrorString(e), oom_note().c_str());
        }
        cudaFree(c->table);
        c->table = nt;
    } else {
        cudaFree(c->table);
        c->table = nullptr; c->table_slots = 0;
        CUDA_TRY(rimu_malloc(&c->table, slots * (c->W == 1 ? 16 : 32)));
    }
    c->table_slots = slots;
    TRY(table_fill(c, slots));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int ensure_scratch(rimu_ctx *c, u64 parents) {
    if (parents <= c->scratch_parents) return 0;
    u64 cap = parents + parents / 4 + 1024;
    cudaFree(c->local_off); cudaFree(c->block_tot); cudaFree(c->block_base);
    c->local_off = c->block_tot = c->block_base = nullptr; c->scratch_parents = 0;
    u64 nblk = (cap + RIMU_TPB - 1) / RIMU_TPB;
    CUDA_TRY(rimu_malloc(&c->local_off, cap * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&c->block_tot, nblk * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&c->block_base, (nblk + 1) * sizeof(u64)));
    c->scratch_parents = cap;
    return 0;
}
static int ensure_stage(rimu_ctx *c, u64 n) {
    if (n <= c->stage_cap) return 0;
    u64 cap = n + n / 4 + 1024;
    cudaFree(c->stage_keys); cudaFree(c->stage_vals);
    c->stage_keys = nullptr; c->stage_vals = nullptr; c->stage_cap = 0;
    CUDA_TRY(rimu_malloc(&c->stage_keys, cap * c->W * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&c->stage_vals, cap * sizeof(u64)));
    c->stage_cap = cap;
    return 0;
}

// ---- partitioned-step working memory
int ensure_seg(rimu_vec *v, u32 nb) {
    if (nb <= v->seg_cap) return 0;
    u64 cap = (u64)nb + nb / 2 + 64;
    cudaFree(v->seg_start); cudaFree(v->seg_len);
    v->seg_start = nullptr; v->seg_len = nullptr; v->seg_cap = 0; v->nb = 0;
    CUDA_TRY(rimu_malloc(&v->seg_start, cap * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&v->seg_len, cap * sizeof(u32)));
    v->seg_cap = cap;
    return 0;
}
static int p2p_setup(rimu_ctx *c);
static void p2p_teardown(rimu_ctx *c);
extern "C" int rimu_comm_allreduce_f64(rimu_ctx *c, double *buf, int n);
// Size the record streams for nb buckets.  shared = the step streams of a multi-GPU context in direct mode: every rank
// calls this with the same nb (it is derived from all-reduced quantities), so (re)allocation and the exchange of the
// CUDA IPC handles are collective.  Otherwise: private streams with one sub-stream per bucket.
static int ensure_part_impl(rimu_ctx *c, PartDev &pt, u64 &nb_cap, u32 nb, bool shared, u32 nlane = 1, bool keep_lanes = false) {
    const u32 capi = part_cap_items(c->W);
    if (keep_lanes && pt.rec && pt.nlane > nlane) nlane = pt.nlane; // local operations reuse a 3-lane layout (lane 0 only)
    const u32 nranks = shared ? (u32)c->nranks : 1u;
    const u32 nsrc = nranks * nlane;
    // capacity of a sub-stream.  One rank: a whole bucket.  R ranks: a bucket's records arrive in R sub-streams of ~1/R each;
    // Poisson spread around the largest mean the bucket-count policy allows (~0.6 * capacity, all of it records in the
    // worst case) + 8 sigma.  Deliberately NOT a power of two: with 8 ranks the old 2 * cap / R = 512 records (8 KiB
    // stride, 1 KiB of it used) mapped every open stream onto one seventh of the L2 sets, and the spawn kernel slowed
    // down from 0.36 to 0.60 ms as 140 000 half-written lines fought over them (profiles/r2_multi_gpu.md).
    u32 rcap = capi;
    if (nranks > 1) {
        const double mean = 0.6 * capi / nranks;
        rcap = (u32)(mean + 8.0 * sqrt(mean) + 16.0);
        rcap = (rcap + 15u) / 16u * 16u + 16u; // multiple of 16 records, odd multiple of 256 bytes for W = 1
        if ((rcap / 16u) % 2u == 0) rcap += 16u;
    }
    if (pt.nlane != nlane && pt.rec) nb_cap = 0; // the sub-stream layout changes: reallocate (collective in shared mode)
    pt.nsrc = nsrc; pt.nlane = nlane; pt.me = shared ? (u32)c->rank * nlane : 0u; pt.rcap = rcap; pt.direct = shared ? 1 : 0;
    if (nb <= nb_cap) { pt.nb = nb; return 0; }
    // (the sub-stream stride depends on nb, not on the allocated bucket count, so a smaller nb reuses the buffers as they are)
    const u64 cap = (u64)nb + nb / 4 + 16;
    const size_t rw = c->W == 1 ? 2 : 4;
    if (shared) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        double zero = 0.0;
        TRY(rimu_comm_allreduce_f64(c, &zero, 1)); // nobody is still storing into a peer's old streams
        p2p_teardown(c);
    }
    cudaFree(pt.rec); cudaFree(pt.rcnt);
    pt.rec = nullptr; pt.rcnt = nullptr; nb_cap = 0;
    CUDA_TRY(rimu_malloc(&pt.rec, cap * nsrc * rcap * rw * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&pt.rcnt, (size_t)nsrc * cap * sizeof(u32)));
    CUDA_TRY(cudaMemsetAsync(pt.rcnt, 0, (size_t)nsrc * cap * sizeof(u32), c->stream));
    if (&pt == &c->part) c->rcnt_clean = 0; // (conservative: the step clears what it uses)
    nb_cap = cap; pt.nb = nb;
    if (shared) TRY(p2p_setup(c));
    return 0;
}
int ensure_part(rimu_ctx *c, u32 nb, u32 nlane) { return ensure_part_impl(c, c->part, c->part_nb_cap, nb, c->nranks > 1 && c->direct, nlane); }
// streams for local record->vector operations
static PartDev &local_part(rimu_ctx *c) { return (c->nranks > 1 && c->direct) ? c->lpart : c->part; }
static u64 &local_part_cap(rimu_ctx *c) { return (c->nranks > 1 && c->direct) ? c->lpart_nb_cap : c->part_nb_cap; }
static int ensure_local_part(rimu_ctx *c, u32 nb) { return ensure_part_impl(c, local_part(c), local_part_cap(c), nb, false, 1, true); }
int ensure_heavy(rimu_ctx *c, u64 parents) {
    c->heavy.packed = (u64 *)&c->d_stats->heavy_packed; // cleared together with the statistics block
    if (parents <= c->heavy.cap) return 0;
    u64 cap = parents + parents / 4 + 1024;
    cudaFree(c->heavy.items); c->heavy.items = nullptr; c->heavy.cap = 0;
    CUDA_TRY(rimu_malloc(&c->heavy.items, cap * sizeof(HeavyItem)));
    c->heavy.cap = cap;
    return 0;
}
static int ensure_bucket_tmp(rimu_ctx *c, u32 nb) {
    if (nb <= c->bucket_tmp_cap) return 0;
    u64 cap = (u64)nb + nb / 2 + 64;
    cudaFree(c->bucket_tmp); c->bucket_tmp = nullptr; c->bucket_tmp_cap = 0;
    CUDA_TRY(rimu_malloc(&c->bucket_tmp, 2 * cap * sizeof(u32)));
    c->bucket_tmp_cap = cap;
    return 0;
}

// ---------------------------------------------------------------- communicator
extern "C" int rimu_comm_unique_id(void *id128) {
    TRY(nccl_load());
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return 0;
}
// ---- peer-direct exchange: map every peer's receive buffers into this process (CUDA IPC over NVLink)
static void p2p_teardown(rimu_ctx *c) {
    for (int k = 0; k < 2; k++)
        for (int r = 0; r < RIMU_MAX_RANKS; r++)
            if (c->peer_open[k][r]) { cudaIpcCloseMemHandle(c->peer_open[k][r]); c->peer_open[k][r] = nullptr; }
    cudaGetLastError();
    memset(c->part.peer_rec, 0, sizeof(c->part.peer_rec));
    memset(c->part.peer_rcnt, 0, sizeof(c->part.peer_rcnt));
}
// all-gather the CUDA IPC handles of two buffers and map every peer's copy; returns 0 and sets *ok
static int ipc_map_all(rimu_ctx *c, void *buf0, void *buf1, void *peers0[], void *peers1[], int *ok) {
    const int R = c->nranks, me = c->rank;
    const size_t HS = sizeof(cudaIpcMemHandle_t);
    double bad = 0.0;
    if (!c->d_ipc) {
        CUDA_TRY(rimu_malloc(&c->d_ipc, (size_t)RIMU_MAX_RANKS * 2 * HS));
        CUDA_TRY(cudaMallocHost(&c->h_ipc, (size_t)RIMU_MAX_RANKS * 2 * HS));
    }
    cudaIpcMemHandle_t mine[2];
    memset(mine, 0, sizeof(mine));
    if (cudaIpcGetMemHandle(&mine[0], buf0) != cudaSuccess || cudaIpcGetMemHandle(&mine[1], buf1) != cudaSuccess) {
        cudaGetLastError();
        bad = 1.0;
    }
    TRY(rimu_comm_allreduce_f64(c, &bad, 1));
    if (bad > 0.0) { *ok = 0; return 0; }
    CUDA_TRY(cudaMemcpyAsync(c->d_ipc + (size_t)me * 2 * HS, mine, 2 * HS, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(g_nccl.AllGather(c->d_ipc + (size_t)me * 2 * HS, c->d_ipc, 2 * HS, 0 /* ncclInt8 */, c->comm, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->h_ipc, c->d_ipc, (size_t)R * 2 * HS, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < R && bad == 0.0; r++) {
        if (r == me) { peers0[r] = buf0; peers1[r] = buf1; continue; }
        for (int k = 0; k < 2; k++) {
            cudaIpcMemHandle_t hnd;
            memcpy(&hnd, c->h_ipc + ((size_t)r * 2 + k) * HS, HS);
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); bad = 1.0; break; }
            c->peer_open[k][r] = ptr;
            if (k == 0) peers0[r] = ptr; else peers1[r] = ptr;
        }
    }
    TRY(rimu_comm_allreduce_f64(c, &bad, 1));
    *ok = bad == 0.0;
    return 0;
}
// direct mode: map every rank's step streams (records + sub-stream fills).  Collective.
static int p2p_setup(rimu_ctx *c) {
    p2p_teardown(c);
    int ok = 0;
    void *pr[RIMU_MAX_RANKS] = {}, *pc[RIMU_MAX_RANKS] = {};
    TRY(ipc_map_all(c, c->part.rec, c->part.rcnt, pr, pc, &ok));
    if (!ok) { p2p_teardown(c); return fail(RIMU_ERR_CUDA, "CUDA IPC mapping of the peers' record streams failed (set RIMU_B200_P2P=0 to use NCCL send/recv)"); }
    for (int r = 0; r < c->nranks; r++) { c->part.peer_rec[r] = (const u64 *)pr[r]; c->part.peer_rcnt[r] = (const u32 *)pc[r]; }
    return 0;
}
// can the ranks map each other's memory at all?  (decides direct vs staged exchange once, at communicator set-up)
static int p2p_probe(rimu_ctx *c) {
    const char *env = getenv("RIMU_B200_P2P");
    double off = (env && !strcmp(env, "0")) ? 1.0 : 0.0;
    TRY(rimu_comm_allreduce_f64(c, &off, 1));
    c->direct = 0;
    if (off > 0.0) return 0;
    u64 *probe0 = nullptr, *probe1 = nullptr;
    CUDA_TRY(rimu_malloc(&probe0, 4096));
    CUDA_TRY(rimu_malloc(&probe1, 4096));
    int ok = 0;
    void *pr[RIMU_MAX_RANKS] = {}, *pc[RIMU_MAX_RANKS] = {};
    int rc = ipc_map_all(c, probe0, probe1, pr, pc, &ok);
    p2p_teardown(c);
    { double zero = 0.0; int rc2 = rimu_comm_allreduce_f64(c, &zero, 1); if (!rc) rc = rc2; } // every mapping is closed before the probes are freed
    cudaFree(probe0); cudaFree(probe1);
    if (rc) return rc;
    c->direct = ok;
    return 0;
}
extern "C" int rimu_comm_p2p(rimu_ctx *c, int *enabled) { *enabled = c->direct; return 0; }
extern "C" int rimu_comm_detach(rimu_ctx *c) {
    if (!c || c->dead || c->nranks == 1 || !c->comm || c->detached) return 0;
    TRY(enter_ctx(c));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double zero = 0.0;
    TRY(rimu_comm_allreduce_f64(c, &zero, 1)); // every rank's kernels that read peer streams have finished
    p2p_teardown(c);                           // close this rank's imports
    TRY(rimu_comm_allreduce_f64(c, &zero, 1)); // every import is closed: exporters may free
    cudaFree(c->part.rec); cudaFree(c->part.rcnt);
    c->part.rec = nullptr; c->part.rcnt = nullptr; c->part_nb_cap = 0;
    c->detached = 1;
    return 0;
}

// staging buffers of the NCCL send/recv exchange (staged mode, and the table method in any mode); allocated on first use
static int ensure_xch(rimu_ctx *c) {
    if (c->xch.keys) return 0;
    const u64 per_peer = c->xch_want < 1024 ? 1024 : c->xch_want;
    CUDA_TRY(rimu_malloc(&c->xch.keys, (u64)c->nranks * per_peer * c->W * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&c->xch.vals, (u64)c->nranks * per_peer * sizeof(u64)));
    c->recv_cap = (u64)c->nranks * per_peer;
    CUDA_TRY(rimu_malloc(&c->recv_keys, c->recv_cap * c->W * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&c->recv_vals, c->recv_cap * sizeof(u64)));
    c->xch.cap = per_peer;
    return 0;
}
extern "C" int rimu_comm_init(rimu_ctx *c, const void *id128, int rank, int nranks, uint64_t per_peer) {
    if (nranks < 1 || nranks > RIMU_MAX_RANKS || rank < 0 || rank >= nranks) return fail(RIMU_ERR_INVALID, "bad rank/nranks");
    if (c->comm) return fail(RIMU_ERR_INVALID, "communicator already attached");
    TRY(enter_ctx(c));
    c->rank = rank; c->nranks = nranks;
    if (nranks == 1) return 0;
    TRY(nccl_load());
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NCCL_TRY(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    if (per_peer < 1024) per_peer = 1024;
    CUDA_TRY(rimu_malloc(&c->xch.counts, RIMU_MAX_RANKS * sizeof(u64)));
    CUDA_TRY(cudaMemset(c->xch.counts, 0, RIMU_MAX_RANKS * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&c->d_allcounts, (u64)nranks * nranks * sizeof(u64)));
    CUDA_TRY(cudaMallocHost(&c->h_allcounts, (u64)nranks * nranks * sizeof(u64)));
    memset(c->h_allcounts, 0, (size_t)nranks * nranks * sizeof(u64));
    c->xch_want = per_peer;
    TRY(p2p_probe(c));
    if (!c->direct) TRY(ensure_xch(c)); // staged exchange: per-peer send segments + receive buffer for NCCL send/recv
    return 0;
}
// grow the per-peer exchange buffers (every rank must call it with the same size; contents are scratch)
extern "C" int rimu_comm_reserve(rimu_ctx *c, uint64_t per_peer) {
    if (c->nranks == 1 || per_peer <= c->xch.cap) return 0;
    TRY(enter_ctx(c));
    if (!c->xch.keys) { if (per_peer > c->xch_want) c->xch_want = per_peer; return 0; } // not allocated yet (direct mode)
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(c->xch.keys); cudaFree(c->xch.vals); cudaFree(c->recv_keys); cudaFree(c->recv_vals);
    c->xch.keys = c->xch.vals = c->recv_keys = c->recv_vals = nullptr; c->xch.cap = 0; c->recv_cap = 0;
    c->xch_want = per_peer;
    return ensure_xch(c);
}
extern "C" int rimu_comm_capacity(rimu_ctx *c, uint64_t *per_peer, uint64_t *needed) {
    *per_peer = c->xch.cap; *needed = c->xch_worst;
    return 0;
}
extern "C" int rimu_comm_rank(rimu_ctx *c, int *rank, int *nranks) { *rank = c->rank; *nranks = c->nranks; return 0; }
extern "C" int rimu_comm_allreduce_f64(rimu_ctx *c, double *buf, int n) {
    if (c->nranks == 1) return 0;
    if (n > 64) return fail(RIMU_ERR_INVALID, "allreduce of at most 64 doubles");
    TRY(enter_ctx(c));
    CUDA_TRY(cudaMemcpyAsync(c->d_reduce, buf, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(g_nccl.AllReduce(c->d_reduce, c->d_reduce, n, ncclFloat64, ncclSum, c->comm, c->stream));
    CUDA_TRY(cudaMemcpyAsync(buf, c->d_reduce, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" uint64_t rimu_addr_hash(const uint64_t *key, int words) {
    return words == 1 ? addr_hash<1>((const u64 *)key) : addr_hash<2>((const u64 *)key);
}
extern "C" int rimu_addr_owner(const uint64_t *key, int words, int nranks) {
    return addr_owner(rimu_addr_hash(key, words), nranks);
}
extern "C" void rimu_step_key(uint64_t seed, uint64_t step, uint32_t key_out[2]) {
    u64 k = splitmix64(seed ^ splitmix64(step));
    key_out[0] = (u32)k; key_out[1] = (u32)(k >> 32);
}
extern "C" void rimu_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr, key[0], key[1], out);
}

// ---------------------------------------------------------------- Hamiltonian
extern "C" int rimu_ham_create(const rimu_ham_desc *d, rimu_ham **out) {
    if (!d || !out) return fail(RIMU_ERR_INVALID, "null argument");
    HamHostImage img; // validation + scalars + constant tables: pure host code (ham_host.h), also what the CPU tests check
    if (ham_build_host(d, &img)) return fail(RIMU_ERR_INVALID, "%s", img.error.c_str());
    rimu_ham *h = new rimu_ham();
    memset(h, 0, sizeof(*h));
    h->desc = *d; h->hk = img.hk; h->W = img.W;
    static u64 next_uid = 1;
    h->uid = next_uid++;
    h->dev = img.dev;
    cudaError_t e = cudaGetDevice(&h->device);
    if (e == cudaSuccess) e = rimu_malloc(&h->d_tables, img.tables.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_tables, img.tables.data(), img.tables.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !img.nbr.empty()) {
        e = rimu_malloc(&h->d_nbr, img.nbr.size());
        if (e == cudaSuccess) e = cudaMemcpy(h->d_nbr, img.nbr.data(), img.nbr.size(), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        cudaFree(h->d_tables); cudaFree(h->d_nbr);
        delete h;
        cudaGetLastError();
        return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? RIMU_ERR_NO_DEVICE : RIMU_ERR_CUDA,
                    "rimu_ham_create: %s", cudaGetErrorString(e));
    }
    ham_set_tables(&h->dev, h->d_tables, h->d_nbr);
    *out = h;
    return 0;
}
extern "C" int rimu_ham_destroy(rimu_ham *h) {
    if (!h) return 0;
    cudaFree(h->d_tables); cudaFree(h->d_nbr);
    delete h;
    return 0;
}
extern "C" int rimu_ham_words(const rimu_ham *h) { return h->W; }

// per-kind entry points (step_hk.cu)
static const HkOps *hk_ops(const rimu_ham *h) {
#ifdef RIMU_TUNE_ONLY_MOM1D // kernel-tuning builds (scratch/): one kind, seconds to compile; never shipped
    if (h->hk == HK_MOM1D_BOSE_PLAIN && h->W == 1) return rimu_hk_ops_9();
    fail(RIMU_ERR_INVALID, "tuning build: only HubbardMom1D/BoseFS one-word addresses are compiled in");
    return nullptr;
#else
    switch (h->hk) {
    case HK_REAL1D_BOSE: return rimu_hk_ops_0();
    case HK_MOM1D_BOSE: return rimu_hk_ops_1();
    case HK_MOM1D_F2C: return rimu_hk_ops_2();
    case HK_RS_BOSE: return rimu_hk_ops_3();
    case HK_RS_FERMI: return rimu_hk_ops_4();
    case HK_RS_F2C: return rimu_hk_ops_5();
    case HK_TC_F2C: return rimu_hk_ops_6();
    case HK_RS_COMP: return rimu_hk_ops_7();
    case HK_REAL1D_BOSE_PLAIN: return rimu_hk_ops_8();
    case HK_MOM1D_BOSE_PLAIN: return rimu_hk_ops_9();
    }
    fail(RIMU_ERR_INVALID, "unknown Hamiltonian kind");
    return nullptr;
#endif
}

static int check_ctx_ham(rimu_ctx *c, const rimu_ham *h) {
    if (!c || !h) return fail(RIMU_ERR_INVALID, "null handle");
    if (c->dead) return fail(RIMU_ERR_INVALID, "the context of these vectors has been destroyed");
    if (c->W != h->W) return fail(RIMU_ERR_INVALID, "context built for %d-word addresses, Hamiltonian needs %d", c->W, h->W);
    if (c->device != h->device) return fail(RIMU_ERR_INVALID, "Hamiltonian tables live on device %d, context on %d", h->device, c->device);
    TRY(enter_ctx(c));
    return 0;
}

extern "C" int rimu_ham_diagonal(rimu_ctx *c, const rimu_ham *h, const uint64_t *keys, int64_t n, double *out) {
    TRY(check_ctx_ham(c, h));
    if (n <= 0) return 0;
    TRY(ensure_stage(c, (u64)n));
    CUDA_TRY(cudaMemcpyAsync(c->stage_keys, keys, n * c->W * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    double *d_out = (double *)c->stage_vals;
    const HkOps *ops = hk_ops(h);
    if (!ops) return RIMU_ERR_INVALID;
    TRY(ops->diag(c, h, c->stage_keys, n, d_out, nullptr));
    CUDA_TRY(cudaMemcpyAsync(out, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int rimu_ham_num_offdiagonals(rimu_ctx *c, const rimu_ham *h, const uint64_t *keys, int64_t n, int64_t *out) {
    TRY(check_ctx_ham(c, h));
    if (n <= 0) return 0;
    TRY(ensure_stage(c, (u64)n));
    CUDA_TRY(cudaMemcpyAsync(c->stage_keys, keys, n * c->W * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    i64 *d_out = (i64 *)c->stage_vals;
    const HkOps *ops = hk_ops(h);
    if (!ops) return RIMU_ERR_INVALID;
    TRY(ops->diag(c, h, c->stage_keys, n, nullptr, d_out));
    CUDA_TRY(cudaMemcpyAsync(out, d_out, n * sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int rimu_ham_offdiagonals(rimu_ctx *c, const rimu_ham *h, const uint64_t *key, int64_t first, int64_t count,
                                     uint64_t *keys_out, double *vals_out) {
    TRY(check_ctx_ham(c, h));
    if (count <= 0) return 0;
    if (first < 1) return fail(RIMU_ERR_INVALID, "off-diagonal indices are 1-based");
    TRY(ensure_stage(c, (u64)count + 1));
    u64 *d_key = c->stage_keys + (u64)count * c->W;
    CUDA_TRY(cudaMemcpyAsync(d_key, key, c->W * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    double *d_vals = (double *)c->stage_vals;
    const HkOps *ops = hk_ops(h);
    if (!ops) return RIMU_ERR_INVALID;
    TRY(ops->offdiag(c, h, d_key, first - 1, count, c->stage_keys, d_vals));
    CUDA_TRY(cudaMemcpyAsync(keys_out, c->stage_keys, count * c->W * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(vals_out, d_vals, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------- vectors
extern "C" int rimu_vec_create(rimu_ctx *c, int val_type, uint64_t capacity, rimu_vec **out) {
    if (!c || !out) return fail(RIMU_ERR_INVALID, "null argument");
    if (val_type != RIMU_VAL_F64 && val_type != RIMU_VAL_I64) return fail(RIMU_ERR_INVALID, "bad value type");
    TRY(enter_ctx(c));
    rimu_vec *v = new rimu_vec();
    v->ctx = c; v->vt = val_type; v->n = 0; v->cap = capacity < 256 ? 256 : capacity;
    v->keys = nullptr; v->vals = nullptr;
    v->nb = 0; v->seg_cap = 0; v->seg_start = nullptr; v->seg_len = nullptr;
    v->diag = nullptr; v->diag_cap = 0; v->diag_uid = 0;
    CUDA_TRY(rimu_malloc(&v->keys, v->cap * c->W * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&v->vals, v->cap * sizeof(u64)));
    c->live_vecs++;
    v->uid = ++c->next_vec_uid; // (0 = "no vector": last_dst_uid starts at 0)
    *out = v;
    return 0;
}
extern "C" int rimu_vec_destroy(rimu_vec *v) {
    if (!v) return 0;
    rimu_ctx *c = v->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(v->keys); cudaFree(v->vals); cudaFree(v->seg_start); cudaFree(v->seg_len); cudaFree(v->diag);
    cudaGetLastError();
    delete v;
    if (--c->live_vecs <= 0 && c->dead) ctx_free_struct(c);
    return 0;
}
extern "C" int rimu_vec_reserve(rimu_vec *v, uint64_t capacity) {
    if (capacity <= v->cap) return 0;
    rimu_ctx *c = v->ctx;
    TRY(enter_ctx(c));
    u64 *nk = nullptr; void *nv = nullptr;
    CUDA_TRY(rimu_malloc(&nk, capacity * c->W * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&nv, capacity * sizeof(u64)));
    if (v->n > 0) {
        CUDA_TRY(cudaMemcpyAsync(nk, v->keys, v->n * c->W * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(nv, v->vals, v->n * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
    }
    double *nd = nullptr;
    if (v->diag) {
        CUDA_TRY(rimu_malloc(&nd, capacity * sizeof(double)));
        if (v->n > 0 && v->diag_uid) CUDA_TRY(cudaMemcpyAsync(nd, v->diag, v->n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(v->keys); cudaFree(v->vals); cudaFree(v->diag);
    v->keys = nk; v->vals = nv; v->cap = capacity;
    v->diag = nd; v->diag_cap = nd ? capacity : 0;
    return 0;
}
int ensure_diag(rimu_vec *v) {
    if (v->diag && v->diag_cap >= v->cap) return 0;
    cudaFree(v->diag); v->diag = nullptr; v->diag_cap = 0; v->diag_uid = 0;
    CUDA_TRY(rimu_malloc(&v->diag, v->cap * sizeof(double)));
    v->diag_cap = v->cap;
    return 0;
}
extern "C" int rimu_vec_clear(rimu_vec *v) { v->version++; v->n = 0; v->nb = 0; v->diag_uid = 0; return 0; }
extern "C" int rimu_vec_length(rimu_vec *v, int64_t *out) { *out = v->n; return 0; }
extern "C" int rimu_vec_capacity(rimu_vec *v, uint64_t *out) { *out = v->cap; return 0; }

static StepDev null_step(rimu_ctx *c) {
    StepDev p;
    memset(&p, 0, sizeof(p));
    p.rank = c->rank; p.nranks = c->nranks;
    return p;
}

static u64 pick_slots(rimu_ctx *c, u64 expected_entries) {
    u64 s = next_pow2(expected_entries * 2 + 1024);
    if (s > c->table_slots) s = c->table_slots;
    return s;
}

// build dst from device-resident records (sum by key, drop zeros) via the working table.
// Retries with more slots on table overflow; grows dst when needed.
static int records_to_vec(rimu_ctx *c, rimu_vec *dst, const u64 *d_keys, const void *d_vals, i64 n,
                          const u64 *d_keys2, const void *d_vals2, i64 n2, double a1, double a2, int use_scale) {
    TRY(enter_ctx(c));
    dst->nb = 0; dst->diag_uid = 0; // the global-table path produces an unsegmented vector
    u64 slots = pick_slots(c, (u64)(n + n2));
    const bool aliased = (const u64 *)dst->keys == d_keys || (const u64 *)dst->keys == d_keys2;
    for (;;) {
        CUDA_TRY(cudaMemsetAsync(c->d_stats, 0, sizeof(StatsDev), c->stream));
        TableDev tab{c->table, slots - 1};
        TRY(dispatch_wv(c->W, dst->vt, [&](auto tag, auto vtag) {
            typedef decltype(vtag) VT;
            constexpr int W = decltype(tag)::w;
            if (n > 0)
                insert_records_kernel<W, VT><<<grid_for(n, c->sm_count), RIMU_TPB, 0, c->stream>>>(
                    d_keys, (const VT *)d_vals, n, a1, use_scale, c->rank, c->nranks, tab, c->d_stats);
            if (n2 > 0)
                insert_records_kernel<W, VT><<<grid_for(n2, c->sm_count), RIMU_TPB, 0, c->stream>>>(
                    d_keys2, (const VT *)d_vals2, n2, a2, use_scale, c->rank, c->nranks, tab, c->d_stats);
            return 0;
        }));
        CUDA_TRY(cudaGetLastError());
        // the table must hold EVERYTHING before it is drained: dst may alias an input, and a partial drain would
        // overwrite records that the retry still has to read
        CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(StatsDev), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (c->h_stats->overflow_table) {
            TRY(table_fill(c, slots)); // forget the partial contents
            if (slots >= c->table_slots) {
                if (c->table_slots >= (1ull << 34)) return fail(RIMU_ERR_TABLE_FULL, "working table (%llu slots) too small for %lld records", (unsigned long long)c->table_slots, (long long)(n + n2));
                TRY(rimu_ctx_resize_table(c, c->table_slots * 4));
            }
            slots = slots * 4 > c->table_slots ? c->table_slots : slots * 4;
            continue;
        }
        TRY(dispatch_wv(c->W, dst->vt, [&](auto tag, auto vtag) {
            typedef decltype(vtag) VT;
            constexpr int W = decltype(tag)::w;
            return compact_into<W, VT>(c, dst, slots, null_step(c));
        }));
        CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(StatsDev), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (c->h_stats->out_count > dst->cap) {
            // the table has been drained; grow and redo (inputs are untouched unless dst aliases them)
            if (aliased)
                return fail(RIMU_ERR_VECTOR_FULL, "destination (aliasing an input) too small: need %llu", (unsigned long long)c->h_stats->out_count);
            dst->n = 0;
            TRY(rimu_vec_reserve(dst, c->h_stats->out_count + c->h_stats->out_count / 4));
            continue;
        }
        dst->n = (i64)c->h_stats->out_count;
        return 0;
    }
}

// ---- the same through the partitioned working memory: records -> bucket streams -> shared-memory merge (MODE 1)
static int records_to_vec_part(rimu_ctx *c, rimu_vec *dst, const u64 *d_keys, const void *d_vals, i64 n,
                               const u64 *d_keys2, const void *d_vals2, i64 n2, double a1, double a2, int use_scale) {
    TRY(enter_ctx(c));
    dst->diag_uid = 0; dst->nb = 0;
    const double cap = (double)part_cap_items(c->W);
    u32 nb = (u32)ceil(((double)(n + n2) * 1.1 + 256.0) / (0.65 * cap));
    if (nb < 1) nb = 1;
    for (int attempt = 0;; attempt++) {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const double recb = c->W == 1 ? 16.0 : 32.0;
        if (nb > local_part_cap(c) && (double)nb * 1.3 * cap * recb > 0.8 * ((double)free_b + (double)local_part_cap(c) * cap * recb))
            return records_to_vec(c, dst, d_keys, d_vals, n, d_keys2, d_vals2, n2, a1, a2, use_scale);
        TRY(ensure_local_part(c, nb));
        PartDev &lp = local_part(c);
        if (&lp == &c->part) c->rcnt_clean = 0; // the step's streams are borrowed: the step clears them again itself
        TRY(ensure_seg(dst, nb));
        CUDA_TRY(cudaMemsetAsync(c->d_stats, 0, sizeof(StatsDev), c->stream));
        CUDA_TRY(cudaMemsetAsync(lp.rcnt, 0, (size_t)lp.nsrc * nb * sizeof(u32), c->stream));
        TRY(dispatch_wv(c->W, dst->vt, [&](auto tag, auto vtag) {
            typedef decltype(vtag) VT;
            constexpr int W = decltype(tag)::w;
            static bool attr_set = false;
            if (!attr_set) {
                CUDA_TRY(cudaFuncSetAttribute(merge_kernel<0, W, VT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem_bytes(W)));
                attr_set = true;
            }
            if (n > 0)
                append_records_kernel<W, VT><<<grid_for(n, c->sm_count, 16), RIMU_TPB, 0, c->stream>>>(
                    d_keys, (const VT *)d_vals, n, a1, use_scale, c->rank, c->nranks, lp, 0u, c->d_stats);
            if (n2 > 0)
                append_records_kernel<W, VT><<<grid_for(n2, c->sm_count, 16), RIMU_TPB, 0, c->stream>>>(
                    d_keys2, (const VT *)d_vals2, n2, a2, use_scale, c->rank, c->nranks, lp, 0u, c->d_stats);
            HamDev hd; memset(&hd, 0, sizeof(hd));
            SegSrc ss{nullptr, nullptr, nullptr, nullptr, nullptr};
            SegDst sd{dst->keys, (u64 *)dst->vals, dst->seg_start, dst->seg_len, dst->cap, nullptr};
            const int mgrid = (int)(nb < c->merge_grid_cap ? nb : c->merge_grid_cap);
            merge_kernel<0, W, VT, 1><<<mgrid, PART_NT, part_smem_bytes(W), c->stream>>>(hd, null_step(c), ss, 1.0, lp, sd, c->d_stats);
            return 0;
        }));
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(StatsDev), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const StatsDev &l = *c->h_stats;
        if (l.overflow_table) {
            if (attempt > 10 || l.max_fill > (u64)(64 * cap)) // one address too hot for a bucket: the table handles it
                return records_to_vec(c, dst, d_keys, d_vals, n, d_keys2, d_vals2, n2, a1, a2, use_scale);
            double need = ceil((double)l.records * 1.15 / (0.6 * cap));
            nb = need > (double)nb * 1.5 ? (u32)need : nb * 2 + 1;
            continue;
        }
        if (l.out_count > dst->cap) {
            if ((const u64 *)dst->keys == d_keys || (const u64 *)dst->keys == d_keys2)
                return fail(RIMU_ERR_VECTOR_FULL, "destination (aliasing an input) too small: need %llu", (unsigned long long)l.out_count);
            dst->n = 0;
            TRY(rimu_vec_reserve(dst, l.out_count + l.out_count / 4));
            continue;
        }
        dst->n = (i64)l.out_count;
        dst->nb = nb;
        return 0;
    }
}
static int records_to_vec_auto(rimu_ctx *c, rimu_vec *dst, const u64 *d_keys, const void *d_vals, i64 n,
                               const u64 *d_keys2, const void *d_vals2, i64 n2, double a1, double a2, int use_scale) {
    const bool aliased = (const u64 *)dst->keys == d_keys || (const u64 *)dst->keys == d_keys2;
    if (c->method == RIMU_ANNIHILATE_PARTITION && !aliased) // (the merge writes dst while it may still be read: aliasing goes through the table, which buffers everything first)
        return records_to_vec_part(c, dst, d_keys, d_vals, n, d_keys2, d_vals2, n2, a1, a2, use_scale);
    return records_to_vec(c, dst, d_keys, d_vals, n, d_keys2, d_vals2, n2, a1, a2, use_scale);
}

extern "C" int rimu_vec_upload(rimu_vec *v, const uint64_t *keys, const void *vals, int64_t n) {
    if (v) v->version++;
    rimu_ctx *c = v->ctx;
    TRY(enter_ctx(c));
    if (n <= 0) { v->n = 0; v->nb = 0; v->diag_uid = 0; return 0; }
    TRY(ensure_stage(c, (u64)n));
    CUDA_TRY(cudaMemcpyAsync(c->stage_keys, keys, n * c->W * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->stage_vals, vals, n * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    return records_to_vec_auto(c, v, c->stage_keys, c->stage_vals, n, nullptr, nullptr, 0, 1.0, 1.0, 0);
}
extern "C" int rimu_vec_assign(rimu_vec *v, const uint64_t *keys, const void *vals, int64_t n) {
    if (v) v->version++;
    rimu_ctx *c = v->ctx;
    TRY(enter_ctx(c));
    if (n < 0) return fail(RIMU_ERR_INVALID, "negative length");
    v->nb = 0; v->diag_uid = 0;
    if ((u64)n > v->cap) { v->n = 0; TRY(rimu_vec_reserve(v, (u64)n)); }
    if (n > 0) {
        CUDA_TRY(cudaMemcpyAsync(v->keys, keys, n * c->W * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(v->vals, vals, n * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    }
    v->n = n;
    return 0;
}
extern "C" int rimu_vec_download(rimu_vec *v, uint64_t *keys_out, void *vals_out, int64_t cap, int64_t *n_out) {
    rimu_ctx *c = v->ctx;
    TRY(enter_ctx(c));
    if (n_out) *n_out = v->n;
    if (cap < v->n) return fail(RIMU_ERR_VECTOR_FULL, "download buffer too small: need %lld", (long long)v->n);
    if (v->n > 0) {
        CUDA_TRY(cudaMemcpyAsync(keys_out, v->keys, v->n * c->W * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(vals_out, v->vals, v->n * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int rimu_vec_copy(rimu_vec *dst, rimu_vec *src) {
    if (dst) dst->version++;
    if (dst == src) return 0;
    rimu_ctx *c = src->ctx;
    if (dst->ctx != c) return fail(RIMU_ERR_INVALID, "copy between vectors of different contexts");
    TRY(enter_ctx(c));
    if ((u64)src->n > dst->cap) { dst->n = 0; TRY(rimu_vec_reserve(dst, (u64)src->n)); }
    if (src->n > 0) {
        CUDA_TRY(cudaMemcpyAsync(dst->keys, src->keys, src->n * c->W * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
        if (dst->vt == src->vt) {
            CUDA_TRY(cudaMemcpyAsync(dst->vals, src->vals, src->n * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
        } else { // eltype conversion Int64 <-> Float64 on the device
            CUDA_TRY(cudaMemsetAsync(&c->d_stats->overflow_vec, 0, sizeof(i64), c->stream));
            if (src->vt == RIMU_VAL_I64) convert_vals_kernel<i64, double><<<grid_for(src->n, c->sm_count), RIMU_TPB, 0, c->stream>>>((const i64 *)src->vals, (double *)dst->vals, src->n, &c->d_stats->overflow_vec);
            else convert_vals_kernel<double, i64><<<grid_for(src->n, c->sm_count), RIMU_TPB, 0, c->stream>>>((const double *)src->vals, (i64 *)dst->vals, src->n, &c->d_stats->overflow_vec);
            CUDA_TRY(cudaGetLastError());
            if (dst->vt == RIMU_VAL_I64) { // Float64 -> Int64 must be exact (the reference throws InexactError)
                CUDA_TRY(cudaMemcpyAsync(&c->h_stats->overflow_vec, &c->d_stats->overflow_vec, sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
                CUDA_TRY(cudaStreamSynchronize(c->stream));
                if (c->h_stats->overflow_vec) {
                    dst->n = 0; dst->nb = 0; dst->diag_uid = 0;
                    return fail(RIMU_ERR_INVALID, "InexactError: a Float64 vector with non-integral values cannot be copied into an Int64 vector");
                }
            }
        }
    }
    dst->n = src->n;
    dst->nb = 0; dst->diag_uid = 0;
    if (src->diag_uid && src->n > 0) { // keep the diagonal-element cache
        TRY(ensure_diag(dst));
        CUDA_TRY(cudaMemcpyAsync(dst->diag, src->diag, src->n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        dst->diag_uid = src->diag_uid;
    }
    if (src->nb) { // keep the bucket segmentation
        TRY(ensure_seg(dst, src->nb));
        CUDA_TRY(cudaMemcpyAsync(dst->seg_start, src->seg_start, src->nb * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(dst->seg_len, src->seg_len, src->nb * sizeof(u32), cudaMemcpyDeviceToDevice, c->stream));
        dst->nb = src->nb;
    }
    return 0;
}
extern "C" int rimu_vec_get(rimu_vec *v, const uint64_t *key, void *val_out) {
    rimu_ctx *c = v->ctx;
    TRY(enter_ctx(c));
    u64 *d_out = (u64 *)c->d_reduce;
    CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(u64), c->stream));
    if (v->n > 0) {
        TRY(dispatch_wv(c->W, v->vt, [&](auto tag, auto vtag) {
            typedef decltype(vtag) VT;
            find_key_kernel<decltype(tag)::w, VT><<<grid_for(v->n, c->sm_count), RIMU_TPB, 0, c->stream>>>(
                v->keys, (const VT *)v->vals, v->n, key[0], c->W == 2 ? key[1] : 0, (VT *)d_out);
            return 0;
        }));
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaMemcpyAsync(val_out, d_out, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int rimu_vec_norm(rimu_vec *v, int p, double *out) {
    rimu_ctx *c = v->ctx;
    TRY(enter_ctx(c));
    CUDA_TRY(cudaMemsetAsync(c->d_stats, 0, sizeof(StatsDev), c->stream));
    if (v->n > 0) {
        if (v->vt == RIMU_VAL_F64) norm_kernel<double><<<grid_for(v->n, c->sm_count), RIMU_TPB, 0, c->stream>>>((const double *)v->vals, v->n, c->d_stats);
        else norm_kernel<i64><<<grid_for(v->n, c->sm_count), RIMU_TPB, 0, c->stream>>>((const i64 *)v->vals, v->n, c->d_stats);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(StatsDev), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double r[2] = {c->h_stats->norm1, c->h_stats->norm2};
    if (c->nranks > 1) {
        TRY(rimu_comm_allreduce_f64(c, r, 2));
        if (p == 0) { // max over ranks through NCCL
            CUDA_TRY(cudaMemcpyAsync(c->d_reduce, &c->h_stats->norminf, sizeof(double), cudaMemcpyHostToDevice, c->stream));
            NCCL_TRY(g_nccl.AllReduce(c->d_reduce, c->d_reduce, 1, ncclFloat64, ncclMax, c->comm, c->stream));
            CUDA_TRY(cudaMemcpyAsync(&c->h_stats->norminf, c->d_reduce, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
    }
    if (p == 1) *out = r[0];
    else if (p == 2) *out = sqrt(r[1]);
    else if (p == 0) *out = c->h_stats->norminf;
    else return fail(RIMU_ERR_INVALID, "norm p must be 1, 2 or 0 (=inf)");
    return 0;
}
extern "C" int rimu_vec_scale(rimu_vec *v, double alpha) {
    if (v) v->version++;
    rimu_ctx *c = v->ctx;
    TRY(enter_ctx(c));
    if (alpha == 0.0) { v->n = 0; v->nb = 0; v->diag_uid = 0; return 0; } // zero values are never stored
    if (v->n > 0) {
        if (v->vt == RIMU_VAL_F64) scale_kernel<double><<<grid_for(v->n, c->sm_count), RIMU_TPB, 0, c->stream>>>((double *)v->vals, v->n, alpha, nullptr);
        else {
            // Int64 vectors: the scaled values must be integers (InexactError in the reference); the kernel leaves an
            // offending entry untouched and raises the flag, so a failed call does not corrupt the vector half-way only
            // when alpha is integral -- which is the one case that can succeed
            if (alpha != rint(alpha)) return fail(RIMU_ERR_INVALID, "InexactError: an Int64 vector can only be scaled by an integer");
            CUDA_TRY(cudaMemsetAsync(&c->d_stats->overflow_vec, 0, sizeof(i64), c->stream));
            scale_kernel<i64><<<grid_for(v->n, c->sm_count), RIMU_TPB, 0, c->stream>>>((i64 *)v->vals, v->n, alpha, &c->d_stats->overflow_vec);
            CUDA_TRY(cudaMemcpyAsync(&c->h_stats->overflow_vec, &c->d_stats->overflow_vec, sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            if (c->h_stats->overflow_vec) return fail(RIMU_ERR_INVALID, "InexactError: scaling overflows the exactly representable integer range");
        }
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}
extern "C" int rimu_vec_dot(rimu_vec *x, rimu_vec *y, double *out) {
    rimu_ctx *c = x->ctx;
    if (y->ctx != c || x->vt != y->vt) return fail(RIMU_ERR_INVALID, "dot of incompatible vectors");
    TRY(enter_ctx(c));
    rimu_vec *build = x->n <= y->n ? x : y, *probe = x->n <= y->n ? y : x; // table from the shorter one
    double r = 0.0;
    if (build->n > 0) {
        u64 slots = pick_slots(c, (u64)build->n);
        for (;;) {
            CUDA_TRY(cudaMemsetAsync(c->d_stats, 0, sizeof(StatsDev), c->stream));
            TableDev tab{c->table, slots - 1};
            TRY(dispatch_wv(c->W, x->vt, [&](auto tag, auto vtag) {
                typedef decltype(vtag) VT;
                constexpr int W = decltype(tag)::w;
                insert_records_kernel<W, VT><<<grid_for(build->n, c->sm_count), RIMU_TPB, 0, c->stream>>>(
                    build->keys, (const VT *)build->vals, build->n, 1.0, 0, 0, 1, tab, c->d_stats);
                lookup_dot_kernel<W, VT><<<grid_for(probe->n, c->sm_count), RIMU_TPB, 0, c->stream>>>(
                    probe->keys, (const VT *)probe->vals, probe->n, tab, c->d_stats);
                return 0;
            }));
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(StatsDev), cudaMemcpyDeviceToHost, c->stream));
            // re-empty the table (only the active prefix was touched)
            TRY(table_fill(c, slots));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            if (c->h_stats->overflow_table) {
                if (slots >= c->table_slots) return fail(RIMU_ERR_TABLE_FULL, "working table too small for dot");
                slots = slots * 4 > c->table_slots ? c->table_slots : slots * 4;
                continue;
            }
            r = c->h_stats->dot;
            break;
        }
    }
    if (c->nranks > 1) TRY(rimu_comm_allreduce_f64(c, &r, 1));
    *out = r;
    return 0;
}
extern "C" int rimu_vec_dot_sparse(rimu_vec *v, const uint64_t *keys, const double *values, int64_t n, double *out) {
    rimu_ctx *c = v->ctx;
    if (!out || n < 0 || (n > 0 && (!keys || !values))) return fail(RIMU_ERR_INVALID, "dot_sparse: bad arguments");
    TRY(enter_ctx(c));
    double r = 0.0;
    if (n > 0 && v->n > 0) {
        TRY(ensure_stage(c, (u64)n));
        CUDA_TRY(cudaMemcpyAsync(c->stage_keys, keys, n * c->W * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->stage_vals, values, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemsetAsync(&c->d_stats->dot, 0, sizeof(double), c->stream));
        const int grid = (int)(n < (i64)c->sm_count * 8 ? n : (i64)c->sm_count * 8);
        TRY(dispatch_wv(c->W, v->vt, [&](auto tag, auto vtag) {
            typedef decltype(vtag) VT;
            dot_sparse_kernel<decltype(tag)::w, VT><<<grid, RIMU_TPB, 0, c->stream>>>(
                c->stage_keys, (const double *)c->stage_vals, n, v->keys, (const VT *)v->vals, v->n,
                v->nb ? v->seg_start : nullptr, v->nb ? v->seg_len : nullptr, v->nb, c->rank, c->nranks, &c->d_stats->dot);
            return 0;
        }));
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&c->h_stats->dot, &c->d_stats->dot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        r = c->h_stats->dot;
    }
    if (c->nranks > 1) TRY(rimu_comm_allreduce_f64(c, &r, 1)); // collective even when this rank holds nothing
    *out = r;
    return 0;
}
extern "C" int rimu_vec_axpby(double alpha, rimu_vec *x, double beta, rimu_vec *y, rimu_vec *out) {
    if (out) out->version++;
    rimu_ctx *c = out->ctx;
    if (x->ctx != c || y->ctx != c || x->vt != out->vt || y->vt != out->vt) return fail(RIMU_ERR_INVALID, "axpby of incompatible vectors");
    if (x->n + y->n > 0 && (u64)(x->n + y->n) > out->cap && (out == x || out == y)) TRY(rimu_vec_reserve(out, (u64)(x->n + y->n)));
    return records_to_vec_auto(c, out, x->keys, x->vals, x->n, y->keys, y->vals, y->n, alpha, beta, 1);
}

extern "C" int rimu_annihilate_device(rimu_vec *dst, const uint64_t *d_keys, const void *d_vals, int64_t n, int method, float *ms_out) {
    if (dst) dst->version++;
    rimu_ctx *c = dst->ctx;
    TRY(enter_ctx(c));
    if (method != RIMU_ANNIHILATE_HASH && method != RIMU_ANNIHILATE_SORT && method != RIMU_ANNIHILATE_PARTITION)
        return fail(RIMU_ERR_INVALID, "annihilation method %d unknown", method);
    if (method == RIMU_ANNIHILATE_SORT && c->W != 1)
        return fail(RIMU_ERR_INVALID, "RIMU_ANNIHILATE_SORT is implemented for one-word addresses only");
    if (method == RIMU_ANNIHILATE_SORT && c->nranks > 1)
        return fail(RIMU_ERR_INVALID, "RIMU_ANNIHILATE_SORT does not filter by owner rank");
    CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    if (method == RIMU_ANNIHILATE_HASH) TRY(records_to_vec(c, dst, (const u64 *)d_keys, d_vals, n, nullptr, nullptr, 0, 1.0, 1.0, 0));
    else if (method == RIMU_ANNIHILATE_PARTITION) TRY(records_to_vec_part(c, dst, (const u64 *)d_keys, d_vals, n, nullptr, nullptr, 0, 1.0, 1.0, 0));
    else {
        dst->nb = 0; dst->diag_uid = 0;
        if (n <= 0) dst->n = 0;
        for (int attempt = 0; n > 0; attempt++) {
            CUDA_TRY(cudaMemsetAsync(c->d_stats, 0, sizeof(StatsDev), c->stream));
            if (rimu_sort_annihilate_w1(c->stream, (SortScratch *)c->sort_scratch, (const u64 *)d_keys, (const u64 *)d_vals, n,
                                        dst->vt == RIMU_VAL_I64, 64, dst->keys, (u64 *)dst->vals, dst->cap, &c->d_stats->out_count))
                return fail(RIMU_ERR_CUDA, "sort-based annihilation failed: %s", cudaGetErrorString(cudaGetLastError()));
            CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(StatsDev), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            if (c->h_stats->out_count > dst->cap) {
                if (attempt > 2) return fail(RIMU_ERR_VECTOR_FULL, "destination vector cannot be grown");
                dst->n = 0;
                TRY(rimu_vec_reserve(dst, c->h_stats->out_count + c->h_stats->out_count / 4));
                continue;
            }
            dst->n = (i64)c->h_stats->out_count;
            break;
        }
    }
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev[1]));
    if (ms_out) CUDA_TRY(cudaEventElapsedTime(ms_out, c->ev[0], c->ev[1]));
    return 0;
}
extern "C" int rimu_annihilate(rimu_vec *dst, const uint64_t *keys, const void *vals, int64_t n, int method) {
    rimu_ctx *c = dst->ctx;
    TRY(enter_ctx(c));
    if (n <= 0) { dst->n = 0; dst->nb = 0; dst->diag_uid = 0; return 0; }
    TRY(ensure_stage(c, (u64)n));
    CUDA_TRY(cudaMemcpyAsync(c->stage_keys, keys, n * c->W * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->stage_vals, vals, n * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    return rimu_annihilate_device(dst, (const uint64_t *)c->stage_keys, c->stage_vals, n, method, nullptr);
}

// ---------------------------------------------------------------- the step
int exchange_spawns(rimu_ctx *c, int vt, u64 slots, i64 *sent_out, bool to_streams) {
    const int R = c->nranks, me = c->rank;
    c->p2p_used = 0;
    if (to_streams && c->direct) {
        // Direct mode: nothing is copied.  Every rank's spawn kernels bucketed the records for every destination in their
        // own memory; the owners' merge kernels read them in place over NVLink.  All that is needed here is the barrier
        // that orders everybody's spawn kernels before anybody's merge: one tiny all-gather.
        NCCL_TRY(g_nccl.AllGather(c->xch.counts, c->d_allcounts, R, ncclUint64, c->comm, c->stream));
        CUDA_TRY(cudaEventRecord(c->ev[6], c->stream));
        c->p2p_used = 1;
        *sent_out = 0;
        return 0;
    }
    TRY(ensure_xch(c));
    NCCL_TRY(g_nccl.AllGather(c->xch.counts, c->d_allcounts, R, ncclUint64, c->comm, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev[6], c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->h_allcounts, c->d_allcounts, (size_t)R * R * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->xch.counts, 0, RIMU_MAX_RANKS * sizeof(u64), c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    u64 worst = 0, total_recv = 0, sent = 0;
    for (int s = 0; s < R; s++)
        for (int d = 0; d < R; d++) { u64 n = c->h_allcounts[s * R + d]; if (n > worst) worst = n; }
    if (worst > c->xch.cap) { // identical decision on every rank
        c->xch_worst = worst;
        return RIMU_ERR_EXCHANGE_FULL;
    }
    for (int s = 0; s < R; s++) if (s != me) total_recv += c->h_allcounts[s * R + me];
    for (int d = 0; d < R; d++) if (d != me) sent += c->h_allcounts[me * R + d];
    *sent_out = (i64)sent;
    NCCL_TRY(g_nccl.GroupStart());
    u64 off = 0;
    for (int r = 0; r < R; r++) {
        if (r == me) continue;
        u64 ns = c->h_allcounts[me * R + r], nr = c->h_allcounts[r * R + me];
        if (ns) {
            NCCL_TRY(g_nccl.Send(c->xch.keys + (u64)r * c->xch.cap * c->W, ns * c->W, ncclUint64, r, c->comm, c->stream));
            NCCL_TRY(g_nccl.Send(c->xch.vals + (u64)r * c->xch.cap, ns, ncclUint64, r, c->comm, c->stream));
        }
        if (nr) {
            NCCL_TRY(g_nccl.Recv(c->recv_keys + off * c->W, nr * c->W, ncclUint64, r, c->comm, c->stream));
            NCCL_TRY(g_nccl.Recv(c->recv_vals + off, nr, ncclUint64, r, c->comm, c->stream));
        }
        off += nr;
    }
    NCCL_TRY(g_nccl.GroupEnd());
    if (total_recv) {
        TableDev tab{c->table, slots - 1};
        TRY(dispatch_wv(c->W, vt, [&](auto tag, auto vtag) {
            typedef decltype(vtag) VT;
            if (to_streams) // partitioned step: received records join this rank's bucket streams
                append_records_kernel<decltype(tag)::w, VT><<<grid_for((i64)total_recv, c->sm_count), RIMU_TPB, 0, c->stream>>>(
                    c->recv_keys, (const VT *)c->recv_vals, (i64)total_recv, 1.0, 0, c->rank, c->nranks, c->part, c->part.me, c->d_stats);
            else
                insert_records_kernel<decltype(tag)::w, VT><<<grid_for((i64)total_recv, c->sm_count), RIMU_TPB, 0, c->stream>>>(
                    c->recv_keys, (const VT *)c->recv_vals, (i64)total_recv, 1.0, 0, 0, 1, tab, c->d_stats);
            c->launches += 1;
            return 0;
        }));
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

// ---- partitioned step (partition.cuh)
// re-segment a vector for `nb` buckets (count, scan, scatter into fresh arrays); contents are unchanged
int rebucket(rimu_vec *v, u32 nb) {
    rimu_ctx *c = v->ctx;
    TRY(ensure_seg(v, nb));
    if (v->n == 0) {
        CUDA_TRY(cudaMemsetAsync(v->seg_start, 0, nb * sizeof(u64), c->stream));
        CUDA_TRY(cudaMemsetAsync(v->seg_len, 0, nb * sizeof(u32), c->stream));
        v->nb = nb;
        return 0;
    }
    TRY(ensure_bucket_tmp(c, nb));
    u32 *counts = c->bucket_tmp, *fill = c->bucket_tmp + c->bucket_tmp_cap;
    CUDA_TRY(cudaMemsetAsync(counts, 0, nb * sizeof(u32), c->stream));
    const bool keep_diag = v->diag && v->diag_uid != 0 && v->diag_cap >= v->cap;
    if (c->spare_cap != v->cap || (keep_diag && !c->spare_diag)) { // (re)allocate the ping-pong partner for this capacity
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->spare_keys); cudaFree(c->spare_vals); cudaFree(c->spare_diag);
        c->spare_keys = c->spare_vals = nullptr; c->spare_diag = nullptr; c->spare_cap = 0;
        CUDA_TRY(rimu_malloc(&c->spare_keys, v->cap * c->W * sizeof(u64)));
        CUDA_TRY(rimu_malloc(&c->spare_vals, v->cap * sizeof(u64)));
        if (v->diag) CUDA_TRY(rimu_malloc(&c->spare_diag, v->cap * sizeof(double)));
        c->spare_cap = v->cap;
    }
    u64 *nk = c->spare_keys, *nv = c->spare_vals;
    double *nd = keep_diag ? c->spare_diag : nullptr;
    const int grid = grid_for(v->n, c->sm_count, 16);
    if (c->W == 1) bucket_count_kernel<1><<<grid, RIMU_TPB, 0, c->stream>>>(v->keys, v->n, c->nranks, nb, counts);
    else bucket_count_kernel<2><<<grid, RIMU_TPB, 0, c->stream>>>(v->keys, v->n, c->nranks, nb, counts);
    bucket_scan_kernel<<<1, 1024, 0, c->stream>>>(counts, nb, v->seg_start, v->seg_len, fill);
    if (c->W == 1) bucket_scatter_kernel<1><<<grid, RIMU_TPB, 0, c->stream>>>(v->keys, (const u64 *)v->vals, keep_diag ? v->diag : nullptr, v->n, c->nranks, nb, v->seg_start, fill, nk, nv, nd);
    else bucket_scatter_kernel<2><<<grid, RIMU_TPB, 0, c->stream>>>(v->keys, (const u64 *)v->vals, keep_diag ? v->diag : nullptr, v->n, c->nranks, nb, v->seg_start, fill, nk, nv, nd);
    CUDA_TRY(cudaGetLastError());
    c->launches += 3;
    // swap: the vector now lives in the former spare buffers, its old buffers are the next spare (all stream-ordered)
    c->spare_keys = v->keys; c->spare_vals = (u64 *)v->vals;
    v->keys = nk; v->vals = nv; v->nb = nb;
    if (keep_diag) { c->spare_diag = v->diag; v->diag = nd; v->diag_cap = v->cap; } // the cached H_aa moved with their entries
    else v->diag_uid = 0;
    return 0;
}

// testing / tuning hook: re-segment a vector for `nb` buckets (0 drops the segmentation)
extern "C" int rimu_vec_rebucket(rimu_vec *v, uint32_t nb) {
    if (v) v->version++;
    CUDA_TRY(cudaSetDevice(v->ctx->device));
    if (nb == 0) { v->nb = 0; return 0; }
    return rebucket(v, nb);
}
extern "C" int rimu_vec_buckets(rimu_vec *v, uint32_t *nb) { *nb = v->nb; return 0; }
extern "C" int rimu_vec_segments(rimu_vec *v, uint64_t *start_out, uint32_t *len_out) {
    rimu_ctx *c = v->ctx;
    TRY(enter_ctx(c));
    if (v->nb) {
        CUDA_TRY(cudaMemcpyAsync(start_out, v->seg_start, v->nb * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(len_out, v->seg_len, v->nb * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// bucket count for a step on `n` local parents
// `parents` = local parents (one rank) or the per-rank share of the global length (multi-GPU: the same number on every rank)
static u32 choose_buckets(rimu_ctx *c, const rimu_vec *src, double parents) {
    // The merge kernel works through a bucket in rounds of PART_NT items, and a round costs the same whether it is full or
    // not (measured: 0.586 / 0.656 / 0.544 ms for mean fills of 1157 / 1282 / 1389 items -- the middle one spills a few
    // items into a sixth round).  So the bucket count aims the MEAN fill just below a round boundary: `rounds` full rounds
    // minus 2.5 sigma of the Poisson spread.  RIMU_B200_ROUNDS overrides the number of rounds (tuning knob).
    static const int rounds_env = [] { const char *e = getenv("RIMU_B200_ROUNDS"); return e ? atoi(e) : 0; }();
    const double cap = (double)part_cap_items(c->W);
    const int max_rounds = (int)(cap / PART_NT);
    int rounds = rounds_env > 0 ? rounds_env : (5 * max_rounds) / 8; // 5 of 8 (measured flat between 5 and 7): room for growth and skew
    if (rounds > max_rounds - 1) rounds = max_rounds - 1;
    if (rounds < 1) rounds = 1;
    const double expected = parents * (1.0 + c->rec_per_parent) * 1.02 + 64.0;
    // Small problems: the merge takes as long as the busiest CTA's rounds, and with 5-round buckets a vector of a few thousand
    // determinants occupies a handful of the GPU's 4 x 148 CTA slots (config 1: 4 buckets, 19 us).  Use as few rounds per bucket
    // as it takes to give every SM a bucket.  (Identical on every rank: all inputs are.)
    const int base_rounds = rounds;
    if (rounds_env <= 0) {
        const double all_slots = (double)c->sm_count * PART_NT; // (one CTA per SM: every bucket has a fixed cost -- table clear, barriers)
        int r = (int)ceil(expected / all_slots);
        if (r < 1) r = 1;
        if (r < rounds) rounds = r;
    }
    // `upper` = the largest mean fill whose Poisson spread (1.5 sigma) still fits the rounds.  A fresh segmentation starts at
    // 0.9 * upper, and the bucket count is kept while the mean stays in [0.7, 1] * upper: a growing population is re-segmented
    // every ~5 % of growth (one step that carries the diagonal deposits as records), a steady one sits just below the
    // round boundary, where every round of every bucket is nearly full.
    const double slots = (double)rounds * PART_NT, upper = slots - 1.5 * sqrt(slots);
    // a count that overflowed is not tried again (nor anything within 20 % of it) until the vector has shrunk by a fifth
    double floor_nb = 1.0;
    if (c->ovf_nb) {
        if (expected >= 0.8 * c->ovf_expected) floor_nb = ceil(1.2 * (double)c->ovf_nb);
        else c->ovf_nb = 0;
    }
    if (src->nb && (double)src->nb >= floor_nb) {
        double fill = expected / src->nb;
        if (fill >= 0.7 * upper && fill <= upper) return src->nb;
        if (rounds < base_rounds) { // a segmentation made for one round more is kept (no flip-flop where the round count changes)
            const double s1 = (double)(rounds + 1) * PART_NT, u1 = s1 - 1.5 * sqrt(s1);
            if (fill >= 0.7 * u1 && fill <= u1) return src->nb;
        }
        if (c->ovf_nb && fill <= upper && (double)src->nb <= 1.5 * floor_nb) return src->nb; // the retry's count: keep it
    }
    double nb = ceil(expected / (0.9 * upper));
    if (nb < floor_nb) nb = floor_nb;
    return nb < 1.0 ? 1u : (u32)nb;
}

extern "C" int rimu_step(rimu_ctx *c, const rimu_ham *h, const rimu_step_params *prm, rimu_vec *src, rimu_vec *dst,
                         rimu_step_stats *out) {
    TRY(check_ctx_ham(c, h));
    const HkOps *ops = hk_ops(h);
    if (!ops) return RIMU_ERR_INVALID;
    if (!prm || !src || !dst) return fail(RIMU_ERR_INVALID, "null argument");
    if (src == dst) return fail(RIMU_ERR_INVALID, "source and target must not alias (Interfaces/dictvectors.jl:115-117)");
    if (src->ctx != c || dst->ctx != c) return fail(RIMU_ERR_INVALID, "vectors belong to another context");
    if (c->detached) return fail(RIMU_ERR_INVALID, "this context was detached from its peers (rimu_comm_detach): no further steps");
    if (src->vt != dst->vt) return fail(RIMU_ERR_INVALID, "source and target value types differ");
    const bool is_int = prm->style == RIMU_STYLE_INTEGER;
    if (is_int != (src->vt == RIMU_VAL_I64))
        return fail(RIMU_ERR_INVALID, "IsStochasticInteger needs Int64 vectors; the other styles need Float64 vectors");
    if (prm->style < 0 || prm->style > 3) return fail(RIMU_ERR_INVALID, "unknown stochastic style %d", prm->style);
    if (is_int && prm->proj_threshold != 0.0) return fail(RIMU_ERR_INVALID, "Thresholding not supported for integer spawns");
    StepDev p;
    memset(&p, 0, sizeof(p));
    p.style = prm->style; p.plain_h = prm->plain_h;
    p.shift = prm->shift; p.dtau = prm->time_step; p.boost = prm->boost;
    p.proj_thr = prm->proj_threshold; p.rel_thr = prm->rel_threshold; p.abs_thr = prm->abs_threshold;
    p.compress_thr = prm->compress_threshold;
    uint32_t key[2];
    rimu_step_key(prm->seed, prm->step, key);
    p.k0 = key[0]; p.k1 = key[1];
    p.rank = c->rank; p.nranks = c->nranks;
    p.init_rule = prm->initiator_rule; p.init_thr = prm->initiator_threshold;
    p.ordered = prm->ordered != 0 && !is_int; // integer sums are exact in any order
    if (p.ordered && (c->method != RIMU_ANNIHILATE_PARTITION || p.init_rule))
        return fail(RIMU_ERR_INVALID, "ordered summation needs the partitioned method and no initiator rule");
    if (p.init_rule < 0 || p.init_rule > 3) return fail(RIMU_ERR_INVALID, "unknown initiator rule %d", p.init_rule);
    if (p.init_rule && prm->plain_h == 0 && !(p.init_thr >= 0.0)) return fail(RIMU_ERR_INVALID, "initiator threshold must be >= 0");
    if (p.init_rule && c->method != RIMU_ANNIHILATE_PARTITION)
        return fail(RIMU_ERR_INVALID, "initiator rules need the partitioned method (the table method has no value lanes)");
    if (p.init_rule && c->nranks > 1 && !c->direct)
        return fail(RIMU_ERR_INVALID, "initiator rules on several GPUs need the direct exchange (peer access); the staged NCCL exchange carries no lanes");

    bool use_part = c->method == RIMU_ANNIHILATE_PARTITION;
    u64 slots = prm->table_slots ? next_pow2(prm->table_slots) : pick_slots(c, (u64)src->n * 2 + (u64)dst->n);
    if (slots > c->table_slots) slots = c->table_slots;
    // multi-GPU: the bucket count must be the same on every rank (peers store into each other's sub-streams), so it
    // is derived from the GLOBAL vector length: known from the previous step's statistics when `src` is that step's
    // untouched result, otherwise all-reduced here
    double parents = (double)src->n, g_len = (double)src->n;
    if (c->nranks > 1) {
        if (c->last_dst_uid == src->uid && c->last_dst_version == src->version) g_len = c->last_g_len;
        else TRY(rimu_comm_allreduce_f64(c, &g_len, 1));
        parents = ceil(g_len / c->nranks * 1.02) + 64.0;
    }
    u32 nb = use_part ? choose_buckets(c, src, parents) : 0;
    const bool multi = c->nranks > 1;
    // record-stream memory budget: beyond it this step falls back to the global HBM table
    size_t free_b = 0, total_b = 0;
    const double rec_bytes_per_bucket = (double)part_cap_items(c->W) * (c->W == 1 ? 16.0 : 32.0);
    i64 sent = 0;
    for (int attempt = 0;; attempt++) {
        if (use_part && nb > c->part_nb_cap && !multi) { // (multi-GPU: a per-rank fallback would desynchronise the ranks)
            CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
            double have = (double)free_b + (double)c->part_nb_cap * rec_bytes_per_bucket;
            if ((double)nb * 1.3 * rec_bytes_per_bucket * (p.init_rule ? 3.0 : 1.0) > 0.8 * have) {
                if (p.init_rule) return fail(RIMU_ERR_WORKMEM, "record streams of an initiator step do not fit in device memory");
                use_part = false;
            }
        }
        if (multi && !(use_part && c->direct)) TRY(ensure_xch(c)); // staged exchange buffers (table method / no peer access)
        int r = ops->step(c, h, p, src, dst, use_part, is_int, nb, slots, &sent);
        if (r == RIMU_ERR_EXCHANGE_FULL) {
            // nothing was sent; drain what this rank deposited locally, then report
            if (!use_part) TRY(table_fill(c, slots));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            return fail(RIMU_ERR_EXCHANGE_FULL, "per-peer exchange buffer (%llu records) too small: a rank produced %llu records for one peer",
                        (unsigned long long)c->xch.cap, (unsigned long long)c->xch_worst);
        }
        if (r) return r;
        CUDA_TRY(cudaMemcpyAsync(c->h_stats_local, c->d_stats, sizeof(StatsDev), cudaMemcpyDeviceToHost, c->stream));
        if (c->nranks > 1) {
            // ONE all-reduce per step: the integer block travels as doubles (counts stay far below 2^53, so the sums
            // are exact) next to the floating-point block (reference: Allreduce of a MultiScalar, pdvec.jl:896-902), the
            // merged-record count and the "my target vector is too small" flag.  It is also the barrier that keeps a fast
            // rank from overwriting its record streams (next step) while a peer's merge is still reading them.
            if (!c->d_red) CUDA_TRY(rimu_malloc(&c->d_red, RIMU_STATS_NPACK * sizeof(double)));
            pack_stats_kernel<<<1, 32, 0, c->stream>>>(c->d_stats, c->d_red, 0, dst->cap);
            CUDA_TRY(cudaGetLastError());
            NCCL_TRY(g_nccl.AllReduce(c->d_red, c->d_red, RIMU_STATS_NPACK, ncclFloat64, ncclSum, c->comm, c->stream));
            if (!c->h_red) CUDA_TRY(cudaMallocHost(&c->h_red, RIMU_STATS_NPACK * sizeof(double)));
            CUDA_TRY(cudaMemcpyAsync(c->h_red, c->d_red, RIMU_STATS_NPACK * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        }
        CUDA_TRY(cudaEventRecord(c->ev[5], c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        *c->h_stats = *c->h_stats_local;
        if (c->nranks > 1) unpack_stats_host(c->h_stats, c->h_red); // the summed blocks replace the local ones (host side: no second kernel)
        const StatsDev &g = *c->h_stats, &l = *c->h_stats_local;
        if (c->nranks > 1 && use_part) sent = l.sent; // (the table method counts in exchange_spawns)
        if (g.overflow_table) { // some rank ran out of room: every rank retries with more working memory
            if (attempt > 12) return fail(use_part ? RIMU_ERR_WORKMEM : RIMU_ERR_TABLE_FULL, "step working memory cannot be grown further");
            if (use_part) {
                // size the bucket count from what this attempt saw (records are counted even when dropped)
                const double cap = (double)part_cap_items(c->W);
                const double recs = multi ? (double)g.records / c->nranks * 1.02 : (double)l.records; // g.records: summed over ranks
                double need = ceil((parents + recs) * 1.15 / (0.6 * cap)); // (diagonal records may be counted twice: harmless)
                u32 nb2;
                if (!multi) {
                    // The fullest bucket = the uniform load (mean fill) + whatever a hot address piles on top of it (e.g. the
                    // reference determinant receiving a record from each of its thousands of neighbours).  More buckets only
                    // thin out the uniform part: size the retry so that the hot part + the thinned mean fits, and go straight
                    // to the table method when the hot part alone does not fit a bucket.
                    const double mean = (parents + recs) / (double)nb;
                    const double hot = (double)l.max_fill > mean ? (double)l.max_fill - mean : 0.0;
                    const double room = 0.85 * cap - hot;
                    if (room < 0.05 * cap) {
                        if (p.init_rule) return fail(RIMU_ERR_WORKMEM, "one address receives more records in a step (%llu) than a bucket holds; initiator steps cannot fall back to the table method", (unsigned long long)l.max_fill);
                        use_part = false;
                        continue;
                    }
                    const double fit = ceil((parents + recs) * 1.05 / room);
                    nb2 = fit > (double)nb * 1.25 ? (u32)fit : (u32)(nb + nb / 4 + 1);
                    if ((double)nb2 < need) nb2 = (u32)need;
                } else {
                    // several ranks: the decision must be identical everywhere, and only summed quantities are
                    nb2 = need > (double)nb * 1.5 ? (u32)need : (u32)(nb + nb / 2 + 1);
                }
                c->ovf_nb = nb; c->ovf_expected = parents * (1.0 + c->rec_per_parent) * 1.02 + 64.0;
                if (!multi && nb2 > (1u << 26)) use_part = false;
                if (multi && nb2 > (1u << 26)) return fail(RIMU_ERR_WORKMEM, "bucket streams cannot be grown further");
                nb = nb2;
                continue;
            }
            if (slots >= c->table_slots)
                return fail(RIMU_ERR_TABLE_FULL, "working table (%llu slots) too small for this step", (unsigned long long)c->table_slots);
            slots = slots * 4 > c->table_slots ? c->table_slots : slots * 4;
            continue;
        }
        // destination capacity: decided globally (the flag was summed by the step's all-reduce) so that all ranks retry together
        const u64 need_local = l.out_count;
        const bool grow = c->nranks > 1 ? g.grow_flag != 0 : need_local > dst->cap;
        if (grow) {
            dst->n = 0; dst->nb = 0;
            if (need_local > dst->cap) TRY(rimu_vec_reserve(dst, need_local + need_local / 4 + 1024));
            if (attempt > 12) return fail(RIMU_ERR_VECTOR_FULL, "destination vector cannot be grown");
            continue;
        }
        if (use_part && c->nranks == 1) c->rcnt_clean = 1; // the merge consumed and cleared every counter
        dst->n = (i64)l.out_count;
        dst->nb = use_part ? nb : 0;
        dst->diag_uid = use_part ? h->uid : 0;
        dst->version++;
        c->last_dst_uid = dst->uid; c->last_dst_version = dst->version; c->last_g_len = (double)g.len;
        if (use_part) {
            const double rec_all = multi ? (double)g.records : (double)l.records, par_all = multi ? g_len : (double)src->n;
            if (par_all > 0) { // identical on every rank in multi-GPU runs (all-reduced inputs)
                double rec = rec_all - (src->nb == nb ? 0.0 : par_all); // unsegmented sources add one diagonal record per parent
                c->rec_per_parent = 0.5 * c->rec_per_parent + 0.5 * (rec > 0 ? rec : 0.0) / par_all;
            }
            c->last_max_fill = l.max_fill;
        }
        if (out) {
            memset(out, 0, sizeof(*out));
            out->exact_steps = g.exact_steps; out->inexact_steps = g.inexact_steps; out->spawn_attempts = g.spawn_attempts;
            out->len_before = g.len_before; out->len = g.len;
            out->spawns = g.spawns; out->deaths = g.deaths; out->clones = g.clones; out->zombies = g.zombies; out->norm1 = g.norm1;
            out->ispawns = g.ispawns; out->ideaths = g.ideaths; out->iclones = g.iclones; out->izombies = g.izombies; out->inorm1 = g.inorm1;
            out->local_len = (i64)l.out_count; out->sent_records = sent; out->deposits = g.deposits;
            cudaEventElapsedTime(&out->ms_diag, c->ev[0], c->ev[4]);
            cudaEventElapsedTime(&out->ms_spawn, c->ev[4], c->ev[1]);
            cudaEventElapsedTime(&out->ms_exchange, c->ev[1], c->ev[2]);
            cudaEventElapsedTime(&out->ms_compact, c->ev[2], c->ev[3]);
            cudaEventElapsedTime(&out->ms_total, c->ev[0], c->ev[3]);
            cudaEventElapsedTime(&out->ms_reduce, c->ev[3], c->ev[5]);
            static const bool timing = getenv("RIMU_B200_TIMING") != nullptr;
            if (timing && c->nranks > 1 && (prm->step % 64) == 0) {
                float g = 0; cudaEventElapsedTime(&g, c->ev[1], c->ev[6]);
                fprintf(stderr, "[rimu_b200 r%d step %llu] spawn %.3f gather(+skew) %.3f append %.3f merge %.3f reduce(+skew) %.3f ms\n", c->rank,
                        (unsigned long long)prm->step, out->ms_spawn, g, out->ms_exchange - g, out->ms_compact, out->ms_reduce);
            }
            out->buckets = use_part ? (int64_t)nb : 0; out->max_bucket_fill = (int64_t)l.max_fill;
        }
        return 0;
    }
}

// ---------------------------------------------------------------- a batch of steps (advance!, fciqmc.jl:126-181)
// update_shift_parameters! on the host (strategies_and_params/shiftstrategy.jl:77-215): the same formulas as advance_ctl_kernel
static bool host_shift_update(rimu_shift_params *sp, double dtau, double tnorm, int64_t len) {
    bool proceed = true;
    if (len <= 0) return proceed;
    switch (sp->strategy) {
    case RIMU_SHIFT_DONT_UPDATE: proceed = tnorm < sp->target_walkers; break;
    case RIMU_SHIFT_LOG_UPDATE: sp->shift -= sp->zeta / dtau * log(tnorm / sp->pnorm); sp->pnorm = tnorm; break;
    case RIMU_SHIFT_LOG_UPDATE_AFTER_TARGET:
        if (sp->shift_mode || tnorm > sp->target_walkers) { sp->shift_mode = 1; sp->shift -= sp->zeta / dtau * log(tnorm / sp->pnorm); }
        sp->pnorm = tnorm; break;
    case RIMU_SHIFT_DOUBLE_LOG_UPDATE:
        sp->shift -= sp->xi / dtau * log(tnorm / sp->target_walkers) + sp->zeta / dtau * log(tnorm / sp->pnorm);
        sp->pnorm = tnorm; break;
    default:
        if (sp->shift_mode || tnorm > sp->target_walkers) {
            sp->shift_mode = 1;
            sp->shift -= sp->xi / dtau * log(tnorm / sp->target_walkers) + sp->zeta / dtau * log(tnorm / sp->pnorm);
        }
        sp->pnorm = tnorm; break;
    }
    return proceed;
}

static void stats_to_abi(const StatsDev &g, rimu_step_stats *out, u32 nb) {
    memset(out, 0, sizeof(*out));
    out->exact_steps = g.exact_steps; out->inexact_steps = g.inexact_steps; out->spawn_attempts = g.spawn_attempts;
    out->len_before = g.len_before; out->len = g.len;
    out->spawns = g.spawns; out->deaths = g.deaths; out->clones = g.clones; out->zombies = g.zombies; out->norm1 = g.norm1;
    out->ispawns = g.ispawns; out->ideaths = g.ideaths; out->iclones = g.iclones; out->izombies = g.izombies; out->inorm1 = g.inorm1;
    out->local_len = (i64)g.out_count; out->deposits = g.deposits;
    out->buckets = (int64_t)nb; out->max_bucket_fill = (int64_t)g.max_fill;
}

static int ensure_advance_buffers(rimu_ctx *c, const rimu_vec *cur) {
    if (!c->d_ring) {
        CUDA_TRY(rimu_malloc(&c->d_ring, RIMU_ADVANCE_CHUNK * sizeof(StatsDev)));
        CUDA_TRY(cudaMallocHost(&c->h_ring, RIMU_ADVANCE_CHUNK * sizeof(StatsDev)));
        CUDA_TRY(rimu_malloc(&c->d_ctl, sizeof(StepCtl)));
        CUDA_TRY(cudaMallocHost(&c->h_ctl, 2 * sizeof(StepCtl))); // [0]: upload image, [1]: read-back
        CUDA_TRY(rimu_malloc(&c->d_shiftlog, 2 * RIMU_ADVANCE_CHUNK * sizeof(double)));
        CUDA_TRY(cudaMallocHost(&c->h_shiftlog, 2 * RIMU_ADVANCE_CHUNK * sizeof(double)));
        CUDA_TRY(rimu_malloc(&c->d_projlog, RIMU_ADVANCE_CHUNK * RIMU_MAX_PROJECTORS * sizeof(double)));
        CUDA_TRY(cudaMallocHost(&c->h_projlog, RIMU_ADVANCE_CHUNK * RIMU_MAX_PROJECTORS * sizeof(double)));
    }
    if (c->snap_cap < (u64)cur->n) {
        cudaFree(c->snap_keys); cudaFree(c->snap_vals); cudaFree(c->snap_diag);
        c->snap_keys = c->snap_vals = nullptr; c->snap_diag = nullptr; c->snap_cap = 0;
        const u64 cap = (u64)cur->n + (u64)cur->n / 2 + 1024;
        CUDA_TRY(rimu_malloc(&c->snap_keys, cap * c->W * sizeof(u64)));
        CUDA_TRY(rimu_malloc(&c->snap_vals, cap * sizeof(u64)));
        CUDA_TRY(rimu_malloc(&c->snap_diag, cap * sizeof(double)));
        c->snap_cap = cap;
    }
    if (c->snap_nb_cap < cur->nb) {
        cudaFree(c->snap_seg_start); cudaFree(c->snap_seg_len);
        c->snap_seg_start = nullptr; c->snap_seg_len = nullptr; c->snap_nb_cap = 0;
        const u64 cap = (u64)cur->nb + cur->nb / 2 + 64;
        CUDA_TRY(rimu_malloc(&c->snap_seg_start, cap * sizeof(u64)));
        CUDA_TRY(rimu_malloc(&c->snap_seg_len, cap * sizeof(u32)));
        c->snap_nb_cap = cap;
    }
    return 0;
}

extern "C" int rimu_advance(rimu_ctx *c, const rimu_ham *h, const rimu_step_params *prm, rimu_shift_params *sp, rimu_vec *v, rimu_vec *w,
                            int64_t nsteps, const rimu_projector *projectors, int32_t nproj, rimu_step_stats *stats_out, double *shift_out,
                            double *proj_out, int64_t *steps_done, int32_t *result_in_w) {
    if (!c || !h || !prm || !sp || !v || !w || !steps_done || !result_in_w) return fail(RIMU_ERR_INVALID, "null argument");
    if (nproj < 0 || nproj > RIMU_MAX_PROJECTORS) return fail(RIMU_ERR_INVALID, "at most %d projectors per call", RIMU_MAX_PROJECTORS);
    if (nproj > 0 && (!projectors || !proj_out)) return fail(RIMU_ERR_INVALID, "projectors without an output array");
    for (int j = 0; j < nproj; j++)
        if (projectors[j].n < 0 || (projectors[j].n > 0 && (!projectors[j].keys || !projectors[j].values)))
            return fail(RIMU_ERR_INVALID, "projector %d: bad arguments", j);
    if (sp->strategy < RIMU_SHIFT_DONT_UPDATE || sp->strategy > RIMU_SHIFT_DOUBLE_LOG_UPDATE_AFTER_TARGET)
        return fail(RIMU_ERR_INVALID, "unknown shift strategy %d", sp->strategy);
    if (prm->plain_h) return fail(RIMU_ERR_INVALID, "rimu_advance steps with FirstOrderTransitionOperator (plain_h must be 0)");
    if (!(prm->time_step > 0.0)) return fail(RIMU_ERR_INVALID, "time_step must be positive");
    const HkOps *ops = hk_ops(h);
    if (!ops) return RIMU_ERR_INVALID;
    const bool is_int = prm->style == RIMU_STYLE_INTEGER;
    if (v == w) return fail(RIMU_ERR_INVALID, "the two vectors must not alias (Interfaces/dictvectors.jl:115-117)");
    if (v->ctx != c || w->ctx != c) return fail(RIMU_ERR_INVALID, "vectors belong to another context");
    if (v->vt != w->vt || is_int != (v->vt == RIMU_VAL_I64))
        return fail(RIMU_ERR_INVALID, "IsStochasticInteger needs Int64 vectors; the other styles need Float64 vectors");
    if (prm->style < 0 || prm->style > 3) return fail(RIMU_ERR_INVALID, "unknown stochastic style %d", prm->style);
    if (is_int && prm->proj_threshold != 0.0) return fail(RIMU_ERR_INVALID, "Thresholding not supported for integer spawns");
    if (prm->initiator_rule < 0 || prm->initiator_rule > 3) return fail(RIMU_ERR_INVALID, "unknown initiator rule %d", prm->initiator_rule);
    if (c->detached) return fail(RIMU_ERR_INVALID, "this context was detached from its peers (rimu_comm_detach): no further steps");
    if (nsteps < 0) return fail(RIMU_ERR_INVALID, "nsteps must not be negative");
    rimu_vec *cur = v, *oth = w;
    int64_t done = 0;
    bool ended = false;
    *steps_done = 0; *result_in_w = 0;
    rimu_step_params q = *prm;
    // one step through rimu_step (validation, sizing, re-segmentation, retries) + the host's shift update
    auto single = [&]() -> int {
        q.shift = sp->shift; q.step = prm->step + (uint64_t)done;
        rimu_step_stats st;
        TRY(rimu_step(c, h, &q, cur, oth, &st));
        std::swap(cur, oth);
        const double tnorm = is_int ? (double)st.inorm1 : st.norm1;
        const bool proceed = host_shift_update(sp, prm->time_step, tnorm, st.len);
        if (stats_out) stats_out[done] = st;
        if (shift_out) shift_out[done] = sp->shift;
        for (int j = 0; j < nproj; j++)
            TRY(rimu_vec_dot_sparse(cur, projectors[j].keys, projectors[j].values, projectors[j].n, &proj_out[done * nproj + j]));
        done++;
        if (st.len == 0 || (sp->max_length > 0 && st.len > sp->max_length) || !proceed) ended = true;
        return 0;
    };
    // the projectors stay on the device for the whole call
    u64 proj_off[RIMU_MAX_PROJECTORS + 1] = {0};
    bool proj_resident = false;
    auto upload_projectors = [&]() -> int {
        if (proj_resident || nproj == 0) return 0;
        TRY(enter_ctx(c));
        for (int j = 0; j < nproj; j++) proj_off[j + 1] = proj_off[j] + (u64)projectors[j].n;
        const u64 tot = proj_off[nproj];
        if (tot > c->proj_cap) {
            cudaFree(c->proj_keys); cudaFree(c->proj_vals); c->proj_keys = nullptr; c->proj_vals = nullptr; c->proj_cap = 0;
            CUDA_TRY(rimu_malloc(&c->proj_keys, (tot + 64) * c->W * sizeof(u64)));
            CUDA_TRY(rimu_malloc(&c->proj_vals, (tot + 64) * sizeof(double)));
            c->proj_cap = tot + 64;
        }
        for (int j = 0; j < nproj; j++) {
            if (!projectors[j].n) continue;
            CUDA_TRY(cudaMemcpyAsync(c->proj_keys + proj_off[j] * c->W, projectors[j].keys, (size_t)projectors[j].n * c->W * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(cudaMemcpyAsync(c->proj_vals + proj_off[j], projectors[j].values, (size_t)projectors[j].n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        }
        CUDA_TRY(cudaStreamSynchronize(c->stream)); // (pageable host arrays: borrowed for the duration of the call only)
        proj_resident = true;
        return 0;
    };
    // (the loop runs inside a lambda so that *steps_done / *result_in_w describe the state on EVERY exit path: after an error
    // the steps taken so far stay taken, and the caller must know which vector is current)
    auto run = [&]() -> int {
    while (done < nsteps && !ended) {
        const bool batchable = c->nranks == 1 && c->method == RIMU_ANNIHILATE_PARTITION && !prm->ordered && cur->nb != 0 && cur->n > 0 &&
                               (u64)cur->n <= RIMU_ADVANCE_MAX_N && cur->diag && cur->diag_uid == h->uid && nsteps - done >= 2 &&
                               choose_buckets(c, cur, (double)cur->n) == cur->nb;
        if (!batchable) { TRY(single()); continue; }
        // ---- a chunk of K steps, enqueued back to back
        TRY(enter_ctx(c));
        const int K = (int)std::min<int64_t>(nsteps - done, RIMU_ADVANCE_CHUNK);
        const u32 nb = cur->nb;
        const u32 nlane = prm->initiator_rule ? 3u : 1u;
        // room: both vectors must hold whatever a step of this chunk produces; sized for twice the current length (a step that
        // outgrows it stops the chunk, which is then repeated step by step with rimu_step's own growth logic)
        const u64 want = (u64)cur->n * 2 + 4096;
        if (cur->cap < want) TRY(rimu_vec_reserve(cur, want));
        if (oth->cap < want) { oth->n = 0; TRY(rimu_vec_reserve(oth, want)); }
        TRY(ensure_seg(oth, nb));
        TRY(ensure_diag(oth));
        if (!cur->diag || cur->diag_cap < cur->cap) { TRY(single()); continue; } // (cannot happen after a reserve; be safe)
        TRY(ensure_part(c, nb, nlane));
        TRY(ensure_heavy(c, std::max(cur->cap, oth->cap)));
        TRY(ensure_advance_buffers(c, cur));
        TRY(upload_projectors());
        if (nproj) CUDA_TRY(cudaMemsetAsync(c->d_projlog, 0, (size_t)K * nproj * sizeof(double), c->stream));
        // snapshot of the chunk's source
        const i64 n0 = cur->n;
        CUDA_TRY(cudaMemcpyAsync(c->snap_keys, cur->keys, (size_t)n0 * c->W * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->snap_vals, cur->vals, (size_t)n0 * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->snap_diag, cur->diag, (size_t)n0 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->snap_seg_start, cur->seg_start, (size_t)nb * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->snap_seg_len, cur->seg_len, (size_t)nb * sizeof(u32), cudaMemcpyDeviceToDevice, c->stream));
        const rimu_shift_params sp0 = *sp;
        rimu_vec *const cur0 = cur;
        // controller + statistics ring
        StepCtl &hc = c->h_ctl[0];
        hc.shift = sp->shift; hc.pnorm = sp->pnorm; hc.shift_mode = sp->shift_mode; hc.stop = 0; hc.steps_done = 0; hc.n = (unsigned long long)n0;
        CUDA_TRY(cudaMemcpyAsync(c->d_ctl, &hc, sizeof(StepCtl), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemsetAsync(c->d_ring, 0, (size_t)K * sizeof(StatsDev), c->stream));
        if (!c->rcnt_clean) CUDA_TRY(cudaMemsetAsync(c->part.rcnt, 0, (size_t)c->part.nsrc * nb * sizeof(u32), c->stream));
        c->rcnt_clean = 0;
        StepDev p;
        memset(&p, 0, sizeof(p));
        p.style = prm->style; p.dtau = prm->time_step; p.boost = prm->boost;
        p.proj_thr = prm->proj_threshold; p.rel_thr = prm->rel_threshold; p.abs_thr = prm->abs_threshold;
        p.compress_thr = prm->compress_threshold;
        p.rank = 0; p.nranks = 1;
        p.init_rule = prm->initiator_rule; p.init_thr = prm->initiator_threshold;
        p.ctl = c->d_ctl;
        c->part.ppc = spawn_chunk_parents(n0, c->sm_count);
        c->adv_grid = (u32)((n0 + n0 / 2) / c->part.ppc + 2); // (a vector that grows beyond this is covered by the kernel's chunk loop)
        AdvanceDev a;
        a.strategy = sp->strategy; a.is_int = is_int ? 1 : 0; a.target_walkers = sp->target_walkers; a.zeta = sp->zeta; a.xi = sp->xi;
        a.dtau = prm->time_step; a.max_length = sp->max_length; a.heavy_cap = c->heavy.cap;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        rimu_vec *src = cur, *dst = oth;
        for (int k = 0; k < K; k++) {
            uint32_t key[2];
            rimu_step_key(prm->seed, prm->step + (uint64_t)done + (uint64_t)k, key);
            p.k0 = key[0]; p.k1 = key[1];
            TRY(ops->enqueue(c, h, p, src, dst, is_int, nb, c->d_ring + k));
            a.dst_cap = dst->cap;
            advance_ctl_kernel<<<1, 1, 0, c->stream>>>(c->d_ctl, c->d_ring + k, a, c->d_shiftlog + 2 * k);
            c->launches += 1;
            std::swap(src, dst);
            for (int j = 0; j < nproj; j++) { // dot(::FrozenDVec, v) on the step's result: per-key lookups in the key's bucket segment
                const i64 nq = projectors[j].n;
                if (!nq) continue;
                const int grid = (int)(nq < (i64)c->sm_count * 8 ? nq : (i64)c->sm_count * 8);
                TRY(dispatch_wv(c->W, src->vt, [&](auto tag, auto vtag) {
                    typedef decltype(vtag) VT;
                    dot_sparse_kernel<decltype(tag)::w, VT><<<grid, RIMU_TPB, 0, c->stream>>>(
                        c->proj_keys + proj_off[j] * c->W, c->proj_vals + proj_off[j], nq, src->keys, (const VT *)src->vals, 0,
                        src->seg_start, src->seg_len, nb, 0, 1, c->d_projlog + (size_t)k * nproj + j);
                    return 0;
                }));
                c->launches += 1;
            }
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->h_ring, c->d_ring, (size_t)K * sizeof(StatsDev), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->h_shiftlog, c->d_shiftlog, (size_t)K * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(&c->h_ctl[1], c->d_ctl, sizeof(StepCtl), cudaMemcpyDeviceToHost, c->stream));
        if (nproj) CUDA_TRY(cudaMemcpyAsync(c->h_projlog, c->d_projlog, (size_t)K * nproj * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const StepCtl &rc = c->h_ctl[1];
        static const bool dbg = getenv("RIMU_B200_DEBUG_ADVANCE") != nullptr;
        if (dbg) fprintf(stderr, "[rimu_advance] chunk at step %lld: K=%d nb=%u n0=%lld -> stop=%d steps_done=%lld n=%llu shift=%.17g\n",
                         (long long)done, K, nb, (long long)n0, rc.stop, rc.steps_done, rc.n, rc.shift);
        if (rc.stop == 2) {
            // some step of the chunk outgrew the working memory or a vector: back to the snapshot, repeat the chunk step by step
            // (same seeds and steps: the same trajectory)
            CUDA_TRY(cudaMemcpyAsync(cur0->keys, c->snap_keys, (size_t)n0 * c->W * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
            CUDA_TRY(cudaMemcpyAsync(cur0->vals, c->snap_vals, (size_t)n0 * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
            CUDA_TRY(cudaMemcpyAsync(cur0->diag, c->snap_diag, (size_t)n0 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            CUDA_TRY(cudaMemcpyAsync(cur0->seg_start, c->snap_seg_start, (size_t)nb * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
            CUDA_TRY(cudaMemcpyAsync(cur0->seg_len, c->snap_seg_len, (size_t)nb * sizeof(u32), cudaMemcpyDeviceToDevice, c->stream));
            cur = cur0; oth = cur0 == v ? w : v;
            cur->n = n0; cur->nb = nb; cur->diag_uid = h->uid; cur->version++;
            oth->n = 0; oth->nb = 0; oth->diag_uid = 0; oth->version++;
            *sp = sp0;
            for (int k = 0; k < K && !ended; k++) TRY(single());
            continue;
        }
        const int kd = (int)rc.steps_done; // steps of this chunk that were taken (all of them unless the run ended)
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]);
        for (int k = 0; k < kd; k++) {
            if (stats_out) {
                stats_to_abi(c->h_ring[k], &stats_out[done + k], nb);
                stats_out[done + k].ms_total = ms / (float)(kd > 0 ? kd : 1);
            }
            if (shift_out) shift_out[done + k] = c->h_shiftlog[2 * k];
            for (int j = 0; j < nproj; j++) proj_out[(done + k) * nproj + j] = c->h_projlog[(size_t)k * nproj + j];
        }
        if (kd > 0) {
            const StatsDev &last = c->h_ring[kd - 1];
            if (kd & 1) std::swap(cur, oth);
            cur->n = (i64)last.out_count; cur->nb = nb; cur->diag_uid = h->uid; cur->version++;
            oth->version++; // (holds the step before: still a valid segmented vector, but nobody should rely on it)
            if (kd >= 2) { oth->n = (i64)c->h_ring[kd - 2].out_count; oth->nb = nb; oth->diag_uid = h->uid; }
            sp->shift = rc.shift; sp->pnorm = rc.pnorm; sp->shift_mode = rc.shift_mode;
            c->last_dst_uid = cur->uid; c->last_dst_version = cur->version; c->last_g_len = (double)last.len;
            const double par = kd >= 2 ? (double)c->h_ring[kd - 2].out_count : (double)n0;
            if (par > 0) c->rec_per_parent = 0.5 * c->rec_per_parent + 0.5 * (double)last.records / par;
            c->last_max_fill = last.max_fill;
            c->rcnt_clean = 1; // every merge that ran cleared the counters it consumed; stopped steps appended nothing
        } else c->rcnt_clean = 1;
        done += kd;
        if (rc.stop == 1) ended = true;
    }
    return 0;
    };
    const int status = run();
    *steps_done = done;
    *result_in_w = cur == w ? 1 : 0;
    return status;
}

