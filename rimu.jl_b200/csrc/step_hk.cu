// step_hk.cu -- the kernels of ONE Hamiltonian kind (compiled once per kind with -DRIMU_HK=n): one attempt of the FCIQMC
// step (partitioned method and table method, both value types, every address width the kind supports) and the
// element-wise Hamiltonian hooks.  rimu_step (api.cu) owns validation, sizing and the retry loop.
#include "internal.cuh"

#ifndef RIMU_HK
#error "compile with -DRIMU_HK=<HamKind>"
#endif

template <int HK, int W, class VT>
static int step_once(rimu_ctx *c, const rimu_ham *h, const StepDev &p, rimu_vec *src, rimu_vec *dst, u64 slots, i64 *sent) {
    const i64 n = src->n;
    TableDev tab{c->table, slots - 1};
    CUDA_TRY(cudaMemsetAsync(c->d_stats, 0, sizeof(StatsDev), c->stream));
    CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    if (n > 0) {
        TRY(ensure_scratch(c, (u64)n));
        const i64 nblk = (n + RIMU_TPB - 1) / RIMU_TPB;
        diag_count_kernel<HK, W, VT><<<(unsigned)nblk, RIMU_TPB, 0, c->stream>>>(
            h->dev, p, src->keys, (const VT *)src->vals, n, tab, c->xch, c->local_off, c->block_tot, c->d_stats);
        scan_blocks_kernel<<<1, 1024, 0, c->stream>>>(c->block_tot, nblk, c->block_base, c->d_stats);
        CUDA_TRY(cudaEventRecord(c->ev[4], c->stream));
        spawn_kernel<HK, W, VT><<<c->sm_count * 8, RIMU_TPB, 0, c->stream>>>(
            h->dev, p, src->keys, (const VT *)src->vals, n, c->block_base, c->local_off, tab, c->xch, c->d_stats);
        CUDA_TRY(cudaGetLastError());
        c->launches += 3;
    } else {
        CUDA_TRY(cudaEventRecord(c->ev[4], c->stream));
    }
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    *sent = 0;
    if (c->nranks > 1) {
        int r = exchange_spawns(c, dst->vt, slots, sent, false);
        if (r) return r;
    }
    CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
    TRY((compact_into<W, VT>(c, dst, slots, p)));
    c->launches += 1;
    CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
    dst->nb = 0; dst->diag_uid = 0;
    return 0;
}

template <int HK, int W, class VT>
static int step_part_once(rimu_ctx *c, const rimu_ham *h, const StepDev &p, rimu_vec *src, rimu_vec *dst, u32 nb, i64 *sent) {
    const i64 n = src->n;
    // Parents are read by bucket segment.  A source that is not segmented for this bucket count (fresh upload, changed count):
    //  * one rank: its diagonal deposits travel through the bucket streams like spawns (one extra record per parent);
    //  * several ranks: the sub-streams are sized for 1/R of a bucket's records, so the parents would not fit the one local
    //    sub-stream -- the source is re-segmented in place instead (count / scan / scatter; the dictionary is unchanged).
    if (c->nranks > 1 && c->direct && src->nb != nb && n > 0) TRY(rebucket(src, nb));
    const bool seg = src->nb == nb && n > 0;
    TRY(ensure_seg(dst, nb));
    TRY(ensure_diag(dst));
    const double *src_diag = (src->diag_uid == h->uid && src->diag) ? src->diag : nullptr;
    const u32 nlane = p.init_rule ? 3u : 1u;
    TRY(ensure_part(c, nb, nlane));
    TRY(ensure_heavy(c, (u64)n));
    const size_t smem_init = (size_t)part_cap_items(W) * 8; // unsafe lane of the initiator rules
    static bool attr_set[HK_COUNT][3][2] = {};
    if (!attr_set[HK][W][std::is_integral<VT>::value]) {
        CUDA_TRY(cudaFuncSetAttribute(merge_kernel<HK, W, VT, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem_bytes(W)));
        CUDA_TRY(cudaFuncSetAttribute(merge_kernel<HK, W, VT, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(part_smem_bytes(W) + smem_init)));
        if constexpr (!std::is_integral<VT>::value)
            CUDA_TRY(cudaFuncSetAttribute(merge_kernel<HK, W, VT, 0, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem_bytes(W)));
        attr_set[HK][W][std::is_integral<VT>::value] = true;
    }
    const bool fs = fixed_step_ok(p, std::is_integral<VT>::value); // the default production run: kernels with its parameters folded in
    CUDA_TRY(cudaMemsetAsync(c->d_stats, 0, sizeof(StatsDev), c->stream)); // statistics + heavy-parent queue head
    if (!c->rcnt_clean) // fills of every (destination, lane) sub-stream; on one rank the previous merge has already cleared them
        CUDA_TRY(cudaMemsetAsync(c->part.rcnt, 0, (size_t)c->part.nsrc * nb * sizeof(u32), c->stream));
    c->rcnt_clean = 0; // until this step has merged successfully
    CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    CUDA_TRY(cudaEventRecord(c->ev[4], c->stream));
    if (n > 0) {
        c->part.ppc = spawn_chunk_parents(n, c->sm_count);
        const i64 nchunks = (n + c->part.ppc - 1) / c->part.ppc;
        const int grid = (int)(nchunks < (i64)c->sm_count * 8 ? nchunks : (i64)c->sm_count * 8);
        bool launched = false;
        if constexpr (!std::is_integral<VT>::value) {
            if (fs) {
                spawn_part_kernel<HK, W, VT, true><<<grid, SPAWN_NT, 0, c->stream>>>(
                    h->dev, p, src->keys, (const VT *)src->vals, n, c->part, c->xch, c->heavy, c->d_stats);
                launched = true;
            }
        }
        if (!launched)
            spawn_part_kernel<HK, W, VT><<<grid, SPAWN_NT, 0, c->stream>>>(
                h->dev, p, src->keys, (const VT *)src->vals, n, c->part, c->xch, c->heavy, c->d_stats);
        spawn_heavy_kernel<HK, W, VT><<<c->sm_count * 4, SPAWN_NT, 0, c->stream>>>(
            h->dev, p, src->keys, (const VT *)src->vals, c->part, c->xch, c->heavy, c->d_stats);
        c->launches += 2;
        if (!seg) {
            diag_append_kernel<HK, W, VT><<<grid_for(n, c->sm_count, 16), RIMU_TPB, 0, c->stream>>>(
                h->dev, p, src->keys, (const VT *)src->vals, src_diag, n, c->part, c->d_stats);
            c->launches += 1;
        }
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    *sent = 0;
    if (c->nranks > 1) {
        int r = exchange_spawns(c, dst->vt, 0, sent, true);
        if (r) return r;
    }
    CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
    SegSrc ss{src->keys, (const u64 *)src->vals, seg ? src->seg_start : nullptr, seg ? src->seg_len : nullptr, src_diag};
    SegDst sd{dst->keys, (u64 *)dst->vals, dst->seg_start, dst->seg_len, dst->cap, dst->diag};
    const int mgrid = (int)(nb < c->merge_grid_cap ? nb : c->merge_grid_cap);
    if constexpr (!std::is_integral<VT>::value) {
        if (p.ordered) { // audit mode: sorted, order-deterministic annihilation + fixed-order walker number
            static bool ord_attr[HK_COUNT][3] = {};
            if (!ord_attr[HK][W]) {
                CUDA_TRY(cudaFuncSetAttribute(merge_kernel<HK, W, VT, 0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem_bytes(W)));
                ord_attr[HK][W] = true;
            }
            const u32 nparts = nb; // one partial walker number per bucket: independent of the launch geometry
            if (c->ord_cap < nparts) {
                cudaFree(c->d_ord); c->d_ord = nullptr; c->ord_cap = 0;
                CUDA_TRY(rimu_malloc(&c->d_ord, (size_t)nparts * sizeof(double)));
                c->ord_cap = nparts;
            }
            merge_kernel<HK, W, VT, 0, false, true><<<mgrid, PART_NT, part_smem_bytes(W), c->stream>>>(h->dev, p, ss, 1.0, c->part, sd, c->d_stats, c->d_ord);
            ordered_sum_kernel<<<1, 1, 0, c->stream>>>(c->d_ord, nparts, &c->d_stats->norm1);
            CUDA_TRY(cudaGetLastError());
            c->launches += 2;
            CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
            return 0;
        }
    }
    bool merged = false;
    if constexpr (!std::is_integral<VT>::value) {
        if (fs) {
            merge_kernel<HK, W, VT, 0, false, false, true><<<mgrid, PART_NT, part_smem_bytes(W), c->stream>>>(h->dev, p, ss, 1.0, c->part, sd, c->d_stats);
            merged = true;
        }
    }
    if (merged) {}
    else if (p.init_rule) merge_kernel<HK, W, VT, 0, true><<<mgrid, PART_NT, part_smem_bytes(W) + smem_init, c->stream>>>(h->dev, p, ss, 1.0, c->part, sd, c->d_stats);
    else merge_kernel<HK, W, VT, 0, false><<<mgrid, PART_NT, part_smem_bytes(W), c->stream>>>(h->dev, p, ss, 1.0, c->part, sd, c->d_stats);
    CUDA_TRY(cudaGetLastError());
    c->launches += 1;
    CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
    return 0;
}


// batch mode: the kernels of one step, nothing else (rimu_advance sized the buffers, cleared the statistics ring and the
// fill counters, and follows every step with the controller kernel)
template <int HK, int W, class VT>
static int step_part_enqueue(rimu_ctx *c, const rimu_ham *h, const StepDev &p, rimu_vec *src, rimu_vec *dst, u32 nb, StatsDev *st) {
    const size_t smem_init = (size_t)part_cap_items(W) * 8;
    static bool attr_set[3][2] = {};
    if (!attr_set[W][std::is_integral<VT>::value]) {
        CUDA_TRY(cudaFuncSetAttribute(merge_kernel<HK, W, VT, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem_bytes(W)));
        CUDA_TRY(cudaFuncSetAttribute(merge_kernel<HK, W, VT, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(part_smem_bytes(W) + smem_init)));
        attr_set[W][std::is_integral<VT>::value] = true;
    }
    HeavyDev hv = c->heavy;
    hv.packed = (u64 *)&st->heavy_packed;
    // the source length is known on the device only: chunk size and grid from the length at the start of the batch (c->part.ppc,
    // set by rimu_advance); the kernel's chunk loop covers whatever the vector has grown to
    const i64 nchunks = ((i64)src->cap + c->part.ppc - 1) / c->part.ppc;
    const i64 want = (i64)c->adv_grid;
    const int grid = (int)std::max<i64>(1, std::min<i64>(std::min<i64>(nchunks, want), (i64)c->sm_count * 8));
    spawn_part_kernel<HK, W, VT><<<grid, SPAWN_NT, 0, c->stream>>>(h->dev, p, src->keys, (const VT *)src->vals, 0, c->part, c->xch, hv, st);
    spawn_heavy_kernel<HK, W, VT><<<c->sm_count * 4, SPAWN_NT, 0, c->stream>>>(h->dev, p, src->keys, (const VT *)src->vals, c->part, c->xch, hv, st);
    SegSrc ss{src->keys, (const u64 *)src->vals, src->seg_start, src->seg_len, src->diag};
    SegDst sd{dst->keys, (u64 *)dst->vals, dst->seg_start, dst->seg_len, dst->cap, dst->diag};
    const int mgrid = (int)(nb < c->merge_grid_cap ? nb : c->merge_grid_cap);
    if (p.init_rule) merge_kernel<HK, W, VT, 0, true><<<mgrid, PART_NT, part_smem_bytes(W) + smem_init, c->stream>>>(h->dev, p, ss, 1.0, c->part, sd, st);
    else merge_kernel<HK, W, VT, 0, false><<<mgrid, PART_NT, part_smem_bytes(W), c->stream>>>(h->dev, p, ss, 1.0, c->part, sd, st);
    CUDA_TRY(cudaGetLastError());
    c->launches += 3;
    return 0;
}

#define RIMU_CAT_(a, b) a##b
#define RIMU_CAT(a, b) RIMU_CAT_(a, b)
#define HKNAME(x) RIMU_CAT(RIMU_CAT(x, _hk), RIMU_HK)
static constexpr int HKC = RIMU_HK;
// address widths a kind is compiled for: the bosonic models need two words beyond 63 bits; fermionic addresses fit one word
#ifdef RIMU_TUNE_ONLY_MOM1D
static constexpr bool HAS_W2 = false;
#else
static constexpr int HKB = HkBase<HKC>::value;
static constexpr bool HAS_W2 = HKB == HK_REAL1D_BOSE || HKB == HK_MOM1D_BOSE || HKB == HK_RS_BOSE || HKB == HK_RS_COMP;
#endif

template <int W> static int step_w(rimu_ctx *c, const rimu_ham *h, const StepDev &p, rimu_vec *src, rimu_vec *dst, bool use_part, bool is_int,
                            u32 nb, u64 slots, i64 *sent) {
    if (use_part) {
        if (is_int) return step_part_once<HKC, W, i64>(c, h, p, src, dst, nb, sent);
        return step_part_once<HKC, W, double>(c, h, p, src, dst, nb, sent);
    }
    if (is_int) return step_once<HKC, W, i64>(c, h, p, src, dst, slots, sent);
    return step_once<HKC, W, double>(c, h, p, src, dst, slots, sent);
}
static int HKNAME(step_entry)(rimu_ctx *c, const rimu_ham *h, const StepDev &p, rimu_vec *src, rimu_vec *dst, bool use_part, bool is_int, u32 nb,
               u64 slots, i64 *sent) {
    if (h->W == 1) return step_w<1>(c, h, p, src, dst, use_part, is_int, nb, slots, sent);
    if constexpr (HAS_W2) return step_w<2>(c, h, p, src, dst, use_part, is_int, nb, slots, sent);
    return fail(RIMU_ERR_INVALID, "this Hamiltonian kind is compiled for one-word addresses only");
}
static int HKNAME(enqueue_entry)(rimu_ctx *c, const rimu_ham *h, const StepDev &p, rimu_vec *src, rimu_vec *dst, bool is_int, u32 nb, StatsDev *st) {
    if (h->W == 1) return is_int ? step_part_enqueue<HKC, 1, i64>(c, h, p, src, dst, nb, st) : step_part_enqueue<HKC, 1, double>(c, h, p, src, dst, nb, st);
    if constexpr (HAS_W2) return is_int ? step_part_enqueue<HKC, 2, i64>(c, h, p, src, dst, nb, st) : step_part_enqueue<HKC, 2, double>(c, h, p, src, dst, nb, st);
    return fail(RIMU_ERR_INVALID, "this Hamiltonian kind is compiled for one-word addresses only");
}
static int HKNAME(diag_entry)(rimu_ctx *c, const rimu_ham *h, const u64 *d_keys, i64 n, double *d_out, i64 *d_nod) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (h->W == 1) ham_diag_kernel<HKC, 1><<<grid, 256, 0, c->stream>>>(h->dev, d_keys, n, d_out, d_nod);
    else if constexpr (HAS_W2) ham_diag_kernel<HKC, 2><<<grid, 256, 0, c->stream>>>(h->dev, d_keys, n, d_out, d_nod);
    else return fail(RIMU_ERR_INVALID, "this Hamiltonian kind is compiled for one-word addresses only");
    CUDA_TRY(cudaGetLastError());
    return 0;
}
static int HKNAME(offdiag_entry)(rimu_ctx *c, const rimu_ham *h, const u64 *d_key, i64 first0, i64 count, u64 *d_keys_out, double *d_vals) {
    const unsigned grid = (unsigned)((count + 255) / 256);
    if (h->W == 1) ham_offdiag_kernel<HKC, 1><<<grid, 256, 0, c->stream>>>(h->dev, d_key, first0, count, d_keys_out, d_vals);
    else if constexpr (HAS_W2) ham_offdiag_kernel<HKC, 2><<<grid, 256, 0, c->stream>>>(h->dev, d_key, first0, count, d_keys_out, d_vals);
    else return fail(RIMU_ERR_INVALID, "this Hamiltonian kind is compiled for one-word addresses only");
    CUDA_TRY(cudaGetLastError());
    return 0;
}
static int HKNAME(sector_mul_entry)(rimu_ctx *c, const rimu_ham *h, const SectorDev *s, const u64 *d_keys, const double *d_x, double *d_y, u64 dim) {
    if (h->W != 1) return fail(RIMU_ERR_INVALID, "dense sectors need one-word addresses");
    sector_mul_kernel<HKC><<<(unsigned)((dim + 255) / 256), 256, 0, c->stream>>>(h->dev, *s, d_keys, d_x, d_y, dim);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
static const HkOps HKNAME(OPS) = {HKNAME(step_entry), HKNAME(diag_entry), HKNAME(offdiag_entry), HKNAME(sector_mul_entry), HKNAME(enqueue_entry)};
const HkOps *RIMU_CAT(rimu_hk_ops_, RIMU_HK)() { return &HKNAME(OPS); }
