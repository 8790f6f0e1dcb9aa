// sector.cu -- host side of the dense-indexed deterministic path (sector.cuh): combinadic tables, the basis array, rank /
// unrank hooks, dense vector operations and the conversions from / to the dictionary vectors.
#include "internal.cuh"
#include "sector.cuh"

struct rimu_sector {
    rimu_ctx *ctx;
    const rimu_ham *ham;
    SectorDev dev;
    u64 dim;
    u64 *d_keys;        // address of every rank, ascending rank
    u64 *d_tab[2];
    double *d_scalar;   // reduction scratch
    double *h_scalar;
};

// ---------------------------------------------------------------- combinatorics (host)
static bool binom_table(int maxn, std::vector<std::vector<u64>> &C) { // C[n][k], false on overflow beyond 2^62
    C.assign(maxn + 1, std::vector<u64>(maxn + 2, 0));
    for (int n = 0; n <= maxn; n++) {
        C[n][0] = 1;
        for (int k = 1; k <= n; k++) {
            const u64 a = C[n - 1][k - 1], b = k <= n - 1 ? C[n - 1][k] : 0;
            if (a > (1ull << 62) || b > (1ull << 62)) C[n][k] = ~0ull; // saturate: never used for a sector that fits memory
            else C[n][k] = a + b;
        }
    }
    return true;
}
// T[chunk][byte][below] = sum over the set bits q of `byte` (t-th set bit, t = 0, 1, ...) of C(8 chunk + q, below + t + 1)
static std::vector<u64> rank_table(int bits, int ones, const std::vector<std::vector<u64>> &C) {
    const int nchunk = (bits + 7) / 8, stride = ones + 1;
    std::vector<u64> T((size_t)nchunk * 256 * stride, 0);
    for (int ch = 0; ch < nchunk; ch++)
        for (int b = 0; b < 256; b++)
            for (int below = 0; below <= ones; below++) {
                u64 r = 0;
                int t = 0;
                for (int q = 0; q < 8; q++)
                    if ((b >> q) & 1) {
                        const int pos = 8 * ch + q, j = below + t + 1;
                        if (pos < bits && j <= ones && j <= pos) r += C[pos][j]; // (C(p, j) = 0 for j > p)
                        t++;
                    }
                T[((size_t)ch * 256 + b) * stride + below] = r;
            }
    return T;
}

// ---------------------------------------------------------------- kernels
// address of rank i: per component, the largest position p with C(p, j) <= r for j = ones .. 1
__global__ void sector_unrank_kernel(SectorDev s, const u64 *__restrict__ binom, int bstride, u64 dim, u64 *__restrict__ keys) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    u64 r[2];
    if (s.ncomp == 2) { r[0] = i / s.dim[1]; r[1] = i - r[0] * s.dim[1]; } else { r[0] = i; r[1] = 0; }
    u64 key = 0;
    for (int c = 0; c < s.ncomp; c++) {
        u64 rem = r[c];
        int p = s.bits[c] - 1;
        for (int j = s.ones[c]; j >= 1; j--) {
            while (binom[p * bstride + j] > rem) p--;
            rem -= binom[p * bstride + j];
            key |= 1ull << (s.shift[c] + p);
            p--;
        }
    }
    keys[i] = key;
}
__global__ void sector_rank_kernel(SectorDev s, const u64 *__restrict__ keys, i64 n, i64 *__restrict__ out) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (i64)sector_rank(s, keys[i]);
}
__global__ void sector_scatter_kernel(SectorDev s, const u64 *__restrict__ keys, const double *__restrict__ vals, i64 n, double *__restrict__ d) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) d[sector_rank(s, keys[i])] = vals[i];
}
__global__ void sector_set_kernel(const i64 *__restrict__ idx, const double *__restrict__ vals, i64 n, double *__restrict__ d) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[idx[i]] = vals[i];
}
__global__ void sector_axpby_kernel(double a, const double *__restrict__ x, double b, double *__restrict__ y, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) y[i] = a * x[i] + b * y[i];
}
__global__ void sector_dot_kernel(const double *__restrict__ x, const double *__restrict__ y, u64 n, double *out) {
    double acc = 0.0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) acc += x[i] * y[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(out, acc);
}
__global__ void sector_compact_kernel(const u64 *__restrict__ keys, const double *__restrict__ d, u64 dim, u64 *__restrict__ okeys,
                                      double *__restrict__ ovals, u64 cap, unsigned long long *cursor) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const double v = i < dim ? d[i] : 0.0;
    const bool keep = v != 0.0;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
        const u64 at = base + __popc(m & ((1u << lane) - 1u));
        if (at < cap) { okeys[at] = keys[i]; ovals[at] = v; }
    }
}

// ---------------------------------------------------------------- API
extern "C" int rimu_sector_create(rimu_ctx *c, const rimu_ham *h, rimu_sector **out) {
    if (!c || !h || !out) return fail(RIMU_ERR_INVALID, "null argument");
    if (c->W != 1 || h->W != 1) return fail(RIMU_ERR_INVALID, "dense sectors need one-word addresses");
    if (h->hk == HK_RS_COMP) return fail(RIMU_ERR_INVALID, "dense sectors are numbered for BoseFS, FermiFS and FermiFS2C addresses; use the dictionary vector for a general CompositeFS");
    if (h->hk == HK_TC_F2C) return fail(RIMU_ERR_INVALID, "the dense H*v is a gather over a real symmetric matrix; Transcorrelated1D is not Hermitian");
    TRY(enter_ctx(c));
    const rimu_ham_desc &d = h->desc;
    rimu_sector *s = new rimu_sector();
    memset(s, 0, sizeof(*s));
    s->ctx = c; s->ham = h;
    SectorDev &sd = s->dev;
    const int M = d.num_modes;
    if (d.addr_kind == RIMU_ADDR_BOSE) {
        sd.ncomp = 1; sd.bits[0] = d.num_particles[0] + M - 1; sd.ones[0] = d.num_particles[0]; sd.shift[0] = 0;
    } else if (d.addr_kind == RIMU_ADDR_FERMI) {
        sd.ncomp = 1; sd.bits[0] = M; sd.ones[0] = d.num_particles[0]; sd.shift[0] = 0;
    } else {
        sd.ncomp = 2;
        for (int k = 0; k < 2; k++) { sd.bits[k] = M; sd.ones[k] = d.num_particles[k]; sd.shift[k] = k * M; }
    }
    std::vector<std::vector<u64>> C;
    binom_table(64, C);
    double dimf = 1.0;
    for (int k = 0; k < sd.ncomp; k++) {
        sd.nchunk[k] = (sd.bits[k] + 7) / 8;
        sd.dim[k] = C[sd.bits[k]][sd.ones[k]];
        dimf *= (double)sd.dim[k];
    }
    if (sd.ncomp == 1) sd.dim[1] = 1;
    if (dimf > 4.0e9) { delete s; return fail(RIMU_ERR_INVALID, "sector of %.3g addresses is too large for a dense vector", dimf); }
    s->dim = sd.ncomp == 2 ? sd.dim[0] * sd.dim[1] : sd.dim[0];
    for (int k = 0; k < sd.ncomp; k++) {
        std::vector<u64> T = rank_table(sd.bits[k], sd.ones[k], C);
        CUDA_TRY(rimu_malloc(&s->d_tab[k], T.size() * sizeof(u64)));
        CUDA_TRY(cudaMemcpy(s->d_tab[k], T.data(), T.size() * sizeof(u64), cudaMemcpyHostToDevice));
        sd.tab[k] = s->d_tab[k];
    }
    // the basis: unrank every index once
    std::vector<u64> flat(65 * 66);
    for (int n = 0; n <= 64; n++) for (int k = 0; k <= 65; k++) flat[n * 66 + k] = k < (int)C[n].size() ? C[n][k] : 0;
    u64 *d_binom = nullptr;
    CUDA_TRY(rimu_malloc(&d_binom, flat.size() * sizeof(u64)));
    CUDA_TRY(cudaMemcpyAsync(d_binom, flat.data(), flat.size() * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(rimu_malloc(&s->d_keys, s->dim * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&s->d_scalar, 2 * sizeof(double)));
    CUDA_TRY(cudaMallocHost(&s->h_scalar, 2 * sizeof(double)));
    sector_unrank_kernel<<<(unsigned)((s->dim + 255) / 256), 256, 0, c->stream>>>(sd, d_binom, 66, s->dim, s->d_keys);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(d_binom);
    *out = s;
    return 0;
}
extern "C" int rimu_sector_destroy(rimu_sector *s) {
    if (!s) return 0;
    cudaSetDevice(s->ctx->device);
    cudaFree(s->d_keys); cudaFree(s->d_tab[0]); cudaFree(s->d_tab[1]); cudaFree(s->d_scalar); cudaFreeHost(s->h_scalar);
    cudaGetLastError();
    delete s;
    return 0;
}
extern "C" int rimu_sector_dim(const rimu_sector *s, uint64_t *dim_out) { *dim_out = s->dim; return 0; }

extern "C" int rimu_sector_rank(rimu_sector *s, const uint64_t *keys, int64_t n, int64_t *index_out) {
    rimu_ctx *c = s->ctx;
    TRY(enter_ctx(c));
    if (n <= 0) return 0;
    u64 *dk = nullptr; i64 *di = nullptr;
    CUDA_TRY(rimu_malloc(&dk, n * sizeof(u64)));
    CUDA_TRY(rimu_malloc(&di, n * sizeof(i64)));
    CUDA_TRY(cudaMemcpyAsync(dk, keys, n * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    sector_rank_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(s->dev, dk, n, di);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(index_out, di, n * sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(dk); cudaFree(di);
    return 0;
}
extern "C" int rimu_sector_keys(rimu_sector *s, int64_t first, int64_t count, uint64_t *keys_out) {
    rimu_ctx *c = s->ctx;
    TRY(enter_ctx(c));
    if (first < 0 || count < 0 || (u64)(first + count) > s->dim) return fail(RIMU_ERR_INVALID, "index range outside the sector");
    if (count) CUDA_TRY(cudaMemcpy(keys_out, s->d_keys + first, count * sizeof(u64), cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int rimu_sector_vec_create(rimu_sector *s, double **d_out) {
    TRY(enter_ctx(s->ctx));
    CUDA_TRY(rimu_malloc(d_out, s->dim * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(*d_out, 0, s->dim * sizeof(double), s->ctx->stream));
    return 0;
}
extern "C" int rimu_sector_vec_destroy(rimu_sector *s, double *d) {
    if (!d) return 0;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    cudaFree(d);
    return 0;
}
extern "C" int rimu_sector_vec_set(rimu_sector *s, double *d, const int64_t *index, const double *vals, int64_t n) {
    rimu_ctx *c = s->ctx;
    TRY(enter_ctx(c));
    for (int64_t i = 0; i < n; i++) if (index[i] < 0 || (u64)index[i] >= s->dim) return fail(RIMU_ERR_INVALID, "index outside the sector");
    CUDA_TRY(cudaMemsetAsync(d, 0, s->dim * sizeof(double), c->stream));
    if (n <= 0) return 0;
    i64 *di = nullptr; double *dv = nullptr;
    CUDA_TRY(rimu_malloc(&di, n * sizeof(i64)));
    CUDA_TRY(rimu_malloc(&dv, n * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(di, index, n * sizeof(i64), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(dv, vals, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    sector_set_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(di, dv, n, d);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(di); cudaFree(dv);
    return 0;
}
extern "C" int rimu_sector_vec_get(rimu_sector *s, const double *d, int64_t first, int64_t count, double *out) {
    TRY(enter_ctx(s->ctx));
    if (first < 0 || count < 0 || (u64)(first + count) > s->dim) return fail(RIMU_ERR_INVALID, "index range outside the sector");
    CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    if (count) CUDA_TRY(cudaMemcpy(out, d + first, count * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int rimu_sector_vec_gather(rimu_sector *s, const double *d, const int64_t *index, int64_t n, double *out) {
    TRY(enter_ctx(s->ctx));
    CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    for (int64_t i = 0; i < n; i++) {
        if (index[i] < 0 || (u64)index[i] >= s->dim) return fail(RIMU_ERR_INVALID, "index outside the sector");
        CUDA_TRY(cudaMemcpy(out + i, d + index[i], sizeof(double), cudaMemcpyDeviceToHost));
    }
    return 0;
}

extern "C" int rimu_sector_mul(rimu_sector *s, const double *d_x, double *d_y, float *ms_out) {
    rimu_ctx *c = s->ctx;
    TRY(enter_ctx(c));
    if (d_x == d_y) return fail(RIMU_ERR_INVALID, "source and target must not alias");
    const HkOps *ops = nullptr;
    switch (s->ham->hk) {
#ifndef RIMU_TUNE_ONLY_MOM1D
    case HK_REAL1D_BOSE: ops = rimu_hk_ops_0(); break;
    case HK_REAL1D_BOSE_PLAIN: ops = rimu_hk_ops_8(); break;
    case HK_MOM1D_F2C: ops = rimu_hk_ops_2(); break;
    case HK_RS_BOSE: ops = rimu_hk_ops_3(); break;
    case HK_RS_FERMI: ops = rimu_hk_ops_4(); break;
    case HK_RS_F2C: ops = rimu_hk_ops_5(); break;
    case HK_MOM1D_BOSE: ops = rimu_hk_ops_1(); break;
#endif
    case HK_MOM1D_BOSE_PLAIN: ops = rimu_hk_ops_9(); break;
    default: return fail(RIMU_ERR_INVALID, "this Hamiltonian kind has no dense H*v");
    }
    CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    TRY(ops->sector_mul(c, s->ham, &s->dev, s->d_keys, d_x, d_y, s->dim));
    c->launches += 1;
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    if (ms_out) {
        CUDA_TRY(cudaEventSynchronize(c->ev[1]));
        CUDA_TRY(cudaEventElapsedTime(ms_out, c->ev[0], c->ev[1]));
    }
    return 0;
}
extern "C" int rimu_sector_axpby(rimu_sector *s, double a, const double *d_x, double b, double *d_y) {
    rimu_ctx *c = s->ctx;
    TRY(enter_ctx(c));
    sector_axpby_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(a, d_x, b, d_y, s->dim);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int rimu_sector_dot(rimu_sector *s, const double *d_x, const double *d_y, double *out) {
    rimu_ctx *c = s->ctx;
    TRY(enter_ctx(c));
    CUDA_TRY(cudaMemsetAsync(s->d_scalar, 0, sizeof(double), c->stream));
    sector_dot_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_x, d_y, s->dim, s->d_scalar);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(s->h_scalar, s->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out = s->h_scalar[0];
    return 0;
}
// dictionary vector -> dense (Float64 vectors; absent addresses are zero)
extern "C" int rimu_sector_from_vec(rimu_sector *s, rimu_vec *v, double *d_out) {
    rimu_ctx *c = s->ctx;
    if (v->ctx != c || v->vt != RIMU_VAL_F64) return fail(RIMU_ERR_INVALID, "dense sectors take Float64 vectors of the same context");
    TRY(enter_ctx(c));
    CUDA_TRY(cudaMemsetAsync(d_out, 0, s->dim * sizeof(double), c->stream));
    if (v->n > 0) {
        sector_scatter_kernel<<<grid_for(v->n, c->sm_count, 16), 256, 0, c->stream>>>(s->dev, v->keys, (const double *)v->vals, v->n, d_out);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}
// dense -> dictionary vector (non-zero entries, unsegmented)
extern "C" int rimu_sector_to_vec(rimu_sector *s, const double *d, rimu_vec *v) {
    rimu_ctx *c = s->ctx;
    if (v->ctx != c || v->vt != RIMU_VAL_F64) return fail(RIMU_ERR_INVALID, "dense sectors give Float64 vectors of the same context");
    TRY(enter_ctx(c));
    v->version++; v->nb = 0; v->diag_uid = 0;
    for (int attempt = 0;; attempt++) {
        unsigned long long *cursor = (unsigned long long *)&c->d_stats->out_count;
        CUDA_TRY(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), c->stream));
        sector_compact_kernel<<<(unsigned)((s->dim + 255) / 256), 256, 0, c->stream>>>(s->d_keys, d, s->dim, v->keys, (double *)v->vals, v->cap, cursor);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&c->h_stats->out_count, cursor, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const u64 n = c->h_stats->out_count;
        if (n > v->cap) {
            if (attempt) return fail(RIMU_ERR_VECTOR_FULL, "destination vector cannot be grown to %llu entries", (unsigned long long)n);
            v->n = 0;
            TRY(rimu_vec_reserve(v, n));
            continue;
        }
        v->n = (i64)n;
        return 0;
    }
}
