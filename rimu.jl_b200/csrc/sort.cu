// sort.cu -- RIMU_ANNIHILATE_SORT: annihilation of a spawn list by radix sort + segmented reduce.
// The measured alternative the task statement asks for (the sort itself is CUB's onesweep radix sort from the
// CUDA toolkit; the reduce-by-key and zero-dropping compaction around it are ours).  Summation order is the
// sorted record order, i.e. deterministic for Float64.  One-word addresses only.
#include "common.cuh"
#include <cub/cub.cuh>

struct SortScratch {
    void *tmp; size_t tmp_bytes;
    u64 *k_sorted, *v_sorted, *k_unique, *v_sum; u64 cap;
    u64 *d_runs;
};

template <class VT> struct AddBits {
    __device__ __forceinline__ u64 operator()(const u64 &a, const u64 &b) const {
        union { u64 b; VT v; } x, y, z; x.b = a; y.b = b; z.v = x.v + y.v; return z.b;
    }
};

template <class VT>
__global__ void compact_nonzero_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ sums, const u64 *__restrict__ nruns,
                                       u64 *__restrict__ out_keys, u64 *__restrict__ out_vals, u64 out_cap, u64 *cursor) {
    const u64 n = *nruns;
    const int lane = threadIdx.x & 31;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 base = (u64)blockIdx.x * blockDim.x; base < n; base += stride) {
        const u64 i = base + threadIdx.x;
        bool keep = false; u64 k = 0, v = 0;
        if (i < n) { k = keys[i]; v = sums[i]; union { u64 b; VT x; } cv; cv.b = v; keep = cv.x != (VT)0; }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            u64 at = 0;
            if (lane == 0) at = atomicAdd((unsigned long long *)cursor, (unsigned long long)__popc(m));
            at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1));
            if (keep && at < out_cap) { out_keys[at] = k; out_vals[at] = v; }
        }
    }
}

extern "C" int rimu_sort_scratch_bytes(void) { return (int)sizeof(SortScratch); }

// returns 0 ok, 1 = CUDA error (message via cudaGetLastError by the caller)
extern "C" int rimu_sort_annihilate_w1(cudaStream_t stream, SortScratch *s, const u64 *keys, const u64 *vals, long long n, int is_int,
                                       int key_bits, u64 *out_keys, u64 *out_vals, u64 out_cap, u64 *d_cursor) {
    if ((u64)n > s->cap) {
        cudaFree(s->k_sorted); cudaFree(s->v_sorted); cudaFree(s->k_unique); cudaFree(s->v_sum);
        s->cap = 0;
        u64 cap = (u64)n + (u64)n / 4 + 1024;
        if (cudaMalloc(&s->k_sorted, cap * 8) || cudaMalloc(&s->v_sorted, cap * 8) || cudaMalloc(&s->k_unique, cap * 8) || cudaMalloc(&s->v_sum, cap * 8)) return 1;
        s->cap = cap;
    }
    if (!s->d_runs && cudaMalloc(&s->d_runs, 8)) return 1;
    size_t need1 = 0, need2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need1, keys, s->k_sorted, vals, s->v_sorted, (int)n, 0, key_bits, stream);
    if (is_int) cub::DeviceReduce::ReduceByKey(nullptr, need2, s->k_sorted, s->k_unique, s->v_sorted, s->v_sum, s->d_runs, AddBits<long long>(), (int)n, stream);
    else cub::DeviceReduce::ReduceByKey(nullptr, need2, s->k_sorted, s->k_unique, s->v_sorted, s->v_sum, s->d_runs, AddBits<double>(), (int)n, stream);
    size_t need = need1 > need2 ? need1 : need2;
    if (need > s->tmp_bytes) {
        cudaFree(s->tmp); s->tmp = nullptr; s->tmp_bytes = 0;
        if (cudaMalloc(&s->tmp, need + need / 4)) return 1;
        s->tmp_bytes = need + need / 4;
    }
    size_t tb = s->tmp_bytes;
    if (cub::DeviceRadixSort::SortPairs(s->tmp, tb, keys, s->k_sorted, vals, s->v_sorted, (int)n, 0, key_bits, stream) != cudaSuccess) return 1;
    tb = s->tmp_bytes;
    cudaError_t e;
    if (is_int) e = cub::DeviceReduce::ReduceByKey(s->tmp, tb, s->k_sorted, s->k_unique, s->v_sorted, s->v_sum, s->d_runs, AddBits<long long>(), (int)n, stream);
    else e = cub::DeviceReduce::ReduceByKey(s->tmp, tb, s->k_sorted, s->k_unique, s->v_sorted, s->v_sum, s->d_runs, AddBits<double>(), (int)n, stream);
    if (e != cudaSuccess) return 1;
    if (is_int) compact_nonzero_kernel<long long><<<148 * 8, 256, 0, stream>>>(s->k_unique, s->v_sum, s->d_runs, out_keys, out_vals, out_cap, d_cursor);
    else compact_nonzero_kernel<double><<<148 * 8, 256, 0, stream>>>(s->k_unique, s->v_sum, s->d_runs, out_keys, out_vals, out_cap, d_cursor);
    return cudaGetLastError() != cudaSuccess;
}
extern "C" void rimu_sort_scratch_free(SortScratch *s) {
    cudaFree(s->tmp); cudaFree(s->k_sorted); cudaFree(s->v_sorted); cudaFree(s->k_unique); cudaFree(s->v_sum); cudaFree(s->d_runs);
    memset(s, 0, sizeof(*s));
}
