// microbenchmark: cost of the scatter-append (bucket counter atomic + record store) at FCIQMC sizes
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
__device__ __forceinline__ u64 fmix64(u64 h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }
template <int MODE>
__global__ void k(const u64 *keys, const u64 *vals, long long n, u32 nb, u32 rcap, u32 *cnt, u64 *rk, u64 *rv, ulonglong2 *rr, u64 *sink) {
    u64 acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        u64 key = keys[i], v = vals[i];
        u64 h = fmix64(key ^ 0x9E3779B97F4A7C15ULL);
        u32 b = __umulhi((u32)(h >> 32), nb);
        if (MODE == 0) { acc += b; continue; }                       // read + hash only
        u32 pos = atomicAdd(&cnt[b], 1u);
        if (MODE == 1) { acc += pos; continue; }                     // + atomic
        if (pos < rcap) {
            u64 at = (u64)b * rcap + pos;
            if (MODE == 2) { rk[at] = key; rv[at] = v; }             // SoA: two 8-byte stores
            if (MODE == 3) { rr[at] = make_ulonglong2(key, v); }     // AoS: one 16-byte store
        }
        if (MODE == 4) { u64 at = (u64)b * rcap + (i % rcap); rr[at] = make_ulonglong2(key, v); } // no atomic, scattered AoS store
    }
    if (acc == 0x1234567) *sink = acc;
}
int main() {
    const long long n = 10000000; const u32 rcap = 2048; const u32 nb = 14000;
    u64 *keys, *vals, *rk, *rv, *sink; ulonglong2 *rr; u32 *cnt;
    cudaMalloc(&keys, n * 8); cudaMalloc(&vals, n * 8); cudaMalloc(&rk, (size_t)nb * rcap * 8); cudaMalloc(&rv, (size_t)nb * rcap * 8);
    cudaMalloc(&rr, (size_t)nb * rcap * 16); cudaMalloc(&cnt, nb * 4); cudaMalloc(&sink, 8);
    u64 *h = (u64 *)malloc(n * 8); u64 x = 88172645463325252ULL;
    for (long long i = 0; i < n; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = x; }
    cudaMemcpy(keys, h, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(vals, h, n * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[] = {"read+hash", "+atomic", "+atomic+SoA 2x8B", "+atomic+AoS 16B", "no atomic, AoS 16B"};
    for (int grid_mul : {4, 8, 16}) for (int mode = 0; mode < 5; mode++) {
        float best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            cudaMemset(cnt, 0, nb * 4);
            cudaEventRecord(e0);
            int g = 148 * grid_mul;
            switch (mode) {
                case 0: k<0><<<g, 256>>>(keys, vals, n, nb, rcap, cnt, rk, rv, rr, sink); break;
                case 1: k<1><<<g, 256>>>(keys, vals, n, nb, rcap, cnt, rk, rv, rr, sink); break;
                case 2: k<2><<<g, 256>>>(keys, vals, n, nb, rcap, cnt, rk, rv, rr, sink); break;
                case 3: k<3><<<g, 256>>>(keys, vals, n, nb, rcap, cnt, rk, rv, rr, sink); break;
                case 4: k<4><<<g, 256>>>(keys, vals, n, nb, rcap, cnt, rk, rv, rr, sink); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("grid %2dx148  %-22s %.3f ms  (%.1f ps/record)\n", grid_mul, names[mode], best, best * 1e9 / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
