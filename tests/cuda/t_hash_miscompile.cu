#include "../../rimu.jl_b200/csrc/partition.cuh"
#include <cstdio>
#include <vector>
__global__ void k1(const u64 *keys, int n, int nranks, u32 nb, u64 *hout, u32 *bout) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        u64 h = hash_bits(keys[i]);
        hout[i] = h;
        bout[i] = bucket_of(h, nranks, nb);
    }
}
__global__ void k2(const u64 *keys, int n, int nranks, u32 nb, u32 *bout) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        bout[i] = bucket_of(hash_bits(keys[i]), nranks, nb);
}
int main() {
    const int n = 16; const u32 nb = 5;
    std::vector<u64> hk(n);
    u64 s = 12345;
    for (int i = 0; i < n; i++) { s = splitmix64(s); hk[i] = s >> 2; }
    u64 *k, *ho; u32 *b1, *b2;
    cudaMalloc(&k, n * 8); cudaMalloc(&ho, n * 8); cudaMalloc(&b1, n * 4); cudaMalloc(&b2, n * 4);
    cudaMemcpy(k, hk.data(), n * 8, cudaMemcpyHostToDevice);
    k1<<<1, 64>>>(k, n, 1, nb, ho, b1);
    k2<<<1, 64>>>(k, n, 1, nb, b2);
    cudaDeviceSynchronize();
    std::vector<u64> hh(n); std::vector<u32> hb1(n), hb2(n);
    cudaMemcpy(hh.data(), ho, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hb1.data(), b1, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb2.data(), b2, n * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; i++) {
        u64 w[1] = {hk[i]}; u64 h = addr_hash<1>(w); u32 x = (u32)(h >> 32); u32 b = (u32)(((u64)x * nb) >> 32);
        printf("key %016llx host h %016llx b %u | dev h %016llx b1 %u b2 %u\n", hk[i], h, b, hh[i], hb1[i], hb2[i]);
    }
    return 0;
}
