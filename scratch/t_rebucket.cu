#include "../rimu.jl_b200/csrc/partition.cuh"
#include <cstdio>
#include <vector>
#include <algorithm>
int main() {
    const int n = 80; const u32 nb = 5;
    std::vector<u64> hk(n), hv(n);
    u64 s = 12345;
    for (int i = 0; i < n; i++) { s = splitmix64(s); hk[i] = s >> 2; hv[i] = i + 1; }
    u64 *k, *v, *ok, *ov, *seg_start; u32 *seg_len, *tmp;
    cudaMalloc(&k, n * 8); cudaMalloc(&v, n * 8); cudaMalloc(&ok, n * 8); cudaMalloc(&ov, n * 8);
    cudaMalloc(&seg_start, 64 * 8); cudaMalloc(&seg_len, 64 * 4); cudaMalloc(&tmp, 2 * 64 * 4);
    cudaMemcpy(k, hk.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(v, hv.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemset(ov, 0, n * 8);
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    u32 *counts = tmp, *fill = tmp + 64;
    cudaMemsetAsync(counts, 0, nb * 4, st);
    cudaMemsetAsync(fill, 0xff, nb * 4, st);
    bucket_count_kernel<1><<<1, RIMU_TPB, 0, st>>>(k, n, 1, nb, counts);
    bucket_scan_kernel<<<1, 1024, 0, st>>>(counts, nb, seg_start, seg_len, fill);
    bucket_scatter_kernel<1><<<1, RIMU_TPB, 0, st>>>(k, v, n, 1, nb, seg_start, fill, ok, ov);
    cudaStreamSynchronize(st);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    std::vector<u64> out(n); std::vector<u32> hf(nb), hc(nb); std::vector<u64> hs(nb);
    cudaMemcpy(out.data(), ov, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hf.data(), fill, nb * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc.data(), counts, nb * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs.data(), seg_start, nb * 8, cudaMemcpyDeviceToHost);
    { std::vector<u32> hb(nb, 0); for (int i = 0; i < n; i++) { u64 w[1] = {hk[i]}; u64 h = addr_hash<1>(w); u32 x = (u32)(h >> 32); hb[(u32)(((u64)x * nb) >> 32)]++; }
      for (u32 b = 0; b < nb; b++) printf("host b %u count %u\n", b, hb[b]); }
    for (u32 b = 0; b < nb; b++) printf("b %u count %u start %llu fill %u\n", b, hc[b], hs[b], hf[b]);
    for (int i = 0; i < n; i++) printf("%llu ", out[i]);
    printf("\n");
    std::sort(out.begin(), out.end());
    int distinct = std::unique(out.begin(), out.end()) - out.begin();
    printf("distinct %d\n", distinct);
    return 0;
}
