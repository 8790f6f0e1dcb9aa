// common.cuh -- shared device/host primitives: Philox4x32-10, address hashing, bit utilities.
//
// RNG design (replaces Julia's task-local Xoshiro256++, pmc_simulation.jl:100-103): every random
// draw is a pure function of (step key, address hash, attempt index, stream id), so a step's
// result does not depend on launch geometry or on how determinant space is partitioned over GPUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;
typedef unsigned __int128 u128;

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#else
#define HD inline
#define DEV inline
#endif

HD u32 mulhi32(u32 a, u32 b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * (u64)b) >> 32);
#endif
}

HD void philox4x32_10(const u32 ctr[4], u32 k0, u32 k1, u32 out[4]) {
    u32 c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
#pragma unroll
    for (int r = 0; r < 10; r++) {
        u32 hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        u32 hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        u32 n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

HD u64 fmix64(u64 h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    return h;
}

template <int W> HD u64 addr_hash(const u64 *w) {
    u64 h = 0x9E3779B97F4A7C15ULL;
#pragma unroll
    for (int j = 0; j < W; j++) h = fmix64(h ^ w[j]);
#ifdef __CUDA_ARCH__
    // nvcc 12.9 (sm_100a, -O3) miscompiles `(u32)(h >> 32) * m` when it fuses it into the final 64-bit
    // multiply of fmix64 (scratch/t_hash.cu reproduces it: the fused form drops the cross terms of the
    // product).  Materialising h in a register pair keeps the hash and its consumers separate.
    asm volatile("" : "+l"(h));
#endif
    return h;
}

// owner rank: fastrange of the high 32 hash bits (reference: fastrange_hash pdvec.jl:6-9,
// target rank communicators.jl:77-81).  The table slot uses the LOW bits, so the two are independent.
HD int addr_owner(u64 h, int nranks) { return (int)(((h >> 32) * (u64)nranks) >> 32); }

enum { STREAM_SPAWN = 0, STREAM_DIAG = 1, STREAM_COMPRESS = 2 };

HD double u53(u32 a, u32 b) { return (double)(((u64)a << 21) ^ ((u64)b >> 11)) * (1.0 / 9007199254740992.0); }

HD void rng_draw(u64 h, u64 k, int stream, u32 k0, u32 k1, u32 out[4]) {
    u32 ctr[4] = {(u32)h, (u32)(h >> 32), (u32)k, ((u32)stream << 28) | (u32)((k >> 32) & 0x0fffffffu)};
    philox4x32_10(ctr, k0, k1, out);
}

HD u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    u64 z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// ---------------------------------------------------------------- bit utilities (device)
// RIMU_HOST_EMULATION: tests/cuda/host_ham.cpp compiles this section and hamiltonians.cuh with g++ (the CUDA intrinsics are
// supplied by that file) so that the address arithmetic of the kernels can be checked against the oracle without a GPU.
#if defined(__CUDACC__) || defined(RIMU_HOST_EMULATION)
template <int W> struct BitsT;
template <> struct BitsT<1> { typedef u64 type; };
template <> struct BitsT<2> { typedef u128 type; };

DEV int popc_(u64 x) { return __popcll(x); }
DEV int popc_(u128 x) { return __popcll((u64)x) + __popcll((u64)(x >> 64)); }
DEV int ctz_(u64 x) { return __clzll((long long)__brevll(x)); } // 64 for x == 0
DEV int ctz_(u128 x) {
    u64 lo = (u64)x;
    if (lo) return __ffsll((long long)lo) - 1;
    u64 hi = (u64)(x >> 64);
    return hi ? 64 + __ffsll((long long)hi) - 1 : 128;
}
template <class B> DEV int cto_(B x) { return ctz_((B)~x); }
template <class B> DEV B lowmask(int p) { return (((B)1) << p) - (B)1; }

// position of the k-th (0-based) set bit within a byte: SELECT8.v[byte * 8 + k]
struct Select8Table { unsigned char v[2048]; };
constexpr Select8Table make_select8() {
    Select8Table t{};
    for (int b = 0; b < 256; b++)
        for (int k = 0; k < 8; k++) {
            int c = 0, pos = 0;
            for (int q = 0; q < 8; q++)
                if ((b >> q) & 1) { if (c == k) pos = q; c++; }
            t.v[b * 8 + k] = (unsigned char)pos;
        }
    return t;
}
#ifdef __CUDACC__
__device__ const Select8Table SELECT8 = make_select8();
#else
static const Select8Table SELECT8 = make_select8();
#endif

// position of the k-th (0-based) set bit of a 32-bit word; requires popc(x) > k.
// Branch-free, one POPC: per-byte counts by SWAR arithmetic, their prefix sums by one multiply, the byte that holds the
// answer by a byte-parallel compare, the position inside the byte from the table.  (The binary descent this replaces
// cost 5 POPCs on the quarter-rate XU pipe and, in its two-word form, ran both sides of a divergent branch: 29 % of the
// spawn kernel's instructions -- profiles/r2_kernels_ncu.md.)
DEV int select32(u32 x, int k) {
    u32 s = x - ((x >> 1) & 0x55555555u);
    s = (s & 0x33333333u) + ((s >> 2) & 0x33333333u);
    s = (s + (s >> 4)) & 0x0f0f0f0fu;
    const u32 ps = s * 0x01010101u;                                                     // byte j: set bits in bytes 0..j
    const u32 ge = ((ps | 0x80808080u) - (u32)(k + 1) * 0x01010101u) & 0x80808080u;     // flag in byte j iff ps_j > k
    const int sh = (4 - __popc(ge)) * 8;                                                // bytes with ps_j <= k come first
    const u32 before = ((ps << 8) >> sh) & 0xffu;                                       // set bits below that byte
    return sh + (int)SELECT8.v[((x >> sh) & 0xffu) * 8u + ((u32)k - before)];
}
DEV int select_(u64 v, int k) {
    const u32 lo = (u32)v, hi = (u32)(v >> 32);
    const int c = __popc(lo);
    const bool up = k >= c;
    return (up ? 32 : 0) + select32(up ? hi : lo, up ? k - c : k);
}
DEV int select_(u128 v, int k) {
    const u64 lo = (u64)v, hi = (u64)(v >> 64);
    const int c = __popcll(lo);
    const bool up = k >= c;
    return (up ? 64 : 0) + select_(up ? hi : lo, up ? k - c : k);
}

template <int W> DEV typename BitsT<W>::type load_key(const u64 *p);
template <> DEV u64 load_key<1>(const u64 *p) { return *p; }
template <> DEV u128 load_key<2>(const u64 *p) {
    ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
    return ((u128)v.y << 64) | (u128)v.x;
}
template <int W> DEV void store_key(u64 *p, typename BitsT<W>::type x);
template <> DEV void store_key<1>(u64 *p, u64 x) { *p = x; }
template <> DEV void store_key<2>(u64 *p, u128 x) {
    *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2((u64)x, (u64)(x >> 64));
}
DEV u64 hash_bits(u64 x) { u64 w[1] = {x}; return addr_hash<1>(w); }
DEV u64 hash_bits(u128 x) { u64 w[2] = {(u64)x, (u64)(x >> 64)}; return addr_hash<2>(w); }
#endif
