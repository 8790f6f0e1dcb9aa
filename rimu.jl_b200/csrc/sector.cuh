// sector.cuh -- dense indexing of a complete particle-number sector (BASELINE config 3: matrix-free deterministic H*v for a
// Lanczos sweep over the whole 4x4 half-filled Hubbard sector, 165 636 900 determinants).
//
// Why a second vector layout: once a Krylov vector fills the whole sector, a dictionary is the wrong container -- every
// address is present, and the push formulation (spawn a record per off-diagonal, annihilate) moves 64 records of 16 bytes
// per determinant through HBM (106 GB per H*v; the record streams do not even fit, so round 1 fell back to the HBM hash
// table: 555 ms per product).  With the addresses of the sector numbered by their combinadic rank the vector is a plain
// array of doubles and y = H x is a GATHER: y[i] = H_ii x[i] + sum_k H_{i,c_k} x[rank(c_k)] over the off-diagonals c_k of
// address i (H real symmetric).  No atomics, no hashing, no records; consecutive threads hold consecutive ranks of the last
// component, so their neighbours' values are consecutive too.  The reference has the same split: dictionary vectors for
// FCIQMC, a basis-ordered dense vector for exact diagonalisation (ExactDiagonalization/basis_set_representation.jl:35-60);
// here the matrix is never stored -- the off-diagonals come from the same device functions the FCIQMC step uses.
//
// Rank of a component with set bits at positions p_1 < ... < p_n (0-based): sum_j C(p_j, j), j = 1..n (colexicographic
// order).  Evaluated byte-wise from a table: T[chunk][byte][ones below the chunk].  Index = rank(comp 0) * dim(comp 1) +
// rank(comp 1) for two-component addresses.
#pragma once
#include "hamiltonians.cuh"

struct SectorDev {
    int ncomp;          // 1 (BoseFS / FermiFS bit string) or 2 (two fermion components)
    int bits[2];        // positions per component (BoseFS: N + M - 1; FermiFS: M)
    int ones[2];        // set bits per component (particles)
    int shift[2];       // first bit of the component inside the key
    int nchunk[2];      // byte chunks per component
    u64 dim[2];         // C(bits, ones)
    const u64 *tab[2];  // [nchunk][256][ones + 1]
};

#if defined(__CUDACC__)
DEV u64 sector_comp_rank(const SectorDev &s, int c, u64 compbits) {
    u64 r = 0;
    int below = 0;
    const int stride = s.ones[c] + 1;
    for (int ch = 0; ch < s.nchunk[c]; ch++) {
        const u32 b = (u32)(compbits >> (8 * ch)) & 0xffu;
        r += s.tab[c][(ch * 256 + (int)b) * stride + below];
        below += __popc(b);
    }
    return r;
}
DEV u64 sector_rank(const SectorDev &s, u64 key) {
    const u64 m0 = s.bits[0] >= 64 ? ~0ull : ((1ull << s.bits[0]) - 1ull);
    u64 r = sector_comp_rank(s, 0, (key >> s.shift[0]) & m0);
    if (s.ncomp == 2) {
        const u64 m1 = (1ull << s.bits[1]) - 1ull;
        r = r * s.dim[1] + sector_comp_rank(s, 1, (key >> s.shift[1]) & m1);
    }
    return r;
}

// y = H x on the complete sector: one thread per address (row)
template <int HK>
__global__ void __launch_bounds__(256)
sector_mul_kernel(const HamDev h, const SectorDev s, const u64 *__restrict__ keys, const double *__restrict__ x,
                  double *__restrict__ y, u64 dim) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    const u64 key = keys[i];
    double acc = ham_diagonal<HK, u64>(h, key) * x[i];
    const long long L = ham_num_offdiagonals<HK, u64>(h, key);
    for (long long k = 0; k < L; k++) {
        u64 child;
        const double m = ham_offdiagonal<HK, u64>(h, key, k, child);
        if (m != 0.0) acc += m * x[sector_rank(s, child)];
    }
    y[i] = acc;
}
#endif
