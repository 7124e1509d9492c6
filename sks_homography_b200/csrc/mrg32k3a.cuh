// cuRAND-compatible MRG32k3a stream, hand-written.
//
// The reference draws its minimal-sample list with cuRAND's HOST API
// (GPU.cu:1443-1446: curandCreateGenerator(CURAND_RNG_PSEUDO_MRG32K3A), seed 11,
// curandGenerate(gen, p_d, 4*numsOfH)) and feeds it to get_rand_list (GPU.cu:52-78).
// cuRAND (libcurand 10.3, CUDA 12.9) is a closed third-party library that is not
// part of /root/reference, so this file restates its PUBLISHED algorithm (L'Ecuyer's
// MRG32k3a as laid out in CUDA's public header curand_kernel.h: seeding :1276-1292,
// one step :1061-1150, 32-bit output :1161-1166) plus the host API's output ORDER,
// which was measured on a B200 against libcurand itself (tools/curand_dump.py):
//     out[n] = draw number floor(n / 81920) of subsequence (n mod 81920),
//     subsequence s starts 2^76 * s steps into the seeded stream.
// tests/test_gpu_parity.py compares this kernel with the live library on the GPU box,
// tests/golden/curand_mrg32k3a.npz holds library output for three seeds.
//
// One thread owns one subsequence: it jumps to its start with a square-and-multiply
// over precomputed powers of the 2^76-step transition matrices (host-computed once,
// passed by value), then emits its draws at stride 81920, so every warp store is
// coalesced.  A sample list is a few MB once per batch: nothing here is hot.
#pragma once
#include <cstdint>

namespace sksb {

constexpr uint32_t kMrgM1 = 4294967087u, kMrgM2 = 4294944443u;
constexpr uint32_t kMrgA12 = 1403580u, kMrgA13n = 810728u, kMrgA21 = 527612u, kMrgA23n = 1370589u;
constexpr int kMrgStreams = 81920;     // cuRAND host API: subsequences interleaved in the output
constexpr int kMrgPowBits = 17;        // 2^17 > 81920

struct MrgJumpTable {
    // pw[c][j] = (A_c ^ (2^76)) ^ (2^j), row-major 3x3, component c = 0 (mod m1), 1 (mod m2)
    uint32_t pw[2][kMrgPowBits][9];
};

__host__ __device__ inline void mrg_matvec(const uint32_t (&A)[9], uint32_t (&v)[3], uint32_t m)
{
    uint32_t r[3];
    for (int i = 0; i < 3; ++i) {
        unsigned long long acc = 0;
        for (int k = 0; k < 3; ++k)
            acc = (acc + (unsigned long long)A[3 * i + k] * v[k] % m) % m;
        r[i] = (uint32_t)acc;
    }
    v[0] = r[0]; v[1] = r[1]; v[2] = r[2];
}

__host__ __device__ inline void mrg_matmul(const uint32_t (&A)[9], const uint32_t (&B)[9],
                                           uint32_t (&Cm)[9], uint32_t m)
{
    uint32_t r[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            unsigned long long acc = 0;
            for (int k = 0; k < 3; ++k)
                acc = (acc + (unsigned long long)A[3 * i + k] * B[3 * k + j] % m) % m;
            r[3 * i + j] = (uint32_t)acc;
        }
    for (int i = 0; i < 9; ++i) Cm[i] = r[i];
}

// seeded state (curand_init with subsequence 0, offset 0)
__host__ __device__ inline void mrg_seed(uint64_t seed, uint32_t (&s1)[3], uint32_t (&s2)[3])
{
    s1[0] = s1[1] = s1[2] = 12345u;
    s2[0] = s2[1] = s2[2] = 12345u;
    if (seed != 0) {
        const uint32_t x1 = (uint32_t)seed ^ 0x55555555u;
        const uint32_t x2 = (uint32_t)(seed >> 32) ^ 0xAAAAAAAAu;
        s1[0] = (uint32_t)((unsigned long long)x1 * 12345u % kMrgM1);
        s1[1] = (uint32_t)((unsigned long long)x2 * 12345u % kMrgM1);
        s1[2] = s1[0];
        s2[0] = (uint32_t)((unsigned long long)x2 * 12345u % kMrgM2);
        s2[1] = (uint32_t)((unsigned long long)x1 * 12345u % kMrgM2);
        s2[2] = s2[0];
    }
}

// one step; returns z in [1, m1]
__host__ __device__ inline uint32_t mrg_step(uint32_t (&s1)[3], uint32_t (&s2)[3])
{
    const unsigned long long p1 =
        ((unsigned long long)kMrgA12 * s1[1] + (unsigned long long)kMrgA13n * (kMrgM1 - s1[0])) % kMrgM1;
    const unsigned long long p2 =
        ((unsigned long long)kMrgA21 * s2[2] + (unsigned long long)kMrgA23n * (kMrgM2 - s2[0])) % kMrgM2;
    s1[0] = s1[1]; s1[1] = s1[2]; s1[2] = (uint32_t)p1;
    s2[0] = s2[1]; s2[1] = s2[2]; s2[2] = (uint32_t)p2;
    const uint32_t d = (uint32_t)p1 - (uint32_t)p2;
    return p1 <= p2 ? d + kMrgM1 : d;
}

// curand(): (unsigned int)(z * 1.000000048662) in double; z = m1 (p1 == p2) lands just
// above 2^32, where the device conversion saturates
__host__ __device__ inline uint32_t mrg_bits(uint32_t z)
{
    const double d = (double)z * 1.000000048662;
    return d >= 4294967296.0 ? 0xFFFFFFFFu : (uint32_t)d;
}

// host: the jump table.  A1 = [[0,1,0],[0,0,1],[-a13n,a12,0]], A2 = [[0,1,0],[0,0,1],[-a23n,0,a21]]
inline void mrg_build_jump_table(MrgJumpTable& t)
{
    uint32_t A[2][9] = { { 0, 1, 0, 0, 0, 1, kMrgM1 - kMrgA13n, kMrgA12, 0 },
                         { 0, 1, 0, 0, 0, 1, kMrgM2 - kMrgA23n, 0, kMrgA21 } };
    const uint32_t mod[2] = { kMrgM1, kMrgM2 };
    for (int c = 0; c < 2; ++c) {
        for (int s = 0; s < 76; ++s)               // A^(2^76)
            mrg_matmul(A[c], A[c], A[c], mod[c]);
        for (int j = 0; j < kMrgPowBits; ++j) {
            for (int i = 0; i < 9; ++i) t.pw[c][j][i] = A[c][i];
            mrg_matmul(A[c], A[c], A[c], mod[c]);
        }
    }
}

__global__ void __launch_bounds__(256)
k_mrg32k3a(uint32_t* __restrict__ out, int64_t n, uint64_t seed, const MrgJumpTable tab)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= kMrgStreams || t >= n)
        return;
    uint32_t s1[3], s2[3];
    mrg_seed(seed, s1, s2);
    for (int j = 0; j < kMrgPowBits; ++j)
        if ((t >> j) & 1) {
            mrg_matvec(tab.pw[0][j], s1, kMrgM1);
            mrg_matvec(tab.pw[1][j], s2, kMrgM2);
        }
    for (int64_t i = t; i < n; i += kMrgStreams)
        out[i] = mrg_bits(mrg_step(s1, s2));
}

}  // namespace sksb
