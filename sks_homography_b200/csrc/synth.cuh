// Device-side synthetic inputs and the minimal-sample gather.
//
// The generator is counter based: value = mix(key(seed) ^ (counter*G + lane)),
// so any quadruple can be regenerated anywhere (other GPU, CPU oracle) from
// (seed, index) alone and multi-GiB inputs never cross PCIe.  Every floating
// point step is a single strictly-rounded operation, which makes the output
// bit-identical to oracle_synth_quads_* (oracle/sks_oracle.c).
// Distributions: SURVEY.md 8(d); the "deep" family mirrors the reference's
// getInput/getTar (PY.py:9-21).
#pragma once
#include <cstdint>

#include "solvers.cuh"

namespace sksb {

constexpr uint64_t kGolden = 0x9E3779B97F4A7C15ULL;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t seed_key(uint64_t seed) { return mix64(seed + kGolden); }
__device__ __forceinline__ uint64_t rng_u64(uint64_t key, uint64_t ctr, uint32_t lane)
{
    return mix64(key ^ (ctr * kGolden + (uint64_t)lane));
}
__device__ __forceinline__ uint32_t bounded(uint64_t r, uint32_t m)
{
    return (uint32_t)(((r >> 32) * (uint64_t)m) >> 32);
}
template <typename T>
__device__ __forceinline__ T u01(uint64_t r);
template <>
__device__ __forceinline__ float u01<float>(uint64_t r)
{
    return __fmul_rn(__uint2float_rn((uint32_t)(r >> 40)), 0x1p-24f);
}
template <>
__device__ __forceinline__ double u01<double>(uint64_t r)
{
    return __dmul_rn(__ull2double_rn(r >> 11), 0x1p-53);
}

// One synthetic quadruple (source s[8], target t[8]) for index q.
template <typename T>
__device__ __forceinline__ void synth_quad(uint64_t key, uint64_t q, int dist, T (&s)[8], T (&t)[8])
{
    using S = Strict<T>;
    if (dist == 1) {   // "image": jittered 256-px square in a 1024x768 frame
        const S bx = S(T(128)) + S(T(512)) * S(u01<T>(rng_u64(key, q, 0)));
        const S by = S(T(128)) + S(T(256)) * S(u01<T>(rng_u64(key, q, 1)));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const S px = bx + S(T(256 * (k & 1))), py = by + S(T(256 * (k >> 1)));
            const S jx = (S(u01<T>(rng_u64(key, q, 2 + 2 * k))) - S(T(0.5))) * S(T(128));
            const S jy = (S(u01<T>(rng_u64(key, q, 3 + 2 * k))) - S(T(0.5))) * S(T(128));
            const S sx = px + jx, sy = py + jy;
            const S ox = (S(u01<T>(rng_u64(key, q, 10 + 2 * k))) - S(T(0.5))) * S(T(64));
            const S oy = (S(u01<T>(rng_u64(key, q, 11 + 2 * k))) - S(T(0.5))) * S(T(64));
            s[2 * k] = sx.v;
            s[2 * k + 1] = sy.v;
            t[2 * k] = (sx + ox).v;
            t[2 * k + 1] = (sy + oy).v;
        }
    } else {   // "deep" (0: continuous offsets, 2: integer offsets)
        const T mx = T(10 + bounded(rng_u64(key, q, 0), 20));
        const T my = T(10 + bounded(rng_u64(key, q, 1), 20));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const S sx = S(mx) + S(T(128 * (k & 1))), sy = S(my) + S(T(128 * (k >> 1)));
            const uint64_t rx = rng_u64(key, q, 2 + 2 * k), ry = rng_u64(key, q, 3 + 2 * k);
            const S ox = dist == 2 ? S(T(bounded(rx, 32))) : S(T(32)) * S(u01<T>(rx));
            const S oy = dist == 2 ? S(T(bounded(ry, 32))) : S(T(32)) * S(u01<T>(ry));
            s[2 * k] = sx.v;
            s[2 * k + 1] = sy.v;
            t[2 * k] = (sx + ox).v;
            t[2 * k + 1] = (sy + oy).v;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_synth_quads(T* __restrict__ src, T* __restrict__ tar, int64_t begin, int64_t count, uint64_t key,
              int dist, int layout, int64_t ld)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count)
        return;
    T s[8], t[8];
    synth_quad<T>(key, (uint64_t)(begin + j), dist, s, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t o = layout == 0 ? j * 8 + k : k * ld + j;
        src[o] = s[k];
        tar[o] = t[k];
    }
}

// RANSAC scene: corr[pair][pt] = (x,y,X,Y).  Ground truth per pair = the ACA
// homography of a "deep" quadruple; inliers follow it with uniform +-noise,
// outliers are uniform in the frame.
__global__ void __launch_bounds__(256)
k_synth_corr(float4* __restrict__ corr, int64_t pair_begin, int64_t n_pairs, int32_t n_pts,
             uint64_t key, int inlier_permille, float noise)
{
    using S = Strict<float>;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_pairs * (int64_t)n_pts)
        return;
    const int64_t p = g / n_pts + pair_begin;
    const uint32_t i = (uint32_t)(g % n_pts);
    float s[8], t[8], h[9];
    synth_quad<float>(key ^ 0xA5A5A5A5A5A5A5A5ULL, (uint64_t)p, 0, s, t);
    aca_solve<float>(s, t, h, true);
    const uint64_t ctr = ((uint64_t)p << 32) | i;
    const bool inl = (int)bounded(rng_u64(key, ctr, 0), 1000) < inlier_permille;
    const S x = S(10.f) + S(148.f) * S(u01<float>(rng_u64(key, ctr, 1)));
    const S y = S(10.f) + S(148.f) * S(u01<float>(rng_u64(key, ctr, 2)));
    const float a = u01<float>(rng_u64(key, ctr, 3)), b = u01<float>(rng_u64(key, ctr, 4));
    S X, Y;
    if (inl) {
        const S w = (S(h[6]) * x + S(h[7]) * y) + S(h[8]);
        X = ((S(h[0]) * x + S(h[1]) * y) + S(h[2])) / w + S(noise) * (S(2.f) * S(a) - S(1.f));
        Y = ((S(h[3]) * x + S(h[4]) * y) + S(h[5])) / w + S(noise) * (S(2.f) * S(b) - S(1.f));
    } else {
        X = S(10.f) + S(180.f) * S(a);
        Y = S(10.f) + S(180.f) * S(b);
    }
    corr[g] = make_float4(x.v, y.v, X.v, Y.v);
}

// Minimal-sample gather, the reference's get_rand_list (GPU.cu:52-78): four
// pool indices r_k % pool_size per hypothesis (repeats allowed, no distinct
// check), r_k from a [4][n] uint32 list (the reference's cuRAND buffer layout)
// or from the counter RNG.
template <typename T>
__global__ void __launch_bounds__(256)
k_gather_samples(const T* __restrict__ pool, uint32_t pool_size, const uint32_t* __restrict__ rand4,
                 uint64_t key, T* __restrict__ src, T* __restrict__ tar, int64_t n, int layout,
                 int64_t ld)
{
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n)
        return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t r = rand4 != nullptr ? rand4[(int64_t)k * n + id]
                                            : (uint32_t)(rng_u64(key, (uint64_t)id, k) >> 32);
        const T* c = pool + 4 * (size_t)(r % pool_size);
        const T x = c[0], y = c[1], X = c[2], Y = c[3];
        if (layout == 0) {
            src[id * 8 + 2 * k] = x;
            src[id * 8 + 2 * k + 1] = y;
            tar[id * 8 + 2 * k] = X;
            tar[id * 8 + 2 * k + 1] = Y;
        } else {
            src[(2 * k) * ld + id] = x;
            src[(2 * k + 1) * ld + id] = y;
            tar[(2 * k) * ld + id] = X;
            tar[(2 * k + 1) * ld + id] = Y;
        }
    }
}

}  // namespace sksb
