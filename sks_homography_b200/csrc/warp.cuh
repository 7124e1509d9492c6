// Consumer on the far side of the path: the sampling grid a
// deep-homography pipeline builds from each predicted homography before it warps the
// image.  ML/ACA_rect.m:33-35 notes that this step does not need the h33 normalisation,
// so the fused form solves ACA-rect up to scale in registers and never writes H:
//     tar[n][8] (4 predicted corners) -> grid[n][gh][gw] of (X, Y) = (u, v) / w,
//     (u,v,w) = H * (x, y, 1),  x = x0 + i*dx, y = y0 + j*dy.
// Nothing like it exists in the reference (parity unpinned); the arithmetic is this
// project's definition, mirrored by oracle_warp_grid_f32:
//     x = fma(i, dx, x0); y = fma(j, dy, y0)
//     u = fma(h0,x,fma(h1,y,h2)); v = fma(h3,x,fma(h4,y,h5)); w = fma(h6,x,fma(h7,y,h8))
//     r = 1 / w (IEEE); X = u * r; Y = v * r   (reciprocal-multiply, the reference's own
//                                               normalisation style, MOD/ACA_SKS.cpp:94-98)
// Bound: HBM writes, 8 bytes per grid point (the 32-byte corner read per sample and the
// 47-flop solve are noise next to gh*gw*8 bytes).  Every thread of a sample's group re-solves
// the sample's homography (cheaper than a shared-memory broadcast plus barrier) and emits four
// neighbouring grid points as one 32-byte store per iteration.
#pragma once
#include <cstdint>

#include "ptx.cuh"
#include "stream_kernels.cuh"

namespace sksb {

struct GridSpec {
    float x0, y0, dx, dy;
    int32_t gw, gh;
};

__device__ __forceinline__ float2 warp_point(const float (&h)[9], float x, float y)
{
    const float u = __fmaf_rn(h[0], x, __fmaf_rn(h[1], y, h[2]));
    const float v = __fmaf_rn(h[3], x, __fmaf_rn(h[4], y, h[5]));
    const float w = __fmaf_rn(h[6], x, __fmaf_rn(h[7], y, h[8]));
    const float r = __frcp_rn(w);        // one IEEE reciprocal, two multiplies (as MOD/ACA_SKS.cpp:94-98)
    return make_float2(__fmul_rn(u, r), __fmul_rn(v, r));
}

// Work split: a sample's gh*gw points are covered by a GROUP of `group` threads (a power of
// two, 1..256; `parts` > 1 CTAs per sample only for grids beyond 256 threads x 32 points), so
// that every thread emits roughly 16-32 points whatever the grid size: small grids pack many
// samples into one CTA, large grids loop.  A thread takes FOUR consecutive points per
// iteration: one 32-bit division locates the run, the row terms h1*y+h2, h4*y+h5, h7*y+h8
// are computed once per row, and the 32 bytes leave as one 256-bit store (sm_100 STG.256)
// when every sample block is 32-byte aligned (gh*gw a multiple of 4).
struct WarpSplit {
    int32_t group;        // threads per sample inside a CTA (divides 256)
    int32_t parts;        // CTAs per sample (1 unless group == 256)
    uint32_t step_i, step_j;   // (4 * group * parts) mod gw, div gw
    int32_t vec;          // 1: gh*gw is a multiple of 4 and out is 32-byte aligned -> 256-bit stores
};

struct RowTerms {
    float au, av, aw;
};

__device__ __forceinline__ RowTerms warp_row(const float (&h)[9], float y)
{
    return RowTerms{__fmaf_rn(h[1], y, h[2]), __fmaf_rn(h[4], y, h[5]), __fmaf_rn(h[7], y, h[8])};
}

// same bits as warp_point(h, x, y) with the y terms hoisted
__device__ __forceinline__ float2 warp_point_row(const float (&h)[9], const RowTerms& r, float x)
{
    const float u = __fmaf_rn(h[0], x, r.au);
    const float v = __fmaf_rn(h[3], x, r.av);
    const float w = __fmaf_rn(h[6], x, r.aw);
    const float q = __frcp_rn(w);
    return make_float2(__fmul_rn(u, q), __fmul_rn(v, q));
}

// FUSED: homographies come from ACA-rect on tar (H == nullptr), else they are read from H[n][9]
template <bool FUSED>
__global__ void __launch_bounds__(256)
k_warp_grid(const float* __restrict__ H, const float* __restrict__ tar, const float* __restrict__ M,
            RectParams<float> rp, GridSpec g, WarpSplit ws, float* __restrict__ out, int64_t n)
{
    const int tid = threadIdx.x;
    const int per_cta = 256 / ws.group;
    const int64_t s = ((int64_t)(blockIdx.x / ws.parts)) * per_cta + tid / ws.group;
    if (s >= n)
        return;
    const uint32_t lane = (uint32_t)(blockIdx.x % ws.parts) * 256u + (uint32_t)(tid % ws.group);
    const uint32_t m = (uint32_t)g.gw * (uint32_t)g.gh, gw = (uint32_t)g.gw;
    const uint32_t stride = 4u * (uint32_t)ws.group * (uint32_t)ws.parts;
    uint32_t p = 4u * lane;
    if (p >= m)
        return;
    float h[9];
    if constexpr (FUSED) {
        float t[8];
        const float4 a = __ldg(reinterpret_cast<const float4*>(tar) + 2 * s);
        const float4 b = __ldg(reinterpret_cast<const float4*>(tar) + 2 * s + 1);
        t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w;
        t[4] = b.x; t[5] = b.y; t[6] = b.z; t[7] = b.w;
        const float mx = M != nullptr ? __ldg(M + 2 * s) : rp.mx;
        const float my = M != nullptr ? __ldg(M + 2 * s + 1) : rp.my;
        aca_rect_solve<float>(t, mx, my, rp.width, rp.ratio, h, false);
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k)
            h[k] = __ldg(H + 9 * s + k);
    }
    float* o = out + 2 * (int64_t)m * s;
    const bool vec = ws.vec != 0;        // every sample block then starts 32-byte aligned
    uint32_t j = p / gw, i = p - j * gw;
    for (; p < m; p += stride) {
        RowTerms r = warp_row(h, __fmaf_rn((float)j, g.dy, g.y0));
        float2 q[4];
        uint32_t ii = i, jj = j;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            q[k] = warp_point_row(h, r, __fmaf_rn((float)ii, g.dx, g.x0));
            if (++ii == gw) {            // run crosses into the next row
                ii = 0;
                ++jj;
                r = warp_row(h, __fmaf_rn((float)jj, g.dy, g.y0));
            }
        }
        if (vec) {
            Chunk32 c;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                c.w[2 * k] = __float_as_uint(q[k].x);
                c.w[2 * k + 1] = __float_as_uint(q[k].y);
            }
            stg_stream32(o + 2 * (size_t)p, c);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (p + k < m)
                    *reinterpret_cast<float2*>(o + 2 * (size_t)(p + k)) = q[k];
        }
        i += ws.step_i;
        j += ws.step_j;
        if (i >= gw) {
            i -= gw;
            ++j;
        }
    }
}

}  // namespace sksb
