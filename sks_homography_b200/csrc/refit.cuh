// Post-RANSAC consumer: least-squares re-estimation of each pair's
// homography from ALL its inliers (the 4-point winner only used four of them).  Nothing like
// it exists in the reference (parity unpinned); the arithmetic is this project's definition,
// mirrored operation for operation -- including the reduction order -- by
// oracle_ransac_refit_f32:
//   one warp per image pair, fp64 throughout;
//   pass 1: centroids of the inlier sources / targets;  pass 2: mean distance to them
//           -> Hartley scales s = sqrt(2) / mean distance;
//   pass 3: normal equations N = sum(aX aX^T + aY aY^T), g = sum(aX U + aY V) of the DLT rows
//           aX = (u, v, 1, 0, 0, 0, -uU, -vU), aY = (0, 0, 0, u, v, 1, -uV, -vV) in normalised
//           coordinates (h33 = 1);
//   lane partial sums over matches lane, lane+32, ... in index order, then the xor-shuffle
//   tree 16, 8, 4, 2, 1 (every lane ends with the same bits);
//   8x8 solve by the LU of solvers.cuh, de-normalisation H = T2^-1 Hn T1, division by h33.
// Pairs with fewer than 4 inliers or a non-finite result keep the input model H_in.
#pragma once
#include <cstdint>

#include "solvers.cuh"

namespace sksb {

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

__global__ void __launch_bounds__(128)
k_ransac_refit(const float4* __restrict__ corr, int64_t n_pairs, int32_t n_pts,
               const uint8_t* __restrict__ mask, const float* __restrict__ H_in,
               float* __restrict__ H_out, uint32_t* __restrict__ n_used)
{
    using S = Strict<double>;
    const int lane = threadIdx.x & 31;
    const int64_t pair = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= n_pairs)
        return;
    const float4* c = corr + (size_t)pair * n_pts;
    const uint8_t* m = mask + (size_t)pair * n_pts;

    // pass 1: count and centroids
    double sx = 0, sy = 0, sX = 0, sY = 0, cnt = 0;
    for (int i = lane; i < n_pts; i += 32)
        if (m[i]) {
            const float4 p = __ldg(c + i);
            sx = __dadd_rn(sx, (double)p.x); sy = __dadd_rn(sy, (double)p.y);
            sX = __dadd_rn(sX, (double)p.z); sY = __dadd_rn(sY, (double)p.w);
            cnt = __dadd_rn(cnt, 1.0);
        }
    sx = warp_sum(sx); sy = warp_sum(sy); sX = warp_sum(sX); sY = warp_sum(sY); cnt = warp_sum(cnt);
    const bool enough = cnt >= 4.0;
    const double cx = __ddiv_rn(sx, cnt), cy = __ddiv_rn(sy, cnt);
    const double cX = __ddiv_rn(sX, cnt), cY = __ddiv_rn(sY, cnt);

    // pass 2: mean distance to the centroid -> isotropic scales
    double d1 = 0, d2 = 0;
    for (int i = lane; i < n_pts; i += 32)
        if (m[i]) {
            const float4 p = __ldg(c + i);
            const double ax = __dsub_rn((double)p.x, cx), ay = __dsub_rn((double)p.y, cy);
            const double bx = __dsub_rn((double)p.z, cX), by = __dsub_rn((double)p.w, cY);
            d1 = __dadd_rn(d1, __dsqrt_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay))));
            d2 = __dadd_rn(d2, __dsqrt_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by))));
        }
    d1 = warp_sum(d1); d2 = warp_sum(d2);
    const double r2 = 1.4142135623730951;
    const double s1 = __ddiv_rn(__dmul_rn(r2, cnt), d1), s2 = __ddiv_rn(__dmul_rn(r2, cnt), d2);

    // pass 3: normal equations, upper triangle (36) + right-hand side (8)
    double N[36], g[8];
#pragma unroll
    for (int k = 0; k < 36; ++k) N[k] = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = 0;
    for (int i = lane; i < n_pts; i += 32)
        if (m[i]) {
            const float4 p = __ldg(c + i);
            const double u = __dmul_rn(__dsub_rn((double)p.x, cx), s1), v = __dmul_rn(__dsub_rn((double)p.y, cy), s1);
            const double U = __dmul_rn(__dsub_rn((double)p.z, cX), s2), V = __dmul_rn(__dsub_rn((double)p.w, cY), s2);
            const double aX[8] = { u, v, 1.0, 0.0, 0.0, 0.0, __dmul_rn(-u, U), __dmul_rn(-v, U) };
            const double aY[8] = { 0.0, 0.0, 0.0, u, v, 1.0, __dmul_rn(-u, V), __dmul_rn(-v, V) };
            int k = 0;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
#pragma unroll
                for (int b = a; b < 8; ++b, ++k)
                    N[k] = __dadd_rn(N[k], __dadd_rn(__dmul_rn(aX[a], aX[b]), __dmul_rn(aY[a], aY[b])));
                g[a] = __dadd_rn(g[a], __dadd_rn(__dmul_rn(aX[a], U), __dmul_rn(aY[a], V)));
            }
        }
    S A[8][8], b[8];
    {
        int k = 0;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
#pragma unroll
            for (int bb = a; bb < 8; ++bb, ++k) {
                const double t = warp_sum(N[k]);
                A[a][bb] = S(t);
                A[bb][a] = S(t);
            }
            b[a] = S(warp_sum(g[a]));
        }
    }
    lu8_solve<double>(A, b);      // hn (h33 = 1) in normalised coordinates, left in b

    // H = T2^-1 * Hn * T1 with T1 = [s1 0 -s1 cx; 0 s1 -s1 cy; 0 0 1], T2^-1 = [1/s2 0 cX; 0 1/s2 cY; 0 0 1]
    const double hn[9] = { b[0].v, b[1].v, b[2].v, b[3].v, b[4].v, b[5].v, b[6].v, b[7].v, 1.0 };
    double M[9];   // Hn * T1
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        M[3 * r] = __dmul_rn(hn[3 * r], s1);
        M[3 * r + 1] = __dmul_rn(hn[3 * r + 1], s1);
        M[3 * r + 2] = __dsub_rn(__dsub_rn(hn[3 * r + 2], __dmul_rn(M[3 * r], cx)), __dmul_rn(M[3 * r + 1], cy));
    }
    const double is2 = __ddiv_rn(1.0, s2);
    double Hd[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        Hd[k] = __dadd_rn(__dmul_rn(M[k], is2), __dmul_rn(cX, M[6 + k]));
        Hd[3 + k] = __dadd_rn(__dmul_rn(M[3 + k], is2), __dmul_rn(cY, M[6 + k]));
        Hd[6 + k] = M[6 + k];
    }
    bool ok = enough;
    float out[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        out[k] = (float)__ddiv_rn(Hd[k], Hd[8]);
        ok = ok && finite_val<float>(out[k]);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k)
            H_out[pair * 9 + k] = ok ? out[k] : H_in[pair * 9 + k];
        if (n_used != nullptr)
            n_used[pair] = ok ? (uint32_t)cnt : 0u;
    }
}

}  // namespace sksb
