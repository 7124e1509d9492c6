// Per-quadruple closed-form solvers (device functions).  One thread solves one
// correspondence quadruple entirely in registers: 16 (or 8) coordinates in,
// 9 homography entries out.  The operation ORDER is the reference's, because
// parity is bit-exact (see strict.cuh); the code is organised around the
// geometry (2-D cross products, planar rotations, row products) rather than
// the reference's flat scalar listing.
//
//   aca_solve      <- sks::runKernel_ACA[_double]   MOD/ACA_SKS.cpp:24-102, :104-179
//                     cal_Homo_ACA (no normalise)   GPU.cu:81-151
//   sks_solve      <- sks::runKernel_SKS[_double]   MOD/ACA_SKS.cpp:189-303, :305-418
//                     cal_Homo_SKS (no normalise)   GPU.cu:153-240
//   aca_rect_solve <- ACA_rect                      ML/ACA_rect.m:25-36
//                     TensorACA_rect                PY.py:296-302
//   ge_solve       <- cv::runKernel_GE (competitor) MOD/GE.cpp:44-188
//                     cal_Homo_GE (fp64)            GPU.cu:359-507
//   gpt_solve      <- cal_Homo_GPT (competitor)     GPU.cu:242-357
#pragma once
#include "strict.cuh"

namespace sksb {

template <typename T>
struct Vec2 {
    Strict<T> x, y;
};

// a.x*b.y - a.y*b.x, both products rounded before the subtraction
template <typename T>
__device__ __forceinline__ Strict<T> cross(Vec2<T> a, Vec2<T> b)
{
    return a.x * b.y - a.y * b.x;
}

// h[0..7] *= 1/h[8]; h[8] = 1   (MOD/ACA_SKS.cpp:94-98): one IEEE division,
// eight multiplies, h33 forced to one even when the reciprocal is not finite.
template <typename T>
__device__ __forceinline__ void scale_by_h33(T (&h)[9])
{
    const Strict<T> inv = Strict<T>(T(1)) / Strict<T>(h[8]);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        h[k] = (Strict<T>(h[k]) * inv).v;
    h[8] = T(1);
}

// Degeneracy flag, SURVEY.md A.3: the reference signals a singular quadruple
// only through non-finite output.
template <typename T>
__device__ __forceinline__ bool is_degenerate(const T (&h)[9], bool normalized)
{
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        ok = ok && finite_val<T>(h[k]);
    if (!normalized)
        ok = ok && finite_val<T>(h[8]) && h[8] != T(0);
    return !ok;
}

// ---------------------------------------------------------------------- ACA
// Affine frame of one plane: u = N-M, v = P-M, q = Q-M; f = u x v is twice the
// triangle area and g = (q x v, u x q) the affine coordinates of Q scaled by f.
template <typename T>
struct AffineFrame {
    Vec2<T> u, v;
    Strict<T> f, gx, gy;
};

template <typename T>
__device__ __forceinline__ AffineFrame<T> affine_frame(const T (&p)[8])
{
    using S = Strict<T>;
    AffineFrame<T> F;
    F.u = { S(p[2]) - S(p[0]), S(p[3]) - S(p[1]) };
    F.v = { S(p[4]) - S(p[0]), S(p[5]) - S(p[1]) };
    const Vec2<T> q = { S(p[6]) - S(p[0]), S(p[7]) - S(p[1]) };
    F.f = cross(F.u, F.v);
    F.gx = cross(q, F.v);
    F.gy = cross(F.u, q);
    return F;
}

template <typename T>
__device__ __forceinline__ void aca_solve(const T (&s)[8], const T (&t)[8], T (&h)[9],
                                          bool normalize)
{
    using S = Strict<T>;
    const AffineFrame<T> A = affine_frame<T>(s);   // source plane
    const AffineFrame<T> B = affine_frame<T>(t);   // target plane

    // core transformation: diag(c1,c2,c3) with last row (c1-c3, c2-c3, c3)
    const S k1 = (A.f - A.gx) - A.gy;
    const S c1 = (A.gy * B.gx) * k1;
    const S c2 = (A.gx * B.gy) * k1;
    const S c3 = (A.gx * A.gy) * ((B.f - B.gx) - B.gy);

    // rows (a_r, b_r, c_r) of H_A2^-1 * H_C
    const S mx = S(t[0]) * c3, my = S(t[1]) * c3;
    const S a[3] = { S(t[2]) * c1 - mx, S(t[3]) * c1 - my, c1 - c3 };
    const S b[3] = { S(t[4]) * c2 - mx, S(t[5]) * c2 - my, c2 - c3 };
    const S c[3] = { mx, my, c3 };

    // times H_A1 = [[v.y,-v.x,0],[-u.y,u.x,0],[0,0,f]] * T(-M1): the third
    // column re-uses the first two
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const S e0 = a[r] * A.v.y - b[r] * A.u.y;
        const S e1 = b[r] * A.u.x - a[r] * A.v.x;
        h[3 * r + 0] = e0.v;
        h[3 * r + 1] = e1.v;
        h[3 * r + 2] = ((c[r] * A.f - e0 * S(s[0])) - e1 * S(s[1])).v;
    }
    if (normalize)
        scale_by_h33<T>(h);
}

// ---------------------------------------------------------------------- SKS
// Similarity frame of the anchor pair (M,N): midpoint o, half vector w with
// its y component flipped, squared length f.  MOD/ACA_SKS.cpp:192-206.
template <typename T>
struct SimFrame {
    Strict<T> ox, oy, wx, wy, f;
};

template <typename T>
__device__ __forceinline__ SimFrame<T> sim_frame(const T (&p)[8])
{
    using S = Strict<T>;
    SimFrame<T> F;
    F.ox = S(T(0.5)) * (S(p[0]) + S(p[2]));
    F.oy = S(T(0.5)) * (S(p[1]) + S(p[3]));
    F.wx = F.ox - S(p[0]);
    F.wy = S(p[1]) - F.oy;
    F.f = F.wx * F.wx + F.wy * F.wy;
    return F;
}

// P and Q carried through similarity, elementary and translation steps
// (MOD/ACA_SKS.cpp:217-232 source plane, :238-253 target plane).
template <typename T>
struct KernelPts {
    Strict<T> px, py, qx, qy, qf;
};

template <typename T>
__device__ __forceinline__ KernelPts<T> carry_points(const SimFrame<T>& F, const T (&p)[8])
{
    using S = Strict<T>;
    KernelPts<T> K;
    const S ax = S(p[4]) - F.ox, ay = S(p[5]) - F.oy;
    const S p3x = F.wx * ax - F.wy * ay;
    const S p3y = F.wy * ax + F.wx * ay;
    const S inv = S(T(1)) / p3y;
    K.px = inv * p3x;
    K.py = inv * F.f;
    const S bx = S(p[6]) - F.ox, by = S(p[7]) - F.oy;
    const S q3x = F.wx * bx - F.wy * by;
    const S q3y = F.wy * bx + F.wx * by;
    K.qx = p3y * q3x - p3x * q3y;
    K.qy = (p3y - q3y) * F.f;
    K.qf = p3y * q3y;
    return K;
}

template <typename T>
__device__ __forceinline__ void sks_solve(const T (&s)[8], const T (&t)[8], T (&h)[9],
                                          bool normalize)
{
    using S = Strict<T>;
    const SimFrame<T> F1 = sim_frame<T>(s);
    const SimFrame<T> F2 = sim_frame<T>(t);
    const KernelPts<T> A = carry_points<T>(F1, s);
    const KernelPts<T> B = carry_points<T>(F2, t);

    // hyperbolic similarity parameters (MOD/ACA_SKS.cpp:263-270)
    const S n1 = A.qx * B.qx - A.qy * B.qy;
    const S n2 = A.qx * B.qy - A.qy * B.qx;
    S d = A.qx * A.qx - A.qy * A.qy;
    d = A.qf / (d * B.qf);
    const S aK = n1 * d, bK = n2 * d;
    const S uK = (B.px - aK * A.px) - bK * A.py;
    const S vK = (B.py - aK * A.py) - bK * A.px;

    // H_L = H_S2^-1 * H_K (:276-278)
    const S L[9] = {
        bK * F2.ox + aK * F2.wx, (F2.wy + F2.ox * vK) + uK * F2.wx, aK * F2.ox + bK * F2.wx,
        bK * F2.oy - aK * F2.wy, (F2.wx + F2.oy * vK) - uK * F2.wy, aK * F2.oy - bK * F2.wy,
        bK, vK, aK
    };
    // last column of H_S1 (:281-282)
    const S s13 = F1.wy * F1.oy - F1.wx * F1.ox;
    const S s23 = (-F1.wy) * F1.ox - F1.wx * F1.oy;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const S a = L[3 * r], b = L[3 * r + 1], c = L[3 * r + 2];
        h[3 * r + 0] = (a * F1.wx + b * F1.wy).v;
        h[3 * r + 1] = (b * F1.wx - a * F1.wy).v;
        h[3 * r + 2] = ((c * F1.f + a * s13) + b * s23).v;
    }
    if (normalize)
        scale_by_h33<T>(h);
}

// ----------------------------------------------------------------- ACA-rect
// Source = axis-aligned rectangle, top-left (mx,my), width, ratio = w/h.
template <typename T>
__device__ __forceinline__ void aca_rect_solve(const T (&t)[8], T mx_, T my_, T width_,
                                               T ratio_, T (&h)[9], bool normalize)
{
    using S = Strict<T>;
    const S mx(mx_), my(my_), width(width_), ratio(ratio_);
    const S dx1 = S(t[2]) - S(t[0]), dx2 = S(t[4]) - S(t[0]), dx3 = S(t[6]) - S(t[0]);
    const S dy1 = S(t[3]) - S(t[1]), dy2 = S(t[5]) - S(t[1]), dy3 = S(t[7]) - S(t[1]);
    // c = cross(d_y, d_x), y row first (ML/ACA_rect.m:26)
    const S c1 = dy2 * dx3 - dy3 * dx2;
    const S c2 = dy3 * dx1 - dy1 * dx3;
    const S c3 = dy1 * dx2 - dy2 * dx1;
    const S sc = (c1 + c2) + c3;
    const S one(T(1));
    const S b[3] = { sc * S(t[0]), sc * S(t[1]), sc * one };
    const S n[3] = { S(t[2]), S(t[3]), one };
    const S p[3] = { S(t[4]), S(t[5]), one };
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const S e0 = n[r] * c1 - b[r];
        const S e1 = ratio * (p[r] * c2 - b[r]);
        h[3 * r + 0] = e0.v;
        h[3 * r + 1] = e1.v;
        h[3 * r + 2] = ((width * b[r] - mx * e0) - my * e1).v;
    }
    if (normalize) {   // H ./ H(3,3): element-wise division (ML/ACA_rect.m:36)
        const S den(h[8]);
#pragma unroll
        for (int k = 0; k < 9; ++k)
            h[k] = (S(h[k]) / den).v;
    }
}

// ----------------------------------------------------------------------- GE
// RHO-GE, the competitor the paper measures against (Bilaniuk et al. 2014; 221
// flops, h33 fixed to 1 by construction, no pivoting: an axis-aligned source
// square makes its first pivot x0 - x2 zero and the result non-finite, exactly
// as in the reference).  Column elimination on
//   A[2][4]: source coordinates with point 2 as origin (A[.][2] = p2 itself)
//   B[3][8]: column j = X-equation of point j, column 4+j = its Y-equation,
//            rows (x*T, y*T, T) relative to point 2, the first two negated.
// Pivot order, the two reciprocals and the output permutation: MOD/GE.cpp:75-186.
template <typename T>
__device__ __forceinline__ void ge_combine(Strict<T> (&B)[3][8], int a, int b, Strict<T> s1,
                                           Strict<T> s2)
{
#pragma unroll
    for (int half = 0; half < 8; half += 4)
#pragma unroll
        for (int r = 0; r < 3; ++r)
            B[r][half + a] = B[r][half + a] * s1 - B[r][half + b] * s2;
}

template <typename T>
__device__ __forceinline__ void ge_solve(const T (&s)[8], const T (&t)[8], T (&h)[9])
{
    using S = Strict<T>;
    S A[2][4], B[3][8];
#pragma unroll
    for (int k = 0; k < 2; ++k) {                 // k = 0: X-equations, 1: Y-equations
        const S T2 = S(t[4 + k]);
        const S xT2 = S(s[4]) * T2, yT2 = S(s[5]) * T2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j == 2) {
                B[0][4 * k + j] = -xT2;
                B[1][4 * k + j] = -yT2;
                B[2][4 * k + j] = T2;
            } else {
                const S Tj = S(t[2 * j + k]);
                B[0][4 * k + j] = xT2 - S(s[2 * j]) * Tj;
                B[1][4 * k + j] = yT2 - S(s[2 * j + 1]) * Tj;
                B[2][4 * k + j] = Tj - T2;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        A[0][j] = (j == 2) ? S(s[4]) : S(s[2 * j]) - S(s[4]);
        A[1][j] = (j == 2) ? S(s[5]) : S(s[2 * j + 1]) - S(s[5]);
    }
    // x out of points 1 and 3 (pivot: point 0), then y out of points 3 and 0 (pivot: point 1)
    S s1 = A[0][0], s2 = A[0][1];
    A[1][1] = A[1][1] * s1 - A[1][0] * s2;
    ge_combine<T>(B, 1, 0, s1, s2);
    s2 = A[0][3];
    A[1][3] = A[1][3] * s1 - A[1][0] * s2;
    ge_combine<T>(B, 3, 0, s1, s2);
    s1 = A[1][1];
    s2 = A[1][3];
    ge_combine<T>(B, 3, 1, s1, s2);
    s2 = A[1][0];
    A[0][0] = A[0][0] * s1;
    ge_combine<T>(B, 0, 1, s1, s2);
    // unit pivots
    s1 = S(T(1)) / A[0][0];
    s2 = S(T(1)) / A[1][1];
#pragma unroll
    for (int half = 0; half < 8; half += 4)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            B[r][half] = B[r][half] * s1;
            B[r][half + 1] = B[r][half + 1] * s2;
        }
    // origin column
    s1 = A[0][2];
    s2 = A[1][2];
#pragma unroll
    for (int half = 0; half < 8; half += 4)
#pragma unroll
        for (int r = 0; r < 3; ++r)
            B[r][half + 2] = B[r][half + 2] - (B[r][half] * s1 + B[r][half + 1] * s2);
    // back-substitution of the two hollowed-out rows (columns 7, then 3)
    s1 = B[0][7];
    B[1][7] = B[1][7] / s1;
    B[2][7] = B[2][7] / s1;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const S f = B[0][c];
        B[1][c] = B[1][c] - f * B[1][7];
        B[2][c] = B[2][c] - f * B[2][7];
    }
    s1 = B[1][3];
    B[2][3] = B[2][3] / s1;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (c != 3)
            B[2][c] = B[2][c] - B[1][c] * B[2][3];
    h[0] = B[2][0].v; h[1] = B[2][1].v; h[2] = B[2][2].v;
    h[3] = B[2][4].v; h[4] = B[2][5].v; h[5] = B[2][6].v;
    h[6] = B[2][7].v; h[7] = B[2][3].v; h[8] = T(1);
}

// A x = b for an 8x8 system by LU with partial pivoting, in place; x is left in b.  The
// operation order is that of the reference's cal_Homo_GPT (GPU.cu:242-292, :345-355).
template <typename T>
__device__ __forceinline__ void lu8_solve(Strict<T> (&A)[8][8], Strict<T> (&b)[8])
{
    using S = Strict<T>;
    // The forward substitution L y = b is carried along as a ninth column: the reference
    // computes y_k = (((b_k - L_k0 y_0) - L_k1 y_1) - ...) / L_kk after the factorisation
    // (GPU.cu:283-292); subtracting L_ri * y_i from b_r at step i performs the same operations
    // in the same order, and a row's partial sum travels with the row through later swaps.
    // Columns left of the diagonal are then dead, so each swap only touches columns >= i.
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // partial pivoting: first row with the largest |A[r][i]|, r >= i (NaN never wins)
        int piv = i;
        T best = fabs(A[i][i].v);
#pragma unroll
        for (int r = i + 1; r < 8; ++r) {
            const T v = fabs(A[r][i].v);
            if (best < v) {
                piv = r;
                best = v;
            }
        }
#pragma unroll
        for (int r = i + 1; r < 8; ++r) {          // branch-free exchange with the chosen row
            const bool sw = (r == piv);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j >= i) {                      // constant after unrolling: dead columns drop out
                    const T ai = A[i][j].v, ar = A[r][j].v;
                    A[i][j] = S(sw ? ar : ai);
                    A[r][j] = S(sw ? ai : ar);
                }
            const T bi = b[i].v, br = b[r].v;
            b[i] = S(sw ? br : bi);
            b[r] = S(sw ? bi : br);
        }
        b[i] = b[i] / A[i][i];
#pragma unroll
        for (int j = i + 1; j < 8; ++j)
            A[i][j] = A[i][j] / A[i][i];
#pragma unroll
        for (int r = i + 1; r < 8; ++r) {
#pragma unroll
            for (int j = i + 1; j < 8; ++j)
                A[r][j] = A[r][j] - A[r][i] * A[i][j];
            b[r] = b[r] - A[r][i] * b[i];
        }
    }
#pragma unroll
    for (int k = 6; k >= 0; --k) {         // U x = y (unit diagonal)
        S acc = b[k];
#pragma unroll
        for (int j = 7; j > k; --j)
            acc = acc - A[k][j] * b[j];
        b[k] = acc;
    }
}

// ---------------------------------------------------------------------- GPT
// GPT-LU, the other competitor the paper times on the GPU: the 8x8 DLT system A h = b
// (h33 = 1) by LU with partial pivoting, in the arithmetic of the reference's kernel
// cal_Homo_GPT (GPU.cu:242-357).  The reference keeps A in a 64-double local array and
// indexes it dynamically; here every index is a compile-time constant after unrolling, so A
// and b live in registers and the data-dependent row swap is a chain of predicated
// exchanges with the candidate rows below the diagonal.
template <typename T>
__device__ __forceinline__ void gpt_solve(const T (&s)[8], const T (&t)[8], T (&h)[9])
{
    using S = Strict<T>;
    S A[8][8], b[8];
    const S zero(T(0)), one(T(1));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const S x(s[2 * i]), y(s[2 * i + 1]), X(t[2 * i]), Y(t[2 * i + 1]);
        A[i][0] = x; A[i][1] = y; A[i][2] = one; A[i][3] = zero; A[i][4] = zero; A[i][5] = zero;
        A[i][6] = (-x) * X; A[i][7] = (-y) * X;
        A[i + 4][0] = zero; A[i + 4][1] = zero; A[i + 4][2] = zero; A[i + 4][3] = x; A[i + 4][4] = y;
        A[i + 4][5] = one; A[i + 4][6] = (-x) * Y; A[i + 4][7] = (-y) * Y;
        b[i] = X;
        b[i + 4] = Y;
    }
    lu8_solve<T>(A, b);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        h[k] = b[k].v;
    h[8] = T(1);
}

}  // namespace sksb
