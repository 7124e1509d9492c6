/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the batched SKS / ACA / ACA-rect
 * path and the ACA-RANSAC scorer.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (libsks_cuda.so) never links, loads or calls it.
 *
 * Parity status:
 *   SKS / ACA (fp32, fp64): PINNED -- bit-compared against the reference's own
 *     C++ (MOD/ACA_SKS.cpp compiled in place into oracle/_ref/libsks_ref.so by
 *     oracle/Makefile) in tests/test_oracle.py, against golden vectors produced
 *     by that library (tests/golden/), and against the two veri_4Pts.m
 *     known-answer cases.
 *   ACA-rect: PINNED to ML/ACA_rect.m via the veri_4Pts.m rectangle KAT (fp64)
 *     and to golden vectors obtained by executing
 *     "PyTorch Codes/Modules_Runtime_Test.py":296-302 on torch-CPU
 *     (tests/golden/make_golden.py).  No C++ version exists in the reference.
 *   RHO-GE (competitor solver): fp32 PINNED -- bit-compared
 *     against the reference's MOD/GE.cpp compiled in place (same _ref library);
 *     fp64 is the same type-generic body (the reference's only fp64 GE is its
 *     FMA-contracted CUDA kernel GPU.cu:359-507; recompiled with -fmad=false into
 *     oracle/_ref/libsks_refgpu.so it is bit-compared on the GPU box).
 *   GPT-LU (competitor, fp64): the arithmetic of the reference's CUDA kernel cal_Homo_GPT
 *     (GPU.cu:242-357); pinned on the GPU box against that kernel compiled with -fmad=false.
 *     (Its CPU form calls OpenCV's getPerspectiveTransform, which is outside the reference tree.)
 *   RANSAC scoring: PARITY UNPINNED -- the reference has no inlier test or model
 *     selection (only the sampler precedent GPU.cu:52-78).  Hypotheses are the
 *     pinned ACA; the scoring rule below is this project's own definition.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
 */
#include "sks_oracle.h"

#include <math.h>
#include <string.h>

#define REAL float
#define SUF(x) x##_f32
#include "solver_body.inc"
#undef REAL
#undef SUF

#define REAL double
#define SUF(x) x##_f64
#include "solver_body.inc"
#undef REAL
#undef SUF

/* ------------------------------------------------------------------ batches */
/* AoS batches: src[n][8], tar[n][8] -> H[n][9]  (the CPU reference's layout,
 * MOD/ACA_SKS.cpp:24; the caller loop is CPU/main.cpp:87-114). */

#define DEF_BATCH(NAME, T, ONE)                                                          \
    void NAME(const T *src, const T *tar, T *H, int64_t n, int normalize)               \
    {                                                                                    \
        for (int64_t i = 0; i < n; ++i)                                                  \
            ONE(src + 8 * i, tar + 8 * i, H + 9 * i, normalize);                         \
    }

DEF_BATCH(oracle_aca_f32, float, oracle_aca_one_f32)
DEF_BATCH(oracle_aca_f64, double, oracle_aca_one_f64)
DEF_BATCH(oracle_sks_f32, float, oracle_sks_one_f32)
DEF_BATCH(oracle_sks_f64, double, oracle_sks_one_f64)
DEF_BATCH(oracle_ge_f32, float, oracle_ge_one_f32)
DEF_BATCH(oracle_ge_f64, double, oracle_ge_one_f64)
DEF_BATCH(oracle_gpt_f32, float, oracle_gpt_one_f32)
DEF_BATCH(oracle_gpt_f64, double, oracle_gpt_one_f64)

/* M == NULL: one shared rectangle corner (mx,my); else per-sample M[n][2]
 * (PyTorch Codes/Modules_Runtime_Test.py:302 reads M per sample while
 * width/ratio come from sample 0, :33-35). */
#define DEF_RECT(NAME, T, ONE)                                                           \
    void NAME(const T *tar, const T *M, T mx, T my, T width, T ratio, T *H, int64_t n,   \
              int normalize)                                                             \
    {                                                                                    \
        for (int64_t i = 0; i < n; ++i) {                                                \
            T ax = M ? M[2 * i] : mx, ay = M ? M[2 * i + 1] : my;                        \
            ONE(tar + 8 * i, ax, ay, width, ratio, H + 9 * i, normalize);                \
        }                                                                                \
    }

DEF_RECT(oracle_aca_rect_f32, float, oracle_aca_rect_one_f32)
DEF_RECT(oracle_aca_rect_f64, double, oracle_aca_rect_one_f64)

/* Degeneracy flag: the reference never reports failure, it
 * emits non-finite entries.  normalised: any of h[0..7] non-finite;
 * up-to-scale: any of h[0..8] non-finite or h[8] == 0. */
#define DEF_FLAGS(NAME, T)                                                               \
    void NAME(const T *H, uint8_t *flag, int64_t n, int normalized)                      \
    {                                                                                    \
        for (int64_t i = 0; i < n; ++i) {                                                \
            const T *h = H + 9 * i;                                                      \
            int bad = 0;                                                                 \
            for (int k = 0; k < 8; ++k)                                                  \
                bad |= !isfinite(h[k]);                                                  \
            if (!normalized)                                                             \
                bad |= !isfinite(h[8]) || h[8] == (T)0;                                  \
            flag[i] = (uint8_t)bad;                                                      \
        }                                                                                \
    }

DEF_FLAGS(oracle_degenerate_f32, float)
DEF_FLAGS(oracle_degenerate_f64, double)

/* ------------------------------------------------- synthetic quad generator */
/* Counter-based generator shared (as a specification) with the device
 * generator in sks_homography_b200/csrc/synth.cu: value = mix(key(seed) ^
 * (quad_index * G + lane)).  Distributions follow SURVEY.md 8(d):
 *   dist 0  "deep"    : M = (U{10..29}, U{10..29}), 128-px source square in
 *                       order TL,TR,BL,BR, target = source + 32*U[0,1)
 *                       (PyTorch Codes/Modules_Runtime_Test.py:9-21, continuous)
 *   dist 1  "image"   : 256-px square jittered by +-64 px inside a 1024x768
 *                       frame, target = source +- 32 px (ML/veri_4Pts.m:35-36)
 *   dist 2  "deep-int": dist 0 with integer offsets U{0..31} (exactly the
 *                       reference's torch.randint generator). */

static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

uint64_t oracle_rng_u64(uint64_t seed, uint64_t ctr, uint32_t lane)
{
    const uint64_t key = mix64(seed + 0x9E3779B97F4A7C15ULL);
    return mix64(key ^ (ctr * 0x9E3779B97F4A7C15ULL + (uint64_t)lane));
}

static inline uint32_t bounded(uint64_t r, uint32_t m)
{
    return (uint32_t)(((r >> 32) * (uint64_t)m) >> 32);
}

static inline float u01_f32(uint64_t r) { return (float)(r >> 40) * 0x1p-24f; }
static inline double u01_f64(uint64_t r) { return (double)(r >> 11) * 0x1p-53; }

#define DEF_SYNTH(NAME, T, U01)                                                          \
    void NAME(T *src, T *tar, int64_t begin, int64_t count, uint64_t seed, int dist)     \
    {                                                                                    \
        static const int cx[4] = { 0, 1, 0, 1 }, cy[4] = { 0, 0, 1, 1 };                 \
        for (int64_t j = 0; j < count; ++j) {                                            \
            const uint64_t q = (uint64_t)(begin + j);                                    \
            T *s = src + 8 * j, *t = tar + 8 * j;                                        \
            if (dist == 1) {                                                             \
                const T bx = (T)128 + (T)512 * U01(oracle_rng_u64(seed, q, 0));          \
                const T by = (T)128 + (T)256 * U01(oracle_rng_u64(seed, q, 1));          \
                for (int k = 0; k < 4; ++k) {                                            \
                    const T px = bx + (T)(256 * cx[k]), py = by + (T)(256 * cy[k]);      \
                    const T jx = (U01(oracle_rng_u64(seed, q, 2 + 2 * k)) - (T)0.5) * (T)128; \
                    const T jy = (U01(oracle_rng_u64(seed, q, 3 + 2 * k)) - (T)0.5) * (T)128; \
                    s[2 * k] = px + jx;                                                  \
                    s[2 * k + 1] = py + jy;                                              \
                    const T ox = (U01(oracle_rng_u64(seed, q, 10 + 2 * k)) - (T)0.5) * (T)64; \
                    const T oy = (U01(oracle_rng_u64(seed, q, 11 + 2 * k)) - (T)0.5) * (T)64; \
                    t[2 * k] = s[2 * k] + ox;                                            \
                    t[2 * k + 1] = s[2 * k + 1] + oy;                                    \
                }                                                                        \
            } else {                                                                     \
                const T mx = (T)(10 + bounded(oracle_rng_u64(seed, q, 0), 20));          \
                const T my = (T)(10 + bounded(oracle_rng_u64(seed, q, 1), 20));          \
                for (int k = 0; k < 4; ++k) {                                            \
                    s[2 * k] = mx + (T)(128 * cx[k]);                                    \
                    s[2 * k + 1] = my + (T)(128 * cy[k]);                                \
                    const uint64_t rx = oracle_rng_u64(seed, q, 2 + 2 * k);              \
                    const uint64_t ry = oracle_rng_u64(seed, q, 3 + 2 * k);              \
                    const T ox = dist == 2 ? (T)bounded(rx, 32) : (T)32 * U01(rx);       \
                    const T oy = dist == 2 ? (T)bounded(ry, 32) : (T)32 * U01(ry);       \
                    t[2 * k] = s[2 * k] + ox;                                            \
                    t[2 * k + 1] = s[2 * k + 1] + oy;                                    \
                }                                                                        \
            }                                                                            \
        }                                                                                \
    }

DEF_SYNTH(oracle_synth_quads_f32, float, u01_f32)
DEF_SYNTH(oracle_synth_quads_f64, double, u01_f64)

/* ------------------------------------------------------- post-RANSAC refit */
/* Our own definition (parity unpinned); mirrors csrc/refit.cuh operation for operation,
 * including the reduction order: 32 "lanes" take matches lane, lane+32, ... in index order and
 * are combined by the xor tree 16, 8, 4, 2, 1 (addition is commutative, so every lane of the
 * device warp ends with these bits). */
static double refit_tree(double v[32])
{
    for (int off = 16; off > 0; off >>= 1) {
        double w[32];
        for (int l = 0; l < 32; ++l)
            w[l] = v[l] + v[l ^ off];
        memcpy(v, w, sizeof w);
    }
    return v[0];
}

void oracle_ransac_refit_f32(const float *corr, int64_t n_pairs, int32_t n_pts, const uint8_t *mask,
                             const float *H_in, float *H_out, uint32_t *n_used)
{
    for (int64_t p = 0; p < n_pairs; ++p) {
        const float *c = corr + 4 * (size_t)n_pts * (size_t)p;
        const uint8_t *m = mask + (size_t)n_pts * (size_t)p;
        double l0[5][32];
        memset(l0, 0, sizeof l0);
        for (int l = 0; l < 32; ++l)
            for (int i = l; i < n_pts; i += 32)
                if (m[i]) {
                    for (int k = 0; k < 4; ++k)
                        l0[k][l] = l0[k][l] + (double)c[4 * i + k];
                    l0[4][l] = l0[4][l] + 1.0;
                }
        const double sx = refit_tree(l0[0]), sy = refit_tree(l0[1]), sX = refit_tree(l0[2]),
                     sY = refit_tree(l0[3]), cnt = refit_tree(l0[4]);
        const double cx = sx / cnt, cy = sy / cnt, cX = sX / cnt, cY = sY / cnt;
        double l1[2][32];
        memset(l1, 0, sizeof l1);
        for (int l = 0; l < 32; ++l)
            for (int i = l; i < n_pts; i += 32)
                if (m[i]) {
                    const double ax = (double)c[4 * i] - cx, ay = (double)c[4 * i + 1] - cy;
                    const double bx = (double)c[4 * i + 2] - cX, by = (double)c[4 * i + 3] - cY;
                    l1[0][l] = l1[0][l] + sqrt(ax * ax + ay * ay);
                    l1[1][l] = l1[1][l] + sqrt(bx * bx + by * by);
                }
        const double d1 = refit_tree(l1[0]), d2 = refit_tree(l1[1]);
        const double r2 = 1.4142135623730951;
        const double s1 = (r2 * cnt) / d1, s2 = (r2 * cnt) / d2;
        static double N[36][32], g[8][32];      /* not re-entrant: test infrastructure */
        memset(N, 0, sizeof N);
        memset(g, 0, sizeof g);
        for (int l = 0; l < 32; ++l)
            for (int i = l; i < n_pts; i += 32)
                if (m[i]) {
                    const double u = ((double)c[4 * i] - cx) * s1, v = ((double)c[4 * i + 1] - cy) * s1;
                    const double U = ((double)c[4 * i + 2] - cX) * s2, V = ((double)c[4 * i + 3] - cY) * s2;
                    const double aX[8] = { u, v, 1.0, 0.0, 0.0, 0.0, -u * U, -v * U };
                    const double aY[8] = { 0.0, 0.0, 0.0, u, v, 1.0, -u * V, -v * V };
                    int k = 0;
                    for (int a = 0; a < 8; ++a) {
                        for (int b = a; b < 8; ++b, ++k)
                            N[k][l] = N[k][l] + (aX[a] * aX[b] + aY[a] * aY[b]);
                        g[a][l] = g[a][l] + (aX[a] * U + aY[a] * V);
                    }
                }
        double A[8][8], b[8];
        {
            int k = 0;
            for (int a = 0; a < 8; ++a) {
                for (int bb = a; bb < 8; ++bb, ++k) {
                    const double t = refit_tree(N[k]);
                    A[a][bb] = t;
                    A[bb][a] = t;
                }
                b[a] = refit_tree(g[a]);
            }
        }
        oracle_lu8_solve_f64(A, b);
        const double hn[9] = { b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], 1.0 };
        double M[9], Hd[9];
        for (int r = 0; r < 3; ++r) {
            M[3 * r] = hn[3 * r] * s1;
            M[3 * r + 1] = hn[3 * r + 1] * s1;
            M[3 * r + 2] = (hn[3 * r + 2] - M[3 * r] * cx) - M[3 * r + 1] * cy;
        }
        const double is2 = 1.0 / s2;
        for (int k = 0; k < 3; ++k) {
            Hd[k] = M[k] * is2 + cX * M[6 + k];
            Hd[3 + k] = M[3 + k] * is2 + cY * M[6 + k];
            Hd[6 + k] = M[6 + k];
        }
        int ok = cnt >= 4.0;
        float out[9];
        for (int k = 0; k < 9; ++k) {
            out[k] = (float)(Hd[k] / Hd[8]);
            ok = ok && isfinite(out[k]);
        }
        for (int k = 0; k < 9; ++k)
            H_out[9 * p + k] = ok ? out[k] : H_in[9 * p + k];
        if (n_used)
            n_used[p] = ok ? (uint32_t)cnt : 0u;
    }
}

/* ------------------------------------------------------------ warp grid */
/* Our own definition (parity unpinned: the reference only remarks that warping
 * needs no normalisation, ML/ACA_rect.m:33-35); mirrors csrc/warp.cuh. */
void oracle_warp_grid_f32(const float *H, int64_t n, float x0, float y0, float dx, float dy,
                          int32_t gw, int32_t gh, float *out)
{
    for (int64_t s = 0; s < n; ++s) {
        const float *h = H + 9 * s;
        for (int32_t j = 0; j < gh; ++j)
            for (int32_t i = 0; i < gw; ++i) {
                const float x = fmaf((float)i, dx, x0), y = fmaf((float)j, dy, y0);
                const float u = fmaf(h[0], x, fmaf(h[1], y, h[2]));
                const float v = fmaf(h[3], x, fmaf(h[4], y, h[5]));
                const float w = fmaf(h[6], x, fmaf(h[7], y, h[8]));
                float *o = out + 2 * (((int64_t)s * gh + j) * gw + i);
                const float r = 1.0f / w;
                o[0] = u * r;
                o[1] = v * r;
            }
    }
}

/* ------------------------------------------------- cuRAND MRG32K3A sample list */
/* The reference fills its sample list with cuRAND's host API (GPU.cu:1443-1446:
 * CURAND_RNG_PSEUDO_MRG32K3A, seed 11, curandGenerate).  cuRAND is a third-party
 * library outside /root/reference (libcurand 10.3.10, CUDA 12.9); this restates its
 * published algorithm -- L'Ecuyer's MRG32k3a as in CUDA's public curand_kernel.h
 * (seeding :1276-1292, step :1061-1150, output scaling :1161-1166) -- and the host
 * API's output order measured against the library on a B200 (tools/curand_dump.py):
 * out[n] = draw floor(n/81920) of subsequence n mod 81920, subsequences 2^76 steps
 * apart.  Pinned by tests/golden/curand_mrg32k3a.npz (library output, 3 seeds). */
#define MRG_M1 4294967087ULL
#define MRG_M2 4294944443ULL

static void mrg_mat_mul(uint64_t C[9], const uint64_t A[9], const uint64_t B[9], uint64_t m)
{
    uint64_t r[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            uint64_t acc = 0;
            for (int k = 0; k < 3; ++k)
                acc = (acc + A[3 * i + k] * B[3 * k + j] % m) % m;
            r[3 * i + j] = acc;
        }
    memcpy(C, r, sizeof r);
}

static void mrg_mat_vec(const uint64_t A[9], uint64_t v[3], uint64_t m)
{
    uint64_t r[3];
    for (int i = 0; i < 3; ++i)
        r[i] = ((A[3 * i] * v[0] % m + A[3 * i + 1] * v[1] % m) % m + A[3 * i + 2] * v[2] % m) % m;
    memcpy(v, r, sizeof r);
}

void oracle_curand_mrg32k3a_u32(uint32_t *out, int64_t n, uint64_t seed)
{
    const int64_t T = 81920;
    uint64_t J1[9] = { 0, 1, 0, 0, 0, 1, MRG_M1 - 810728ULL, 1403580ULL, 0 };
    uint64_t J2[9] = { 0, 1, 0, 0, 0, 1, MRG_M2 - 1370589ULL, 0, 527612ULL };
    for (int s = 0; s < 76; ++s) {               /* one subsequence = 2^76 steps */
        mrg_mat_mul(J1, J1, J1, MRG_M1);
        mrg_mat_mul(J2, J2, J2, MRG_M2);
    }
    uint64_t s1[3] = { 12345, 12345, 12345 }, s2[3] = { 12345, 12345, 12345 };
    if (seed != 0) {
        const uint64_t x1 = (uint32_t)seed ^ 0x55555555u, x2 = (uint32_t)(seed >> 32) ^ 0xAAAAAAAAu;
        s1[0] = s1[2] = x1 * 12345 % MRG_M1;
        s1[1] = x2 * 12345 % MRG_M1;
        s2[0] = s2[2] = x2 * 12345 % MRG_M2;
        s2[1] = x1 * 12345 % MRG_M2;
    }
    for (int64_t t = 0; t < T && t < n; ++t) {
        uint64_t a[3] = { s1[0], s1[1], s1[2] }, b[3] = { s2[0], s2[1], s2[2] };
        for (int64_t i = t; i < n; i += T) {
            const uint64_t p1 = (1403580ULL * a[1] + 810728ULL * (MRG_M1 - a[0])) % MRG_M1;
            const uint64_t p2 = (527612ULL * b[2] + 1370589ULL * (MRG_M2 - b[0])) % MRG_M2;
            a[0] = a[1]; a[1] = a[2]; a[2] = p1;
            b[0] = b[1]; b[1] = b[2]; b[2] = p2;
            const uint64_t z = p1 > p2 ? p1 - p2 : p1 + MRG_M1 - p2;        /* in [1, m1] */
            const double d = (double)z * 1.000000048662;
            out[i] = d >= 4294967296.0 ? 0xFFFFFFFFu : (uint32_t)d;
        }
        mrg_mat_vec(J1, s1, MRG_M1);
        mrg_mat_vec(J2, s2, MRG_M2);
    }
}

/* ------------------------------------------------------------ ACA-RANSAC */
/* Our own definition (parity unpinned, see header).
 *   sample    : 4 indices per hypothesis, idx_k = u32(seed, pair*2^32 + hyp, k)
 *               % n_pts -- "r % size" with repeats allowed, as GPU.cu:55-58;
 *               or taken from an explicit sample list [n_hyp][4].
 *   hypothesis: the pinned fp32 ACA, h33-normalised (MOD/ACA_SKS.cpp:24-102).
 *   score     : forward transfer error |proj(x) - X|^2 < thr^2, division-free
 *               (both sides multiplied by w^2) with the threshold folded into
 *               the operands once per hypothesis / match instead of once per
 *               evaluation (11 instead of 12 FP32 operations per hypothesis x
 *               match); every rounding step explicit so CPU and GPU agree bit
 *               for bit:
 *                 it = 1 / sqrtf(thr2)                       (once)
 *                 g_k = h_k * it (k = 0..5), g_6..8 = h_6..8 (once per hypothesis)
 *                 Xs = X * it, Ys = Y * it                   (once per match)
 *                 u = fma(g0,x,fma(g1,y,g2)); v = fma(g3,x,fma(g4,y,g5));
 *                 w = fma(g6,x,fma(g7,y,g8));
 *                 du = fma(-Xs,w,u); dv = fma(-Ys,w,v);
 *                 e = fma(dv,dv,du*du); acc = fma(-w,w,e);
 *                 inlier <=> acc < 0   (NaN is never an inlier)
 *   select    : key = count<<32 | (0xFFFFFFFF - hyp); best = max key
 *               (highest count, lowest hypothesis id on ties). */

uint32_t oracle_ransac_count_f32(const float *H, const float *corr, int32_t n_pts, float thr2)
{
    const float it = 1.0f / sqrtf(thr2);
    float g[9];
    for (int k = 0; k < 9; ++k)
        g[k] = k < 6 ? H[k] * it : H[k];
    uint32_t cnt = 0;
    for (int32_t i = 0; i < n_pts; ++i) {
        const float x = corr[4 * i], y = corr[4 * i + 1];
        const float Xs = corr[4 * i + 2] * it, Ys = corr[4 * i + 3] * it;
        const float u = fmaf(g[0], x, fmaf(g[1], y, g[2]));
        const float v = fmaf(g[3], x, fmaf(g[4], y, g[5]));
        const float w = fmaf(g[6], x, fmaf(g[7], y, g[8]));
        const float du = fmaf(-Xs, w, u);
        const float dv = fmaf(-Ys, w, v);
        const float e = fmaf(dv, dv, du * du);
        const float acc = fmaf(-w, w, e);
        cnt += (acc < 0.0f) ? 1u : 0u;
    }
    return cnt;
}

void oracle_ransac_sample(uint64_t seed, int64_t pair, uint32_t hyp, int32_t n_pts,
                          uint32_t idx[4])
{
    const uint64_t ctr = ((uint64_t)pair << 32) | (uint64_t)hyp;
    for (uint32_t k = 0; k < 4; ++k)
        idx[k] = (uint32_t)(oracle_rng_u64(seed, ctr, k) >> 32) % (uint32_t)n_pts;
}

void oracle_ransac_hypothesis_f32(const float *corr, const uint32_t idx[4], float *H)
{
    float s[8], t[8];
    for (int k = 0; k < 4; ++k) {
        const float *c = corr + 4 * (size_t)idx[k];
        s[2 * k] = c[0];
        s[2 * k + 1] = c[1];
        t[2 * k] = c[2];
        t[2 * k + 1] = c[3];
    }
    oracle_aca_one_f32(s, t, H, 1);
}

/* corr / samples / best_key / counts_out point at the first pair of a shard whose global pair
 * ids are [pair_begin, pair_begin + n_pairs): the id only keys the sampler. */
void oracle_ransac_aca_shard_f32(const float *corr, int64_t pair_begin, int64_t n_pairs, int32_t n_pts,
                                 const uint32_t *samples, uint32_t hyp_begin, uint32_t hyp_count,
                                 uint32_t hyp_stride, uint64_t seed, float thr2, uint64_t *best_key,
                                 uint32_t *counts_out)
{
    for (int64_t p = 0; p < n_pairs; ++p) {
        const float *c = corr + 4 * (size_t)n_pts * (size_t)p;
        uint64_t best = 0;
        for (uint32_t j = 0; j < hyp_count; ++j) {
            const uint32_t hyp = hyp_begin + j;
            uint32_t idx[4];
            if (samples) {   /* reduced modulo n_pts like the reference's get_rand_list (GPU.cu:55-58) */
                memcpy(idx, samples + 4 * ((size_t)p * hyp_stride + hyp), sizeof idx);
                for (int k = 0; k < 4; ++k)
                    idx[k] %= (uint32_t)n_pts;
            } else
                oracle_ransac_sample(seed, pair_begin + p, hyp, n_pts, idx);
            float H[9];
            oracle_ransac_hypothesis_f32(c, idx, H);
            const uint32_t cnt = oracle_ransac_count_f32(H, c, n_pts, thr2);
            if (counts_out)
                counts_out[(size_t)p * hyp_count + j] = cnt;
            const uint64_t key = ((uint64_t)cnt << 32) | (uint64_t)(0xFFFFFFFFu - hyp);
            if (key > best)
                best = key;
        }
        best_key[p] = best;
    }
}

void oracle_ransac_aca_f32(const float *corr, int64_t n_pairs, int32_t n_pts,
                           const uint32_t *samples, uint32_t hyp_begin, uint32_t hyp_count,
                           uint32_t hyp_stride, uint64_t seed, float thr2, uint64_t *best_key,
                           uint32_t *counts_out)
{
    oracle_ransac_aca_shard_f32(corr, 0, n_pairs, n_pts, samples, hyp_begin, hyp_count, hyp_stride, seed,
                                thr2, best_key, counts_out);
}
