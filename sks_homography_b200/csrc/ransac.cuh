// Fused ACA-RANSAC: hypothesis generation, scoring and best-model selection in
// one kernel; no per-hypothesis homography ever touches HBM.
//
// Nothing like this exists in the reference; the sampler follows its precedent
// (r % size with repeats allowed, GPU.cu:52-78) and every hypothesis is the
// bit-exact fp32 ACA (MOD/ACA_SKS.cpp:24-102).  The scoring rule is this
// project's definition and is mirrored operation for operation by
// oracle_ransac_count_f32 (oracle/sks_oracle.c) so inlier counts are bit-exact:
//     it = 1/sqrt(thr2);  g_k = h_k*it (k<6), g_6..8 = h_6..8;  Xs = X*it, Ys = Y*it
//     u = fma(g0,x,fma(g1,y,g2));  v = fma(g3,x,fma(g4,y,g5));
//     w = fma(g6,x,fma(g7,y,g8));
//     du = fma(-Xs,w,u); dv = fma(-Ys,w,v); e = fma(dv,dv,du*du)
//     acc = fma(-w, w, e);  inlier <=> acc < 0
//   (division-free forward transfer error |proj(x) - X|^2 < thr^2, both sides
//   multiplied by (w/thr)^2; the threshold is folded into the operands once per
//   hypothesis and once per match, which leaves 11 FP32 operations per hypothesis
//   x match instead of 12).  The integer scorers (MODE 0-2) read "acc < 0" off the
//   sign bit and add it to the count with one LEA.HI: acc is never -0 (e >= +0 and
//   an exact cancellation rounds to +0) and a NaN acc is the canonical positive NaN
//   on NVIDIA hardware, so the sign bit is exactly the IEEE comparison the oracle
//   evaluates.  The default scorer (MODE 3) keeps the count on the FP32 pipe:
//   cnt = fma_rd(acc, 2^-149, cnt) steps a float counter down by exactly one ulp
//   when acc < 0 (see ransac_inlier2_fp) -- one packed FFMA2.RM per two evaluations
//   where the integer form needs two LEA.HI at two issue cycles each.
//
// Mapping: a CTA owns one image pair and a contiguous chunk of hypothesis ids.
// The pair's correspondences (x,y,X,Y) are pulled into shared memory once by
// the bulk-copy engine (64 KiB for 4096 points) and re-laid out in place as
// pairs; every thread then carries HPT hypotheses in registers and walks the
// tile with warp-uniform (broadcast) 16-byte shared loads, scoring two matches
// per instruction with sm_100a's packed FFMA2 / FMUL2, so the inner loop is pure
// FP32 pipe work: 12 packed instructions per two hypothesis x match evaluations.
// The best (count, lowest id) is reduced with warp shuffles, then one 64-bit
// atomicMax per CTA.  Bound: the FP32 pipe (ncu: 93 % busy) together with the
// register file's read bandwidth (two operands per cycle per scheduler:
// tools/ubench/count_ops.cu, a packed FMA that needs a third register from one
// bank takes three cycles instead of two), not HBM.
#pragma once
#include <cstdint>

#include "ptx.cuh"
#include "solvers.cuh"
#include "synth.cuh"

// inner-loop shape of the packed scorer (swept in tools/ransac_sweep.py)
#ifndef SKS_RANSAC_UNROLL
#define SKS_RANSAC_UNROLL 8      // match PAIRS per unrolled iteration (measured best: 8, hypothesis-major)
#endif
#ifndef SKS_RANSAC_FP_UNROLL
#define SKS_RANSAC_FP_UNROLL 2   // match pairs per unrolled iteration of the FP-count scorer (MODE 3); measured best of 1/2/3/4/6/8
                                 // (profiles/r02_ransac_fp_sweep.log)
#endif
// resident CTAs per SM the register allocation must allow: a 64 KiB tile leaves room for three
#ifndef SKS_RANSAC_MIN_CTAS
#define SKS_RANSAC_MIN_CTAS(T) ((T) <= 256 ? 3 : (T) <= 384 ? 2 : 1)
#endif
#ifndef SKS_RANSAC_FP_JOINT
#define SKS_RANSAC_FP_JOINT 0
#endif
#ifndef SKS_RANSAC_HYP_MAJOR
#define SKS_RANSAC_HYP_MAJOR 1
#endif
#ifndef SKS_RANSAC_SCALAR_RESID
#define SKS_RANSAC_SCALAR_RESID 0
#endif
#ifndef SKS_RANSAC_TAIL
#define SKS_RANSAC_TAIL 1        // last step of the packed scorer: 0 = two scalar FFMA(-w,w,e);
                                 // 1 = sign flip of w on the ALU pipe + one FFMA2; 2 = FMUL2 + compare
#endif
#ifndef SKS_RANSAC_H8_ONE
#define SKS_RANSAC_H8_ONE 0      // 1 = packed scorer uses the literal 1.0f for h33 (every hypothesis is
                                 // h33-normalised, MOD/ACA_SKS.cpp:98, so the bits are the same)
#endif

namespace sksb {

constexpr int kRansacMaxTilePts = 8192;    // 128 KiB of shared memory at most

// pair: index into this call's arrays; pair_id_base + pair: the pair's global id, which keys
// the counter RNG (so a rank that holds only a shard of the pairs draws the same samples
// as a rank that holds them all)
__device__ __forceinline__ void ransac_sample(uint64_t key, int64_t pair, int64_t pair_id_base,
                                              uint32_t hyp, const uint32_t* __restrict__ samples,
                                              uint32_t hyp_stride, uint32_t n_pts, uint32_t (&idx)[4])
{
    if (samples != nullptr) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(samples) +
                              ((size_t)pair * hyp_stride + hyp));
        // reduced like the reference's sampler (get_rand_list: r % size, GPU.cu:55-58), so a raw
        // 32-bit stream (e.g. sks_cuda_curand_mrg32k3a_u32) is a valid list; in-range lists unchanged
        idx[0] = v.x % n_pts; idx[1] = v.y % n_pts; idx[2] = v.z % n_pts; idx[3] = v.w % n_pts;
    } else {
        const uint64_t ctr = ((uint64_t)(pair + pair_id_base) << 32) | (uint64_t)hyp;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            idx[k] = (uint32_t)(rng_u64(key, ctr, k) >> 32) % n_pts;
    }
}

__device__ __forceinline__ void ransac_hypothesis(const float4* __restrict__ corr_pair,
                                                  const uint32_t (&idx)[4], float (&h)[9])
{
    float s[8], t[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 c = __ldg(corr_pair + idx[k]);
        s[2 * k] = c.x; s[2 * k + 1] = c.y;
        t[2 * k] = c.z; t[2 * k + 1] = c.w;
    }
    aca_solve<float>(s, t, h, true);
}

// 1/thr, the factor folded into hypotheses and matches (IEEE sqrt and divide)
__device__ __forceinline__ float ransac_inv_thr(float thr2)
{
    return __fdiv_rn(1.0f, __fsqrt_rn(thr2));
}

// g = h with rows 1-2 scaled by 1/thr (once per hypothesis)
__device__ __forceinline__ void ransac_scale_h(float (&h)[9], float it)
{
#pragma unroll
    for (int k = 0; k < 6; ++k)
        h[k] = __fmul_rn(h[k], it);
}

// 1 if the match c = (x, y, -Xs, -Ys) -- target already scaled by 1/thr and negated
// (an exact operation) -- is an inlier of the scaled hypothesis g, else 0 (sign bit
// of acc, see header)
__device__ __forceinline__ uint32_t ransac_inlier(const float (&g)[9], const float4 c)
{
    const float u = __fmaf_rn(g[0], c.x, __fmaf_rn(g[1], c.y, g[2]));
    const float v = __fmaf_rn(g[3], c.x, __fmaf_rn(g[4], c.y, g[5]));
    const float w = __fmaf_rn(g[6], c.x, __fmaf_rn(g[7], c.y, g[8]));
    const float du = __fmaf_rn(c.z, w, u);
    const float dv = __fmaf_rn(c.w, w, v);
    const float e = __fmaf_rn(dv, dv, __fmul_rn(du, du));
    const float acc = __fmaf_rn(-w, w, e);
    return __float_as_uint(acc) >> 31;
}

// Packed form for sm_100a's 2-wide FP32 instructions (FFMA2 / FMUL2): the same
// test on TWO matches at once.  p0 = (x0,x1,y0,y1), p1 = (-Xs0,-Xs1,-Ys0,-Ys1) is
// the pair layout the kernel rewrites the shared-memory tile into; g[k] =
// (g_k,g_k).  Bit-identical to two ransac_inlier calls: each half is an IEEE
// fused multiply-add and (-X)*it is the exact negation of X*it.  The last step
// needs a negated operand, which only the scalar FFMA has, so it is two FFMAs
// (two pipe cycles, where FMUL2 + FFMA2 took four).
__device__ __forceinline__ uint32_t ransac_inlier2(const float2 (&g)[9], const float4 p0,
                                                   const float4 p1)
{
    const float2 x = make_float2(p0.x, p0.y), y = make_float2(p0.z, p0.w);
    const float2 nX = make_float2(p1.x, p1.y), nY = make_float2(p1.z, p1.w);
    const float2 u = __ffma2_rn(g[0], x, __ffma2_rn(g[1], y, g[2]));
    const float2 v = __ffma2_rn(g[3], x, __ffma2_rn(g[4], y, g[5]));
#if SKS_RANSAC_H8_ONE
    const float2 w = __ffma2_rn(g[6], x, __ffma2_rn(g[7], y, make_float2(1.0f, 1.0f)));
#else
    const float2 w = __ffma2_rn(g[6], x, __ffma2_rn(g[7], y, g[8]));
#endif
#if SKS_RANSAC_SCALAR_RESID
    // residuals as four scalar FFMAs: a 3-pair FFMA2 needs six register reads (three
    // cycles), whereas scalar FFMAs with -X / -Y served by the reuse cache run at full rate
    const float2 du = make_float2(__fmaf_rn(nX.x, w.x, u.x), __fmaf_rn(nX.y, w.y, u.y));
    const float2 dv = make_float2(__fmaf_rn(nY.x, w.x, v.x), __fmaf_rn(nY.y, w.y, v.y));
#else
    const float2 du = __ffma2_rn(nX, w, u);
    const float2 dv = __ffma2_rn(nY, w, v);
#endif
    const float2 e = __ffma2_rn(dv, dv, __fmul2_rn(du, du));
#if SKS_RANSAC_TAIL == 0
    const float ax = __fmaf_rn(-w.x, w.x, e.x);
    const float ay = __fmaf_rn(-w.y, w.y, e.y);
    return (__float_as_uint(ax) >> 31) + (__float_as_uint(ay) >> 31);
#elif SKS_RANSAC_TAIL == 1
    const float2 nw = make_float2(__uint_as_float(__float_as_uint(w.x) ^ 0x80000000u),
                                  __uint_as_float(__float_as_uint(w.y) ^ 0x80000000u));
    const float2 acc = __ffma2_rn(nw, w, e);
    return (__float_as_uint(acc.x) >> 31) + (__float_as_uint(acc.y) >> 31);
#else
    const float2 w2 = __fmul2_rn(w, w);    // exact products: e < w*w is the same predicate as
    return (e.x < w2.x ? 1u : 0u) + (e.y < w2.y ? 1u : 0u);   // fma(-w,w,e) < 0 only up to rounding
#endif
}

// Third form: the packed scorer above with the inlier COUNT kept on the FP32 pipe.
// The two integer LEA.HI of ransac_inlier2 run on the 16-lane ALU pipe and cost 1.6-2
// issue cycles each next to FFMA2 (profiles/r01_ubench_issue_mix.log); here the count
// is one more packed FMA with directed rounding:
//     cnt = fma_rd(acc, 2^-149, cnt)          cnt in [2^23, 2^24): ulp(cnt) = 1
// |acc * 2^-149| < 2^-21 for every finite acc, the product is exact inside the FMA and
// non-zero whenever acc != 0, so rounding toward -inf gives cnt - 1 exactly when
// acc < 0 and cnt when acc >= +0 (or -0, which the IEEE test "acc < 0" of the oracle
// rejects as well).  A NaN or infinite acc poisons cnt; the caller sees a non-finite
// counter after the tile and recounts that hypothesis with the integer form, so the
// result is the oracle's count in every case.
__device__ __forceinline__ void ransac_inlier2_fp(const float2 (&g)[9], const float4 p0,
                                                  const float4 p1, float2& cnt)
{
    const float2 x = make_float2(p0.x, p0.y), y = make_float2(p0.z, p0.w);
    const float2 nX = make_float2(p1.x, p1.y), nY = make_float2(p1.z, p1.w);
    const float2 u = __ffma2_rn(g[0], x, __ffma2_rn(g[1], y, g[2]));
    const float2 v = __ffma2_rn(g[3], x, __ffma2_rn(g[4], y, g[5]));
    const float2 w = __ffma2_rn(g[6], x, __ffma2_rn(g[7], y, g[8]));
    const float2 du = __ffma2_rn(nX, w, u);
    const float2 dv = __ffma2_rn(nY, w, v);
    const float2 e = __ffma2_rn(dv, dv, __fmul2_rn(du, du));
    const float2 nw = make_float2(__uint_as_float(__float_as_uint(w.x) ^ 0x80000000u),
                                  __uint_as_float(__float_as_uint(w.y) ^ 0x80000000u));
    const float2 acc = __ffma2_rn(nw, w, e);
    const float tiny = __uint_as_float(1u);                       // 2^-149
    cnt = __ffma2_rd(acc, make_float2(tiny, tiny), cnt);
}
constexpr float kRansacFpCount0 = 16777215.0f;
// HPT hypotheses x one match pair, written phase by phase: every FFMA2 of a phase reads the
// same match operand (y, x, -Xs / -Ys) in the same source slot, so consecutive instructions
// take it from the operand-reuse cache and read at most two even and two odd registers -- a
// packed FMA that needs a third register from one bank occupies the pipe for three cycles
// instead of two (tools/ubench/count_ops.cu: 3.03 vs 2.04 cycles).  Same operations per
// hypothesis x match as ransac_inlier2_fp, so the same bits.
template <int HPT, int NP>
__device__ __forceinline__ void ransac_score_pairs_fp(const float (&g)[HPT][9], const float4 (&c)[2 * NP],
                                                      float2 (&fc)[HPT])
{
    // NP match pairs x HPT hypotheses per phase: each phase is NP*HPT (x3 in the first two)
    // independent instructions, long enough to cover the packed FMA's latency, so a warp keeps
    // issuing across phase boundaries (and keeps its operand-reuse cache) instead of yielding
    float2 u[NP][HPT], v[NP][HPT], w[NP][HPT];
#define SKS_B(j, k) make_float2(g[j][k], g[j][k])
#pragma unroll
    for (int p = 0; p < NP; ++p) {                        // phase 1: y in slot b
        const float2 y = make_float2(c[2 * p].z, c[2 * p].w);
#pragma unroll
        for (int j = 0; j < HPT; ++j) {
            u[p][j] = __ffma2_rn(SKS_B(j, 1), y, SKS_B(j, 2));
            v[p][j] = __ffma2_rn(SKS_B(j, 4), y, SKS_B(j, 5));
            w[p][j] = __ffma2_rn(SKS_B(j, 7), y, SKS_B(j, 8));
        }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {                        // phase 2: x in slot b
        const float2 x = make_float2(c[2 * p].x, c[2 * p].y);
#pragma unroll
        for (int j = 0; j < HPT; ++j) {
            u[p][j] = __ffma2_rn(SKS_B(j, 0), x, u[p][j]);
            v[p][j] = __ffma2_rn(SKS_B(j, 3), x, v[p][j]);
            w[p][j] = __ffma2_rn(SKS_B(j, 6), x, w[p][j]);
        }
    }
#undef SKS_B
#pragma unroll
    for (int p = 0; p < NP; ++p) {                        // phase 3: residuals; w[j] shared in slot b,
        const float2 nX = make_float2(c[2 * p + 1].x, c[2 * p + 1].y);   // -Xs / -Ys shared in slot a
        const float2 nY = make_float2(c[2 * p + 1].z, c[2 * p + 1].w);
#pragma unroll
        for (int j = 0; j < HPT; ++j) {
            if (j & 1) {
                v[p][j] = __ffma2_rn(nY, w[p][j], v[p][j]);
                u[p][j] = __ffma2_rn(nX, w[p][j], u[p][j]);
            } else {
                u[p][j] = __ffma2_rn(nX, w[p][j], u[p][j]);
                v[p][j] = __ffma2_rn(nY, w[p][j], v[p][j]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int j = 0; j < HPT; ++j)
            u[p][j] = __fmul2_rn(u[p][j], u[p][j]);
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int j = 0; j < HPT; ++j)
            u[p][j] = __ffma2_rn(v[p][j], v[p][j], u[p][j]);
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int j = 0; j < HPT; ++j) {
            const float2 nw = make_float2(-w[p][j].x, -w[p][j].y);   // exact; ptxas folds it into FFMA2's operand negation
            u[p][j] = __ffma2_rn(nw, w[p][j], u[p][j]);
        }
    const float tiny = __uint_as_float(1u);
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int j = 0; j < HPT; ++j)
            fc[j] = __ffma2_rd(u[p][j], make_float2(tiny, tiny), fc[j]);
}

template <int HPT>
__device__ __forceinline__ void ransac_score_pair_fp(const float (&g)[HPT][9], const float4 p0,
                                                     const float4 p1, float2 (&fc)[HPT])
{
    const float4 c[2] = {p0, p1};
    ransac_score_pairs_fp<HPT, 1>(g, c, fc);
}

// Second packed form: TWO HYPOTHESES per instruction, one match.  g[k] = (gA_k, gB_k)
// are natural register pairs and the match scalars enter as broadcast operands
// (SASS: Rn.F32), so no FFMA2 reads more than five registers (pair, scalar, pair)
// and consecutive instructions share the scalar through the operand-reuse cache;
// the match-pair form above needs six for its residuals (three pairs), which costs
// a third pipe cycle (tools/ubench/fma_peak.cu).  c = (x, y, -Xs, -Ys).
__device__ __forceinline__ void ransac_inlier_hp(const float2 (&g)[9], const float4 c,
                                                 uint32_t& cnt_a, uint32_t& cnt_b)
{
    const float2 x = make_float2(c.x, c.x), y = make_float2(c.y, c.y);
    const float2 nX = make_float2(c.z, c.z), nY = make_float2(c.w, c.w);
    const float2 u = __ffma2_rn(g[0], x, __ffma2_rn(g[1], y, g[2]));
    const float2 v = __ffma2_rn(g[3], x, __ffma2_rn(g[4], y, g[5]));
    const float2 w = __ffma2_rn(g[6], x, __ffma2_rn(g[7], y, g[8]));
    const float2 du = __ffma2_rn(nX, w, u);
    const float2 dv = __ffma2_rn(nY, w, v);
    const float2 e = __ffma2_rn(dv, dv, __fmul2_rn(du, du));
    // -w: ptxas folds the sign flip into FFMA2's operand negation
    const float2 nw = make_float2(__uint_as_float(__float_as_uint(w.x) ^ 0x80000000u),
                                  __uint_as_float(__float_as_uint(w.y) ^ 0x80000000u));
    const float2 acc = __ffma2_rn(nw, w, e);
    cnt_a += __float_as_uint(acc.x) >> 31;
    cnt_b += __float_as_uint(acc.y) >> 31;
}

__device__ __forceinline__ unsigned long long ransac_key(uint32_t count, uint32_t hyp)
{
    return ((unsigned long long)count << 32) | (unsigned long long)(0xFFFFFFFFu - hyp);
}

// State of a CTA's correspondence tile stream, shared by the scoring passes.
struct RansacTileStream {
    float4* tile;
    uint64_t* bar;
    const float4* corr_pair;
    int32_t n_pts, tile_pts, n_tiles;
    uint32_t phase;
    bool resident;       // a single tile stays in shared memory for all rounds and passes

    __device__ __forceinline__ void load(int tl) const   // one thread: one bulk copy per tile
    {
        const int lo = tl * tile_pts;
        const int cnt = (n_pts - lo < tile_pts) ? (n_pts - lo) : tile_pts;
        mbar_arrive_expect_tx(bar, (uint32_t)cnt * 16u);
        bulk_g2s(tile, corr_pair + lo, (uint32_t)cnt * 16u, bar);
    }
};

// One pass of a CTA over its hypothesis ids [c_lo, c_hi) of one pair: every thread carries HPT
// hypotheses per round and walks the pair's tile(s).  MODE 0: scalar FFMA scorer; 1: two matches
// per FFMA2/FMUL2 (the tile is re-laid out in pairs on arrival); 2: two hypotheses per
// FFMA2/FMUL2 (tile stays one match per 16 bytes); 3: as 1 with the count on the FP32 pipe.
// Returns true if an FP-pipe counter was poisoned by a NaN / infinite score (MODE 3 only): the
// keys of this pass are then lower bounds and the caller repeats the pass with MODE 1.
template <int kRansacHpt, int MODE, int kRansacThreads>
__device__ __forceinline__ bool ransac_pass(RansacTileStream& ts, int64_t pair, int64_t pair_id_base,
                                            const uint32_t* __restrict__ samples, uint32_t hyp_stride,
                                            uint32_t hyp_begin, uint32_t c_lo, uint32_t c_hi, uint64_t key,
                                            float it, unsigned long long& best)
{
    constexpr bool PACKED = (MODE == 1 || MODE == 3);
    constexpr bool FPCOUNT = (MODE == 3);     // inlier count on the FP32 pipe (ransac_score_pair_fp)
    constexpr bool HYPPAIR = (MODE == 2);
    static_assert(!HYPPAIR || kRansacHpt % 2 == 0, "hypothesis pairs");
    const int tid = threadIdx.x;
    float4* tile = ts.tile;
    const int n_pts = ts.n_pts, tile_pts = ts.tile_pts, n_tiles = ts.n_tiles;
    bool poisoned = false;
    for (uint32_t base = c_lo; base < c_hi; base += kRansacThreads * kRansacHpt) {
        float h[kRansacHpt][9];
        float2 h2[(PACKED && !FPCOUNT) ? kRansacHpt : HYPPAIR ? kRansacHpt / 2 : 1][9];
        uint32_t cnt[kRansacHpt], hyp[kRansacHpt];
        bool live[kRansacHpt];
#pragma unroll
        for (int j = 0; j < kRansacHpt; ++j) {
            const uint32_t local = base + (uint32_t)j * kRansacThreads + tid;
            live[j] = local < c_hi;
            hyp[j] = hyp_begin + (live[j] ? local : c_lo);
            uint32_t idx[4];
            ransac_sample(key, pair, pair_id_base, hyp[j], samples, hyp_stride, (uint32_t)n_pts, idx);
            ransac_hypothesis(ts.corr_pair, idx, h[j]);
            ransac_scale_h(h[j], it);
            cnt[j] = 0;
            if constexpr (FPCOUNT) {
                // A hypothesis with a non-finite entry (repeated or collinear sample) can never have an
                // inlier: every entry feeds u, v or w, non-finite values propagate to e = +inf / NaN and
                // acc = fma(-w, w, e) is then +inf or NaN.  Score it as the zero matrix instead (acc = +0
                // for every finite match) so that it does not poison the FP-pipe counter; its key stays
                // (0, id), as in the oracle.
                uint32_t bad = 0;
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    bad |= ((__float_as_uint(h[j][k]) & 0x7f800000u) == 0x7f800000u) ? 1u : 0u;
                if (bad) {
#pragma unroll
                    for (int k = 0; k < 9; ++k)
                        h[j][k] = 0.0f;
                }
            } else if constexpr (PACKED) {
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    h2[j][k] = make_float2(h[j][k], h[j][k]);
            }
        }
        if constexpr (HYPPAIR) {
#pragma unroll
            for (int j = 0; j < kRansacHpt / 2; ++j)
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    h2[j][k] = make_float2(h[2 * j][k], h[2 * j + 1][k]);
        }
        for (int tl = 0; tl < n_tiles; ++tl) {
            const int lo = tl * tile_pts;
            const int np = (n_pts - lo < tile_pts) ? (n_pts - lo) : tile_pts;
            if (!ts.resident) {
                mbar_wait(ts.bar, ts.phase);
                ts.phase ^= 1;
                ts.resident = (n_tiles == 1);
                if constexpr (PACKED) {
                    // re-lay the freshly landed AoS tile out in pairs, in place, with the
                    // targets scaled by 1/thr: (x0,x1,y0,y1)(-Xs0,-Xs1,-Ys0,-Ys1); an odd
                    // tail is padded with a NaN target that can never be an inlier (the FP-pipe
                    // count, which a NaN would poison, takes an odd last match through the integer form)
                    const float qnan = __int_as_float(0x7fffffff);
                    for (int p = tid; 2 * p < np; p += kRansacThreads) {
                        const float4 a = tile[2 * p];
                        const float4 b = (2 * p + 1 < np) ? tile[2 * p + 1]
                                                          : make_float4(0.f, 0.f, qnan, qnan);
                        tile[2 * p] = make_float4(a.x, b.x, a.y, b.y);
                        tile[2 * p + 1] = make_float4(__fmul_rn(-a.z, it), __fmul_rn(-b.z, it),
                                                      __fmul_rn(-a.w, it), __fmul_rn(-b.w, it));
                    }
                } else {
                    for (int p = tid; p < np; p += kRansacThreads) {   // (x, y, -Xs, -Ys)
                        const float4 a = tile[p];
                        tile[p] = make_float4(a.x, a.y, __fmul_rn(-a.z, it), __fmul_rn(-a.w, it));
                    }
                }
                __syncthreads();
            }
            if constexpr (FPCOUNT) {
                const int npair = np >> 1;                // an odd last match goes through the integer form
                float2 fc[kRansacHpt];
#pragma unroll
                for (int j = 0; j < kRansacHpt; ++j)
                    fc[j] = make_float2(kRansacFpCount0, kRansacFpCount0);
                int i = 0;
                for (; i + SKS_RANSAC_FP_UNROLL <= npair; i += SKS_RANSAC_FP_UNROLL) {
                    float4 c[2 * SKS_RANSAC_FP_UNROLL];
#pragma unroll
                    for (int k = 0; k < 2 * SKS_RANSAC_FP_UNROLL; ++k)
                        c[k] = tile[2 * i + k];   // warp-uniform address: broadcast
#if SKS_RANSAC_FP_JOINT     // all pairs of the iteration phase by phase: measured 100.9 vs 96.3 ms (ptxas
                            // re-serialises the pairs under the 80-register cap and schedules them worse)
                    ransac_score_pairs_fp<kRansacHpt, SKS_RANSAC_FP_UNROLL>(h, c, fc);
#else
#pragma unroll
                    for (int k = 0; k < SKS_RANSAC_FP_UNROLL; ++k)
                        ransac_score_pair_fp<kRansacHpt>(h, c[2 * k], c[2 * k + 1], fc);
#endif
                }
                for (; i < npair; ++i)
                    ransac_score_pair_fp<kRansacHpt>(h, tile[2 * i], tile[2 * i + 1], fc);
#pragma unroll
                for (int j = 0; j < kRansacHpt; ++j) {
                    const float dx = __fsub_rn(kRansacFpCount0, fc[j].x);   // exact when healthy
                    const float dy = __fsub_rn(kRansacFpCount0, fc[j].y);
                    const bool ok = dx >= 0.0f && dx <= 8388607.0f && dy >= 0.0f && dy <= 8388607.0f;
                    cnt[j] += ok ? (uint32_t)dx + (uint32_t)dy : 0u;   // poisoned: a lower bound
                    poisoned = poisoned || !ok;
                    if (np & 1) {
                        const float4 a = tile[np - 1], b = tile[np];
                        cnt[j] += ransac_inlier(h[j], make_float4(a.x, a.z, b.x, b.z));
                    }
                }
            } else if constexpr (PACKED) {
                const int npair = (np + 1) >> 1;
                int i = 0;
                for (; i + SKS_RANSAC_UNROLL <= npair; i += SKS_RANSAC_UNROLL) {
                    float4 c[2 * SKS_RANSAC_UNROLL];
#pragma unroll
                    for (int k = 0; k < 2 * SKS_RANSAC_UNROLL; ++k)
                        c[k] = tile[2 * i + k];   // warp-uniform address: broadcast
#if SKS_RANSAC_HYP_MAJOR
#pragma unroll
                    for (int j = 0; j < kRansacHpt; ++j)
#pragma unroll
                        for (int k = 0; k < SKS_RANSAC_UNROLL; ++k)
                            cnt[j] += ransac_inlier2(h2[j], c[2 * k], c[2 * k + 1]);
#else
#pragma unroll
                    for (int k = 0; k < SKS_RANSAC_UNROLL; ++k)
#pragma unroll
                        for (int j = 0; j < kRansacHpt; ++j)
                            cnt[j] += ransac_inlier2(h2[j], c[2 * k], c[2 * k + 1]);
#endif
                }
                for (; i < npair; ++i) {
                    const float4 c0 = tile[2 * i], c1 = tile[2 * i + 1];
#pragma unroll
                    for (int j = 0; j < kRansacHpt; ++j)
                        cnt[j] += ransac_inlier2(h2[j], c0, c1);
                }
            } else if constexpr (HYPPAIR) {
                int i = 0;
                for (; i + SKS_RANSAC_UNROLL <= np; i += SKS_RANSAC_UNROLL) {
                    float4 c[SKS_RANSAC_UNROLL];
#pragma unroll
                    for (int k = 0; k < SKS_RANSAC_UNROLL; ++k)
                        c[k] = tile[i + k];   // warp-uniform address: broadcast
#pragma unroll
                    for (int k = 0; k < SKS_RANSAC_UNROLL; ++k)
#pragma unroll
                        for (int j = 0; j < kRansacHpt / 2; ++j)
                            ransac_inlier_hp(h2[j], c[k], cnt[2 * j], cnt[2 * j + 1]);
                }
                for (; i < np; ++i) {
                    const float4 c = tile[i];
#pragma unroll
                    for (int j = 0; j < kRansacHpt / 2; ++j)
                        ransac_inlier_hp(h2[j], c, cnt[2 * j], cnt[2 * j + 1]);
                }
            } else {
                int i = 0;
                for (; i + 4 <= np; i += 4) {
                    float4 c[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        c[k] = tile[i + k];   // warp-uniform address: broadcast
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int j = 0; j < kRansacHpt; ++j)
                            cnt[j] += ransac_inlier(h[j], c[k]);
                }
                for (; i < np; ++i) {
                    const float4 c = tile[i];
#pragma unroll
                    for (int j = 0; j < kRansacHpt; ++j)
                        cnt[j] += ransac_inlier(h[j], c);
                }
            }
            if (n_tiles > 1) {   // stream the next tile (wraps for the next round)
                __syncthreads();
                const bool more = (tl + 1 < n_tiles) ||
                                  (base + kRansacThreads * kRansacHpt < c_hi);
                if (tid == 0 && more)
                    ts.load((tl + 1) % n_tiles);
            }
        }
#pragma unroll
        for (int j = 0; j < kRansacHpt; ++j)
            if (live[j]) {
                const unsigned long long k = ransac_key(cnt[j], hyp[j]);
                best = k > best ? k : best;
            }
    }
    return poisoned;
}

// grid = (chunks_per_pair, n_pairs); each CTA scores hypothesis ids
// [hyp_begin + chunk*chunk_size, +chunk_size) ∩ [hyp_begin, hyp_begin+hyp_count)
// kRansacHpt: hypotheses carried per thread per round; MODE: see ransac_pass.  MODE 3 keeps its
// hot loop free of cold code: if any FP-pipe counter of the CTA was poisoned (overflowing or
// non-finite matches -- never on sane image coordinates), the CTA repeats its chunk with the
// integer-count scorer of MODE 1 on the same packed tile and the keys are max-combined, which
// gives the oracle's keys in every case.
template <int kRansacHpt, int MODE, int kRansacThreads>
__global__ void __launch_bounds__(kRansacThreads, SKS_RANSAC_MIN_CTAS(kRansacThreads))
k_ransac_aca(const float4* __restrict__ corr, int64_t pair_base, int32_t n_pts, int32_t tile_pts,
             const uint32_t* __restrict__ samples, uint32_t hyp_stride, uint32_t hyp_begin,
             uint32_t hyp_count, uint32_t chunk_size, uint64_t key, float thr2,
             unsigned long long* __restrict__ best_key, int64_t pair_id_base)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ unsigned long long warp_best[kRansacThreads / 32];

    const int tid = threadIdx.x;
    const int64_t pair = pair_base + blockIdx.y;
    const uint32_t c_lo = blockIdx.x * chunk_size;
    if (c_lo >= hyp_count)
        return;
    const uint32_t c_hi = (hyp_count - c_lo < chunk_size) ? hyp_count : c_lo + chunk_size;

    RansacTileStream ts;
    ts.tile = reinterpret_cast<float4*>(smem);
    ts.bar = &bar;
    ts.corr_pair = corr + (size_t)pair * n_pts;
    ts.n_pts = n_pts;
    ts.tile_pts = tile_pts;
    ts.n_tiles = (n_pts + tile_pts - 1) / tile_pts;
    ts.phase = 0;
    ts.resident = false;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0)
        ts.load(0);

    const float it = ransac_inv_thr(thr2);
    unsigned long long best = 0ull;
    const bool poisoned = ransac_pass<kRansacHpt, MODE, kRansacThreads>(
        ts, pair, pair_id_base, samples, hyp_stride, hyp_begin, c_lo, c_hi, key, it, best);
    if constexpr (MODE == 3) {
        if (__syncthreads_or(poisoned ? 1 : 0)) {      // cold: exact repeat with the integer count
            if (!ts.resident && tid == 0)
                ts.load(0);                            // multi-tile stream: start over at tile 0
            ransac_pass<kRansacHpt, 1, kRansacThreads>(ts, pair, pair_id_base, samples, hyp_stride,
                                                       hyp_begin, c_lo, c_hi, key, it, best);
        }
    }
    // CTA-wide max: shuffles inside the warp, shared memory across warps
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
        best = o > best ? o : best;
    }
    if ((tid & 31) == 0)
        warp_best[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int wv = 1; wv < kRansacThreads / 32; ++wv)
            best = warp_best[wv] > best ? warp_best[wv] : best;
        atomicMax(best_key + pair, best);
    }
}

// One CTA per pair: rebuild the winning hypothesis from its id and emit the
// model, its inlier count and (optionally) the inlier mask.
__global__ void __launch_bounds__(256)
k_ransac_finalize(const float4* __restrict__ corr, int32_t n_pts,
                  const uint32_t* __restrict__ samples, uint32_t hyp_stride, uint64_t key,
                  float thr2, const unsigned long long* __restrict__ best_key,
                  float* __restrict__ H_best, uint32_t* __restrict__ inlier_count,
                  uint8_t* __restrict__ inlier_mask, int64_t pair_id_base,
                  const float* __restrict__ H_given)
{
    __shared__ float hs[9];
    __shared__ uint32_t total;
    const int64_t pair = blockIdx.x;
    const float4* corr_pair = corr + (size_t)pair * n_pts;
    if (threadIdx.x == 0) {
        float h[9];
        if (H_given != nullptr) {           // score the caller's models (post-RANSAC polishing)
#pragma unroll
            for (int k = 0; k < 9; ++k)
                h[k] = H_given[pair * 9 + k];
        } else {
            // a key that is still zero (nothing was scored for this pair) decodes to the id
            // 0xFFFFFFFF: no model -- NaN matrix, count 0 -- instead of a fabricated one
            const unsigned long long bk = best_key[pair];
            const uint32_t hyp = 0xFFFFFFFFu - (uint32_t)(bk & 0xFFFFFFFFull);
            if (bk == 0ull || (samples != nullptr && hyp >= hyp_stride)) {
                const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    h[k] = qnan;
            } else {
                uint32_t idx[4];
                ransac_sample(key, pair, pair_id_base, hyp, samples, hyp_stride, (uint32_t)n_pts, idx);
                ransac_hypothesis(corr_pair, idx, h);
            }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            hs[k] = h[k];
            if (H_best != nullptr)
                H_best[pair * 9 + k] = h[k];
        }
        total = 0;
    }
    __syncthreads();
    const float it = ransac_inv_thr(thr2);
    float h[9];
#pragma unroll
    for (int k = 0; k < 9; ++k)
        h[k] = hs[k];
    ransac_scale_h(h, it);
    uint32_t mine = 0;
    for (int i = threadIdx.x; i < n_pts; i += blockDim.x) {
        float4 c = __ldg(corr_pair + i);
        c.z = __fmul_rn(-c.z, it);
        c.w = __fmul_rn(-c.w, it);
        const uint32_t in = ransac_inlier(h, c);
        mine += in;
        if (inlier_mask != nullptr)
            inlier_mask[(size_t)pair * n_pts + i] = in ? 1 : 0;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0)
        atomicAdd(&total, mine);
    __syncthreads();
    if (threadIdx.x == 0 && inlier_count != nullptr)
        inlier_count[pair] = total;
}

}  // namespace sksb
