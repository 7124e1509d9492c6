// Streaming solver kernels: one thread per correspondence quadruple, HBM-bound
// (arithmetic intensity ~1 flop/B, >10x below the B200 ridge, so no tensor
// cores: there is no contraction to feed them).  Four kernels per
// (solver, precision):
//
//   k_aos_ring    persistent CTAs; the AoS input tiles are pulled into a
//                 multi-stage shared-memory ring by the bulk async-copy (TMA)
//                 engine (cp.async.bulk + mbarrier), each thread reads its own
//                 32/64-byte quadruple with conflict-free swizzled 16-byte
//                 shared loads, and the 9-word results are transposed through
//                 shared memory and leave as one bulk shared->global copy per
//                 tile.  No register-staged global traffic at all.
//   k_aos_wring   the same with a PRIVATE ring per warp: lane 0 of every warp is its
//                 own TMA producer, so the kernel has no CTA-wide barrier at all.
//   k_aos_direct  one tile per CTA; each thread pulls its own quadruple with
//                 256-bit loads (sm_100's LDG.256: one full 32-byte sector per
//                 lane) or, for pointers that are only 16-byte aligned, 16-byte
//                 read-only loads; results are transposed through shared memory
//                 into coalesced 256-bit / 16-byte streaming stores.
//   k_soa         the reference GPU layout (GPU.cu:87-95,141-149): 16-byte
//                 coalesced loads of 4 (fp32) / 2 (fp64) consecutive quadruples
//                 per thread per coordinate plane, streaming cache hints.
#pragma once
#include <cstdint>

#include "ptx.cuh"
#include "solvers.cuh"

namespace sksb {

enum { SOLVER_ACA = 0, SOLVER_SKS = 1, SOLVER_RECT = 2, SOLVER_GE = 3, SOLVER_GPT = 4 };

template <typename T>
struct RectParams {
    T mx, my, width, ratio;
};

template <int SOLVER, typename T>
__device__ __forceinline__ void solve_quad(const T (&s)[8], const T (&t)[8], T mx, T my,
                                           const RectParams<T>& rp, T (&h)[9], bool normalize)
{
    if constexpr (SOLVER == SOLVER_ACA)
        aca_solve<T>(s, t, h, normalize);
    else if constexpr (SOLVER == SOLVER_SKS)
        sks_solve<T>(s, t, h, normalize);
    else if constexpr (SOLVER == SOLVER_GE)
        ge_solve<T>(s, t, h);          // h33 = 1 by construction, with or without `normalize`
    else if constexpr (SOLVER == SOLVER_GPT)
        gpt_solve<T>(s, t, h);
    else
        aca_rect_solve<T>(t, mx, my, rp.width, rp.ratio, h, normalize);
}

// ---- 16-byte chunk <-> scalars ---------------------------------------------
template <typename T>
struct ChunkTraits;
template <>
struct ChunkTraits<float> {
    static constexpr int EPC = 4;   // elements per 16-byte chunk
    static __device__ __forceinline__ float get(const Chunk16& c, int e) { return __uint_as_float(c.w[e]); }
    static __device__ __forceinline__ void set(Chunk16& c, int e, float v) { c.w[e] = __float_as_uint(v); }
};
template <>
struct ChunkTraits<double> {
    static constexpr int EPC = 2;
    static __device__ __forceinline__ double get(const Chunk16& c, int e)
    {
        return __hiloint2double((int)c.w[2 * e + 1], (int)c.w[2 * e]);
    }
    static __device__ __forceinline__ void set(Chunk16& c, int e, double v)
    {
        c.w[2 * e] = (uint32_t)__double2loint(v);
        c.w[2 * e + 1] = (uint32_t)__double2hiint(v);
    }
};

__device__ __forceinline__ Chunk16 pick(bool second, const Chunk16& a, const Chunk16& b)
{
    Chunk16 r;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        r.w[k] = second ? b.w[k] : a.w[k];
    return r;
}

// One quadruple (8 scalars) straight from global memory, 16 bytes at a time.
template <typename T>
__device__ __forceinline__ void load_quad_global(const T* p, T (&v)[8])
{
    constexpr int EPC = ChunkTraits<T>::EPC, NC = 8 / EPC;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const Chunk16 c = ldg_nc(p + j * EPC);
#pragma unroll
        for (int e = 0; e < EPC; ++e)
            v[j * EPC + e] = ChunkTraits<T>::get(c, e);
    }
}

// Same, with 256-bit loads (one per fp32 quadruple, two per fp64 quadruple).
template <typename T>
__device__ __forceinline__ void load_quad_global_wide(const T* p, T (&v)[8])
{
    constexpr int EPW = 32 / (int)sizeof(T), NW = 8 / EPW;   // elements per 32 bytes
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const Chunk32 c = ldg_stream32(p + j * EPW);
#pragma unroll
        for (int e = 0; e < EPW; ++e) {
            if constexpr (sizeof(T) == 4)
                v[j * EPW + e] = __uint_as_float(c.w[e]);
            else
                v[j * EPW + e] = __hiloint2double((int)c.w[2 * e + 1], (int)c.w[2 * e]);
        }
    }
}

// One quadruple from an AoS shared-memory tile.  Thread `tid` owns bytes
// [tid*8*sizeof(T), +8*sizeof(T)); a plain 16-byte read at that stride is a
// 2-way (fp32) / 4-way (fp64) bank conflict, so each quarter-warp reads its
// chunks in a rotated order that covers all 32 banks exactly once, and the
// rotation is undone in registers with selects.
template <typename T>
__device__ __forceinline__ void load_quad_smem(const unsigned char* tile, int tid, T (&v)[8])
{
    constexpr int EPC = ChunkTraits<T>::EPC, NC = 8 / EPC;
    const unsigned char* mine = tile + (size_t)tid * (8 * sizeof(T));
    Chunk16 d[NC];
    if constexpr (NC == 2) {
        const int r = (tid >> 2) & 1;
        const Chunk16 c0 = lds16(mine + 16 * r);
        const Chunk16 c1 = lds16(mine + 16 * (r ^ 1));
        d[0] = pick(r, c0, c1);
        d[1] = pick(r, c1, c0);
    } else {
        const int r = (tid >> 1) & 3;
        Chunk16 c[4], e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            c[j] = lds16(mine + 16 * ((j + r) & 3));   // c[j] = chunk[(j+r)&3]
#pragma unroll
        for (int j = 0; j < 4; ++j)
            e[j] = pick(r & 1, c[j], c[(j + 3) & 3]);  // e[j] = chunk[(j+(r&2))&3]
#pragma unroll
        for (int j = 0; j < 4; ++j)
            d[j] = pick(r & 2, e[j], e[(j + 2) & 3]);  // d[j] = chunk[j]
    }
#pragma unroll
    for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int e = 0; e < EPC; ++e)
            v[j * EPC + e] = ChunkTraits<T>::get(d[j], e);
}

// per-quadruple rectangle corner (AoS M[i*2+k])
template <typename T>
__device__ __forceinline__ void load_corner_aos(const T* M, int64_t i, T& mx, T& my)
{
    if constexpr (sizeof(T) == 4) {
        const float2 m = __ldg(reinterpret_cast<const float2*>(M) + i);
        mx = m.x;
        my = m.y;
    } else {
        const double2 m = __ldg(reinterpret_cast<const double2*>(M) + i);
        mx = m.x;
        my = m.y;
    }
}

// Coalesced write-out of one tile: the 9*cnt results staged in shared memory are
// contiguous in H (AoS), so they leave as 16-byte (or 256-bit) streaming stores.
template <typename T, int TILE, bool WIDE>
__device__ __forceinline__ void store_tile(const T* stage, T* __restrict__ H, int64_t q0, int cnt,
                                           int tid)
{
    const int total = cnt * 9;
    T* dst = H + q0 * 9;
    if constexpr (WIDE && SKS_WST_HINT != 0) {
        constexpr int EPW = 32 / (int)sizeof(T);
        const int nwide = total / EPW;
        for (int c = tid; c < nwide; c += TILE) {
            const Chunk16 lo = lds16(stage + c * EPW), hi = lds16(stage + c * EPW + EPW / 2);
            Chunk32 w;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                w.w[k] = lo.w[k];
                w.w[4 + k] = hi.w[k];
            }
            stg_stream32(dst + c * EPW, w);
        }
        for (int e = nwide * EPW + tid; e < total; e += TILE)
            dst[e] = stage[e];
    } else {
        constexpr int EPC = ChunkTraits<T>::EPC;
        const int nchunk = total / EPC;
        for (int c = tid; c < nchunk; c += TILE)
            stg_stream(dst + c * EPC, lds16(stage + c * EPC));
        for (int e = nchunk * EPC + tid; e < total; e += TILE)
            dst[e] = stage[e];
    }
}

// ------------------------------------------------------------ k_aos_direct
// WIDE: 256-bit global loads and stores (needs 32-byte aligned base pointers).
template <int SOLVER, typename T, int TILE, bool WIDE>
__global__ void __launch_bounds__(TILE)
k_aos_direct(const T* __restrict__ src, const T* __restrict__ tar, const T* __restrict__ M,
             RectParams<T> rp, T* __restrict__ H, uint8_t* __restrict__ degen, int64_t n,
             bool normalize)
{
    __shared__ __align__(32) T stage[TILE * 9];
    const int tid = threadIdx.x;
    const int64_t q0 = (int64_t)blockIdx.x * TILE;
    const int64_t i = q0 + tid;
    const int cnt = (int)((n - q0) < (int64_t)TILE ? (n - q0) : (int64_t)TILE);
    if (tid < cnt) {
        T s[8], t[8], h[9];
        T mx = rp.mx, my = rp.my;
        if constexpr (WIDE) {
            if constexpr (SOLVER != SOLVER_RECT)
                load_quad_global_wide<T>(src + i * 8, s);
            load_quad_global_wide<T>(tar + i * 8, t);
        } else {
            if constexpr (SOLVER != SOLVER_RECT)
                load_quad_global<T>(src + i * 8, s);
            load_quad_global<T>(tar + i * 8, t);
        }
        if constexpr (SOLVER == SOLVER_RECT)
            if (M != nullptr)
                load_corner_aos<T>(M, i, mx, my);
        solve_quad<SOLVER, T>(s, t, mx, my, rp, h, normalize);
#pragma unroll
        for (int k = 0; k < 9; ++k)
            stage[tid * 9 + k] = h[k];   // stride 9 words: conflict-free
        if (degen != nullptr)
            degen[i] = is_degenerate<T>(h, normalize) ? 1 : 0;
    }
    __syncthreads();
    store_tile<T, TILE, WIDE>(stage, H, q0, cnt, tid);
}

// --------------------------------------------------------- k_rect_planar34
// ACA-rect on the reference's torch tensor convention (PY.py:24-37, :286-302):
// tar34 / src34 are [n][3][4] homogeneous (rows x, y, 1; columns TL,TR,BL,BR).
// Each thread reads the x row and the y row of its sample (two 16-byte loads at
// a 48-byte stride, fp64: four at 96); the constant row of ones is never used.
// The per-sample rectangle corner is src34[i][0..1][0] (PY.py:302).
template <typename T, int TILE>
__global__ void __launch_bounds__(TILE)
k_rect_planar34(const T* __restrict__ tar34, const T* __restrict__ src34, RectParams<T> rp,
                T* __restrict__ H, uint8_t* __restrict__ degen, int64_t n, bool normalize,
                const T* __restrict__ width_dev, const T* __restrict__ ratio_dev)
{
    // the reference passes `scale` and `div` as one-element DEVICE tensors (PY.py:33-35, :301-302);
    // reading them here keeps the call free of a device->host synchronisation
    if (width_dev != nullptr) rp.width = __ldg(width_dev);
    if (ratio_dev != nullptr) rp.ratio = __ldg(ratio_dev);
    __shared__ __align__(32) T stage[TILE * 9];
    constexpr int EPC = ChunkTraits<T>::EPC, NC = 4 / EPC;   // 16-byte chunks per row
    const int tid = threadIdx.x;
    const int64_t q0 = (int64_t)blockIdx.x * TILE;
    const int64_t i = q0 + tid;
    const int cnt = (int)((n - q0) < (int64_t)TILE ? (n - q0) : (int64_t)TILE);
    if (tid < cnt) {
        T xr[4], yr[4], t[8], s[8], h[9];
        const T* base = tar34 + i * 12;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const Chunk16 cx = ldg_nc(base + j * EPC), cy = ldg_nc(base + 4 + j * EPC);
#pragma unroll
            for (int e = 0; e < EPC; ++e) {
                xr[j * EPC + e] = ChunkTraits<T>::get(cx, e);
                yr[j * EPC + e] = ChunkTraits<T>::get(cy, e);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            t[2 * k] = xr[k];
            t[2 * k + 1] = yr[k];
        }
        T mx = rp.mx, my = rp.my;
        if (src34 != nullptr) {
            mx = __ldg(src34 + i * 12);
            my = __ldg(src34 + i * 12 + 4);
        }
        solve_quad<SOLVER_RECT, T>(s, t, mx, my, rp, h, normalize);
#pragma unroll
        for (int k = 0; k < 9; ++k)
            stage[tid * 9 + k] = h[k];
        if (degen != nullptr)
            degen[i] = is_degenerate<T>(h, normalize) ? 1 : 0;
    }
    __syncthreads();
    store_tile<T, TILE, false>(stage, H, q0, cnt, tid);
}

// ----------------------------------------------------------- k_gather_solve
// The reference's GPU flow is get_rand_list -> cal_Homo_* (GPU.cu:1449-1464): the
// gather writes 16 coordinates per hypothesis to HBM and the solver reads them
// straight back (256 B of traffic per fp64 hypothesis before any result).  Fused
// here: each thread draws its four pool indices (r % pool_size, repeats allowed,
// GPU.cu:55-58; from the caller's [4][n] list or the counter RNG), gathers the
// matches from the L1/L2-resident pool, solves, and only H leaves the SM.
// WIDE (fp64, 32-byte aligned pool): a match (x,y,X,Y) is exactly one 32-byte sector and arrives
// with ONE 256-bit load instead of two 16-byte ones -- the kernel is bound by L1 gather
// wavefronts (32 random sectors per warp instruction), so this halves its dominant cost.
template <int SOLVER, typename T, int TILE, bool WIDE>
__global__ void __launch_bounds__(TILE)
k_gather_solve(const T* __restrict__ pool, uint32_t pool_size, const uint32_t* __restrict__ rand4,
               uint64_t key, T* __restrict__ H, uint8_t* __restrict__ degen, int64_t n, int layout,
               int64_t ld, bool normalize)
{
    __shared__ __align__(32) T stage[TILE * 9];
    const int tid = threadIdx.x;
    const int64_t q0 = (int64_t)blockIdx.x * TILE;
    const int64_t i = q0 + tid;
    const int cnt = (int)((n - q0) < (int64_t)TILE ? (n - q0) : (int64_t)TILE);
    if (tid < cnt) {
        T s[8], t[8], h[9];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t r;
            if (rand4 != nullptr)
                r = ldg_stream_u32(rand4 + (int64_t)k * n + i);   // streamed once: do not evict the pool from L1
            else {
                uint64_t z = key ^ ((uint64_t)i * 0x9E3779B97F4A7C15ULL + (uint64_t)k);
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
                r = (uint32_t)((z ^ (z >> 31)) >> 32);
            }
            const T* c = pool + 4 * (size_t)(r % pool_size);
            if constexpr (sizeof(T) == 4) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(c));
                s[2 * k] = m.x; s[2 * k + 1] = m.y; t[2 * k] = m.z; t[2 * k + 1] = m.w;
            } else if constexpr (WIDE) {
                const Chunk32 m = ldg_nc32(c);
                s[2 * k] = __hiloint2double((int)m.w[1], (int)m.w[0]);
                s[2 * k + 1] = __hiloint2double((int)m.w[3], (int)m.w[2]);
                t[2 * k] = __hiloint2double((int)m.w[5], (int)m.w[4]);
                t[2 * k + 1] = __hiloint2double((int)m.w[7], (int)m.w[6]);
            } else {
                const double2 a = __ldg(reinterpret_cast<const double2*>(c));
                const double2 b = __ldg(reinterpret_cast<const double2*>(c) + 1);
                s[2 * k] = a.x; s[2 * k + 1] = a.y; t[2 * k] = b.x; t[2 * k + 1] = b.y;
            }
        }
        RectParams<T> none{};
        solve_quad<SOLVER, T>(s, t, T(0), T(0), none, h, normalize);
        if (layout == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k)
                stage[tid * 9 + k] = h[k];
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k)
                H[k * ld + i] = h[k];
        }
        if (degen != nullptr)
            degen[i] = is_degenerate<T>(h, normalize) ? 1 : 0;
    }
    if (layout == 0) {
        __syncthreads();
        store_tile<T, TILE, false>(stage, H, q0, cnt, tid);
    }
}

// ------------------------------------------------------ k_gather_solve_pool
// Same flow with the match pool held in SHARED memory by persistent CTAs (SoA output, the
// reference GPU layout).  The L1 path above pays one L1 wavefront per lane for every random
// 32-byte match (ncu: L1 71 % busy, the kernel's limiter); a 16-byte shared load serves eight
// lanes per cycle when their banks differ, about three lanes with random indices.  Two
// 512-thread CTAs per SM each copy the pool once (a few tens of KB against tens of MB of
// output) and then walk tiles of 512 hypotheses without any barrier.
template <int SOLVER, typename T>
__global__ void __launch_bounds__(512)
k_gather_solve_pool(const T* __restrict__ pool, uint32_t pool_size, const uint32_t* __restrict__ rand4,
                    uint64_t key, T* __restrict__ H, uint8_t* __restrict__ degen, int64_t n,
                    int64_t ld, bool normalize)
{
    extern __shared__ __align__(128) unsigned char pool_smem[];
    Chunk16* sp = reinterpret_cast<Chunk16*>(pool_smem);
    const int tid = threadIdx.x;
    const uint32_t chunks = pool_size * (uint32_t)(4 * sizeof(T) / 16);
    for (uint32_t c = tid; c < chunks; c += 512)
        sp[c] = ldg_stream(reinterpret_cast<const Chunk16*>(pool) + c);
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * 512 + tid; i < n; i += (int64_t)gridDim.x * 512) {
        T s[8], t[8], h[9];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t r;
            if (rand4 != nullptr)
                r = ldg_stream_u32(rand4 + (int64_t)k * n + i);
            else {
                uint64_t z = key ^ ((uint64_t)i * 0x9E3779B97F4A7C15ULL + (uint64_t)k);
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
                r = (uint32_t)((z ^ (z >> 31)) >> 32);
            }
            const uint32_t e = r % pool_size;
            if constexpr (sizeof(T) == 4) {
                const Chunk16 m = lds16(sp + e);
                s[2 * k] = ChunkTraits<T>::get(m, 0); s[2 * k + 1] = ChunkTraits<T>::get(m, 1);
                t[2 * k] = ChunkTraits<T>::get(m, 2); t[2 * k + 1] = ChunkTraits<T>::get(m, 3);
            } else {
                const Chunk16 a = lds16(sp + 2 * e), b = lds16(sp + 2 * e + 1);
                s[2 * k] = ChunkTraits<T>::get(a, 0); s[2 * k + 1] = ChunkTraits<T>::get(a, 1);
                t[2 * k] = ChunkTraits<T>::get(b, 0); t[2 * k + 1] = ChunkTraits<T>::get(b, 1);
            }
        }
        RectParams<T> none{};
        solve_quad<SOLVER, T>(s, t, T(0), T(0), none, h, normalize);
#pragma unroll
        for (int k = 0; k < 9; ++k)
            H[k * ld + i] = h[k];
        if (degen != nullptr)
            degen[i] = is_degenerate<T>(h, normalize) ? 1 : 0;
    }
}

// -------------------------------------------------------------- k_aos_ring
template <int SOLVER, typename T, int TILE>
struct RingLayout {
    static constexpr int ARR_BYTES = TILE * 8 * (int)sizeof(T);
    static constexpr int NARR = (SOLVER == SOLVER_RECT) ? 1 : 2;
    static constexpr int STAGE_BYTES = ARR_BYTES * NARR;
    static constexpr int OUT_BYTES = TILE * 9 * (int)sizeof(T);
    static __host__ __device__ constexpr int smem_bytes(int stages)
    {
        return stages * STAGE_BYTES + 2 * OUT_BYTES + stages * 8;
    }
};

template <int SOLVER, typename T, int TILE>
__global__ void __launch_bounds__(TILE)
k_aos_ring(const T* __restrict__ src, const T* __restrict__ tar, const T* __restrict__ M,
           RectParams<T> rp, T* __restrict__ H, uint8_t* __restrict__ degen, int64_t n,
           int stages, bool normalize)
{
    using L = RingLayout<SOLVER, T, TILE>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ring = smem;
    T* out_base = reinterpret_cast<T*>(smem + stages * L::STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * L::STAGE_BYTES + 2 * L::OUT_BYTES);

    const int tid = threadIdx.x;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t my_tiles = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;

    // producer side (thread 0): arm the stage's barrier with the byte count and
    // hand both AoS tiles to the bulk-copy engine
    auto issue = [&](int64_t it, int s) {
        const int64_t q0 = (first + it * stride) * TILE;
        const int64_t left = n - q0;
        const uint32_t cnt = (uint32_t)(left < (int64_t)TILE ? left : (int64_t)TILE);
        const uint32_t bytes = cnt * 8u * (uint32_t)sizeof(T);
        unsigned char* dst = ring + s * L::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], bytes * L::NARR);
        if constexpr (SOLVER != SOLVER_RECT) {
            bulk_g2s(dst, src + q0 * 8, bytes, &full[s]);
            bulk_g2s(dst + L::ARR_BYTES, tar + q0 * 8, bytes, &full[s]);
        } else {
            bulk_g2s(dst, tar + q0 * 8, bytes, &full[s]);
        }
    };

    if (tid == 0) {
        for (int s = 0; s < stages; ++s)
            mbar_init(&full[s], 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < stages && s < my_tiles; ++s)
            issue(s, s);

    int s = 0;
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
        const int64_t q0 = (first + it * stride) * TILE;
        const int64_t left = n - q0;
        const int cnt = (int)(left < (int64_t)TILE ? left : (int64_t)TILE);
        const bool active = tid < cnt;
        const int64_t i = q0 + tid;

        T mx = rp.mx, my = rp.my;
        if constexpr (SOLVER == SOLVER_RECT)
            if (M != nullptr && active)
                load_corner_aos<T>(M, i, mx, my);

        mbar_wait(&full[s], parity);   // this stage's bytes have landed
        T sv[8], tv[8];
        if (active) {
            const unsigned char* st = ring + s * L::STAGE_BYTES;
            if constexpr (SOLVER != SOLVER_RECT) {
                load_quad_smem<T>(st, tid, sv);
                load_quad_smem<T>(st + L::ARR_BYTES, tid, tv);
            } else {
                load_quad_smem<T>(st, tid, tv);
            }
        }
        // the output buffer about to be overwritten was handed to the copy
        // engine two tiles ago; make sure it has been read out
        if (tid == 0)
            bulk_wait_read<1>();
        __syncthreads();   // stage s fully consumed, out[it&1] free
        if (tid == 0 && it + stages < my_tiles)
            issue(it + stages, s);

        T* out = out_base + (size_t)(it & 1) * (TILE * 9);
        if (active) {
            T h[9];
            solve_quad<SOLVER, T>(sv, tv, mx, my, rp, h, normalize);
#pragma unroll
            for (int k = 0; k < 9; ++k)
                out[tid * 9 + k] = h[k];
            if (degen != nullptr)
                degen[i] = is_degenerate<T>(h, normalize) ? 1 : 0;
        }
        fence_async_smem();
        __syncthreads();
        const uint32_t out_bytes = (uint32_t)cnt * 9u * (uint32_t)sizeof(T);
        if ((out_bytes & 15u) == 0) {
            if (tid == 0) {
                bulk_s2g(H + q0 * 9, out, out_bytes);
                bulk_commit();
            }
        } else {   // ragged last tile: size is not a 16-byte multiple
            for (int e = tid; e < cnt * 9; e += TILE)
                H[q0 * 9 + e] = out[e];
            if (tid == 0)
                bulk_commit();   // empty group keeps the one-group-per-tile count
        }
        if (++s == stages) {
            s = 0;
            parity ^= 1;
        }
    }
    if (tid == 0)
        bulk_wait_all<0>();
}

// -------------------------------------------------------------- k_aos_wring
// Warp-private TMA ring (variant 3): the k_aos_ring idea without a single CTA-wide barrier.
// Every WARP of a persistent CTA owns a private multi-stage ring in shared memory, its own
// `full` mbarriers and two private output buffers, and walks tiles of QPW quadruples
// (tile t belongs to global warp t mod G).  Lane 0 is the warp's producer: it arms a stage's
// barrier and hands the tile's AoS bytes to the bulk-copy (TMA) engine; once the warp has read
// a stage (a __syncwarp away) the same lane refills it, so no `empty` barrier and no producer
// warp are needed and the copy engine always has STAGES tiles per warp in flight.  Results are
// staged at stride 9 words and leave as one bulk shared->global copy per tile (double-buffered
// on the bulk async-group of lane 0).
template <int SOLVER, typename T, int QPW, int STAGES>
struct WringLayout {
    static constexpr int NARR = (SOLVER == SOLVER_RECT) ? 1 : 2;
    static constexpr int ARR_BYTES = QPW * 8 * (int)sizeof(T);
    static constexpr int STAGE_BYTES = ARR_BYTES * NARR;
    static constexpr int OUT_BYTES = QPW * 9 * (int)sizeof(T);
    static constexpr int OUT_PAD = (OUT_BYTES + 127) & ~127;
    static constexpr int WARP_BYTES = STAGES * STAGE_BYTES + 2 * OUT_PAD + 128;   // + barriers
    static __host__ __device__ constexpr int smem_bytes(int warps) { return warps * WARP_BYTES; }
};

template <int SOLVER, typename T, int WARPS, int QPW, int STAGES>
__global__ void __launch_bounds__(WARPS * 32)
k_aos_wring(const T* __restrict__ src, const T* __restrict__ tar, const T* __restrict__ M,
            RectParams<T> rp, T* __restrict__ H, uint8_t* __restrict__ degen, int64_t n, bool normalize)
{
    using L = WringLayout<SOLVER, T, QPW, STAGES>;
    static_assert(QPW % 32 == 0, "whole quadruples per lane");
    constexpr int PER_LANE = QPW / 32;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem + (size_t)warp * L::WARP_BYTES;
    unsigned char* ring = mine;
    unsigned char* outs = mine + STAGES * L::STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(outs + 2 * L::OUT_PAD);

    const int64_t n_tiles = (n + QPW - 1) / QPW;
    const int64_t G = (int64_t)gridDim.x * WARPS;
    const int64_t first = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t my_tiles = first < n_tiles ? (n_tiles - first + G - 1) / G : 0;

    auto issue = [&](int64_t it, int s) {       // lane 0 only
        const int64_t q0 = (first + it * G) * QPW;
        const int64_t left = n - q0;
        const uint32_t cnt = (uint32_t)(left < (int64_t)QPW ? left : (int64_t)QPW);
        const uint32_t bytes = cnt * 8u * (uint32_t)sizeof(T);
        unsigned char* dst = ring + s * L::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], bytes * L::NARR);
        if constexpr (SOLVER != SOLVER_RECT) {
            bulk_g2s(dst, src + q0 * 8, bytes, &full[s]);
            bulk_g2s(dst + L::ARR_BYTES, tar + q0 * 8, bytes, &full[s]);
        } else {
            bulk_g2s(dst, tar + q0 * 8, bytes, &full[s]);
        }
    };

    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s)
            mbar_init(&full[s], 1);
        mbar_init_fence();
        for (int s = 0; s < STAGES && s < my_tiles; ++s)
            issue(s, s);
    }
    __syncwarp();

    int s = 0;
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
        const int64_t q0 = (first + it * G) * QPW;
        const int64_t left = n - q0;
        const int cnt = (int)(left < (int64_t)QPW ? left : (int64_t)QPW);
        mbar_wait(&full[s], parity);            // this stage's bytes have landed
        T sv[PER_LANE][8], tv[PER_LANE][8];
        const unsigned char* st = ring + s * L::STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
            const int q = lane + 32 * j;
            if (q < cnt) {
                if constexpr (SOLVER != SOLVER_RECT) {
                    load_quad_smem<T>(st, q, sv[j]);
                    load_quad_smem<T>(st + L::ARR_BYTES, q, tv[j]);
                } else {
                    load_quad_smem<T>(st, q, tv[j]);
                }
            }
        }
        // the output buffer about to be overwritten was handed to the copy engine two tiles ago
        if (lane == 0)
            bulk_wait_read<1>();
        __syncwarp();                            // stage s fully read by the warp, out[it&1] free
        if (lane == 0 && it + STAGES < my_tiles) {
            fence_async_smem();                  // generic-proxy reads of the stage precede its refill
            issue(it + STAGES, s);
        }
        T* out = reinterpret_cast<T*>(outs + (size_t)(it & 1) * L::OUT_PAD);
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
            const int q = lane + 32 * j;
            if (q < cnt) {
                const int64_t i = q0 + q;
                T mx = rp.mx, my = rp.my;
                if constexpr (SOLVER == SOLVER_RECT)
                    if (M != nullptr)
                        load_corner_aos<T>(M, i, mx, my);
                T h[9];
                solve_quad<SOLVER, T>(sv[j], tv[j], mx, my, rp, h, normalize);
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    out[q * 9 + k] = h[k];       // stride 9 words: conflict-free
                if (degen != nullptr)
                    degen[i] = is_degenerate<T>(h, normalize) ? 1 : 0;
            }
        }
        fence_async_smem();                      // results visible to the async proxy
        __syncwarp();
        const uint32_t out_bytes = (uint32_t)cnt * 9u * (uint32_t)sizeof(T);
        if ((out_bytes & 15u) == 0) {
            if (lane == 0) {
                bulk_s2g(H + q0 * 9, out, out_bytes);
                bulk_commit();
            }
        } else {                                 // ragged last tile: not a 16-byte multiple
            for (int e = lane; e < cnt * 9; e += 32)
                H[q0 * 9 + e] = out[e];
            if (lane == 0)
                bulk_commit();                   // empty group keeps the one-group-per-tile count
            __syncwarp();
        }
        if (++s == STAGES) {
            s = 0;
            parity ^= 1;
        }
    }
    if (lane == 0)
        bulk_wait_all<0>();
}

// ------------------------------------------------------------------- k_soa
// VEC: 16-byte vector path (Q = 4 fp32 / 2 fp64 consecutive quadruples per
// thread); !VEC: one quadruple per thread, scalar accesses (any ld/alignment).
template <int SOLVER, typename T, int THREADS, bool VEC>
__global__ void __launch_bounds__(THREADS)
k_soa(const T* __restrict__ src, const T* __restrict__ tar, const T* __restrict__ M,
      RectParams<T> rp, T* __restrict__ H, uint8_t* __restrict__ degen, int64_t n, int64_t ld,
      bool normalize)
{
    constexpr int Q = VEC ? ChunkTraits<T>::EPC : 1;
    const int64_t i0 = ((int64_t)blockIdx.x * THREADS + threadIdx.x) * Q;
    if (i0 >= n)
        return;
    if (VEC && i0 + Q <= n) {
        Chunk16 cs[8], ct[8], cm[2];
        if constexpr (SOLVER != SOLVER_RECT) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                cs[k] = ldg_stream(src + k * ld + i0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            ct[k] = ldg_stream(tar + k * ld + i0);
        const bool per_m = (SOLVER == SOLVER_RECT) && M != nullptr;
        if (per_m) {
            cm[0] = ldg_stream(M + i0);
            cm[1] = ldg_stream(M + ld + i0);
        }
        Chunk16 co[9];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            T s[8], t[8], h[9];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if constexpr (SOLVER != SOLVER_RECT)
                    s[k] = ChunkTraits<T>::get(cs[k], q);
                t[k] = ChunkTraits<T>::get(ct[k], q);
            }
            T mx = rp.mx, my = rp.my;
            if (per_m) {
                mx = ChunkTraits<T>::get(cm[0], q);
                my = ChunkTraits<T>::get(cm[1], q);
            }
            solve_quad<SOLVER, T>(s, t, mx, my, rp, h, normalize);
#pragma unroll
            for (int k = 0; k < 9; ++k)
                ChunkTraits<T>::set(co[k], q, h[k]);
            if (degen != nullptr)
                degen[i0 + q] = is_degenerate<T>(h, normalize) ? 1 : 0;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k)
            stg_stream(H + k * ld + i0, co[k]);
    } else {
        const int64_t end = (i0 + Q < n) ? (i0 + Q) : n;
        for (int64_t i = i0; i < end; ++i) {
            T s[8], t[8], h[9];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if constexpr (SOLVER != SOLVER_RECT)
                    s[k] = __ldg(src + k * ld + i);
                t[k] = __ldg(tar + k * ld + i);
            }
            T mx = rp.mx, my = rp.my;
            if constexpr (SOLVER == SOLVER_RECT)
                if (M != nullptr) {
                    mx = __ldg(M + i);
                    my = __ldg(M + ld + i);
                }
            solve_quad<SOLVER, T>(s, t, mx, my, rp, h, normalize);
#pragma unroll
            for (int k = 0; k < 9; ++k)
                H[k * ld + i] = h[k];
            if (degen != nullptr)
                degen[i] = is_degenerate<T>(h, normalize) ? 1 : 0;
        }
    }
}

}  // namespace sksb
