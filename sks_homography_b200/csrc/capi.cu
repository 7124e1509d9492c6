// libsks_cuda: the C ABI (include/sks_cuda.h) over the sm_100a kernels.
// Device-pointer entry points only enqueue work; nothing here synchronises,
// allocates per call, or falls back to the CPU.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <cuda_runtime.h>

#include "../../include/sks_cuda.h"
#include "mrg32k3a.cuh"
#include "peer.cuh"
#include "ransac.cuh"
#include "refit.cuh"
#include "stream_kernels.cuh"
#include "synth.cuh"
#include "warp.cuh"

using namespace sksb;

namespace {

std::atomic<int64_t> g_launches{0};
// Kernel tuning knobs (sks_cuda_set_variant / _tuning / _ransac_tuning).  They are PER HOST THREAD:
// a setting made by one thread never changes what another thread's calls launch (round-1 review: the
// knobs used to be process-global).  Worker threads the library starts itself (host_api.cu) inherit
// the caller's settings through sks_tuning_get_internal / sks_tuning_set_internal.
struct Tuning {
    int variant = 0;          // 0 default, 1 direct, 2 ring, 3 warp-private ring
    int tile_small = 0;       // ring tile: 0 = 256/128 (f32/f64), 1 = 128/64
    int stages = 4;
    int ctas_per_sm = 0;      // 0 = whatever the occupancy calculator allows
    int wide = 1;             // direct kernel: 256-bit loads/stores when 32-byte aligned
    int ransac_hpt = 2;       // hypotheses per thread in the RANSAC kernel (2 or 4)
    int ransac_packed = 3;    // scorer: 0 scalar FFMA, 1 FFMA2 over two matches, 2 FFMA2 over two hypotheses,
                              // 3 (default) = 1 with the inlier count on the FP32 pipe (FFMA2.RM)
    int ransac_threads = 256;
    int ransac_rounds = 8;    // rounds per CTA (chunk = rounds * 256 * hpt hypotheses)
};
thread_local Tuning t_tuning;

struct DevInfo {
    int sms = 0;
    int smem_optin = 0;
    bool ok = false;
};

int device_info(DevInfo& out)
{
    static DevInfo cache[64];
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? SKS_ERR_NO_DEVICE : (int)e;
    }
    std::lock_guard<std::mutex> lk(mu);
    DevInfo& d = cache[dev & 63];
    if (!d.ok) {
        if ((e = cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
            return (int)e;
        if ((e = cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess)
            return (int)e;
        d.ok = true;
    }
    out = d;
    return SKS_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

inline int finish_launch()
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SKS_OK : (int)e;
}

template <int SOLVER, typename T, int TILE>
int launch_ring(const T* src, const T* tar, const T* M, RectParams<T> rp, T* H, uint8_t* degen,
                int64_t n, bool normalize, const DevInfo& dev, cudaStream_t st)
{
    using L = RingLayout<SOLVER, T, TILE>;
    auto kern = k_aos_ring<SOLVER, T, TILE>;
    int stages = t_tuning.stages;
    if (stages < 2) stages = 2;
    while (stages > 2 && L::smem_bytes(stages) > dev.smem_optin) --stages;
    const int smem = L::smem_bytes(stages);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TILE, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) occ = 1;
    const int want = t_tuning.ctas_per_sm;
    if (want > 0 && want < occ) occ = want;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    const int64_t grid = n_tiles < (int64_t)dev.sms * occ ? n_tiles : (int64_t)dev.sms * occ;
    kern<<<(unsigned)grid, TILE, smem, st>>>(src, tar, M, rp, H, degen, n, stages, normalize);
    return finish_launch();
}

// variant 3: warp-private TMA ring.  Shapes measured on B200 (tools/wring_sweep.py,
// profiles/r02_wring_sweep.log; QPW quadruples per warp tile x stages x warps per CTA):
// fp32 64 x 2 x 16 (ACA 5.87 TB/s, rect 6.37), fp64 32 x 3 x 8 (SKS 6.57 TB/s) -- all bit-exact,
// all below the direct kernel's 7.0 TB/s, which therefore stays the default.
template <int SOLVER, typename T>
int launch_wring(const T* src, const T* tar, const T* M, RectParams<T> rp, T* H, uint8_t* degen,
                 int64_t n, bool normalize, const DevInfo& dev, cudaStream_t st)
{
#ifdef SKS_WRING_QPW_F32      // sweep builds: one shape for both precisions
    constexpr int QPW = sizeof(T) == 4 ? SKS_WRING_QPW_F32 : (SKS_WRING_QPW_F32 >= 64 ? SKS_WRING_QPW_F32 / 2 : 32),
                  STAGES = SKS_WRING_STAGES, WARPS = SKS_WRING_WARPS;
#else
    constexpr int QPW = sizeof(T) == 4 ? 64 : 32, STAGES = sizeof(T) == 4 ? 2 : 3, WARPS = sizeof(T) == 4 ? 16 : 8;
#endif
    using L = WringLayout<SOLVER, T, QPW, STAGES>;
    auto kern = k_aos_wring<SOLVER, T, WARPS, QPW, STAGES>;
    const int smem = L::smem_bytes(WARPS);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) occ = 1;
    const int want = t_tuning.ctas_per_sm;
    if (want > 0 && want < occ) occ = want;
    const int64_t n_tiles = (n + QPW - 1) / QPW;
    const int64_t ctas_needed = (n_tiles + WARPS - 1) / WARPS;
    const int64_t grid = ctas_needed < (int64_t)dev.sms * occ ? ctas_needed : (int64_t)dev.sms * occ;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(src, tar, M, rp, H, degen, n, normalize);
    return finish_launch();
}

template <int SOLVER, typename T>
int launch_stream(const T* src, const T* tar, const T* M, RectParams<T> rp, T* H, int64_t n,
                  int layout, int64_t ld, int flags, uint8_t* degen, void* stream)
{
    if (n < 0) return SKS_ERR_INVALID_ARG;
    if (layout != SKS_LAYOUT_AOS && layout != SKS_LAYOUT_SOA) return SKS_ERR_INVALID_ARG;
    if (flags & ~SKS_FLAG_NORMALIZE) return SKS_ERR_INVALID_ARG;
    if (n > 0 && (H == nullptr || tar == nullptr)) return SKS_ERR_INVALID_ARG;
    if (n > 0 && SOLVER != SOLVER_RECT && src == nullptr) return SKS_ERR_INVALID_ARG;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n == 0) return SKS_OK;   // empty batch: nothing to enqueue (buffers may be NULL)
    if (!aligned16(H) || !aligned16(tar) || (src && !aligned16(src)) || (M && !aligned16(M)))
        return SKS_ERR_UNALIGNED;
    if (n > ((int64_t)1 << 36)) return SKS_ERR_INVALID_ARG;   // grid.x limit (2^31 - 1 CTAs of >= 32 quadruples)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool normalize = (flags & SKS_FLAG_NORMALIZE) != 0;

    if (layout == SKS_LAYOUT_SOA) {
        if (ld == 0) ld = n;
        if (ld < n) return SKS_ERR_INVALID_ARG;
        constexpr int THREADS = 128;
        constexpr int Q = ChunkTraits<T>::EPC;
        if (ld % Q == 0) {
            const int64_t threads = (n + Q - 1) / Q;
            const int64_t grid = (threads + THREADS - 1) / THREADS;
            k_soa<SOLVER, T, THREADS, true><<<(unsigned)grid, THREADS, 0, st>>>(
                src, tar, M, rp, H, degen, n, ld, normalize);
        } else {
            const int64_t grid = (n + THREADS - 1) / THREADS;
            k_soa<SOLVER, T, THREADS, false><<<(unsigned)grid, THREADS, 0, st>>>(
                src, tar, M, rp, H, degen, n, ld, normalize);
        }
        return finish_launch();
    }

    int variant = t_tuning.variant;
    if (variant == 0) variant = 1;   // measured best on B200 (profiles/): direct > ring
#ifndef SKS_DIRECT_TILE_F32
#define SKS_DIRECT_TILE_F32 256
#endif
    constexpr int BIG = sizeof(T) == 4 ? SKS_DIRECT_TILE_F32 : 128;
    constexpr int SMALL = BIG / 2;
    if (variant == 1) {
        const int64_t grid = (n + BIG - 1) / BIG;
        const bool wide = t_tuning.wide && aligned32(H) && aligned32(tar) && (!src || aligned32(src));
        if (wide)
            k_aos_direct<SOLVER, T, BIG, true><<<(unsigned)grid, BIG, 0, st>>>(src, tar, M, rp, H,
                                                                               degen, n, normalize);
        else
            k_aos_direct<SOLVER, T, BIG, false><<<(unsigned)grid, BIG, 0, st>>>(src, tar, M, rp, H,
                                                                                degen, n, normalize);
        return finish_launch();
    }
    if (variant == 3) {
        if constexpr (SOLVER == SOLVER_GPT)        // 200 registers: stays on the direct kernel
            return SKS_ERR_INVALID_ARG;
        else
            return launch_wring<SOLVER, T>(src, tar, M, rp, H, degen, n, normalize, dev, st);
    }
    if (t_tuning.tile_small)
        return launch_ring<SOLVER, T, SMALL>(src, tar, M, rp, H, degen, n, normalize, dev, st);
    return launch_ring<SOLVER, T, BIG>(src, tar, M, rp, H, degen, n, normalize, dev, st);
}

template <int SOLVER, typename T>
int launch_gather_solve(const T* pool, uint32_t pool_size, const uint32_t* rand4, uint64_t seed, T* H,
                        int64_t n, int layout, int64_t ld, int flags, uint8_t* degen, void* stream)
{
    if (n < 0 || pool_size == 0 || (flags & ~SKS_FLAG_NORMALIZE)) return SKS_ERR_INVALID_ARG;
    if (layout != SKS_LAYOUT_AOS && layout != SKS_LAYOUT_SOA) return SKS_ERR_INVALID_ARG;
    if (n > 0 && (pool == nullptr || H == nullptr)) return SKS_ERR_INVALID_ARG;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n == 0) return SKS_OK;
    if (!aligned16(pool) || !aligned16(H)) return SKS_ERR_UNALIGNED;
    if (ld == 0) ld = n;
    if (ld < n) return SKS_ERR_INVALID_ARG;
    constexpr int TILE = sizeof(T) == 4 ? 256 : 128;
    const int64_t grid = (n + TILE - 1) / TILE;
    // SoA output with a pool that fits twice into an SM's shared memory: persistent CTAs gather
    // from shared memory (variant 2 of sks_cuda_set_variant forces the L1 path for comparison)
    const size_t pool_bytes = (size_t)pool_size * 4 * sizeof(T);
    if (layout == SKS_LAYOUT_SOA && pool_bytes <= 96u * 1024 && n >= (int64_t)dev.sms * 2048 &&
        t_tuning.variant != 2) {
        auto pk = k_gather_solve_pool<SOLVER, T>;
        cudaError_t e = cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pool_bytes);
        if (e != cudaSuccess) return (int)e;
        pk<<<(unsigned)(dev.sms * 2), 512, pool_bytes, static_cast<cudaStream_t>(stream)>>>(
            pool, pool_size, rand4, seed_key(seed), H, degen, n, ld, (flags & SKS_FLAG_NORMALIZE) != 0);
        return finish_launch();
    }
    const bool wide = sizeof(T) == 8 && aligned32(pool) && t_tuning.wide != 0;
    auto kern = wide ? k_gather_solve<SOLVER, T, TILE, true> : k_gather_solve<SOLVER, T, TILE, false>;
    kern<<<(unsigned)grid, TILE, 0, static_cast<cudaStream_t>(stream)>>>(
        pool, pool_size, rand4, seed_key(seed), H, degen, n, layout, ld,
        (flags & SKS_FLAG_NORMALIZE) != 0);
    return finish_launch();
}

template <typename T>
int launch_rect_planar(const T* tar34, const T* src34, RectParams<T> rp, T* H, int64_t n, int flags,
                       uint8_t* degen, void* stream, const T* width_dev = nullptr,
                       const T* ratio_dev = nullptr)
{
    if (n < 0 || (flags & ~SKS_FLAG_NORMALIZE)) return SKS_ERR_INVALID_ARG;
    if (n > 0 && (tar34 == nullptr || H == nullptr)) return SKS_ERR_INVALID_ARG;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n == 0) return SKS_OK;
    if (!aligned16(tar34) || !aligned16(H) || (src34 && !aligned16(src34))) return SKS_ERR_UNALIGNED;
    constexpr int TILE = sizeof(T) == 4 ? 256 : 128;
    const int64_t grid = (n + TILE - 1) / TILE;
    k_rect_planar34<T, TILE><<<(unsigned)grid, TILE, 0, static_cast<cudaStream_t>(stream)>>>(
        tar34, src34, rp, H, degen, n, (flags & SKS_FLAG_NORMALIZE) != 0, width_dev, ratio_dev);
    return finish_launch();
}

template <typename T>
int launch_synth_quads(T* src, T* tar, int64_t begin, int64_t count, uint64_t seed, int dist,
                       int layout, int64_t ld, void* stream)
{
    if (count < 0 || begin < 0 || src == nullptr || tar == nullptr) return SKS_ERR_INVALID_ARG;
    if (dist < 0 || dist > 2 || (layout != SKS_LAYOUT_AOS && layout != SKS_LAYOUT_SOA))
        return SKS_ERR_INVALID_ARG;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (count == 0) return SKS_OK;
    if (ld == 0) ld = count;
    const int64_t grid = (count + 255) / 256;
    k_synth_quads<T><<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, tar, begin, count, seed_key(seed), dist, layout, ld);
    return finish_launch();
}

template <typename T>
int launch_gather(const T* pool, uint32_t pool_size, const uint32_t* rand4, uint64_t seed, T* src,
                  T* tar, int64_t n, int layout, int64_t ld, void* stream)
{
    if (n < 0 || pool == nullptr || pool_size == 0 || src == nullptr || tar == nullptr)
        return SKS_ERR_INVALID_ARG;
    if (layout != SKS_LAYOUT_AOS && layout != SKS_LAYOUT_SOA) return SKS_ERR_INVALID_ARG;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n == 0) return SKS_OK;
    if (ld == 0) ld = n;
    const int64_t grid = (n + 255) / 256;
    k_gather_samples<T><<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pool, pool_size, rand4, seed_key(seed), src, tar, n, layout, ld);
    return finish_launch();
}

}  // namespace

extern "C" {

int sks_cuda_abi_version(void) { return SKS_CUDA_ABI_VERSION; }

const char* sks_cuda_error_string(int status)
{
    switch (status) {
        case SKS_OK: return "success";
        case SKS_ERR_INVALID_ARG: return "invalid argument";
        case SKS_ERR_UNALIGNED: return "pointer not 16-byte aligned";
        case SKS_ERR_NO_DEVICE: return "no CUDA device (libsks_cuda has no CPU fallback)";
        case SKS_ERR_NO_PEER_ACCESS: return "a device cannot access the primary device's memory (no P2P)";
        default: break;
    }
    return status > 0 ? cudaGetErrorString(static_cast<cudaError_t>(status)) : "unknown error";
}

int sks_cuda_device_count(int* count)
{
    if (count == nullptr) return SKS_ERR_INVALID_ARG;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? SKS_ERR_NO_DEVICE : (int)e;
    }
    return SKS_OK;
}

#define SKS_DEFINE_GENERAL(NAME, SOLVER, T)                                                     \
    int NAME(const T* src, const T* tar, T* H, int64_t n, int layout, int64_t ld, int flags,    \
             uint8_t* degenerate, void* stream)                                                 \
    {                                                                                           \
        return launch_stream<SOLVER, T>(src, tar, nullptr, RectParams<T>{}, H, n, layout, ld,   \
                                        flags, degenerate, stream);                             \
    }
SKS_DEFINE_GENERAL(sks_cuda_aca_f32, SOLVER_ACA, float)
SKS_DEFINE_GENERAL(sks_cuda_aca_f64, SOLVER_ACA, double)
SKS_DEFINE_GENERAL(sks_cuda_sks_f32, SOLVER_SKS, float)
SKS_DEFINE_GENERAL(sks_cuda_sks_f64, SOLVER_SKS, double)
SKS_DEFINE_GENERAL(sks_cuda_ge_f32, SOLVER_GE, float)
SKS_DEFINE_GENERAL(sks_cuda_ge_f64, SOLVER_GE, double)
SKS_DEFINE_GENERAL(sks_cuda_gpt_f64, SOLVER_GPT, double)

#define SKS_DEFINE_RECT(NAME, T)                                                                \
    int NAME(const T* tar, const T* M, T mx, T my, T width, T ratio, T* H, int64_t n,           \
             int layout, int64_t ld, int flags, uint8_t* degenerate, void* stream)              \
    {                                                                                           \
        return launch_stream<SOLVER_RECT, T>(nullptr, tar, M, RectParams<T>{mx, my, width, ratio}, \
                                             H, n, layout, ld, flags, degenerate, stream);      \
    }
SKS_DEFINE_RECT(sks_cuda_aca_rect_f32, float)
SKS_DEFINE_RECT(sks_cuda_aca_rect_f64, double)

int sks_cuda_aca_rect_planar_f32(const float* tar34, const float* src34, float mx, float my,
                                 float width, float ratio, float* H, int64_t n, int flags,
                                 uint8_t* degenerate, void* stream)
{
    return launch_rect_planar<float>(tar34, src34, RectParams<float>{mx, my, width, ratio}, H, n, flags,
                                     degenerate, stream);
}
int sks_cuda_aca_rect_planar_f64(const double* tar34, const double* src34, double mx, double my,
                                 double width, double ratio, double* H, int64_t n, int flags,
                                 uint8_t* degenerate, void* stream)
{
    return launch_rect_planar<double>(tar34, src34, RectParams<double>{mx, my, width, ratio}, H, n,
                                      flags, degenerate, stream);
}

int sks_cuda_aca_rect_planar_dev_f32(const float* tar34, const float* src34, const float* width_dev,
                                     const float* ratio_dev, float* H, int64_t n, int flags,
                                     uint8_t* degenerate, void* stream)
{
    if (width_dev == nullptr || ratio_dev == nullptr) return SKS_ERR_INVALID_ARG;
    return launch_rect_planar<float>(tar34, src34, RectParams<float>{0.f, 0.f, 0.f, 0.f}, H, n, flags, degenerate,
                                     stream, width_dev, ratio_dev);
}
int sks_cuda_aca_rect_planar_dev_f64(const double* tar34, const double* src34, const double* width_dev,
                                     const double* ratio_dev, double* H, int64_t n, int flags,
                                     uint8_t* degenerate, void* stream)
{
    if (width_dev == nullptr || ratio_dev == nullptr) return SKS_ERR_INVALID_ARG;
    return launch_rect_planar<double>(tar34, src34, RectParams<double>{0., 0., 0., 0.}, H, n, flags, degenerate,
                                      stream, width_dev, ratio_dev);
}

int sks_cuda_gather_samples_f32(const float* pool, uint32_t pool_size, const uint32_t* rand4,
                                uint64_t seed, float* src, float* tar, int64_t n, int layout,
                                int64_t ld, void* stream)
{
    return launch_gather<float>(pool, pool_size, rand4, seed, src, tar, n, layout, ld, stream);
}
int sks_cuda_gather_samples_f64(const double* pool, uint32_t pool_size, const uint32_t* rand4,
                                uint64_t seed, double* src, double* tar, int64_t n, int layout,
                                int64_t ld, void* stream)
{
    return launch_gather<double>(pool, pool_size, rand4, seed, src, tar, n, layout, ld, stream);
}

#define SKS_DEFINE_GATHER_SOLVE(NAME, SOLVER, T)                                                 \
    int NAME(const T* pool, uint32_t pool_size, const uint32_t* rand4, uint64_t seed, T* H,     \
             int64_t n, int layout, int64_t ld, int flags, uint8_t* degenerate, void* stream)   \
    {                                                                                           \
        return launch_gather_solve<SOLVER, T>(pool, pool_size, rand4, seed, H, n, layout, ld,    \
                                              flags, degenerate, stream);                       \
    }
SKS_DEFINE_GATHER_SOLVE(sks_cuda_gather_aca_f32, SOLVER_ACA, float)
SKS_DEFINE_GATHER_SOLVE(sks_cuda_gather_aca_f64, SOLVER_ACA, double)
SKS_DEFINE_GATHER_SOLVE(sks_cuda_gather_sks_f32, SOLVER_SKS, float)
SKS_DEFINE_GATHER_SOLVE(sks_cuda_gather_sks_f64, SOLVER_SKS, double)

namespace {
int launch_warp_grid(const float* H, const float* tar, const float* M, RectParams<float> rp,
                     int64_t n, float x0, float y0, float dx, float dy, int32_t gw, int32_t gh,
                     float* out, void* stream)
{
    if (n < 0 || gw <= 0 || gh <= 0 || out == nullptr || (H == nullptr && tar == nullptr))
        return SKS_ERR_INVALID_ARG;
    if (!aligned16(out) || (tar != nullptr && !aligned16(tar))) return SKS_ERR_UNALIGNED;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n == 0) return SKS_OK;
    const GridSpec g{x0, y0, dx, dy, gw, gh};
    const int64_t m = (int64_t)gw * gh;
    if (m >= (int64_t)1 << 31) return SKS_ERR_INVALID_ARG;
    // about 16-32 points per thread (4 per iteration), which keeps the per-warp cost of
    // re-solving the sample's homography below 10 % in the fused form: group = largest power
    // of two <= m/16 (1..256); beyond 256 x 32 points, several CTAs per sample
    WarpSplit ws{1, 1, 0, 0, (m % 4 == 0 && aligned32(out)) ? 1 : 0};
    while (ws.group < 256 && (int64_t)ws.group * 32 <= m) ws.group *= 2;
    if (ws.group == 256) ws.parts = (int32_t)((m + 8191) / 8192);
    const uint32_t stride = 4u * (uint32_t)ws.group * (uint32_t)ws.parts;
    ws.step_i = stride % (uint32_t)gw;
    ws.step_j = stride / (uint32_t)gw;
    const int64_t per_cta = 256 / ws.group;
    const int64_t ctas = ((n + per_cta - 1) / per_cta) * ws.parts;
    if (ctas >= (int64_t)1 << 31) return SKS_ERR_INVALID_ARG;
    if (H != nullptr)
        k_warp_grid<false><<<(unsigned)ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            H, nullptr, nullptr, rp, g, ws, out, n);
    else
        k_warp_grid<true><<<(unsigned)ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            nullptr, tar, M, rp, g, ws, out, n);
    if (int rc = finish_launch()) return rc;
    return SKS_OK;
}
}  // namespace

int sks_cuda_warp_grid_f32(const float* H, int64_t n, float x0, float y0, float dx, float dy,
                           int32_t gw, int32_t gh, float* grid_xy, void* stream)
{
    if (H == nullptr) return SKS_ERR_INVALID_ARG;
    return launch_warp_grid(H, nullptr, nullptr, RectParams<float>{}, n, x0, y0, dx, dy, gw, gh, grid_xy,
                            stream);
}

int sks_cuda_aca_rect_warp_grid_f32(const float* tar, const float* M, float mx, float my, float width,
                                    float ratio, int64_t n, float x0, float y0, float dx, float dy,
                                    int32_t gw, int32_t gh, float* grid_xy, void* stream)
{
    if (tar == nullptr) return SKS_ERR_INVALID_ARG;
    return launch_warp_grid(nullptr, tar, M, RectParams<float>{mx, my, width, ratio}, n, x0, y0, dx, dy,
                            gw, gh, grid_xy, stream);
}

int sks_cuda_curand_mrg32k3a_u32(uint32_t* out, int64_t n, uint64_t seed, void* stream)
{
    if (n < 0 || (n > 0 && out == nullptr)) return SKS_ERR_INVALID_ARG;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n == 0) return SKS_OK;
    static const MrgJumpTable table = [] {
        MrgJumpTable t;
        mrg_build_jump_table(t);
        return t;
    }();
    const int64_t threads = n < kMrgStreams ? n : kMrgStreams;
    k_mrg32k3a<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        out, n, seed, table);
    return finish_launch();
}

// Hypothesis ids per CTA.  Pure host arithmetic, exported so that it can be tested without a GPU.
uint32_t sks_cuda_ransac_chunk_plan(int64_t n_pairs, uint32_t hyp_count, int threads, int hyps_per_thread,
                                    int max_rounds, int resident_ctas)
{
    const uint32_t round = (uint32_t)threads * (uint32_t)hyps_per_thread;
    if (max_rounds < 1) max_rounds = 1;
    const double slots = resident_ctas > 0 ? (double)resident_ctas : 1.0;
    uint32_t chunk = round * (uint32_t)max_rounds;
    double best_cost = 0;
    for (uint32_t r = (uint32_t)max_rounds, first = 1; r >= 1; r /= 2, first = 0) {
        const double ctas = (double)(((uint64_t)hyp_count + (uint64_t)r * round - 1) / ((uint64_t)r * round)) *
                            (double)n_pairs;
        double waves = (double)(int64_t)((ctas + slots - 1) / slots);
        if (waves < 1) waves = 1;
        const double cost = waves * ((double)r + 0.03);
        if (first || cost < best_cost * 0.995) {     // prefer more rounds per CTA unless it pays clearly
            best_cost = cost;
            chunk = r * round;
        }
        if (r == 1) break;
    }
    return chunk;
}

int sks_cuda_ransac_aca_shard_f32(const float* corr, int64_t pair_begin, int64_t n_pairs, int32_t n_pts,
                                  const uint32_t* samples, uint32_t hyp_stride, uint32_t hyp_begin,
                                  uint32_t hyp_count, uint64_t seed, float thr2,
                                  unsigned long long* best_key, void* stream)
{
    if (pair_begin < 0) return SKS_ERR_INVALID_ARG;
    if (corr == nullptr || best_key == nullptr || n_pairs < 0 || n_pts <= 0)
        return SKS_ERR_INVALID_ARG;
    if (samples != nullptr && ((uint64_t)hyp_stride < (uint64_t)hyp_begin + hyp_count || !aligned16(samples)))
        return SKS_ERR_INVALID_ARG;
    if (!aligned16(corr)) return SKS_ERR_UNALIGNED;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n_pairs == 0 || hyp_count == 0) return SKS_OK;
    const int32_t tile_pts = n_pts < kRansacMaxTilePts ? n_pts : kRansacMaxTilePts;
    const int smem = ((tile_pts + 1) & ~1) * 16;
    const int hpt = t_tuning.ransac_hpt;                       // 2 or "4" (the HI instantiation)
    const int mode = t_tuning.ransac_packed;
    const int threads = t_tuning.ransac_threads;
    using Kern = void (*)(const float4*, int64_t, int32_t, int32_t, const uint32_t*, uint32_t, uint32_t,
                          uint32_t, uint32_t, uint64_t, float, unsigned long long*, int64_t);
#ifndef SKS_RANSAC_HPT_HI
#define SKS_RANSAC_HPT_HI 4      // the "4 hypotheses per thread" instantiation (sweep builds try 3)
#endif
    constexpr int HI = SKS_RANSAC_HPT_HI, HI2 = (SKS_RANSAC_HPT_HI % 2) ? 2 : SKS_RANSAC_HPT_HI;
#define SKS_RANSAC_PICK(T)                                                                        \
    (mode == 3 ? (hpt == 4 ? k_ransac_aca<HI, 3, T> : k_ransac_aca<2, 3, T>)                      \
     : mode == 2 ? (hpt == 4 ? k_ransac_aca<HI2, 2, T> : k_ransac_aca<2, 2, T>)                   \
     : mode == 1 ? (hpt == 4 ? k_ransac_aca<HI, 1, T> : k_ransac_aca<2, 1, T>)                    \
                 : (hpt == 4 ? k_ransac_aca<HI, 0, T> : k_ransac_aca<2, 0, T>))
    const Kern kern = threads == 512 ? SKS_RANSAC_PICK(512)
                    : threads == 384 ? SKS_RANSAC_PICK(384) : SKS_RANSAC_PICK(256);
#undef SKS_RANSAC_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    // chunk = rounds * (threads * hpt) hypothesis ids per CTA.  CTAs of one launch all take the
    // same time, so the launch runs in ceil(CTAs / resident slots) waves: 1024 pairs x 2 chunks on
    // 444 slots is 4.6 waves of work but 5 waves of time (8 % lost at 8 ranks, measured 13.0 ms =
    // 5 x 2.6 ms).  Pick the rounds per CTA -- the tuning value, halved down to 1 -- that
    // minimises waves x (rounds + fixed cost), the fixed cost being the tile load and re-layout
    // (~10 us against ~330 us per round at 4096 matches, i.e. ~0.03 rounds).
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) occ = 1;
    const uint32_t chunk = sks_cuda_ransac_chunk_plan(n_pairs, hyp_count, threads, hpt == 4 ? HI : 2, t_tuning.ransac_rounds,
                                                      dev.sms * occ);
    const unsigned chunks = (hyp_count + chunk - 1) / chunk;
    // grid.y is limited to 65535: larger batches go out in blocks of pairs; the kernel
    // takes its pair id from pair_base + blockIdx.y so sampling does not depend on blocking
    for (int64_t p0 = 0; p0 < n_pairs; p0 += 65535) {
        const int64_t np = n_pairs - p0 < 65535 ? n_pairs - p0 : 65535;
        dim3 grid(chunks, (unsigned)np);
        kern<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4*>(corr), p0, n_pts, tile_pts, samples, hyp_stride, hyp_begin,
            hyp_count, chunk, seed_key(seed), thr2, best_key, pair_begin);
        if (int rc = finish_launch()) return rc;
    }
    return SKS_OK;
}

int sks_cuda_ransac_aca_f32(const float* corr, int64_t n_pairs, int32_t n_pts,
                            const uint32_t* samples, uint32_t hyp_stride, uint32_t hyp_begin,
                            uint32_t hyp_count, uint64_t seed, float thr2,
                            unsigned long long* best_key, void* stream)
{
    return sks_cuda_ransac_aca_shard_f32(corr, 0, n_pairs, n_pts, samples, hyp_stride, hyp_begin, hyp_count,
                                         seed, thr2, best_key, stream);
}

int sks_cuda_ransac_finalize_f32(const float* corr, int64_t n_pairs, int32_t n_pts,
                                 const uint32_t* samples, uint32_t hyp_stride, uint64_t seed,
                                 float thr2, const unsigned long long* best_key, float* H_best,
                                 uint32_t* inlier_count, uint8_t* inlier_mask, void* stream)
{
    return sks_cuda_ransac_finalize_shard_f32(corr, 0, n_pairs, n_pts, samples, hyp_stride, seed, thr2,
                                              best_key, H_best, inlier_count, inlier_mask, stream);
}

int sks_cuda_ransac_finalize_shard_f32(const float* corr, int64_t pair_begin, int64_t n_pairs,
                                       int32_t n_pts, const uint32_t* samples, uint32_t hyp_stride,
                                       uint64_t seed, float thr2, const unsigned long long* best_key,
                                       float* H_best, uint32_t* inlier_count, uint8_t* inlier_mask,
                                       void* stream)
{
    if (pair_begin < 0) return SKS_ERR_INVALID_ARG;
    if (corr == nullptr || best_key == nullptr || H_best == nullptr || n_pairs < 0 || n_pts <= 0)
        return SKS_ERR_INVALID_ARG;
    if (!aligned16(corr) || (samples && !aligned16(samples))) return SKS_ERR_UNALIGNED;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n_pairs == 0) return SKS_OK;
    k_ransac_finalize<<<(unsigned)n_pairs, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(corr), n_pts, samples, hyp_stride, seed_key(seed), thr2,
        best_key, H_best, inlier_count, inlier_mask, pair_begin, nullptr);
    return finish_launch();
}

int sks_cuda_ransac_score_f32(const float* corr, int64_t n_pairs, int32_t n_pts, const float* H, float thr2,
                              uint32_t* inlier_count, uint8_t* inlier_mask, void* stream)
{
    if (corr == nullptr || H == nullptr || n_pairs < 0 || n_pts <= 0) return SKS_ERR_INVALID_ARG;
    if (!aligned16(corr)) return SKS_ERR_UNALIGNED;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n_pairs == 0) return SKS_OK;
    k_ransac_finalize<<<(unsigned)n_pairs, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(corr), n_pts, nullptr, 0, 0, thr2, nullptr, nullptr, inlier_count,
        inlier_mask, 0, H);
    return finish_launch();
}

int sks_cuda_ransac_refit_f32(const float* corr, int64_t n_pairs, int32_t n_pts, const uint8_t* inlier_mask,
                              const float* H_in, float* H_out, uint32_t* n_used, void* stream)
{
    if (corr == nullptr || inlier_mask == nullptr || H_in == nullptr || H_out == nullptr || n_pairs < 0 ||
        n_pts <= 0)
        return SKS_ERR_INVALID_ARG;
    if (!aligned16(corr)) return SKS_ERR_UNALIGNED;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n_pairs == 0) return SKS_OK;
    k_ransac_refit<<<(unsigned)((n_pairs + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(corr), n_pairs, n_pts, inlier_mask, H_in, H_out, n_used);
    return finish_launch();
}

int sks_cuda_synth_quads_f32(float* src, float* tar, int64_t begin, int64_t count, uint64_t seed,
                             int dist, int layout, int64_t ld, void* stream)
{
    return launch_synth_quads<float>(src, tar, begin, count, seed, dist, layout, ld, stream);
}
int sks_cuda_synth_quads_f64(double* src, double* tar, int64_t begin, int64_t count,
                             uint64_t seed, int dist, int layout, int64_t ld, void* stream)
{
    return launch_synth_quads<double>(src, tar, begin, count, seed, dist, layout, ld, stream);
}

int sks_cuda_synth_corr_f32(float* corr, int64_t pair_begin, int64_t n_pairs, int32_t n_pts,
                            uint64_t seed, int inlier_permille, float noise, void* stream)
{
    if (corr == nullptr || n_pairs < 0 || n_pts <= 0 || pair_begin < 0) return SKS_ERR_INVALID_ARG;
    if (!aligned16(corr)) return SKS_ERR_UNALIGNED;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (n_pairs == 0) return SKS_OK;
    const int64_t total = n_pairs * (int64_t)n_pts;
    k_synth_corr<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float4*>(corr), pair_begin, n_pairs, n_pts, seed_key(seed),
        inlier_permille, noise);
    return finish_launch();
}

// ---- NVLink peer exchange (hand-written max-reduce of the RANSAC keys) ----------
int sks_cuda_peer_alloc(void** block, int64_t n_keys)
{
    if (block == nullptr || n_keys <= 0) return SKS_ERR_INVALID_ARG;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    const size_t bytes = sizeof(PeerBlock) + 2 * (size_t)n_keys * sizeof(unsigned long long);
    cudaError_t e = cudaMalloc(block, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*block, 0, bytes);
    if (e != cudaSuccess) return (int)e;
    const unsigned long long nk = (unsigned long long)n_keys;
    e = cudaMemcpy(&static_cast<PeerBlock*>(*block)->n_keys, &nk, sizeof nk, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaDeviceSynchronize();
}
int sks_cuda_peer_free(void* block) { return (int)cudaFree(block); }
int sks_cuda_peer_export(void* block, void* handle64)
{
    if (block == nullptr || handle64 == nullptr) return SKS_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    return (int)cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), block);
}
int sks_cuda_peer_open(const void* handle64, void** block)
{
    if (block == nullptr || handle64 == nullptr) return SKS_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    return (int)cudaIpcOpenMemHandle(block, h, cudaIpcMemLazyEnablePeerAccess);
}
int sks_cuda_peer_close(void* block) { return (int)cudaIpcCloseMemHandle(block); }

int sks_cuda_peer_push_max(const unsigned long long* local_keys, int64_t n_keys,
                           void* const* peer_blocks, int world, int rank, uint64_t epoch,
                           void* stream)
{
    if (local_keys == nullptr || peer_blocks == nullptr || n_keys <= 0 || world < 1 ||
        world > kMaxPeers || rank < 0 || rank >= world)
        return SKS_ERR_INVALID_ARG;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    PeerTable tab{};
    for (int g = 0; g < world; ++g) {
        if (peer_blocks[g] == nullptr) return SKS_ERR_INVALID_ARG;
        tab.blk[g] = static_cast<PeerBlock*>(peer_blocks[g]);
    }
    const unsigned grid = (unsigned)((n_keys + 255) / 256);
    k_peer_push_max<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(local_keys, n_keys, tab, world,
                                                                         rank, epoch);
    return finish_launch();
}

int sks_cuda_peer_wait(void* own_block, int world, uint64_t epoch, unsigned long long* keys_out,
                       int64_t n_keys, int* status_dev, double timeout_s, void* stream)
{
    if (own_block == nullptr || keys_out == nullptr || n_keys <= 0 || world < 1)
        return SKS_ERR_INVALID_ARG;
    DevInfo dev;
    if (int rc = device_info(dev)) return rc;
    if (timeout_s <= 0) timeout_s = 5.0;
    const long long cycles = (long long)(timeout_s * 1.9e9);
    k_peer_wait_copy<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<PeerBlock*>(own_block), world, epoch, keys_out, n_keys, status_dev, cycles);
    return finish_launch();
}

int sks_cuda_shard_range(int64_t n, int rank, int world, int64_t* begin, int64_t* count)
{
    if (n < 0 || world <= 0 || rank < 0 || rank >= world || begin == nullptr || count == nullptr)
        return SKS_ERR_INVALID_ARG;
    const int64_t base = n / world, rem = n % world;
    *begin = base * rank + (rank < rem ? rank : rem);
    *count = base + (rank < rem ? 1 : 0);
    return SKS_OK;
}

int64_t sks_cuda_launch_count(void) { return g_launches.load(); }
void sks_cuda_reset_launch_count(void) { g_launches.store(0); }

// for library-owned worker threads (host_api.cu): copy the calling thread's knobs
void sks_tuning_get_internal(int* out9)
{
    const Tuning& t = t_tuning;
    const int v[9] = {t.variant, t.tile_small, t.stages, t.ctas_per_sm, t.wide, t.ransac_hpt, t.ransac_packed,
                      t.ransac_threads, t.ransac_rounds};
    for (int i = 0; i < 9; ++i) out9[i] = v[i];
}
void sks_tuning_set_internal(const int* in9)
{
    t_tuning = Tuning{in9[0], in9[1], in9[2], in9[3], in9[4], in9[5], in9[6], in9[7], in9[8]};
}

int sks_cuda_set_variant(int variant)
{
    if (variant < 0 || variant > 3) return SKS_ERR_INVALID_ARG;
    t_tuning.variant = variant;
    return SKS_OK;
}
int sks_cuda_get_variant(void) { return t_tuning.variant; }

int sks_cuda_set_ransac_tuning(int hyps_per_thread, int rounds_per_cta, int packed)
{
    if ((hyps_per_thread != 2 && hyps_per_thread != 4) || rounds_per_cta < 1 || rounds_per_cta > 1024)
        return SKS_ERR_INVALID_ARG;
    t_tuning.ransac_hpt = hyps_per_thread;
    t_tuning.ransac_rounds = rounds_per_cta;
    t_tuning.ransac_packed = packed & 3;
    const int t = packed >> 2;                       // bits 2.. select the CTA size
    t_tuning.ransac_threads = t == 1 ? 384 : t == 2 ? 512 : 256;
    return SKS_OK;
}

int sks_cuda_set_tuning(int small_tile, int stages, int ctas_per_sm)
{
    if (stages < 2 || stages > 16 || ctas_per_sm < 0) return SKS_ERR_INVALID_ARG;
    t_tuning.wide = (small_tile & 2) ? 0 : 1;          // bit 1: force 16-byte accesses in the direct kernel
    t_tuning.tile_small = small_tile & 1;
    t_tuning.stages = stages;
    t_tuning.ctas_per_sm = ctas_per_sm;
    return SKS_OK;
}

}  // extern "C"
