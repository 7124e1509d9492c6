// Thin inline-PTX layer for sm_100a: streaming vector loads/stores, mbarrier,
// and the 1-D bulk asynchronous copy engine (cp.async.bulk, the non-tensor
// form of TMA; SASS UBLKCP).  Nothing here is portable below sm_90.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sksb {

__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- 16-byte streaming global accesses ------------------------------------
struct __align__(16) Chunk16 {
    uint32_t w[4];
};

// Experiment knobs (tools/hint_sweep.py builds the library with -DSKS_LD_HINT=n
// -DSKS_ST_HINT=m); the defaults are the measured best.
#ifndef SKS_LD_HINT
#define SKS_LD_HINT 0
#endif
#ifndef SKS_ST_HINT
#define SKS_ST_HINT 0
#endif

// read-only path, keep the line in L1 (two lanes-halves of one sector are read
// by two consecutive instructions in the AoS direct kernel)
__device__ __forceinline__ Chunk16 ldg_nc(const void* p)
{
    Chunk16 c;
#if SKS_LD_HINT == 1
    asm volatile("ld.global.nc.L2::evict_first.v4.u32 {%0,%1,%2,%3}, [%4];"
#elif SKS_LD_HINT == 2
    asm volatile("ld.global.nc.L1::evict_last.L2::evict_first.v4.u32 {%0,%1,%2,%3}, [%4];"
#elif SKS_LD_HINT == 3
    asm volatile("ld.global.nc.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
#else
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
#endif
                 : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3])
                 : "l"(p));
    return c;
}
// read-once data: do not allocate in L1
__device__ __forceinline__ Chunk16 ldg_stream(const void* p)
{
    Chunk16 c;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3])
                 : "l"(p));
    return c;
}
// read-once 32-bit value (sample lists): keep it out of L1, which holds the gather table
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(void* p, const Chunk16& c)
{
#if SKS_ST_HINT == 1
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(c.w[0]),
#elif SKS_ST_HINT == 2
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(c.w[0]),
#elif SKS_ST_HINT == 3
    asm volatile("st.global.L2::evict_first.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(c.w[0]),
#else
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(c.w[0]),
#endif
                 "r"(c.w[1]), "r"(c.w[2]), "r"(c.w[3])
                 : "memory");
}
// ---- 32-byte (256-bit) global accesses: new on sm_100 (SASS LDG.256 / STG.256).
// One instruction moves a whole 32-byte sector per lane, which is exactly one
// fp32 AoS quadruple: no half-sector re-reads, so nothing needs to stay in L1
// (L1::no_allocate).  The hint variants below were swept on B200: loads are
// insensitive (+-1 %), 256-bit STORES cost 4-10 % on the write-heavy solvers,
// so results leave as 16-byte stores.
struct __align__(32) Chunk32 {
    uint32_t w[8];
};
// measured on B200 (profiles/r01_hint_sweep.log): no_allocate loads + 16-byte stores
#ifndef SKS_WLD_HINT
#define SKS_WLD_HINT 1
#endif
#ifndef SKS_WST_HINT
#define SKS_WST_HINT 0
#endif
__device__ __forceinline__ Chunk32 ldg_stream32(const void* p)
{
    Chunk32 c;
    asm volatile(
#if SKS_WLD_HINT == 0
        "ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#elif SKS_WLD_HINT == 1
        "ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
        "ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
        : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3]), "=r"(c.w[4]), "=r"(c.w[5]),
          "=r"(c.w[6]), "=r"(c.w[7])
        : "l"(p));
    return c;
}
// 256-bit load through the read-only path that DOES allocate in L1: one whole 32-byte sector
// per lane, for gathers from a small table that should stay resident (fp64 match pool)
__device__ __forceinline__ Chunk32 ldg_nc32(const void* p)
{
    Chunk32 c;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3]), "=r"(c.w[4]), "=r"(c.w[5]),
                   "=r"(c.w[6]), "=r"(c.w[7])
                 : "l"(p));
    return c;
}
__device__ __forceinline__ void stg_stream32(void* p, const Chunk32& c)
{
    asm volatile(
#if SKS_WST_HINT == 1
        "st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
#elif SKS_WST_HINT == 2
        "st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
#else
        "st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
#endif
        "r"(c.w[0]), "r"(c.w[1]), "r"(c.w[2]), "r"(c.w[3]), "r"(c.w[4]), "r"(c.w[5]), "r"(c.w[6]),
        "r"(c.w[7])
        : "memory");
}

__device__ __forceinline__ Chunk16 lds16(const void* p)
{
    Chunk16 c;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3])
                 : "r"(smem_addr(p)));
    return c;
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals)
                 : "memory");
}
// make freshly initialised barriers visible to the async proxy
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- 1-D bulk async copies (TMA engine, no tensor map) ----------------------
// global -> shared, completion signalled on an mbarrier as transaction bytes.
// size and both addresses must be multiples of 16 bytes.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}
// shared -> global, tracked by bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_addr(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until at most N committed groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// order generic-proxy shared-memory writes before async-proxy reads of them
__device__ __forceinline__ void fence_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace sksb
