// Micro-benchmark: what a 64 %-read / 36 %-write stream can reach on this B200,
// with no arithmetic at all -- the ceiling the streaming solvers are measured against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dram_ceiling dram_ceiling.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct __align__(32) C32 { uint32_t w[8]; };
struct __align__(16) C16 { uint32_t w[4]; };

__device__ __forceinline__ C32 ld32(const void* p)
{
    C32 c;
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3]), "=r"(c.w[4]), "=r"(c.w[5]),
                   "=r"(c.w[6]), "=r"(c.w[7]) : "l"(p));
    return c;
}
__device__ __forceinline__ void st16(void* p, C16 c)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(c.w[0]), "r"(c.w[1]),
                 "r"(c.w[2]), "r"(c.w[3]) : "memory");
}

// read-only: every thread reads 64 B (two arrays), folds them, one thread in 2^20 writes
__global__ void __launch_bounds__(256) k_read(const C32* a, const C32* b, uint32_t* out, size_t n)
{
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    C32 x = ld32(a + i), y = ld32(b + i);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x.w[k] ^ y.w[k];
    if (s == 0x12345678u) out[0] = s;
}
// write-only: 36 B per thread, coalesced 16-byte stores of a tile
__global__ void __launch_bounds__(256) k_write(C16* h, size_t n)
{
    size_t base = (size_t)blockIdx.x * 576;           // 256 * 36 / 16 chunks per tile
    for (int c = threadIdx.x; c < 576; c += 256) {
        C16 v = {{(uint32_t)c, 1u, 2u, 3u}};
        if (base + c < n * 36 / 16) st16(h + base + c, v);
    }
}
// the solver's traffic shape without the solver: 64 B in, 36 B out through smem
__global__ void __launch_bounds__(256) k_mix(const C32* a, const C32* b, C16* h, size_t n)
{
    __shared__ __align__(16) uint32_t stage[256 * 9];
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        C32 x = ld32(a + i), y = ld32(b + i);
#pragma unroll
        for (int k = 0; k < 8; ++k) stage[threadIdx.x * 9 + k] = x.w[k] + y.w[k];
        stage[threadIdx.x * 9 + 8] = x.w[0] ^ y.w[7];
    }
    __syncthreads();
    size_t base = (size_t)blockIdx.x * 576;
    for (int c = threadIdx.x; c < 576; c += 256)
        if (base + c < n * 36 / 16) st16(h + base + c, *reinterpret_cast<C16*>(stage + 4 * c));
}

// ACA-rect's shape: 32 B in (one array), 36 B out
__global__ void __launch_bounds__(256) k_mix_rect(const C32* a, C16* h, size_t n)
{
    __shared__ __align__(16) uint32_t stage[256 * 9];
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        C32 x = ld32(a + i);
#pragma unroll
        for (int k = 0; k < 8; ++k) stage[threadIdx.x * 9 + k] = x.w[k] + 1u;
        stage[threadIdx.x * 9 + 8] = x.w[0] ^ x.w[7];
    }
    __syncthreads();
    size_t base = (size_t)blockIdx.x * 576;
    for (int c = threadIdx.x; c < 576; c += 256)
        if (base + c < n * 36 / 16) st16(h + base + c, *reinterpret_cast<C16*>(stage + 4 * c));
}
// the fp64 solvers' shape: 128 B in (two arrays of 64 B), 72 B out; 128 quadruples per CTA
__global__ void __launch_bounds__(128) k_mix_f64(const C32* a, const C32* b, C16* h, size_t n)
{
    __shared__ __align__(16) uint32_t stage[128 * 18];
    size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (i < n) {
        C32 x0 = ld32(a + 2 * i), x1 = ld32(a + 2 * i + 1), y0 = ld32(b + 2 * i), y1 = ld32(b + 2 * i + 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            stage[threadIdx.x * 18 + k] = x0.w[k] + y0.w[k];
            stage[threadIdx.x * 18 + 8 + k] = x1.w[k] + y1.w[k];
        }
        stage[threadIdx.x * 18 + 16] = x0.w[0] ^ y1.w[7];
        stage[threadIdx.x * 18 + 17] = x1.w[0] ^ y0.w[7];
    }
    __syncthreads();
    size_t base = (size_t)blockIdx.x * 576;           // 128 * 72 / 16 chunks per tile
    for (int c = threadIdx.x; c < 576; c += 128)
        if (base + c < n * 72 / 16) st16(h + base + c, *reinterpret_cast<C16*>(stage + 4 * c));
}

template <typename F>
float time_ms(F f)
{
    for (int i = 0; i < 3; ++i) f();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 10;
}

int main()
{
    const size_t n = 1ull << 26;
    C32 *a, *b; C16* h; uint32_t* out;
    cudaMalloc(&a, n * 32); cudaMalloc(&b, n * 32); cudaMalloc(&h, n * 36); cudaMalloc(&out, 4);
    cudaMemset(a, 1, n * 32); cudaMemset(b, 2, n * 32);
    const unsigned grid = (unsigned)(n / 256);
    float t;
    t = time_ms([&] { k_read<<<grid, 256>>>(a, b, out, n); });
    printf("read-only   64 B/thread            %7.3f ms  %7.1f GB/s\n", t, n * 64 / t / 1e6);
    t = time_ms([&] { k_write<<<grid, 256>>>(h, n); });
    printf("write-only  36 B/thread            %7.3f ms  %7.1f GB/s\n", t, n * 36 / t / 1e6);
    t = time_ms([&] { k_mix<<<grid, 256>>>(a, b, h, n); });
    printf("64 B in + 36 B out, no arithmetic  %7.3f ms  %7.1f GB/s\n", t, n * 100 / t / 1e6);
    t = time_ms([&] { k_mix_rect<<<grid, 256>>>(a, h, n); });
    printf("32 B in + 36 B out (ACA-rect)      %7.3f ms  %7.1f GB/s\n", t, n * 68 / t / 1e6);
    {
        const size_t n64 = n / 2;                     // 2^25 fp64 quadruples in the same buffers
        t = time_ms([&] { k_mix_f64<<<(unsigned)(n64 / 128), 128>>>(a, b, h, n64); });
        printf("128 B in + 72 B out (fp64 solvers) %7.3f ms  %7.1f GB/s\n", t, n64 * 200 / t / 1e6);
    }
    t = time_ms([&] { cudaMemcpyAsync(h, a, n * 32, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D (read+write bytes)  %7.3f ms  %7.1f GB/s\n", t, 2.0 * n * 32 / t / 1e6);
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
