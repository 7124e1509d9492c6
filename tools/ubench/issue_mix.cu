// Micro-benchmark: do ALU-pipe (LEA.HI) and LSU (LDS) instructions take issue slots away
// from a stream of packed FP32 FFMA2 on B200?  Each FFMA2 keeps the FMA pipe busy for two
// cycles; if the scheduler can slip other pipes' instructions into the second cycle, the
// FFMA2 rate is unaffected.  Shapes follow the RANSAC scorer (csrc/ransac.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

// MODE 0: 8 FFMA2  C[i] = a[i].F32 * B0 + C[i]           (scalar x shared pair + pair)
// MODE 1: + 2 LEA.HI per 8 FFMA2 (count sign bits of results)
// MODE 2: + 4 LEA.HI     MODE 3: + 8 LEA.HI
// MODE 4: + 1 broadcast LDS.128 per 8 FFMA2 feeding B0
// MODE 5: 8 FFMA2  C[i] = A[i] * b0.F32 + C[i]           (pair x shared scalar + pair)
// MODE 6: 8 FFMA2  C[i] = A[i] * b[i].F32 + C[i]         (pair x own scalar + pair, no reuse)
// MODE 7: MODE 5 + 2 LEA.HI + LDS.128 every 11 FFMA2 (the scorer's mix)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, const float4* __restrict__ gsrc)
{
    __shared__ float4 tile[256];
    tile[threadIdx.x] = gsrc[threadIdx.x];
    __syncthreads();
    float a[8], b[8];
    float2 A[8], B[8], C[8];
    uint32_t cnt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed + i * 1e-3f + threadIdx.x * 1e-6f;
        b[i] = seed - i * 1e-4f * (1 + (threadIdx.x & 3));
        A[i] = make_float2(a[i], a[i] + 1e-5f);
        B[i] = make_float2(b[i], b[i] - 1e-5f);
        C[i] = make_float2((float)i, i + 1.f);
        cnt[i] = 0;
    }
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 4 || MODE == 7) {
            const float4 t = tile[it & 255];
            B[0] = make_float2(t.x, t.y);
            b[0] = t.z;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE <= 4) C[i] = __ffma2_rn(make_float2(a[i], a[i]), B[0], C[i]);
            if (MODE == 5 || MODE == 7) C[i] = __ffma2_rn(A[i], make_float2(b[0], b[0]), C[i]);
            if (MODE == 6) C[i] = __ffma2_rn(A[i], make_float2(b[i], b[i]), C[i]);
        }
        const int nlea = MODE == 1 ? 2 : MODE == 2 ? 4 : MODE == 3 ? 8 : MODE == 7 ? 2 : 0;
#pragma unroll
        for (int i = 0; i < nlea; ++i)
            cnt[i] += __float_as_uint(i & 1 ? C[i].y : C[i].x) >> 31;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += C[i].x + C[i].y + (float)cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8;
    float* out; cudaMalloc(&out, blocks * 256 * sizeof(float));
    float4* src; cudaMalloc(&src, 256 * sizeof(float4)); cudaMemset(src, 0, 256 * sizeof(float4));
    k<MODE><<<blocks, 256>>>(out, 1.0f, src);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<blocks, 256>>>(out, 1.0f, src);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double ffma2 = (double)blocks * 256 * ITERS * 8;
    const double cyc_per_ffma2 = (ms * 1e-3) * (clk * 1e3) * sms * 4 / (ffma2 / 32);
    printf("%-58s %8.3f ms  %6.3f SMSP-cycles per FFMA2  (%5.1f FMA/clk/SM of 128)\n", name, ms,
           cyc_per_ffma2, 2 * ffma2 / (ms * 1e-3) / (clk * 1e3) / sms);
    cudaFree(out); cudaFree(src);
}

int main()
{
    run<0>("8 FFMA2 (scalar x shared pair + pair)");
    run<1>("8 FFMA2 + 2 LEA.HI");
    run<2>("8 FFMA2 + 4 LEA.HI");
    run<3>("8 FFMA2 + 8 LEA.HI");
    run<4>("8 FFMA2 + 1 broadcast LDS.128");
    run<5>("8 FFMA2 (pair x shared scalar + pair)");
    run<6>("8 FFMA2 (pair x own scalar + pair)");
    run<7>("8 FFMA2 (pair x shared scalar) + 2 LEA.HI + LDS.128");
    return 0;
}
