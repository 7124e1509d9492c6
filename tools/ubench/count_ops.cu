// Micro-benchmark (round 2): what does it cost to COUNT sign bits next to a packed FP32
// stream on B200, and what do register-bank conflicts of FFMA2 cost with and without the
// operand-reuse cache?  Shapes follow the RANSAC scorer (csrc/ransac.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o count_ops count_ops.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__device__ __forceinline__ float2 ffma2_rd_imm(float2 a, float2 c)
{
    return __ffma2_rd(a, make_float2(__uint_as_float(1u), __uint_as_float(1u)), c);
}

// MODE 0: 8 FFMA2   C[i] = a[i].F32 * B0 + C[i]                   (reuse-friendly baseline)
// MODE 1: 8 FFMA2.RM C[i] = A[i] * imm + C[i]                      (the FP-pipe count op alone)
// MODE 2: 8 FFMA2.RN C[i] = A[i] * imm + C[i]                      (same shape, round to nearest)
// MODE 3: 8 scalar FFMA.RM c[i] = a[i] * imm + c[i]
// MODE 4: 8 FFMA2 (baseline) + 1 FFMA2.RM                           (11:1-like mix, 8:1 here)
// MODE 5: 8 FFMA2 (baseline) + 2 LEA.HI
// MODE 6: 8 FFMA2  C[i] = a[i].F32 * B[i] + C[i]    scalar, own pair, own pair (5 registers, no reuse)
// MODE 7: 8 FFMA2  C[i] = a[i].F32 * B0 + C[i] with B0 re-loaded every iteration (reuse across 8)
// MODE 8: 8 FFMA2  C[i] = A[i] * B[i] + C[i]        three own pairs (6 registers)
// MODE 9: 8 FFMA2  C[i] = A[i] * B0 + C[i]          pair, shared pair (reuse), pair
// MODE 10: 8 FFMA2 + 2 IMAD.HI-style counts (FMA-pipe integer)
// MODE 11: 8 FFMA2 + 2 FSETP/SEL-free predicate counts: cnt += (C < 0) via ISETP on the bits + IADD with predicate
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* __restrict__ in)
{
    float a[8], c[8];
    float2 A[8], B[8], C[8];
    uint32_t cnt[8];
    const float* my = in + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 56;   // runtime data: nothing folds
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = my[i];
        c[i] = my[8 + i];
        A[i] = make_float2(my[16 + i], my[24 + i]);
        B[i] = make_float2(my[32 + i], my[40 + i]);
        C[i] = make_float2(my[48 + i], my[48 + ((i + 3) & 7)]);
        cnt[i] = 0;
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 4 || MODE == 5 || MODE == 7 || MODE == 10 || MODE == 11)
                C[i] = __ffma2_rn(make_float2(a[i], a[i]), B[0], C[i]);
            if (MODE == 1) C[i] = ffma2_rd_imm(A[i], C[i]);
            if (MODE == 2) C[i] = __ffma2_rn(A[i], make_float2(__uint_as_float(1u), __uint_as_float(1u)), C[i]);
            if (MODE == 3) c[i] = __fmaf_rd(a[i], __uint_as_float(1u), c[i]);
            if (MODE == 6) C[i] = __ffma2_rn(make_float2(a[i], a[i]), B[i], C[i]);
            if (MODE == 8) C[i] = __ffma2_rn(A[i], B[i], C[i]);
            if (MODE == 9) C[i] = __ffma2_rn(A[i], B[0], C[i]);
        }
        if (MODE == 4) A[0] = ffma2_rd_imm(C[0], A[0]);
        if (MODE == 5) {
            cnt[0] += __float_as_uint(C[0].x) >> 31;
            cnt[1] += __float_as_uint(C[1].y) >> 31;
        }
        if (MODE == 10) {
            cnt[0] = __umulhi(__float_as_uint(C[0].x), 2u) + cnt[0];
            cnt[1] = __umulhi(__float_as_uint(C[1].y), 2u) + cnt[1];
        }
        if (MODE == 11) {
            if ((int)__float_as_uint(C[0].x) < 0) cnt[0]++;
            if ((int)__float_as_uint(C[1].y) < 0) cnt[1]++;
        }
        if (MODE == 7) B[0].x += 1.0f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i] + C[i].x + C[i].y + A[i].x + (float)cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8;
    float* out; cudaMalloc(&out, blocks * 256 * sizeof(float));
    float* in; cudaMalloc(&in, (size_t)blocks * 256 * 56 * sizeof(float));
    {
        const size_t n = (size_t)blocks * 256 * 56;
        float* h = (float*)malloc(n * sizeof(float));
        for (size_t i = 0; i < n; ++i) h[i] = 1.0f + 1e-3f * (float)((i * 2654435761u) % 1000) - 0.5f;
        cudaMemcpy(in, h, n * sizeof(float), cudaMemcpyHostToDevice);
        free(h);
    }
    k<MODE><<<blocks, 256>>>(out, in);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<blocks, 256>>>(out, in);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double groups = (double)blocks * 256 * ITERS / 32;            // warp-iterations
    const double cyc = (ms * 1e-3) * (clk * 1e3) * sms * 4 / groups;    // SMSP cycles per iteration of 8 ops
    printf("%-66s %8.3f ms  %6.2f SMSP-cycles per iteration (8 ops = 16.0 ideal packed / 8.0 scalar)\n",
           name, ms, cyc);
    cudaFree(out); cudaFree(in);
}

int main()
{
    run<0>("8 FFMA2 scalar x shared pair + pair (baseline)");
    run<1>("8 FFMA2.RM pair x imm + pair");
    run<2>("8 FFMA2.RN pair x imm + pair");
    run<3>("8 FFMA.RM scalar x imm + scalar");
    run<4>("8 FFMA2 + 1 FFMA2.RM(imm)");
    run<5>("8 FFMA2 + 2 LEA.HI");
    run<10>("8 FFMA2 + 2 IMAD.HI counts");
    run<11>("8 FFMA2 + 2 predicated increments");
    run<6>("8 FFMA2 scalar x own pair + own pair (5 registers)");
    run<7>("8 FFMA2 scalar x shared pair (rewritten each iteration) + pair");
    run<8>("8 FFMA2 own pair x own pair + own pair (6 registers)");
    run<9>("8 FFMA2 own pair x shared pair + own pair");
    return 0;
}
