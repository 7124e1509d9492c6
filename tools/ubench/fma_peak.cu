// Micro-benchmark: attainable FP32 FMA rate on B200 for the operand patterns the
// RANSAC scorer uses (scalar FFMA vs packed FFMA2, register-bank pressure).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

// MODE 0: scalar FFMA, 8 chains, acc = fma(a_k, b_k, acc_k), a/b fixed distinct registers
// MODE 1: scalar FFMA, 8 chains, shared multiplicand b (operand reuse)
// MODE 2: FFMA2, 8 chains of pairs, distinct pair operands
// MODE 3: FFMA2, broadcast scalar a (the .F32 form), shared pair b
// MODE 4: FFMA2, acc = fma(acc, acc, c)   (squares: one source register pair)
// MODE 5/6: FFMA2 and scalar FFMA interleaved (do the heavy and lite FMA pipes add up?)
//           (MODE 6 reports FMA/clk as if 2 per op; true count is 2.5 -> multiply by 1.25)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed)
{
    float a[8], b[8], c[8];
    float2 A[8], B[8], C[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed + i * 1e-3f + threadIdx.x * 1e-6f;
        b[i] = seed - i * 1e-4f * (1 + (threadIdx.x & 3));   // per-thread, not foldable
        c[i] = i;
        A[i] = make_float2(a[i], a[i] + 1e-5f);
        B[i] = make_float2(b[i], b[i] - 1e-5f);
        C[i] = make_float2(c[i], c[i] + 1.f);
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) c[i] = __fmaf_rn(a[i], b[i], c[i]);
            if (MODE == 1) c[i] = __fmaf_rn(a[i], b[0], c[i]);
            if (MODE == 2) C[i] = __ffma2_rn(A[i], B[i], C[i]);
            if (MODE == 3) C[i] = __ffma2_rn(make_float2(a[i], a[i]), B[0], C[i]);
            if (MODE == 4) C[i] = __ffma2_rn(C[i], B[0], A[i]);
            if (MODE == 5) { C[i] = __ffma2_rn(make_float2(a[i], a[i]), B[0], C[i]); c[i] = __fmaf_rn(a[i], b[0], c[i]); }
            if (MODE == 6) { C[i] = __ffma2_rn(make_float2(a[i], a[i]), B[0], C[i]); if (i < 4) c[i] = __fmaf_rn(a[i], b[0], c[i]); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i] + C[i].x + C[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int fma_per_op)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8;
    float* out; cudaMalloc(&out, blocks * 256 * sizeof(float));
    k<MODE><<<blocks, 256>>>(out, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<blocks, 256>>>(out, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double fmas = (double)blocks * 256 * ITERS * 8 * fma_per_op;
    const double per_clk_sm = fmas / (ms * 1e-3) / (clk * 1e3) / sms;
    printf("%-44s %8.3f ms  %7.1f FMA/clk/SM (of 128)  %6.2f TFLOP/s\n", name, ms, per_clk_sm,
           2 * fmas / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main()
{
    run<0>("FFMA  3 distinct registers", 1);
    run<1>("FFMA  shared multiplicand (reuse)", 1);
    run<2>("FFMA2 3 distinct register pairs", 2);
    run<3>("FFMA2 broadcast scalar x shared pair", 2);
    run<4>("FFMA2 acc as multiplicand", 2);
    run<5>("mix: 8 FFMA2 + 8 FFMA per iteration (reuse forms)", 3);
    run<6>("mix: 8 FFMA2 + 4 FFMA per iteration (reuse forms)", 2);
    return 0;
}
