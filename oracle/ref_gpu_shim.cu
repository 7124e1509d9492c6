// TEST / BENCH INFRASTRUCTURE ONLY -- same-box GPU comparator.
// Wraps the reference's own CUDA kernels cal_Homo_ACA / cal_Homo_SKS
// ("GPU_Runtime Test.cu":81-240), cal_Homo_GE (:359-507) and cal_Homo_GPT (:242-357), which oracle/Makefile
// extracts AT BUILD TIME
// into the git-ignored oracle/_ref/ref_gpu_kernels.inc (nothing of the reference
// is stored in this repository), and launches them exactly as the reference's
// host wrappers do (GPU.cu:1177-1178,1193 / :1216-1217,1232): fp64, SoA,
// <<<ceil(N/32), 32>>>, un-normalised, default nvcc flags (FMA contraction on,
// so libsks_refgpu.so is NOT a parity oracle -- perf comparator only).  The same file
// built with -fmad=false (libsks_refgpu_nofma.so) rounds like the reference's C++ and
// is bit-compared with our SoA fp64 kernels in tests/test_gpu_parity.py.
#include <cuda_runtime.h>

#include "_ref/ref_gpu_kernels.inc"

extern "C" {

int refgpu_aca_f64(double* d_src, double* d_tar, double* d_H, int n, void* stream)
{
    dim3 block = 32;
    dim3 grid = (n + block.x - 1) / block.x;
    cal_Homo_ACA<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(d_src, d_tar, d_H, n);
    return (int)cudaGetLastError();
}

int refgpu_sks_f64(double* d_src, double* d_tar, double* d_H, int n, void* stream)
{
    dim3 block = 32;
    dim3 grid = (n + block.x - 1) / block.x;
    cal_Homo_SKS<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(d_src, d_tar, d_H, n);
    return (int)cudaGetLastError();
}

int refgpu_ge_f64(double* d_src, double* d_tar, double* d_H, int n, void* stream)
{
    dim3 block = 32;                       // GPU.cu:1286-1287
    dim3 grid = (n + block.x - 1) / block.x;
    cal_Homo_GE<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(d_src, d_tar, d_H, n);
    return (int)cudaGetLastError();
}

int refgpu_gpt_f64(double* d_src, double* d_tar, double* d_H, int n, void* stream)
{
    dim3 block = 32;
    dim3 grid = (n + block.x - 1) / block.x;
    cal_Homo_GPT<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(d_src, d_tar, d_H, n);
    return (int)cudaGetLastError();
}

}  // extern "C"
