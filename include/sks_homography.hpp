// sks_homography.hpp -- C++ host-side drop-in for the reference's solver header.
//
// The reference exposes four free functions in namespace sks
// ("C++ Codes/modules/ACA_SKS.hpp":17-20):
//     int runKernel_ACA(float* src, float* tar, float* result);
//     int runKernel_ACA_double(double*, double*, double*);
//     int runKernel_SKS(float*, float*, float*);
//     int runKernel_SKS_double(double*, double*, double*);
// each solving ONE quadruple (src[8], tar[8] -> result[9], h33-normalised,
// always returning 0) and called in a loop by the harness
// ("C++ Codes/Runtime Test/CPU_Runtime Test/main.cpp":87-114).
//
// This header keeps those names and the first three parameters and adds the
// batch count, so the caller's loop
//     for (k = 0; k < n; ++k) sks::runKernel_ACA(src + 8*k, tar + 8*k, H + 9*k);
// becomes one call
//     sks::runKernel_ACA(src, tar, H, n);
// executed on the GPU by libsks_cuda (host pointers in, host pointers out; see
// sks_host_* in sks_cuda.h).  The 3-argument forms are kept for source
// compatibility (n = 1).  Return value: 0 on success like the reference;
// non-zero = SKS_ERR_* / cudaError_t (there is no CPU fallback).  Degenerate
// quadruples are not errors: as in the reference they yield non-finite H.
//
// New here (no C++ counterpart in the reference): runKernel_ACA_rect[_double],
// the batched form of ACA_rect(TargetPts, M_x, M_y, width, ratio_rec)
// ("Matlab Codes/ACA_rect.m":22) / TensorACA_rect ("PyTorch Codes/
// Modules_Runtime_Test.py":286).
#pragma once
#include <cstdint>

#include "sks_cuda.h"

namespace sks {

inline int runKernel_ACA(float* src, float* tar, float* result, std::int64_t n)
{
    return sks_host_aca_f32(src, tar, result, n, SKS_FLAG_NORMALIZE);
}
inline int runKernel_ACA_double(double* src, double* tar, double* result, std::int64_t n)
{
    return sks_host_aca_f64(src, tar, result, n, SKS_FLAG_NORMALIZE);
}
inline int runKernel_SKS(float* src, float* tar, float* result, std::int64_t n)
{
    return sks_host_sks_f32(src, tar, result, n, SKS_FLAG_NORMALIZE);
}
inline int runKernel_SKS_double(double* src, double* tar, double* result, std::int64_t n)
{
    return sks_host_sks_f64(src, tar, result, n, SKS_FLAG_NORMALIZE);
}

// reference-identical signatures (one quadruple)
inline int runKernel_ACA(float* src, float* tar, float* result) { return runKernel_ACA(src, tar, result, 1); }
inline int runKernel_ACA_double(double* src, double* tar, double* result) { return runKernel_ACA_double(src, tar, result, 1); }
inline int runKernel_SKS(float* src, float* tar, float* result) { return runKernel_SKS(src, tar, result, 1); }
inline int runKernel_SKS_double(double* src, double* tar, double* result) { return runKernel_SKS_double(src, tar, result, 1); }

// Buffers that are used for more than one call can be pinned in place once, which takes the
// batched calls above from the staged (~0.5 G H/s) to the direct-DMA path (~0.79 G H/s per GPU):
//     sks::pin(src, n * 8 * sizeof(float));  ...  sks::unpin(src);
inline int pin(void* buffer, std::int64_t bytes) { return sks_host_register(buffer, bytes); }
inline int unpin(void* buffer) { return sks_host_unregister(buffer); }

// tar[n][8] = target corners TL,TR,BL,BR; shared source rectangle (M_x, M_y,
// width, ratio_rec = width/height); result h33-normalised by division.
inline int runKernel_ACA_rect(float* tar, float M_x, float M_y, float width, float ratio_rec,
                              float* result, std::int64_t n = 1)
{
    return sks_host_aca_rect_f32(tar, nullptr, M_x, M_y, width, ratio_rec, result, n, SKS_FLAG_NORMALIZE);
}
inline int runKernel_ACA_rect_double(double* tar, double M_x, double M_y, double width,
                                     double ratio_rec, double* result, std::int64_t n = 1)
{
    return sks_host_aca_rect_f64(tar, nullptr, M_x, M_y, width, ratio_rec, result, n, SKS_FLAG_NORMALIZE);
}

// New here (nothing like it in the reference): the fused ACA-RANSAC over host buffers.
// corr[n_pairs][n_pts][4] = (x, y, X, Y); the n_hyp minimal samples per pair come from the
// counter RNG keyed by seed; optional outputs may be null.
inline int runRansac_ACA(const float* corr, std::int64_t n_pairs, std::int32_t n_pts, std::uint32_t n_hyp,
                         std::uint64_t seed, float thr2, float* H_best, std::uint32_t* inlier_count = nullptr,
                         std::uint8_t* inlier_mask = nullptr)
{
    return sks_host_ransac_aca_f32(corr, n_pairs, n_pts, nullptr, n_hyp, seed, thr2, H_best, inlier_count,
                                   inlier_mask, nullptr);
}

// The same over `ngpu` GPUs of this process (0 = every visible GPU): hypothesis ids sharded,
// winners merged over NVLink inside libsks_cuda, bit-identical to one GPU.
inline int runRansac_ACA_multi(const float* corr, std::int64_t n_pairs, std::int32_t n_pts, std::uint32_t n_hyp,
                               std::uint64_t seed, float thr2, int ngpu, float* H_best,
                               std::uint32_t* inlier_count = nullptr, std::uint8_t* inlier_mask = nullptr)
{
    return sks_host_ransac_aca_multi_f32(corr, n_pairs, n_pts, nullptr, n_hyp, seed, thr2, ngpu, H_best,
                                         inlier_count, inlier_mask, nullptr);
}

}  // namespace sks

// The reference keeps its competitor solvers in namespace cv; the one that is
// self-contained, RHO-GE ("C++ Codes/modules/GE.hpp":9, void return), is served by
// the same streaming kernel (bit-exact to MOD/GE.cpp).  Define SKS_NO_CV_NAMESPACE
// when OpenCV's own cv:: is in scope and the names would collide.
#ifndef SKS_NO_CV_NAMESPACE
namespace cv {
inline int runKernel_GE(float* src, float* tar, float* result, std::int64_t n)
{
    return sks_host_ge_f32(src, tar, result, n, SKS_FLAG_NORMALIZE);
}
inline void runKernel_GE(float* src, float* tar, float* result) { (void)runKernel_GE(src, tar, result, 1); }
}  // namespace cv
#endif
