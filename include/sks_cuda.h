/*
 * libsks_cuda -- C ABI of the B200-native batched 4-point homography engine.
 *
 * This is the drop-in boundary for the batched SKS / ACA / ACA-rect path of
 * cscvlab/SKS-Homography.  The reference has no FFI layer of its own; its
 * boundary is four C++ free functions, two CUDA kernels with host wrappers,
 * two Python functions and one MATLAB function.  Each entry point below names
 * the reference interface it replaces ("MOD/" = "C++ Codes/modules/",
 * "GPU.cu" = "C++ Codes/Runtime Test/GPU_Runtime Test/GPU_Runtime Test.cu",
 * "PY.py" = "PyTorch Codes/Modules_Runtime_Test.py", "ML/" = "Matlab Codes/").
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary.
 *   - return 0 on success (the reference's solvers always return 0,
 *     MOD/ACA_SKS.cpp:101,178,302,417); <0 = SKS_ERR_*, >0 = cudaError_t.
 *   - numerical degeneracy is NOT an error: exactly like the reference it shows
 *     up as non-finite entries of H and, if requested, in the
 *     optional per-quadruple `degenerate` byte array.
 *   - the sks_cuda_* functions take DEVICE pointers, enqueue on `stream`
 *     (a cudaStream_t passed as void*, NULL = legacy default stream) and
 *     return without synchronising.  The sks_host_* functions take HOST
 *     pointers (what the reference's C++ functions take) and return when the
 *     result is in the caller's buffer.
 *   - there is no CPU fallback and no backend dispatch: without a CUDA device
 *     every compute entry point fails with SKS_ERR_NO_DEVICE / a cudaError_t.
 *   - arithmetic reproduces the reference's operation order with fused
 *     multiply-add contraction disabled, so results are bit-identical to the
 *     reference's C++ built with -ffp-contract=off.
 *   - all base pointers must be 16-byte aligned (cudaMalloc gives 256).
 *
 * Layouts (per quadruple i of n; points M,N,P,Q as x0,y0,x1,y1,x2,y2,x3,y3)
 *   SKS_LAYOUT_AOS  src[i*8+k], tar[i*8+k], H[i*9+k]       (MOD/ACA_SKS.cpp:24)
 *   SKS_LAYOUT_SOA  src[k*ld+i], tar[k*ld+i], H[k*ld+i]    (GPU.cu:87-95,141-149)
 *                   ld = leading stride in elements, 0 means n; 64-bit offsets
 *                   (the reference's `int` offsets overflow beyond 2^31/9).
 */
#ifndef SKS_CUDA_H
#define SKS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKS_CUDA_ABI_VERSION 2

enum { SKS_OK = 0, SKS_ERR_INVALID_ARG = -1, SKS_ERR_UNALIGNED = -2, SKS_ERR_NO_DEVICE = -3,
       SKS_ERR_NO_PEER_ACCESS = -4 /* multi-GPU entry: a device cannot map the primary's memory */ };
enum { SKS_LAYOUT_AOS = 0, SKS_LAYOUT_SOA = 1 };
enum {
    SKS_FLAG_NORMALIZE = 1 /* divide out h33 as MOD/ACA_SKS.cpp:94-98 does; without it the
                              result is up to scale as in GPU.cu:141-149 / PY.py:372-381 */
};
enum { SKS_DIST_DEEP = 0, SKS_DIST_IMAGE = 1, SKS_DIST_DEEP_INT = 2 };

int sks_cuda_abi_version(void);
const char *sks_cuda_error_string(int status);
int sks_cuda_device_count(int *count);

/* ---- streaming solvers, device pointers --------------------------------- */
/* replaces sks::runKernel_ACA            MOD/ACA_SKS.hpp:17, MOD/ACA_SKS.cpp:24-102
 *          cal_Homo_ACA / cal_ACA (SoA)  GPU.cu:81-151, :1166-1206
 *          ACA_vanilla                   PY.py:312-388                        */
int sks_cuda_aca_f32(const float *src, const float *tar, float *H, int64_t n, int layout,
                     int64_t ld, int flags, uint8_t *degenerate, void *stream);
/* replaces sks::runKernel_ACA_double     MOD/ACA_SKS.hpp:18, MOD/ACA_SKS.cpp:104-179 */
int sks_cuda_aca_f64(const double *src, const double *tar, double *H, int64_t n, int layout,
                     int64_t ld, int flags, uint8_t *degenerate, void *stream);
/* replaces sks::runKernel_SKS            MOD/ACA_SKS.hpp:19, MOD/ACA_SKS.cpp:189-303 */
int sks_cuda_sks_f32(const float *src, const float *tar, float *H, int64_t n, int layout,
                     int64_t ld, int flags, uint8_t *degenerate, void *stream);
/* replaces sks::runKernel_SKS_double     MOD/ACA_SKS.hpp:20, MOD/ACA_SKS.cpp:305-418
 *          cal_Homo_SKS / cal_SKS (SoA)  GPU.cu:153-240, :1208-1243              */
int sks_cuda_sks_f64(const double *src, const double *tar, double *H, int64_t n, int layout,
                     int64_t ld, int flags, uint8_t *degenerate, void *stream);

/* Competitor solver in the same harness, bit-exact to
 * cv::runKernel_GE                       MOD/GE.hpp:9, MOD/GE.cpp:44-188 (fp32)
 * cal_Homo_GE (SoA fp64, -fmad=false)    GPU.cu:359-507
 * RHO Gaussian elimination, 221 flops; h33 is 1 by construction, so flags'
 * SKS_FLAG_NORMALIZE bit makes no difference; no pivoting: an axis-aligned source
 * square gives a non-finite result exactly as the reference does. */
int sks_cuda_ge_f32(const float *src, const float *tar, float *H, int64_t n, int layout,
                    int64_t ld, int flags, uint8_t *degenerate, void *stream);
int sks_cuda_ge_f64(const double *src, const double *tar, double *H, int64_t n, int layout,
                    int64_t ld, int flags, uint8_t *degenerate, void *stream);

/* Second competitor, fp64 only: GPT-LU, the 8x8 DLT system by LU with partial pivoting in the
 * arithmetic of cal_Homo_GPT            GPU.cu:242-357 (bit-equal to that kernel built with
 * -fmad=false; its CPU form MOD/GPT.cpp is cv::getPerspectiveTransform, i.e. OpenCV). */
int sks_cuda_gpt_f64(const double *src, const double *tar, double *H, int64_t n, int layout,
                     int64_t ld, int flags, uint8_t *degenerate, void *stream);

/* replaces ACA_rect(TargetPts, M_x, M_y, width, ratio_rec)  ML/ACA_rect.m:22-38
 *          TensorACA_rect(bs, src, tar, scale, div)         PY.py:286-309
 * tar holds the 4 target corners TL,TR,BL,BR (8 values per quadruple, the
 * homogeneous 1-row of the reference's 3x4 matrices is implied).  The source
 * rectangle (top-left (mx,my), width, ratio = width/height) is shared by the
 * batch; if M != NULL it supplies a per-quadruple top-left corner
 * (AoS M[i*2+k], SoA M[k*ld+i]) as PY.py:302 does, and mx,my are ignored.
 * With SKS_FLAG_NORMALIZE every element is DIVIDED by h33 (ML/ACA_rect.m:36). */
int sks_cuda_aca_rect_f32(const float *tar, const float *M, float mx, float my, float width,
                          float ratio, float *H, int64_t n, int layout, int64_t ld, int flags,
                          uint8_t *degenerate, void *stream);
int sks_cuda_aca_rect_f64(const double *tar, const double *M, double mx, double my,
                          double width, double ratio, double *H, int64_t n, int layout,
                          int64_t ld, int flags, uint8_t *degenerate, void *stream);

/* Same solver on the reference's torch tensor convention (PY.py:24-37,286-302):
 * tar34 / src34 are [n][3][4] homogeneous point matrices (rows x, y, 1; columns
 * TL,TR,BL,BR), read in place without a transposing copy; src34 (nullable)
 * supplies the per-sample corner src34[i][0..1][0] as PY.py:302 does.  H is AoS. */
int sks_cuda_aca_rect_planar_f32(const float *tar34, const float *src34, float mx, float my,
                                 float width, float ratio, float *H, int64_t n, int flags,
                                 uint8_t *degenerate, void *stream);
int sks_cuda_aca_rect_planar_f64(const double *tar34, const double *src34, double mx, double my,
                                 double width, double ratio, double *H, int64_t n, int flags,
                                 uint8_t *degenerate, void *stream);
/* The same with the rectangle's width (`scale`) and width/height ratio (`div`) read from
 * one-element DEVICE arrays, as the reference holds them (PY.py:33-35: tensors derived from
 * batch element 0, used at :301-302): no device->host synchronisation to fetch two scalars.
 * src34 supplies the per-sample corner as above (NULL: corner (0, 0)). */
int sks_cuda_aca_rect_planar_dev_f32(const float *tar34, const float *src34, const float *width_dev,
                                     const float *ratio_dev, float *H, int64_t n, int flags,
                                     uint8_t *degenerate, void *stream);
int sks_cuda_aca_rect_planar_dev_f64(const double *tar34, const double *src34,
                                     const double *width_dev, const double *ratio_dev, double *H,
                                     int64_t n, int flags, uint8_t *degenerate, void *stream);

/* ---- host-pointer entry points (what the reference's C++ callers hold) ---- */
/* Drop-in for the loop `for (k...) sks::runKernel_*(src, tar, result)` of
 * CPU/main.cpp:87-114 over n quadruples: AoS host buffers in, AoS host buffer
 * out, always on the current device, chunked and double-buffered so H2D copy,
 * kernel and D2H copy overlap.  Pinned (cudaHostAlloc / cudaHostRegister)
 * buffers are copied directly; pageable ones go through an internal pinned
 * ring.  flags as above. */
int sks_host_aca_f32(const float *src, const float *tar, float *H, int64_t n, int flags);
int sks_host_aca_f64(const double *src, const double *tar, double *H, int64_t n, int flags);
int sks_host_sks_f32(const float *src, const float *tar, float *H, int64_t n, int flags);
int sks_host_sks_f64(const double *src, const double *tar, double *H, int64_t n, int flags);
int sks_host_ge_f32(const float *src, const float *tar, float *H, int64_t n, int flags);
int sks_host_ge_f64(const double *src, const double *tar, double *H, int64_t n, int flags);
int sks_host_aca_rect_f32(const float *tar, const float *M, float mx, float my, float width,
                          float ratio, float *H, int64_t n, int flags);
int sks_host_aca_rect_f64(const double *tar, const double *M, double mx, double my,
                          double width, double ratio, double *H, int64_t n, int flags);
/* Fused ACA-RANSAC for callers that hold their matches in host memory (corr
 * [n_pairs][n_pts][4], optional samples [n_pairs][n_hyp][4]): copies in, scores hypothesis
 * ids [0, n_hyp), rebuilds the winners and copies out H_best [n_pairs][9] and, when
 * non-NULL, inlier_count [n_pairs], inlier_mask [n_pairs][n_pts], best_key [n_pairs].
 * Synchronous; same definitions as sks_cuda_ransac_aca_f32 / _finalize_f32 below. */
int sks_host_ransac_aca_f32(const float *corr, int64_t n_pairs, int32_t n_pts,
                            const uint32_t *samples, uint32_t n_hyp, uint64_t seed, float thr2,
                            float *H_best, uint32_t *inlier_count, uint8_t *inlier_mask,
                            unsigned long long *best_key);
/* The same over `ngpu` GPUs of this process (0 = all visible): one H2D copy to the current
 * device, then the flow of sks_cuda_ransac_aca_multi_f32 below.  sks_host_ransac_aca_f32
 * itself takes this route when sks_host_set_device_count(g != 1) is in effect, which is how
 * sks::runRansac_ACA (include/sks_homography.hpp) reaches all GPUs of a box. */
int sks_host_ransac_aca_multi_f32(const float *corr, int64_t n_pairs, int32_t n_pts,
                                  const uint32_t *samples, uint32_t n_hyp, uint64_t seed,
                                  float thr2, int ngpu, float *H_best, uint32_t *inlier_count,
                                  uint8_t *inlier_mask, unsigned long long *best_key);
/* In-process multi-GPU driver for the host-pointer entry points: shard every batch
 * contiguously over `count` GPUs (devices 0..count-1; 0 = all visible; default 1 =
 * the current device only), one host thread and one PCIe link per GPU, no
 * inter-GPU traffic. */
int sks_host_set_device_count(int count);
/* Pipeline granularity of the host-pointer path: bytes of ONE input array per chunk
 * (upper bound, default 64 MiB; 64 KiB .. 1 GiB; batches are cut into ~8 chunks). */
int sks_host_set_chunk_bytes(int64_t bytes_per_input_array);
/* Staging copies of the pageable path: bit 0 = non-temporal (streaming) stores into the pinned
 * ring, bit 1 = out of it into the caller's result buffer; 0 = plain memcpy both ways.
 * A tuning knob for measurements; every setting gives the same bytes. */
int sks_host_set_staging_copy(int non_temporal);
/* Worker threads one staging copy is spread over (1..64; 0 = default = three quarters of the pool of
 * min(16, cores) workers). */
int sks_host_set_staging_threads(int threads_per_copy);
/* pinned host allocation helpers for callers that want the zero-staging path */
int sks_host_alloc_pinned(void **ptr, int64_t bytes);
int sks_host_free_pinned(void *ptr);
/* ... or pin a buffer the caller already owns (the storage of a std::vector, a numpy array) in
 * place: page-locks [ptr, ptr+bytes) with cudaHostRegister, after which sks_host_* calls DMA from /
 * into it directly (0.79 instead of ~0.5 G H/s per GPU).  Registration costs about as much as one
 * pass over the buffer, so it pays for buffers that are used more than once.  Unregister before
 * freeing the memory. */
int sks_host_register(void *ptr, int64_t bytes);
int sks_host_unregister(void *ptr);

/* ---- minimal-sample gather (hypothesis generation from a match pool) ------ */
/* replaces get_rand_list  GPU.cu:52-78 (+ host setup :1443-1451): for each of
 * n hypotheses pick 4 correspondences r_k % pool_size (repeats allowed) from a
 * pool of (x,y,X,Y) matches and emit src/tar quadruples.  rand4 is
 * [4][n] uint32 as the reference's cuRAND buffer; NULL = counter RNG(seed). */
int sks_cuda_gather_samples_f32(const float *pool_xyXY, uint32_t pool_size,
                                const uint32_t *rand4, uint64_t seed, float *src, float *tar,
                                int64_t n, int layout, int64_t ld, void *stream);
int sks_cuda_gather_samples_f64(const double *pool_xyXY, uint32_t pool_size,
                                const uint32_t *rand4, uint64_t seed, double *src, double *tar,
                                int64_t n, int layout, int64_t ld, void *stream);

/* Fused gather + solve: the reference's GPU flow get_rand_list -> cal_Homo_ACA /
 * cal_Homo_SKS (GPU.cu:1449-1464) in ONE kernel -- the sampled quadruples never
 * touch HBM, only H (AoS or SoA, like the solvers) is written.  Same sampling
 * rule and arguments as sks_cuda_gather_samples_*. */
int sks_cuda_gather_aca_f32(const float *pool_xyXY, uint32_t pool_size, const uint32_t *rand4,
                            uint64_t seed, float *H, int64_t n, int layout, int64_t ld, int flags,
                            uint8_t *degenerate, void *stream);
int sks_cuda_gather_aca_f64(const double *pool_xyXY, uint32_t pool_size, const uint32_t *rand4,
                            uint64_t seed, double *H, int64_t n, int layout, int64_t ld, int flags,
                            uint8_t *degenerate, void *stream);
int sks_cuda_gather_sks_f32(const float *pool_xyXY, uint32_t pool_size, const uint32_t *rand4,
                            uint64_t seed, float *H, int64_t n, int layout, int64_t ld, int flags,
                            uint8_t *degenerate, void *stream);
int sks_cuda_gather_sks_f64(const double *pool_xyXY, uint32_t pool_size, const uint32_t *rand4,
                            uint64_t seed, double *H, int64_t n, int layout, int64_t ld, int flags,
                            uint8_t *degenerate, void *stream);

/* The reference's sample list itself (GPU.cu:1443-1446): n 32-bit draws of cuRAND's
 * host-API generator CURAND_RNG_PSEUDO_MRG32K3A with the given seed (the reference
 * uses 11), offset 0, default ordering -- bit-identical to
 *   curandCreateGenerator(&g, CURAND_RNG_PSEUDO_MRG32K3A);
 *   curandSetPseudoRandomGeneratorSeed(g, seed);  curandGenerate(g, out, n);
 * without linking libcurand (hand-written kernel, csrc/mrg32k3a.cuh).  With
 * n = 4*numsOfH the buffer is the rand4 argument of the gather entry points above,
 * which replays the reference's exact hypothesis set. */
int sks_cuda_curand_mrg32k3a_u32(uint32_t *out, int64_t n, uint64_t seed, void *stream);

/* ---- fused ACA-RANSAC ----------------------------------------------------- */
/* New (nothing like it in the reference; sampler precedent GPU.cu:52-78).
 * corr: [n_pairs][n_pts][4] = (x,y,X,Y) fp32.  Hypothesis ids
 * [hyp_begin, hyp_begin+hyp_count) of every pair are generated (4 indices
 * u32 % n_pts from the counter RNG keyed (seed, pair, hyp), or read from
 * samples[n_pairs][hyp_stride][4] if non-NULL -- entries are reduced % n_pts like the
 * reference's get_rand_list, GPU.cu:55-58, so a raw 32-bit stream such as the output of
 * sks_cuda_curand_mrg32k3a_u32 is a valid list), solved with the bit-exact fp32
 * ACA, scored against the pair's correspondences held in shared memory and
 * reduced to best_key[pair] = max(count<<32 | (0xFFFFFFFF - hyp)).  best_key
 * is MAX-combined into the caller's array (zero it before the first call), so
 * several calls / several GPUs can cover disjoint hypothesis ranges and be
 * merged with an integer max-reduce.  No per-hypothesis H touches HBM.
 * Inlier rule (this project's definition; mirrored bit for bit by the CPU oracle):
 * forward transfer error |proj(x) - X|^2 < thr2, evaluated division-free as
 *   it = 1/sqrtf(thr2); g = H with rows 1-2 scaled by it; Xs = X*it, Ys = Y*it;
 *   u,v,w = fma chains of g on (x, y, 1); du = fma(-Xs,w,u); dv = fma(-Ys,w,v);
 *   inlier <=> fma(-w, w, fma(dv,dv,du*du)) < 0      (NaN / inf are never inliers)
 * thr2 is the squared pixel threshold; keep it within ~1e-30..1e30 so that the folded
 * operands stay finite. */
int sks_cuda_ransac_aca_f32(const float *corr, int64_t n_pairs, int32_t n_pts,
                            const uint32_t *samples, uint32_t hyp_stride, uint32_t hyp_begin,
                            uint32_t hyp_count, uint64_t seed, float thr2,
                            unsigned long long *best_key, void *stream);
/* The same on a SHARD of the image pairs (multi-GPU variant with zero communication:
 * each GPU owns pairs [pair_begin, pair_begin + n_pairs)).  corr / samples / best_key point at
 * the shard's first pair; pair_begin only keys the counter RNG, so a shard draws exactly
 * the samples the same pairs get in an unsharded call. */
int sks_cuda_ransac_aca_shard_f32(const float *corr, int64_t pair_begin, int64_t n_pairs,
                                  int32_t n_pts, const uint32_t *samples, uint32_t hyp_stride,
                                  uint32_t hyp_begin, uint32_t hyp_count, uint64_t seed,
                                  float thr2, unsigned long long *best_key, void *stream);
int sks_cuda_ransac_finalize_shard_f32(const float *corr, int64_t pair_begin, int64_t n_pairs,
                                       int32_t n_pts, const uint32_t *samples,
                                       uint32_t hyp_stride, uint64_t seed, float thr2,
                                       const unsigned long long *best_key, float *H_best,
                                       uint32_t *inlier_count, uint8_t *inlier_mask,
                                       void *stream);
/* Recompute the winning model of every pair from best_key (the sample list is
 * a pure function of the seed, so no H ever travels between GPUs): H_best
 * [n_pairs][9], inlier_count [n_pairs], optional inlier_mask [n_pairs][n_pts].
 * A pair whose key is still 0 (nothing was scored for it), or whose id lies outside an
 * explicit sample list, has no model: H_best = NaN, inlier_count = 0, mask all 0. */
int sks_cuda_ransac_finalize_f32(const float *corr, int64_t n_pairs, int32_t n_pts,
                                 const uint32_t *samples, uint32_t hyp_stride, uint64_t seed,
                                 float thr2, const unsigned long long *best_key, float *H_best,
                                 uint32_t *inlier_count, uint8_t *inlier_mask, void *stream);

/* Multi-GPU driver entry (SURVEY.md 8(b) "sks_cuda_*_multi(..., int ngpu)", 8(e) variant A) for
 * callers that are ONE process, not one process per GPU: the whole estimate -- zero the keys,
 * score hypothesis ids [0, n_hyp) of every pair, merge, rebuild the winners -- over `ngpu`
 * devices (the current one, which owns corr / samples and all outputs, plus the next ngpu-1 in
 * index order; 0 = all visible).  The matches reach the other devices by a binomial-tree
 * broadcast of NVLink peer copies into library-owned buffers (an explicit sample list is read
 * in place over peer access); device k scores the k-th contiguous shard of the hypothesis ids
 * as soon as its copy has landed and max-combines its winners into best_key with system-scope
 * atomics; CUDA events order the devices' streams, so the call only enqueues, like every other
 * sks_cuda_* entry: results are valid when `stream` reaches the end of the enqueued work.  No
 * NCCL, no IPC.  best_key [n_pairs] is overwritten (not max-combined).  H_best may be NULL to
 * skip the finalize step.  Results are bit-identical to ngpu = 1.  Fails with
 * SKS_ERR_NO_PEER_ACCESS where a device cannot map the current device's memory. */
int sks_cuda_ransac_aca_multi_f32(const float *corr, int64_t n_pairs, int32_t n_pts,
                                  const uint32_t *samples, uint32_t n_hyp, uint64_t seed,
                                  float thr2, int ngpu, unsigned long long *best_key,
                                  float *H_best, uint32_t *inlier_count, uint8_t *inlier_mask,
                                  void *stream);

/* Score GIVEN models with the same inlier rule: H [n_pairs][9] (h33-normalised or not) ->
 * inlier_count [n_pairs] and / or inlier_mask [n_pairs][n_pts] (either may be NULL). */
int sks_cuda_ransac_score_f32(const float *corr, int64_t n_pairs, int32_t n_pts, const float *H,
                              float thr2, uint32_t *inlier_count, uint8_t *inlier_mask,
                              void *stream);

/* Post-RANSAC re-estimation hook (new; nothing like it in the reference): least-squares
 * homography of every pair from ALL matches flagged in inlier_mask [n_pairs][n_pts] (as written
 * by sks_cuda_ransac_finalize_f32): Hartley-normalised DLT normal equations accumulated and
 * solved in fp64, one warp per pair.  H_out [n_pairs][9] is h33-normalised; pairs with fewer
 * than 4 inliers or a non-finite solution keep H_in (n_used = 0 for them). */
int sks_cuda_ransac_refit_f32(const float *corr, int64_t n_pairs, int32_t n_pts,
                              const uint8_t *inlier_mask, const float *H_in, float *H_out,
                              uint32_t *n_used, void *stream);

/* ---- consumer after the path: sampling grids for image warping ------------- */
/* New (the reference only remarks that warping does not need the normalisation,
 * ML/ACA_rect.m:33-35).  grid_xy[n][gh][gw][2] = (u, v) * (1/w) with (u,v,w) =
 * H * (x0 + i*dx, y0 + j*dy, 1); H[n][9] may be at any scale.  The fused form runs
 * ACA-rect on tar[n][8] (arguments as sks_cuda_aca_rect_f32) up to scale in
 * registers and never writes H.  HBM-bound on the 8-byte-per-point output. */
int sks_cuda_warp_grid_f32(const float *H, int64_t n, float x0, float y0, float dx, float dy,
                           int32_t gw, int32_t gh, float *grid_xy, void *stream);
int sks_cuda_aca_rect_warp_grid_f32(const float *tar, const float *M, float mx, float my,
                                    float width, float ratio, int64_t n, float x0, float y0,
                                    float dx, float dy, int32_t gw, int32_t gh, float *grid_xy,
                                    void *stream);

/* ---- synthetic inputs (bench / tests), generated on the device ------------ */
/* Counter-based generator, bit-identical to oracle_synth_quads_* for the same
 * (seed, dist); distributions per SURVEY.md 8(d) (PY.py:9-21, ML/veri_4Pts.m). */
int sks_cuda_synth_quads_f32(float *src, float *tar, int64_t begin, int64_t count,
                             uint64_t seed, int dist, int layout, int64_t ld, void *stream);
int sks_cuda_synth_quads_f64(double *src, double *tar, int64_t begin, int64_t count,
                             uint64_t seed, int dist, int layout, int64_t ld, void *stream);
/* RANSAC scene: per pair a ground-truth homography from a SKS_DIST_DEEP quad,
 * inlier_permille of the points follow it with +-noise px uniform noise, the
 * rest are uniform outliers in the 160x160 frame. */
int sks_cuda_synth_corr_f32(float *corr, int64_t pair_begin, int64_t n_pairs, int32_t n_pts,
                            uint64_t seed, int inlier_permille, float noise, void *stream);

/* ---- NVLink peer exchange: hand-written max-reduce of the RANSAC keys ------- */
/* Optional replacement for the one NCCL all-reduce of the multi-GPU RANSAC step
 * (one process per GPU, same node).  Each rank allocates an exchange block for
 * n_keys packed keys, exports its 64-byte CUDA-IPC handle, and opens every
 * peer's.  A reduce is then sks_cuda_peer_push_max (system-scope atomicMax of
 * the local keys into every rank's block over NVLink + arrival signal) followed
 * by sks_cuda_peer_wait (bounded device-side wait for all `world` arrivals, then
 * the reduced keys are copied to keys_out).  On a timeout *status_dev is set to 1 -- it is
 * never cleared by the library, so zero it once and poll it; it may live in pinned host
 * memory -- and keys_out is left UNTOUCHED: a partial max is never returned.  A timeout is
 * fatal for the exchange (the arrival counters no longer line up): free and re-allocate the
 * blocks, or fall back to an all-reduce.  `epoch`
 * counts reduces since allocation (0, 1, 2, ...) and must advance in lock-step
 * on all ranks; peer_blocks[g] is rank g's block as mapped in THIS process
 * (peer_blocks[rank] = the own block). */
int sks_cuda_peer_alloc(void **block, int64_t n_keys);
int sks_cuda_peer_free(void *block);
int sks_cuda_peer_export(void *block, void *handle64);
int sks_cuda_peer_open(const void *handle64, void **block);
int sks_cuda_peer_close(void *block);
int sks_cuda_peer_push_max(const unsigned long long *local_keys, int64_t n_keys,
                           void *const *peer_blocks, int world, int rank, uint64_t epoch,
                           void *stream);
int sks_cuda_peer_wait(void *own_block, int world, uint64_t epoch, unsigned long long *keys_out,
                       int64_t n_keys, int *status_dev, double timeout_s, void *stream);

/* ---- multi-GPU helpers ----------------------------------------------------- */
/* Contiguous shard [begin, begin+count) of n units for `rank` of `world`
 * (SURVEY.md 8(e)); the streaming solvers need no collective at all. */
int sks_cuda_shard_range(int64_t n, int rank, int world, int64_t *begin, int64_t *count);

/* ---- introspection / tuning ------------------------------------------------ */
/* Number of kernels this library has launched in this process since the last
 * reset (bench.py reports it as gpu_launches). */
int64_t sks_cuda_launch_count(void);
void sks_cuda_reset_launch_count(void);
/* The three tuning setters below act on the CALLING HOST THREAD only (thread-local state): a
 * setting never changes what another thread's calls launch.
 * Kernel variant for the AoS streaming solvers: 0 = default (best measured, = 1),
 * 1 = direct vector loads + shared-memory transposed stores,
 * 2 = persistent TMA bulk-copy ring (cp.async.bulk + mbarrier); 2 also keeps the fused
 * gather+solve entry points on their L1 gather path instead of the shared-memory-pool kernel;
 * 3 = warp-private TMA ring (every warp its own producer, no CTA-wide barrier).
 * A tuning knob for bench.py sweeps and tests, not a backend switch: every variant is
 * sm_100a CUDA. */
int sks_cuda_set_variant(int variant);
int sks_cuda_get_variant(void);
/* Kernel tuning: small_tile bit 0 (ring kernel: 0 = 256 fp32 / 128 fp64 quadruples
 * per tile, 1 = half of that), bit 1 (direct kernel: 1 = force 16-byte instead of
 * 256-bit global accesses), stages (2..16 shared-memory ring slots), ctas_per_sm
 * (0 = as many as fit). */
int sks_cuda_set_tuning(int small_tile, int stages, int ctas_per_sm);
/* RANSAC kernel tuning: hypotheses carried per thread (2 or 4), scoring rounds
 * per CTA (a CTA scores rounds*256*hyps_per_thread hypothesis ids), and packed:
 * bits 0-1 = scorer (0: scalar FFMA; 1: sm_100a FFMA2/FMUL2 over two matches;
 * 2: FFMA2/FMUL2 over two hypotheses; 3 (default): as 1 with the inlier count kept on
 * the FP32 pipe by a directed-rounding FFMA2 -- all four give the same bits),
 * bits 2.. = CTA size (0: 256, 1: 384, 2: 512 threads). */
int sks_cuda_set_ransac_tuning(int hyps_per_thread, int rounds_per_cta, int packed);
/* The launch plan of the RANSAC scorer, exported for tests (pure host arithmetic, no GPU): the
 * number of hypothesis ids one CTA scores -- rounds * threads * hyps_per_thread with rounds the
 * tuning value halved down to 1, whichever minimises (waves of resident_ctas CTAs) x (rounds +
 * 0.03): all CTAs of a launch take the same time, so a grid of 4.6 waves costs 5. */
uint32_t sks_cuda_ransac_chunk_plan(int64_t n_pairs, uint32_t hyp_count, int threads,
                                    int hyps_per_thread, int max_rounds, int resident_ctas);
int sks_cuda_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* SKS_CUDA_H */
