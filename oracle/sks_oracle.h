/*
 * TEST INFRASTRUCTURE ONLY -- declarations of the CPU oracle (see sks_oracle.c
 * for the parity status and the reference anchors).  Not part of the product.
 */
#ifndef SKS_ORACLE_H
#define SKS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* one quadruple: s[8], t[8] (x0,y0,...,x3,y3 = M,N,P,Q) -> h[9] row-major */
void oracle_aca_one_f32(const float *s, const float *t, float *h, int normalize);
void oracle_aca_one_f64(const double *s, const double *t, double *h, int normalize);
void oracle_sks_one_f32(const float *s, const float *t, float *h, int normalize);
void oracle_sks_one_f64(const double *s, const double *t, double *h, int normalize);
void oracle_ge_one_f32(const float *s, const float *t, float *h, int normalize);
void oracle_ge_one_f64(const double *s, const double *t, double *h, int normalize);
void oracle_aca_rect_one_f32(const float *t, float mx, float my, float width, float ratio,
                             float *h, int normalize);
void oracle_aca_rect_one_f64(const double *t, double mx, double my, double width, double ratio,
                             double *h, int normalize);

/* AoS batches */
void oracle_aca_f32(const float *src, const float *tar, float *H, int64_t n, int normalize);
void oracle_aca_f64(const double *src, const double *tar, double *H, int64_t n, int normalize);
void oracle_sks_f32(const float *src, const float *tar, float *H, int64_t n, int normalize);
void oracle_sks_f64(const double *src, const double *tar, double *H, int64_t n, int normalize);
void oracle_gpt_one_f32(const float *s, const float *t, float *h, int normalize);
void oracle_gpt_one_f64(const double *s, const double *t, double *h, int normalize);
void oracle_gpt_f32(const float *src, const float *tar, float *H, int64_t n, int normalize);
void oracle_gpt_f64(const double *src, const double *tar, double *H, int64_t n, int normalize);
void oracle_ge_f32(const float *src, const float *tar, float *H, int64_t n, int normalize);
void oracle_ge_f64(const double *src, const double *tar, double *H, int64_t n, int normalize);
void oracle_aca_rect_f32(const float *tar, const float *M, float mx, float my, float width,
                         float ratio, float *H, int64_t n, int normalize);
void oracle_aca_rect_f64(const double *tar, const double *M, double mx, double my, double width,
                         double ratio, double *H, int64_t n, int normalize);
void oracle_degenerate_f32(const float *H, uint8_t *flag, int64_t n, int normalized);
void oracle_degenerate_f64(const double *H, uint8_t *flag, int64_t n, int normalized);

/* synthetic quadruples [begin, begin+count), AoS */
uint64_t oracle_rng_u64(uint64_t seed, uint64_t ctr, uint32_t lane);
void oracle_synth_quads_f32(float *src, float *tar, int64_t begin, int64_t count, uint64_t seed,
                            int dist);
void oracle_synth_quads_f64(double *src, double *tar, int64_t begin, int64_t count,
                            uint64_t seed, int dist);

/* post-RANSAC least-squares refit on the inlier mask (consumer after the path; our definition) */
void oracle_lu8_solve_f32(float A[8][8], float b[8]);
void oracle_lu8_solve_f64(double A[8][8], double b[8]);
void oracle_ransac_refit_f32(const float *corr, int64_t n_pairs, int32_t n_pts, const uint8_t *mask,
                             const float *H_in, float *H_out, uint32_t *n_used);

/* sampling grid of a homography (consumer after the path; our definition) */
void oracle_warp_grid_f32(const float *H, int64_t n, float x0, float y0, float dx, float dy,
                          int32_t gw, int32_t gh, float *out);

/* cuRAND host-API MRG32K3A stream (the reference's sample list, GPU.cu:1443-1446) */
void oracle_curand_mrg32k3a_u32(uint32_t *out, int64_t n, uint64_t seed);

/* ACA-RANSAC scorer; corr is [n_pairs][n_pts][4] = (x,y,X,Y) */
uint32_t oracle_ransac_count_f32(const float *H, const float *corr, int32_t n_pts, float thr2);
void oracle_ransac_sample(uint64_t seed, int64_t pair, uint32_t hyp, int32_t n_pts,
                          uint32_t idx[4]);
void oracle_ransac_hypothesis_f32(const float *corr, const uint32_t idx[4], float *H);
void oracle_ransac_aca_shard_f32(const float *corr, int64_t pair_begin, int64_t n_pairs, int32_t n_pts,
                                 const uint32_t *samples, uint32_t hyp_begin, uint32_t hyp_count,
                                 uint32_t hyp_stride, uint64_t seed, float thr2, uint64_t *best_key,
                                 uint32_t *counts_out);
void oracle_ransac_aca_f32(const float *corr, int64_t n_pairs, int32_t n_pts,
                           const uint32_t *samples, uint32_t hyp_begin, uint32_t hyp_count,
                           uint32_t hyp_stride, uint64_t seed, float thr2, uint64_t *best_key,
                           uint32_t *counts_out);

#ifdef __cplusplus
}
#endif
#endif
