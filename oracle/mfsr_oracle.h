/*
 * mfsr_oracle.h — CPU restatement of the reference burst-SR kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under multi_frame_super_resolution_b200/
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg use it, as the checker.
 *
 * Every function restates one reference kernel (file:line in the .c file) in
 * plain C with the same fp32 operation order; build with -ffp-contract=off.
 * All images are DENSE row-major (the reference's byte pitches carry no
 * arithmetic).  float3 = 3 packed floats, float4 = 4, float2 = 2.
 *
 * Pinning status: see the header of mfsr_oracle.c.
 */
#ifndef MFSR_ORACLE_H_
#define MFSR_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_merge_geom {
    int raw_w, raw_h, scale, out_w, out_h, org_x, org_y;
    int clamp_x0, clamp_x1, clamp_y0, clamp_y1;
} orc_merge_geom;

int  orc_set_threads(int n);           /* OpenMP threads used by the loops below; returns the count in use */

/* DeBayerKernels.cu:244 */
void orc_subsample3(const uint16_t* raw, float* rgb3, float maxVal, int dimX, int dimY, const int cfa[4]);
/* DeBayerKernels.cu:55 / :153 — rgb3 must be zero-initialised by the caller (unwritten border) */
void orc_debayer_green(const float* raw, float* rgb3, int w, int h, const int cfa[4], const float black[3], const float scale[3]);
void orc_debayer_redblue(const float* raw, float* rgb3, int w, int h, const int cfa[4], const float black[3], const float scale[3]);
/* main.cpp:370 gaussin_filter_1D: returns tap count (odd), taps[] must hold >= 99 */
int  orc_gauss_taps(float sigma, float* taps);
/* restated host: luminance + separable blur + quantise */
void orc_tracking_image(const float* rgb3, float* gray, uint8_t* gray_q, int w, int h, float sigma, int track_bits);
void orc_pyramid_down(const uint8_t* in, int in_w, int in_h, uint8_t* out);

/* kernel.cu:265 / :324 — tiles: float [tiles][P][P] */
void orc_tiles_border(const float* img, float* tiles, int w, int h, int M, int T, int tx, int ty, float bsx, float bsy, float rot);
void orc_tiles_preshift(const float* img, float* tiles, const float* pre2, int w, int h, int M, int T, int tx, int ty, float bsx, float bsy, float rot);
/* cross-correlation the absent host got from cuFFT (conj(FFT(a))*FFT(b), kernel.cu:485, scaled 1/P^2),
 * restated as the circular correlation it equals, summed row-major */
void orc_cross_correlation(const float* a_tiles, const float* b_tiles, float* cc, int P, int tiles);
/* kernel.cu:119,149,186,227,512 */
void orc_squared_sum(const float* tiles, float* out, int M, int T, int tiles_n);
void orc_box_x(const float* in, float* out, int M, int T, int tiles_n);
void orc_box_y(const float* in, float* out, int M, int T, int tiles_n);
void orc_normalized_cc(const float* cc, const float* sq, const float* box, float* ssd, int M, int T, int tiles_n);
void orc_find_minimum(const float* ssd, float* coord2, int32_t* argmin2, int M, int tiles_n, float threshold);
/* convenience: the whole chain above for u8 images; out_shift2 = coord + round(pre+base) - base */
void orc_tile_align(const uint8_t* ref, const uint8_t* mov, int w, int h, const float* pre2,
                    float* out_shift2, int32_t* argmin2, float* ssd_out,
                    int T, int M, int tx, int ty, float bsx, float bsy, float rot, float threshold);
/* kernel.cu:642 */
void orc_upsample_shifts(const float* in2, float* out2, int oldLevel, int newLevel, int oldCX, int oldCY,
                         int newCX, int newCY, int oldT, int newT);
/* ShiftMinimizerKernels.cu:29-218 + the batched normal-equation solve of the absent host */
void orc_consolidate_shifts(const float* measured2, const int* pair_from, const int* pair_to, int m,
                            int imageCount, int tilesX, int tilesY, int referenceImage,
                            float* one_to_one2, float* frame_shift2, int32_t* status);
/* opticalFlow.cu:48,28,97,151,190 */
void orc_consolidate_shifts_masked(const float* measured2, const int* pair_from, const int* pair_to, const uint8_t* pair_valid, int m,
                                   int imageCount, int tilesX, int tilesY, int referenceImage,
                                   float* one_to_one2, float* frame_shift2, int32_t* status);
void orc_tile_align_cs(const uint8_t* ref, const uint8_t* mov, int w, int h, const float* pre2,
                       float* out_shift2, int32_t* argmin2, float* ssd_out,
                       int T, int M, int tx, int ty, float bsx, float bsy, float cf, float sf, float threshold);
void orc_flow_from_tiles_cs(const float* tile2, int tilesX, int tilesY, int T, float* flow2, int w, int h,
                            float bsx, float bsy, float cf, float sf);
void orc_prealign_search(const uint8_t* ref, const uint8_t* mov, int w, int h, const float* cs, int idx0, int step, int n_ang,
                         int cx, int cy, int R, int sub, int32_t* out3);
void orc_flow_from_tiles(const float* tile2, int tilesX, int tilesY, int T, float* flow2, int w, int h,
                         float bsx, float bsy, float rot);
void orc_warp(const float* flow2, const float* img, float* out, int w, int h);
void orc_derivatives(const float* src, const float* tgt, float* Ix, float* Iy, float* Iz, int w, int h);
void orc_derivatives2(const float* img, float* Ix, float* Iy, int w, int h);
void orc_lucas_kanade(float* flow2, const float* Ix, const float* Iy, const float* It, int w, int h, int halfWin, float minDet);
void orc_lk_iteration(const float* ref, const float* mov, const float* flow_in2, float* flow_out2, int w, int h, int halfWin, float minDet);
/* kernel.cu:691,718 (+ box mean of the absent host) */
void orc_structure_tensor(const float* Ix, const float* Iy, float* t3, int w, int h);
void orc_box_mean3(const float* in3, float* out3, int w, int h, int r);
void orc_kernel_param(float* k3, int w, int h, float Dth, float Dtr, float kDetail, float kDenoise, float kStretch, float kShrink);
void orc_kernel_params(const float* gray, float* kernel4, int w, int h, int box_r, float Dth, float Dtr,
                       float kDetail, float kDenoise, float kStretch, float kShrink);
/* RobustnessModell.cu:29 — mask4 zero-initialised by caller; flow2 is (2w x 2h) */
void orc_robustness_mask(const float* ref3, const float* mov3, float* mask4, const float* flow2, int flow_w, int flow_h,
                         int w, int h, float alpha, float beta, float thresholdM);
void orc_mask_erode(const float* in4, float* out4, int w, int h, int r);
/* DeBayerKernels.cu:379 (scale 2) / :290 (scale 1), generalised by orc_merge_geom; RMW one frame */
void orc_accumulate(const uint16_t* raw, float* sum3, float* weight3, const float* mask4, const float* kernel4,
                    const float* flow2, const orc_merge_geom* g, const int cfa[4], const float white[3], const float black[3]);
/* kernel.cu:426, :393 */
void orc_apply_weighting(float* inout3, const float* final3, const float* weight3, int w, int h, float threshold);
void orc_gamma_srgb(float* img3, int w, int h);
void orc_fallback_upsample(const float* rgb3, int w, int h, float* out3, const orc_merge_geom* g);

#ifdef __cplusplus
}
#endif
#endif
