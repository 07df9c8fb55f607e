/*
 * mfsr.h — C ABI of the B200-native burst multi-frame super-resolution path.
 *
 * This is the drop-in boundary for the hot path of
 * zhongzisha/multi_frame_super_resolution (the host-less ImageStackAlignator
 * kernels under test_opencv/{kernel,...}.cu).  Conventions follow the reference's own
 * `extern "C"` wrappers (test_opencv/myKernels.cu:114-343): the CALLER owns all
 * device memory, images are raw device pointers + byte pitches + dims, scalars
 * are passed by value, nothing throws, nothing calls exit().  Unlike the
 * reference wrappers every entry point takes an explicit stream and returns an
 * int status (0 = ok, >0 = cudaError_t value, <0 = MFSR_E_*), because the
 * reference only checks errors in addWithCuda (test_opencv/kernel.cu:37-114).
 *
 * No torch / C++ types appear in any signature.  `stream` is a cudaStream_t
 * passed as void* (NULL = legacy default stream, as in the reference).
 *
 * Name map (reference `extern "C" __global__` symbol -> entry point here):
 *   deBayersSubSample3 (DeBayerKernels.cu:244)                    -> mfsr_stage_subsample3
 *   deBayerGreenKernel + deBayerRedBlueKernel (:55, :153)         -> mfsr_stage_demosaic
 *   [absent host: B/W + gaussin_filter_1D blur, main.cpp:370]     -> mfsr_stage_tracking_image
 *   [absent host: NPP resize pyramid]                             -> mfsr_stage_pyramid_down
 *   convertToTilesOverlapBorder/PreShift (kernel.cu:265,324),
 *     cuFFT R2C/C2R + conjugateComplexMulKernel (:485),
 *     squaredSum (:119), boxFilterWithBorderX/Y (:149,:186),
 *     normalizedCC (:227), findMinimum (:512)                     -> mfsr_stage_tile_align
 *   UpSampleShifts (kernel.cu:642)                                -> mfsr_stage_upsample_shifts
 *   concatenateShifts/copyShiftMatrix/setPointers/transposeShifts/
 *     checkForOutliers/getOptimalShifts/separateShifts
 *     (ShiftMinimizerKernels.cu:223,29,51,143,81,179,242)
 *     + cuBLAS batched gemm/matinv of the absent host             -> mfsr_stage_consolidate_shifts
 *   CreateFlowFieldFromTiles (opticalFlow.cu:48)                  -> mfsr_stage_flow_from_tiles
 *   WarpingKernel + ComputeDerivativesKernel + lucasKanadeOptim
 *     (opticalFlow.cu:28,97,190)                                  -> mfsr_stage_lk_iteration
 *   ComputeDerivatives2Kernel (opticalFlow.cu:151) +
 *     ComputeStructureTensor (kernel.cu:691) + [box smooth] +
 *     ComputeKernelParam (kernel.cu:718)                          -> mfsr_stage_kernel_params
 *   ComputeRobustnessMask (RobustnessModell.cu:29) [+ min filter] -> mfsr_stage_robustness
 *   accumulateImagesSuperRes (DeBayerKernels.cu:379) x N frames +
 *     accumulateImages (:290, scale 1) + ApplyWeighting
 *     (kernel.cu:426) + GammasRGB (kernel.cu:393)                 -> mfsr_stage_merge
 *   whole chain, SURVEY §3.2 A..I                                 -> mfsr_create / mfsr_set_frames / mfsr_run
 */
#ifndef MFSR_H_
#define MFSR_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFSR_ABI_VERSION 1

/* Negative status codes (positive values are cudaError_t). */
#define MFSR_OK            0
#define MFSR_E_INVALID   (-1)   /* bad argument / unsupported configuration */
#define MFSR_E_STATE     (-2)   /* call order violated (e.g. run before set_frames) */
#define MFSR_E_NOMEM     (-3)
#define MFSR_E_NODEVICE  (-4)   /* no sm_100 capable device visible */

/* Bayer colours; numeric values equal the reference enum BayerColor
 * (DeBayerKernels.cu:28-37).  Only R/G/B are handled by the reference kernels. */
#define MFSR_RED   0
#define MFSR_GREEN 1
#define MFSR_BLUE  2

/* Pixel formats accepted by mfsr_set_frames. */
#define MFSR_FMT_BAYER_U16 0    /* CFA mosaic, 16-bit container               */
#define MFSR_FMT_GRAY_U16  1    /* monochrome: cfa = {G,G,G,G}, channel .y     */

/* Flags for mfsr_stage_merge / mfsr_params.merge_flags */
#define MFSR_OUT_F32 0           /* output formats of mfsr_run_format */
#define MFSR_OUT_F16 1
#define MFSR_OUT_U8  2

#define MFSR_MERGE_GAMMA        1   /* apply GammasRGB (kernel.cu:393) in the epilogue     */
#define MFSR_MERGE_NO_FALLBACK  2   /* `fallback` is NULL: treat as all-zero reference img */

/*
 * Every tunable of the hot path.  In the reference all of these are kernel
 * arguments with no default anywhere in the repo; the
 * defaults written by mfsr_default_params() are the values shared with the
 * oracle (oracle/mfsr_oracle.c) and documented in DESIGN.md.
 */
typedef struct mfsr_params {
    int   abi_version;        /* must be MFSR_ABI_VERSION                                   */
    int   scale;              /* output zoom s; the reference hard-codes 2                  */
    int   full_frame;         /* 0: reference geometry (output dims == raw dims, 2x zoom of
                                 the central half-FOV, DeBayerKernels.cu:414-423);
                                 1: whole field of view (output = s*raw dims)               */
    int   cfa[4];             /* c_cfaPattern[2][2] row-major (DeBayerKernels.cu:41)        */
    float black_level[3];     /* per colour R,G,B (accumulateImagesSuperRes blackLevel)     */
    float white_level[3];     /* divisor per colour (whiteLevel); demosaic scale = 1/white  */
    /* alignment */
    int   tile_size;          /* T  (kernel.cu tileSize)                                     */
    int   max_shift;          /* M  (kernel.cu maxShift)                                     */
    int   levels;             /* pyramid levels, factor 2 each: l = 2^(levels-1) .. 1        */
    int   pair_span;          /* measured pairs (i,j), 0 < j-i <= pair_span                  */
    int   track_bits;         /* tracking image quantisation (7 -> exact fp32 SSD, T=16)     */
    float track_sigma;        /* gaussin_filter_1D sigma (main.cpp:370,1868) = 0.5           */
    float min_threshold;      /* findMinimum `threshold` (kernel.cu:519): SSD span below which a tile gets a zero shift; default 1024 */
    float base_shift[2];      /* global pre-alignment (kernel.cu:275-276); 0 = disabled      */
    float base_rotation;
    /* optical flow */
    int   lk_iterations;
    int   lk_half_window;     /* lucasKanadeOptim halfWindowSize (opticalFlow.cu:199)        */
    float lk_min_det;         /* minDet (opticalFlow.cu:200)                                 */
    /* kernel estimation (ComputeKernelParam args, kernel.cu:723-728) */
    float Dth, Dtr, kDetail, kDenoise, kStretch, kShrink;
    int   tensor_box_radius;  /* box smoothing of the structure tensor by the absent host    */
    /* robustness (RobustnessModell.cu:38-40) */
    float alpha, beta, thresholdM;
    int   mask_erode_radius;  /* min filter on the mask by the absent host; 0 = none         */
    /* merge */
    float weight_threshold;   /* ApplyWeighting threshold (kernel.cu:433)                    */
    int   merge_flags;        /* MFSR_MERGE_*                                                */
    /* Row-band mode (one very large burst split over GPUs): the frames given to mfsr_set_frames are
     * rows [band_row0, band_row0 + height) of a burst that is band_global_h rows tall; only output rows
     * s*[band_keep_row0, band_keep_row0 + band_keep_rows) (LOCAL raw rows) are produced.  band_row0 must be a
     * multiple of tile_size << (levels - 1) so that the tile and pyramid grids coincide with the full-frame ones;
     * base_rotation must be 0.  All zero = not a band.
     * band_margin > 0: the per-pixel stages (flow, kernel parameters, robustness) run only on the kept rows +- band_margin
     * rows instead of the whole band + halo (the wide halo is needed by the coarse pyramid levels of the tile matcher only);
     * it must cover the stencil footprints plus the largest vertical flow: 15 + 12 + max |flow_y| rows.  0 = whole band. */
    int   band_global_h, band_row0, band_keep_row0, band_keep_rows;
    int   band_margin;
    /* Global pre-alignment (PreAlignment skeleton boxFilterNPP.cpp:102-166): 1 = estimate one shift + rotation per
     * frame against the reference frame (csrc/prealign.cu: exhaustive search over +-20 degrees and +-8 coarse pixels on two
     * levels of the tracking pyramid) and feed it to the tile matcher and the flow field as their baseShift / baseRotation
     * (kernel.cu:275-276, opticalFlow.cu:57-58) per pair / per frame; it then replaces base_shift / base_rotation.  0 = off. */
    int   prealign;
    /* 1: the Lucas-Kanade warp samples the moved frame through the texture unit, exactly as the reference's WarpingKernel does
     * (opticalFlow.cu:36-41; hardware 1.8 fixed-point filter: bit-identical to the reference kernels, ~0.4 ms faster per 12 MP
     * burst).  0 (default): the ALU model of that filter, bit-identical to the CPU restatement and across row bands. */
    int   lk_texture;
    int   reserved[1];
} mfsr_params;

int         mfsr_abi_version(void);
const char* mfsr_error_string(int status);
/* Fills *p with the defaults shared with the oracle. */
int         mfsr_default_params(mfsr_params* p);
/* Number of visible CUDA devices with compute capability 10.x; <0 on error. */
int         mfsr_device_count(void);

/* ------------------------------------------------------------------------- *
 *  Stage entry points.  All pointers are DEVICE pointers, pitches in BYTES.
 *  float3 images are packed x,y,z (12 B / px), float4 16 B / px, float2 8 B.
 * ------------------------------------------------------------------------- */

/* deBayersSubSample3: Bayer quad -> one RGB pixel, G averaged, * 1/maxVal.
 * raw: dense-or-pitched u16, 2*dimX x 2*dimY; rgb_half: float3 dimX x dimY. */
int mfsr_stage_subsample3(const uint16_t* raw, int64_t raw_pitch,
                          float* rgb_half, int64_t rgb_pitch,
                          float maxVal, int dimX, int dimY, const int cfa[4],
                          void* stream);

/* deBayerGreenKernel + deBayerRedBlueKernel fused (identical values: green is
 * 0 outside [2,dim-2) exactly as the reference's unwritten border).
 * raw u16 is converted to float exactly; (raw - black[c]) * scale[c]. */
int mfsr_stage_demosaic(const uint16_t* raw, int64_t raw_pitch,
                        float* rgb, int64_t rgb_pitch,
                        int width, int height, const int cfa[4],
                        const float black[3], const float scale[3],
                        void* stream);

/* Restated host step: luminance 0.25R+0.5G+0.25B of the demosaiced frame,
 * separable Gaussian (taps from gaussin_filter_1D(sigma), main.cpp:370, clamp
 * border), written as float `gray` and as `track_bits`-bit integer `gray_q`
 * (u8).  Either output may be NULL. */
int mfsr_stage_tracking_image(const uint16_t* raw, int64_t raw_pitch,
                              float* gray, int64_t gray_pitch,
                              uint8_t* gray_q, int64_t gray_q_pitch,
                              int width, int height, const int cfa[4],
                              const float black[3], const float scale[3],
                              float sigma, int track_bits, void* stream);

/* 2x2 integer box: out = (a+b+c+d+2)>>2; out dims = floor(in/2). */
int mfsr_stage_pyramid_down(const uint8_t* in, int64_t in_pitch, int in_w, int in_h,
                            uint8_t* out, int64_t out_pitch, void* stream);

/* Tile block matching for one image pair at one pyramid level: the whole
 * convertToTiles -> FFT cross-correlation -> squaredSum/boxFilter/normalizedCC
 * -> findMinimum chain fused.  The SSD map is computed in exact integer
 * arithmetic (every partial sum of the reference's fp32 formula is exactly
 * representable for track_bits<=7, T<=16) so the arg-min is bit-identical.
 *   pre_shift  : float2 [tilesY][tilesX] (pitched) or NULL (= zeros)
 *   out_shift  : float2 [tilesY][tilesX] (pitched) total tile shift
 *                = findMinimum coord + round(pre_shift + base) - base
 *   out_argmin : optional int2 [tilesY][tilesX] dense: (minIdx % S - M, minIdx / S - M)
 *   out_ssd    : optional float [tiles][S*S] dense (normalizedCC output)        */
int mfsr_stage_tile_align(const uint8_t* ref, const uint8_t* mov, int64_t img_pitch,
                          int width, int height,
                          const float* pre_shift, int64_t pre_shift_pitch,
                          float* out_shift, int64_t out_shift_pitch,
                          int32_t* out_argmin, float* out_ssd,
                          int tile_size, int max_shift, int tilesX, int tilesY,
                          float base_shift_x, float base_shift_y, float base_rotation,
                          float threshold, void* stream);

/* UpSampleShifts (kernel.cu:642). */
/* One stage of the global pre-alignment search on one image pair (8-bit tracking images of one pyramid level): candidates
 * angle a = 0 .. n_ang-1 -> (cos, sin) = cs_table[idx0 + a*step] (device, n_table pairs), shift b = (cx, cy) + [-R, R]^2, every
 * sub-th pixel; a reference pixel with centred coordinates c reads the moved image at p + round(R(theta)(c - b) - c) as
 * kernel.cu:299-311 does.  out3 (device int[3]) = (a or -1, bx, by) of the candidate with the smallest mean squared difference. */
int mfsr_stage_prealign_search(const uint8_t* ref, const uint8_t* mov, int64_t pitch, int width, int height,
                               const float* cs_table, int n_table, int idx0, int step, int n_ang, int cx, int cy, int R, int sub,
                               int* out3, void* stream);
int mfsr_stage_upsample_shifts(const float* in_shift, int64_t in_pitch,
                               float* out_shift, int64_t out_pitch,
                               int oldLevel, int newLevel,
                               int oldCountX, int oldCountY, int newCountX, int newCountY,
                               int oldTileSize, int newTileSize, void* stream);

/* Per-tile least squares of m measured pairwise shifts -> n1 = imageCount-1
 * sequential shifts with iterative largest-outlier removal (> 1 px^2,
 * ShiftMinimizerKernels.cu:109), then getOptimalShifts for every frame.
 *   measured  : float2 [tiles][m] dense (concatenateShifts layout, :238)
 *   pair_from/pair_to : HOST int[m], pair k measures frame pair_from[k] -> pair_to[k]
 *   one_to_one: optional float2 [tiles][n1] dense (shiftsOneToOne)
 *   frame_shift: float2 [imageCount][tilesY][tilesX] dense, shift reference->frame
 *   status    : optional int [tiles]: number of removed measurements       */
int mfsr_stage_consolidate_shifts(const float* measured, const int* pair_from, const int* pair_to,
                                  int m, int imageCount, int tilesX, int tilesY,
                                  int referenceImage,
                                  float* one_to_one, float* frame_shift, int32_t* status,
                                  void* stream);

/* CreateFlowFieldFromTiles: dense float2 flow from the tile-shift grid sampled
 * like a normalised-coordinate linear/clamp texture (8-bit fraction model). */
int mfsr_stage_flow_from_tiles(const float* tile_shift, int64_t tile_pitch,
                               int tilesX, int tilesY, int tile_size,
                               float* flow, int64_t flow_pitch, int width, int height,
                               float base_shift_x, float base_shift_y, float base_rotation,
                               void* stream);

/* One Lucas-Kanade refinement sweep: warp `mov` by `flow`, 5-tap derivatives of
 * `ref` and the warped image, windowed normal equations, closed-form 2x2 SVD
 * pseudo-inverse (incl. the fminf(sigma1,sigma1) quirk, opticalFlow.cu:255),
 * flow += UV.  `flow_out` may alias `flow_in` only if scratch is provided by
 * the pipeline; the stage entry requires distinct buffers. */
int mfsr_stage_lk_iteration(const float* ref, const float* mov, int64_t img_pitch,
                            const float* flow_in, float* flow_out, int64_t flow_pitch,
                            int width, int height, int half_window, float min_det,
                            void* stream);
/* The same sweep with the bilinear fetch of the warp on the TEXTURE UNIT, as the reference's WarpingKernel does it
 * (opticalFlow.cu:36-41); `mov` must be 512-byte aligned with a pitch that is a multiple of 32 bytes. */
int mfsr_stage_lk_iteration_tex(const float* ref, const float* mov, int64_t img_pitch,
                            const float* flow_in, float* flow_out, int64_t flow_pitch,
                            int width, int height, int half_window, float min_det,
                            void* stream);

/* ComputeDerivatives2Kernel -> ComputeStructureTensor -> (2r+1)^2 box mean
 * (clamp border) -> ComputeKernelParam, fused.  Output float4 (b22,b11,-b12,0)/det. */
int mfsr_stage_kernel_params(const float* gray, int64_t gray_pitch,
                             float* kernel4, int64_t kernel_pitch,
                             int width, int height, int box_radius,
                             float Dth, float Dtr, float kDetail, float kDenoise,
                             float kStretch, float kShrink, void* stream);

/* ComputeRobustnessMask on half-resolution RGB (incl. the min/max quirk,
 * RobustnessModell.cu:67-70) followed by a (2r+1)^2 min filter on .xyz.
 *   flow: full-resolution float2 flow (2*w x 2*h), sampled like the texture.
 *   mask: float4 w x h; border pixels are 0 (unwritten in the reference).
 *   scratch: float4 w x h work buffer for the two-kernel form (certainty, then min filter); NULL with erode_radius > 0
 *   selects the fused kernel (what mfsr_run uses): same mask, bit for bit, no intermediate image.     */
int mfsr_stage_robustness(const float* rgb_ref, const float* rgb_mov, int64_t rgb_pitch,
                          const float* flow, int64_t flow_pitch,
                          float* mask, int64_t mask_pitch, float* scratch,
                          int w, int h, float alpha, float beta, float thresholdM,
                          int erode_radius, void* stream);

/* Geometry of a merge launch.  Reference geometry for raw dims (dimX,dimY):
 *   scale=2, out_w=dimX, out_h=dimY, org_x=dimX/2, org_y=dimY/2,
 *   clamp = [dimX/4, dimX/4+dimX/2-1] x [dimY/4, dimY/4+dimY/2-1].        */
/* Scale factors: a plain integer s (1, 2, 3, 4 ...; the reference hard-codes 2, DeBayerKernels.cu:414-423), or a rational
 * num / den encoded in the same int (den in the upper half, 0 == 1): MFSR_SCALE_RATIONAL(3, 2) is 1.5x.  With a rational scale
 * every "/ s" of the reference becomes "* den / num" (integer taps: ((X + px + sx) * den) / num, texture coordinate
 * ((X + 0.5) * den) / num, integer shift round(flow * num / den)); integer scales are the den == 1 case, bit for bit. */
#define MFSR_SCALE_RATIONAL(num, den) ((int)(((unsigned)(den) << 16) | (unsigned)(num)))
#define MFSR_SCALE_NUM(s) ((int)((unsigned)(s) & 0xffffu))
#define MFSR_SCALE_DEN(s) ((int)((((unsigned)(s) >> 16) & 0xffffu) ? (((unsigned)(s) >> 16) & 0xffffu) : 1u))

typedef struct mfsr_merge_geom {
    int raw_w, raw_h;        /* dimX, dimY of the raw frames                           */
    int scale;               /* s, or MFSR_SCALE_RATIONAL(num, den)                    */
    int out_w, out_h;        /* output window size in HR pixels                        */
    int org_x, org_y;        /* HR coordinate of output pixel (0,0)                    */
    int clamp_x0, clamp_x1;  /* inclusive raw-coordinate clamp of the taps             */
    int clamp_y0, clamp_y1;
} mfsr_merge_geom;

/* Fused kernel-regression merge over all N frames + normalisation.
 *   raw      : u16 frames, frame f at raw + f*raw_frame_stride (bytes), row pitch raw_pitch
 *   mask     : float4 (raw_w/2 x raw_h/2) per frame, same addressing scheme
 *   flow     : float2 (raw_w x raw_h) per frame
 *   kernel4  : float4 (raw_w x raw_h) inverse covariance (x,y,z used)
 *   fallback : float3 out_w x out_h (ApplyWeighting inOutImg) or NULL w/ NO_FALLBACK
 *   out      : float3 out_w x out_h
 *   sum_out/weight_out : optional float3 dumps of the raw accumulators (may be NULL) */
int mfsr_stage_merge(const uint16_t* raw, int64_t raw_pitch, int64_t raw_frame_stride,
                     const float* mask, int64_t mask_pitch, int64_t mask_frame_stride,
                     const float* flow, int64_t flow_pitch, int64_t flow_frame_stride,
                     const float* kernel4, int64_t kernel_pitch,
                     const float* fallback, int64_t fallback_pitch,
                     float* out, int64_t out_pitch,
                     float* sum_out, float* weight_out, int64_t acc_pitch,
                     int n_frames, const mfsr_merge_geom* geom, const int cfa[4],
                     const float white[3], const float black[3],
                     float threshold, int flags, void* stream);

/* Bilinear x`scale` upsampling of the demosaiced reference to the merge
 * output window (the ApplyWeighting fallback image built by the absent host). */
int mfsr_stage_fallback_upsample(const float* rgb, int64_t rgb_pitch, int width, int height,
                                 float* out, int64_t out_pitch, const mfsr_merge_geom* geom,
                                 void* stream);

/* ------------------------------------------------------------------------- *
 *  Pipeline handle: SURVEY §3.2 A..I for one burst on one GPU.
 *  Mirrors the reference's pull-style use of cv::superres
 *  (finalProject/Project/multi_frame_sr.cpp:165-194): create, configure,
 *  give it the frames, pull the result.
 * ------------------------------------------------------------------------- */
typedef struct mfsr_context* mfsr_handle;

/* device: CUDA ordinal.  max_w/max_h/max_frames size the workspace (one cudaMalloc). */
int mfsr_create(const mfsr_params* params, int device, int max_w, int max_h, int max_frames,
                mfsr_handle* out);
int mfsr_destroy(mfsr_handle h);
/* Output size for a given raw size under the handle's params. */
int mfsr_output_size(mfsr_handle h, int width, int height, int* out_w, int* out_h);
/* Bytes of device workspace the handle owns. */
int64_t mfsr_workspace_bytes(mfsr_handle h);

/* frames[i]: pointer to frame i (u16, row pitch `pitch` bytes).  on_host != 0:
 * pointers are HOST memory (pinned or pageable) and are copied H2D on the
 * handle's stream; otherwise they are device pointers: an evenly spaced stack with a 16-byte aligned base and
 * pitch is used IN PLACE (keep it alive and unchanged until the run has finished), anything else is copied. */
int mfsr_set_frames(mfsr_handle h, const void* const* frames, int n, int width, int height,
                    int64_t pitch, int format, int ref_idx, int on_host);
/* Runs the whole chain on the handle's stream.  `out`: float3 image
 * (out_w x out_h), device pointer, or host pointer when out_on_host != 0
 * (copied D2H and synchronised).  Device output is NOT synchronised. */
int mfsr_run(mfsr_handle h, float* out, int64_t out_pitch, int out_on_host);
/* Same, but never synchronises: with a (pinned) host `out` the D2H copy is only enqueued; call
 * mfsr_synchronize(h) before reading it or re-using the handle's frames.  Two handles driven alternately
 * overlap one burst's PCIe transfers with the other's kernels (the way bench.py measures `e2e`). */
int mfsr_run_async(mfsr_handle h, float* out, int64_t out_pitch, int out_on_host);
/* Same chain, result delivered in `out_format`: MFSR_OUT_F32 (float3, identical to mfsr_run), MFSR_OUT_F16 (half3, round to
 * nearest: within 2.5e-4 of the float image on [0,1]) or MFSR_OUT_U8 (floor(v*255+0.5) saturated, the 8-bit image the reference
 * program writes after GammasRGB, multi_frame_sr.cpp:207 — set MFSR_MERGE_GAMMA in merge_flags for sRGB coding).  `out_pitch`
 * in bytes of the chosen format.  async != 0 behaves like mfsr_run_async.  Cuts the D2H volume of a host result 2x / 4x. */
int mfsr_run_format(mfsr_handle h, void* out, int64_t out_pitch, int out_on_host, int out_format, int async);
int mfsr_synchronize(mfsr_handle h);
void* mfsr_stream(mfsr_handle h);

/* Integer tile arg-min of the finest level for measured pair k (bit-exact
 * contract) and the consolidated reference->frame tile shifts. Host buffers. */
int mfsr_get_tile_grid(mfsr_handle h, int* tilesX, int* tilesY, int* n_pairs);
int mfsr_get_tile_argmin(mfsr_handle h, int pair, int32_t* host_int2);
int mfsr_get_tile_shifts(mfsr_handle h, int frame, float* host_float2);
/* Stage timings of the last mfsr_run in milliseconds (CUDA events on the
 * handle's stream). names: see mfsr_stage_name(i); returns count written. */
int mfsr_get_stage_ms(mfsr_handle h, float* ms, int capacity);
const char* mfsr_stage_name(int i);
/* Device pointers to intermediates of the last run (for stage A/B tests):
 * "flow","mask","kernel","fallback","gray","gray_q0","rgb_half","raw". */
int mfsr_get_buffer(mfsr_handle h, const char* name, void** dev_ptr, int64_t* pitch,
                    int64_t* frame_stride);
/* Number of kernels launched by the last mfsr_run. */
int mfsr_last_launch_count(mfsr_handle h);

#ifdef __cplusplus
}
#endif
#endif /* MFSR_H_ */
