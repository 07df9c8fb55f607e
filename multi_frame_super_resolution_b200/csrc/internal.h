// internal.h — private (non-ABI) launchers shared between translation units.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace mfsr {

constexpr int CONS_MAX_M = 64, CONS_MAX_N = 32;
struct PairTable { int8_t from[CONS_MAX_M], to[CONS_MAX_M]; };

// Batched tile matcher: pair k matches image pt.from[k] (template) against pt.to[k], images of one
// pyramid level stored as a stack (frame f at img + f*frame_stride).  Per-pair strides in BYTES.
struct TileAlignBatch {
    const uint8_t* img; int64_t pitch, frame_stride; int w, h;
    const float2* pre; int64_t pre_pitch, pre_pair_stride;          // may be null
    float2* out; int64_t out_pitch, out_pair_stride;
    int2* argmin; int64_t argmin_pair_stride;                        // may be null (dense [ty][tx] per pair)
    float* ssd; int64_t ssd_pair_stride;                             // may be null
    PairTable pt; int n_pairs;
    int T, M, tx, ty;
    float bsx, bsy, rot, threshold;
    // global pre-alignment (prealign.cu): per pair (shift x, shift y in full-resolution pixels, cos, sin) on the device, or null;
    // when given it replaces bsx / bsy / rot.  pose_scale = 1 / 2^level.
    const float* pair_pose; float pose_scale;
};
int launch_tile_align(const TileAlignBatch& b, cudaStream_t st);

struct UpsampleBatch {
    const float2* in; int64_t in_pitch, in_pair_stride;
    float2* out; int64_t out_pitch, out_pair_stride;
    int n_pairs, oldLevel, newLevel, oldCX, oldCY, newCX, newCY, oldT, newT;
};
int launch_upsample_shifts(const UpsampleBatch& b, cudaStream_t st);

// measured element (tile t, pair k) at measured[t*tile_stride + k*pair_stride] (in float2 units)
int launch_consolidate(const float2* measured, int64_t tile_stride, int64_t pair_stride, const PairTable& pt, int m,
                       int imageCount, int nTiles, int referenceImage, float2* one_to_one, float2* frame_shift,
                       int* status, float* inv0_scratch, cudaStream_t st,       // inv0_scratch: (n-1)^2 + 1 floats or null
                       const unsigned long long* active0 = nullptr);            // device: bit k = pair k takes part (null: all)

// CreateFlowFieldFromTiles for a row band: pixel rows and tile rows are mapped through the FULL frame's normalised texture
// coordinates (global height gh, global tile rows gty; local row 0 is global row gy0, local tile row 0 is global tile row
// gy0 / T) so that a band reproduces the full-frame flow.  gh == 0: plain local mapping.
int launch_flow_from_tiles(const float2* tiles, int64_t tile_pitch, int tilesX, int tilesY, float2* flow, int64_t flow_pitch, int w, int h,
                           float bsx, float bsy, float rot, int gh, int gy0, int gty, int tile_row0, cudaStream_t st,
                           const float* frame_pose = nullptr,      // frame_pose: device (bx, by, cos, sin) replacing bsx / bsy / rot
                           int frames = 1, int64_t tiles_fs = 0, int64_t flow_fs = 0);      // one launch for `frames` frames (strides in bytes)

// per-frame kernels launched once per burst (grid.z = frame, frame strides in bytes)
int launch_subsample3(const uint16_t* raw, int64_t raw_pitch, int64_t raw_fs, float* rgb_half, int64_t rgb_pitch, int64_t rgb_fs, int frames,
                      float maxVal, int dimX, int dimY, const int cfa[4], cudaStream_t stream);
int launch_tracking_image(const uint16_t* raw, int64_t raw_pitch, int64_t raw_fs, float* gray, int64_t gray_pitch, int64_t gray_fs,
                          uint8_t* gray_q, int64_t gray_q_pitch, int64_t gray_q_fs, int frames, int width, int height, const int cfa[4],
                          const float black[3], const float scale[3], float sigma, int track_bits, cudaStream_t stream);
int launch_pyramid_down(const uint8_t* in, int64_t in_pitch, int64_t in_fs, int in_w, int in_h, uint8_t* out, int64_t out_pitch, int64_t out_fs, int frames,
                        cudaStream_t stream);
int launch_robustness(const float* rgb_ref, const float* rgb_mov, int64_t rgb_pitch, int64_t rgb_fs, const float* flow, int64_t flow_pitch,
                      int64_t flow_fs, float* mask, int64_t mask_pitch, int64_t mask_fs, float* scratch, int64_t scratch_fs, int frames,
                      int w, int h, float alpha, float beta, float thresholdM, int erode_radius, cudaStream_t st,
                      float* ref_stats = nullptr);      // ref_stats: w * h * 6 floats of scratch: the reference part of the certainty computed once

// global pre-alignment (prealign.cu)
int launch_prealign_stage(const uint8_t* img, int64_t pitch, int64_t frame_stride, int w, int h, int n_frames, int ref_idx,
                          const float* cs, int zero_idx, void* fs_in, int step, int n_ang, int R, int sub,
                          unsigned long long* ssd, unsigned* cnt, int* result3,
                          void* fs_next, int next_scale, int next_half, int next_step, float* pose, int pose_scale, cudaStream_t st);
int launch_prealign_init(void* fs, int n, int idx0, cudaStream_t st);
int launch_pair_pose(const float* pose, const PairTable& pt, int m, float* pair_pose, unsigned long long* pair_valid, cudaStream_t st);

// one Lucas-Kanade sweep; gh / gy0 as above (0: whole frame)
int launch_lk_iteration(const float* ref, const float* mov, int64_t img_pitch, const float2* flow_in, float2* flow_out, int64_t flow_pitch,
                        int width, int height, int half_window, float min_det, int gh, int gy0, cudaStream_t st, cudaTextureObject_t movtex = 0,
                        int frames = 1, int64_t mov_fs = 0, int64_t flow_fs = 0, int ref_frame = -1);
// movtex != 0 (and not a band): the warp step samples the moved image through this texture (make_gray_texture) instead of the ALU model
int make_gray_texture(const float* img, int64_t pitch, int width, int height, cudaTextureObject_t* out);

}  // namespace mfsr
