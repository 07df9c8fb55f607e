// merge_common.cuh — argument block and epilogue helpers shared by merge.cu (generic kernel, any scale)
// and merge_fast.cu (scale-2 shared-memory kernel).
#pragma once
#include "common.cuh"

namespace mfsr {

struct MergeArgs {
    const uint16_t* raw;  int64_t raw_pitch,  raw_fs;
    const float4*   mask; int64_t mask_pitch, mask_fs;
    const float2*   flow; int64_t flow_pitch, flow_fs;
    const float4*   kern; int64_t kern_pitch;
    const float*    fallback; int64_t fb_pitch;
    float*          out;  int64_t out_pitch;
    float*          sum_out; float* weight_out; int64_t acc_pitch;
    const float*    sum_in;  const float* weight_in;      // internal (frame-chunked scale-2 merge): partial sums of earlier frames, acc_pitch
    int n_frames;
    mfsr_merge_geom g;
    Cfa cfa;
    float white[3], black[3];
    float threshold;
    int flags;
};

// ApplyWeighting (kernel.cu:426) for one channel; `fb` is the reference's inOutImg value.
__device__ __forceinline__ float apply_weighting(float val, float w, float fb, float threshold)
{
    if (w < threshold) { val += fb; w += 1.0f; }
    return (w != 0.0f) ? val / w : 0.0f;
}
__device__ __forceinline__ float finish_px(float v, int flags)
{
    if (flags & MFSR_MERGE_GAMMA) {                // GammasRGB (kernel.cu:393)
        if (isnan(v)) v = 0.0f;
        v = fmaxf(fminf(v, 1.0f), 0.0f);
        v = srgb_gamma(v);
    }
    return v;
}

// internal flag of the frame-chunked scale-2 merge: write the partial sums only, no image
#define MFSR_MERGE_PARTIAL_INTERNAL (1 << 16)

// launcher of the scale-2 fast path (merge_fast.cu); returns MFSR_E_INVALID when the configuration is
// outside what it supports (the caller then runs the generic kernel).
int launch_merge_s2(const MergeArgs& A, cudaStream_t st);

}  // namespace mfsr
