// merge_s2_common.cuh — helpers of the scale-2 merge kernel (merge_dyn.cu): shift evaluation and the per-pixel
// generic tap loop used for clamped taps, alignment outliers and outsized shifts.
#pragma once
#include "merge_common.cuh"
#include "merge_taps.h"
#include <stdio.h>
#include <stdlib.h>

namespace mfsr {
namespace s2 {

constexpr int TW = 128;          // output tile width
constexpr int RWS = 96;          // staged raw window: columns (64 + taps 3 + alignment 3 + shift slack +-13 raw px)
constexpr int MWS = 34;          // staged certainty window: columns (TW/4 + 2)
constexpr int MAXF = 40;         // frames the shared-memory bookkeeping is sized for

struct FrameInfo { int rx0, ry0, need_h, slow; };

struct FastArgs {
    MergeArgs a;
    float inv_white[3];
    float black_ph[4], inv_ph[4];   // black level / reciprocal white level per CFA phase (y parity * 2 + x parity)
    int x_off, y_off;            // tile grid origin: window x of tile column 0 is -x_off (absolute X multiple of 4)
    unsigned ph2c;               // bit 3q + c: CFA phase q has colour c (merge_pf.cu epilogue)
    float cfa_sel[4][3];         // 1 / 0: CFA phase q has colour c (certainty staging of merge_pf.cu)
    float nbi_ph[4];             // -black * 1/white per CFA phase (raw staging of merge_pf.cu)
    int in_x0, in_x1, in_y0, in_y1;   // merge_pf.cu: the tile kernel's pixels (window coordinates); the rest is the clamp band's
    int use_tma;                 // merge_pf.cu: fallback / result rows of a tile move by TMA bulk copies
};

// merge_pf.cu: the predicate-free slot kernel (16-row tiles) and the number of frames it keeps resident
int launch_merge_pf(const FastArgs& F, cudaStream_t st);
int merge_pf_capacity();

// integer HR shift of absolute HR pixel (X, Y) in frame `flow`: round(2 * tex(flow)) with the 1.8 fixed-point
// bilinear model of common.cuh (fractions are exactly .25/.75 at scale 2).
__device__ __forceinline__ float mix25(float left, float right, bool frac75)
{
    // frac75: left*.25 + right*.75 ; else left*.75 + right*.25 (strict rounding: the *.25 product is exact)
    return frac75 ? __fmaf_rn(left, 0.25f, __fmul_rn(right, 0.75f)) : __fmaf_rn(right, 0.25f, __fmul_rn(left, 0.75f));
}

__device__ __forceinline__ int2 shift_global(const MergeArgs& A, int f, int X, int Y)
{
    const mfsr_merge_geom& g = A.g;
    const float2* flow = (const float2*)((const char*)A.flow + A.flow_fs * f);
    const int fx = (X - 1) >> 1, fy = (Y - 1) >> 1;
    const int x0 = clampi(fx, 0, g.raw_w - 1), x1 = clampi(fx + 1, 0, g.raw_w - 1);
    const int y0 = clampi(fy, 0, g.raw_h - 1), y1 = clampi(fy + 1, 0, g.raw_h - 1);
    const float2 s00 = __ldg(row_ptr(flow, A.flow_pitch, y0) + x0), s10 = __ldg(row_ptr(flow, A.flow_pitch, y0) + x1);
    const float2 s01 = __ldg(row_ptr(flow, A.flow_pitch, y1) + x0), s11 = __ldg(row_ptr(flow, A.flow_pitch, y1) + x1);
    const bool ax = !(X & 1), ay = !(Y & 1);
    const float vx = mix25(mix25(s00.x, s10.x, ax), mix25(s01.x, s11.x, ax), ay);
    const float vy = mix25(mix25(s00.y, s10.y, ax), mix25(s01.y, s11.y, ax), ay);
    return make_int2((int)roundf(__fmul_rn(vx, 2.0f)), (int)roundf(__fmul_rn(vy, 2.0f)));
}

// The reference's tap loop for one pixel with a known shift, reading global memory (clamps and all).
// acc/wacc are indexed by the ABSOLUTE CFA phase (y parity * 2 + x parity) of the raw sample.
// w13 / ab are LOCAL-memory temporaries of the caller (never the register-resident accumulators).
static __device__ __noinline__ void generic_pixel(const FastArgs& F, int f, int X, int Y, int sx, int sy, const float* w13, float* ab)
{
    const MergeArgs& A = F.a;
    const mfsr_merge_geom& g = A.g;
    const uint16_t* raw = (const uint16_t*)((const char*)A.raw + A.raw_fs * f);
    const float4* mask = (const float4*)((const char*)A.mask + A.mask_fs * f);
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    for (int py = -2; py <= 2; py++) {
        const int ppsy = min(max((Y + py + sy) / 2, g.clamp_y0), g.clamp_y1);
        const int ppy = min(max((Y + py) / 2, g.clamp_y0), g.clamp_y1);
        const uint16_t* rrow = row_ptr(raw, A.raw_pitch, ppsy);
        const float4* mrow = row_ptr(mask, A.mask_pitch, ppy / 2);
        for (int px = -2; px <= 2; px++) {
            const int ppsx = min(max((X + px + sx) / 2, g.clamp_x0), g.clamp_x1);
            const int ppx = min(max((X + px) / 2, g.clamp_x0), g.clamp_x1);
            const int q = (ppsy & 1) * 2 + (ppsx & 1);
            const int col = A.cfa.c[q];
            const int apx = (py < 0 || (py == 0 && px < 0)) ? -px : px, apy = (py < 0 || (py == 0 && px < 0)) ? -py : py;
            const float wt = w13[apy == 0 ? apx : (apy == 1 ? 5 + apx : 10 + apx)];
            const float4 m = __ldg(mrow + (ppx / 2));
            float cert = col == 0 ? m.x : (col == 1 ? m.y : m.z);
            if (!isfinite(cert)) cert = 0.0f;
            const float rn = ((float)__ldg(rrow + ppsx) - A.black[col]) * F.inv_white[col];
            const float t = wt * cert;
            a[q] += t * rn; b[q] += t;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) { ab[q] = a[q]; ab[4 + q] = b[q]; }
}


}  // namespace s2
}  // namespace mfsr
