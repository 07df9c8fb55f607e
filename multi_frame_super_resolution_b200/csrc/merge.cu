// merge.cu — kernel-regression merge onto the HR grid, fused over all frames.
//
// Replaces, in ONE launch:  N x accumulateImagesSuperRes (DeBayerKernels.cu:379;
// scale 1: accumulateImages :290) + ApplyWeighting (kernel.cu:426) + GammasRGB
// (kernel.cu:393).  The reference keeps float3 sum/weight images in HBM and
// read-modify-writes them once per frame (48 B/px/frame); here the accumulators
// live in registers for the whole frame loop and the image is written once.
//
// Two kernels:
//   merge_generic_kernel  any scale / geometry, one thread per output pixel,
//                         direct global gathers.  Reference semantics, simple.
//   merge_s2_kernel       scale 2 fast path (see below).
#include "merge_common.cuh"
#include <stdlib.h>

namespace mfsr {

// S: compile-time scale (1..4; 0 = read it from the geometry).  With a runtime scale every tap pays four integer divisions
// (~80 of ~90 instructions per tap); as a template constant they are multiply-shift sequences.
template <int S>
__global__ void __launch_bounds__(256)
merge_generic_kernel(const __grid_constant__ MergeArgs A)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const mfsr_merge_geom& g = A.g;
    if (x >= g.out_w || y >= g.out_h) return;

    float acc[3] = {0.f, 0.f, 0.f}, wacc[3] = {0.f, 0.f, 0.f};
    const int s = S ? S : g.scale;
    // the reference skips the 1-px border of the output window (DeBayerKernels.cu:391)
    const bool interior = !(x < 1 || y < 1 || x >= g.out_w - 1 || y >= g.out_h - 1);
    if (interior) {
        const int X = x + g.org_x, Y = y + g.org_y;
        const float u = __fdiv_rn((float)X + 0.5f, (float)s), v = __fdiv_rn((float)Y + 0.5f, (float)s);
        const TexAxis tx = tex_axis(u, g.raw_w), ty = tex_axis(v, g.raw_h);
        // kernel parameter fetch (float4 texture, :401)
        float kx, ky, kz;
        {
            const float4 k00 = row_ptr(A.kern, A.kern_pitch, ty.i0)[tx.i0], k10 = row_ptr(A.kern, A.kern_pitch, ty.i0)[tx.i1];
            const float4 k01 = row_ptr(A.kern, A.kern_pitch, ty.i1)[tx.i0], k11 = row_ptr(A.kern, A.kern_pitch, ty.i1)[tx.i1];
            kx = tex_mix(k00.x, k10.x, k01.x, k11.x, tx.a, ty.a);
            ky = tex_mix(k00.y, k10.y, k01.y, k11.y, tx.a, ty.a);
            kz = tex_mix(k00.z, k10.z, k01.z, k11.z, tx.a, ty.a);
        }
        // 25 regression weights: frame independent, computed once (:427-430)
        float w[25];
#pragma unroll
        for (int py = -2; py <= 2; py++)
#pragma unroll
            for (int px = -2; px <= 2; px++) {
                float q = (float)(px * px) * kx + (float)(2 * px * py) * kz + (float)(py * py) * ky;
                float e = expf(-0.5f * q);
                if (!isfinite(e)) e = (px * py == 0) ? 1.0f : 0.0f;
                w[(py + 2) * 5 + (px + 2)] = e;
            }
        for (int f = 0; f < A.n_frames; f++) {
            const float2* flow = (const float2*)((const char*)A.flow + A.flow_fs * f);
            const uint16_t* raw = (const uint16_t*)((const char*)A.raw + A.raw_fs * f);
            const float4* mask = (const float4*)((const char*)A.mask + A.mask_fs * f);
            const float2 s00 = row_ptr(flow, A.flow_pitch, ty.i0)[tx.i0], s10 = row_ptr(flow, A.flow_pitch, ty.i0)[tx.i1];
            const float2 s01 = row_ptr(flow, A.flow_pitch, ty.i1)[tx.i0], s11 = row_ptr(flow, A.flow_pitch, ty.i1)[tx.i1];
            const int sx = (int)roundf(__fmul_rn(tex_mix(s00.x, s10.x, s01.x, s11.x, tx.a, ty.a), (float)s));
            const int sy = (int)roundf(__fmul_rn(tex_mix(s00.y, s10.y, s01.y, s11.y, tx.a, ty.a), (float)s));
#pragma unroll
            for (int py = -2; py <= 2; py++) {
                const int ppsy = min(max((Y + py + sy) / s, g.clamp_y0), g.clamp_y1);
                const int ppy = min(max((Y + py) / s, g.clamp_y0), g.clamp_y1);
                const uint16_t* rrow = row_ptr(raw, A.raw_pitch, ppsy);
                const float4* mrow = row_ptr(mask, A.mask_pitch, ppy / 2);
#pragma unroll
                for (int px = -2; px <= 2; px++) {
                    const int ppsx = min(max((X + px + sx) / s, g.clamp_x0), g.clamp_x1);
                    const int ppx = min(max((X + px) / s, g.clamp_x0), g.clamp_x1);
                    const int col = A.cfa.c[(ppsy & 1) * 2 + (ppsx & 1)];
                    const float wt = w[(py + 2) * 5 + (px + 2)];
                    const float r = (float)__ldg(rrow + ppsx);
                    const float4 m = __ldg(mrow + (ppx / 2));
                    float cert = col == 0 ? m.x : (col == 1 ? m.y : m.z);
                    if (!isfinite(cert)) cert = 0.0f;
                    // one division per tap (the colour only selects its operands), same expression order as the reference
                    const float bl = col == 0 ? A.black[0] : (col == 1 ? A.black[1] : A.black[2]);
                    const float wh = col == 0 ? A.white[0] : (col == 1 ? A.white[1] : A.white[2]);
                    const float rn = (r - bl) / wh;
                    const float v = rn * wt * cert, wv = wt * cert;
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        if (col == c) { acc[c] += v; wacc[c] += wv; }
                }
            }
        }
    }
    if (A.sum_out) {
        float* so = row_ptr(A.sum_out, A.acc_pitch, y) + 3 * x;
        float* wo = row_ptr(A.weight_out, A.acc_pitch, y) + 3 * x;
        so[0] = acc[0]; so[1] = acc[1]; so[2] = acc[2];
        wo[0] = wacc[0]; wo[1] = wacc[1]; wo[2] = wacc[2];
    }
    float fb[3] = {0.f, 0.f, 0.f};
    if (A.fallback) {
        const float* p = row_ptr(A.fallback, A.fb_pitch, y) + 3 * x;
        fb[0] = p[0]; fb[1] = p[1]; fb[2] = p[2];
    }
    float* o = row_ptr(A.out, A.out_pitch, y) + 3 * x;
#pragma unroll
    for (int c = 0; c < 3; c++) o[c] = finish_px(apply_weighting(acc[c], wacc[c], fb[c], A.threshold), A.flags);
}

}  // namespace mfsr

using namespace mfsr;

extern "C" int mfsr_stage_merge(const uint16_t* raw, int64_t raw_pitch, int64_t raw_frame_stride,
                                const float* mask, int64_t mask_pitch, int64_t mask_frame_stride,
                                const float* flow, int64_t flow_pitch, int64_t flow_frame_stride,
                                const float* kernel4, int64_t kernel_pitch,
                                const float* fallback, int64_t fallback_pitch,
                                float* out, int64_t out_pitch,
                                float* sum_out, float* weight_out, int64_t acc_pitch,
                                int n_frames, const mfsr_merge_geom* geom, const int cfa[4],
                                const float white[3], const float black[3],
                                float threshold, int flags, void* stream)
{
    if (!kernel4 || !out || !geom || !cfa || !white || !black) return MFSR_E_INVALID;
    if (n_frames > 0 && (!raw || !mask || !flow)) return MFSR_E_INVALID;
    if (n_frames < 0 || geom->scale < 1 || geom->out_w <= 0 || geom->out_h <= 0) return MFSR_E_INVALID;
    if (!fallback && !(flags & MFSR_MERGE_NO_FALLBACK)) return MFSR_E_INVALID;
    if ((sum_out == nullptr) != (weight_out == nullptr)) return MFSR_E_INVALID;
    if (geom->clamp_x0 < 0 || geom->clamp_x1 >= geom->raw_w || geom->clamp_y0 < 0 || geom->clamp_y1 >= geom->raw_h ||
        geom->clamp_x0 > geom->clamp_x1 || geom->clamp_y0 > geom->clamp_y1) return MFSR_E_INVALID;
    MergeArgs A;
    A.raw = raw; A.raw_pitch = raw_pitch; A.raw_fs = raw_frame_stride;
    A.mask = (const float4*)mask; A.mask_pitch = mask_pitch; A.mask_fs = mask_frame_stride;
    A.flow = (const float2*)flow; A.flow_pitch = flow_pitch; A.flow_fs = flow_frame_stride;
    A.kern = (const float4*)kernel4; A.kern_pitch = kernel_pitch;
    A.fallback = fallback; A.fb_pitch = fallback_pitch;
    A.out = out; A.out_pitch = out_pitch;
    A.sum_out = sum_out; A.weight_out = weight_out; A.acc_pitch = acc_pitch;
    A.sum_in = nullptr; A.weight_in = nullptr;
    A.n_frames = n_frames; A.g = *geom;
    for (int i = 0; i < 4; i++) A.cfa.c[i] = cfa[i];
    for (int i = 0; i < 3; i++) { A.white[i] = white[i]; A.black[i] = black[i]; }
    A.threshold = threshold; A.flags = flags;
    // scale-2 fast path (merge_fast.cu); MFSR_MERGE_GENERIC=1 forces the generic kernel (A/B tests)
    static const bool force_generic = getenv("MFSR_MERGE_GENERIC") != nullptr;
    if (geom->scale == 2 && n_frames > 0 && !force_generic) {
        const int rc = launch_merge_s2(A, (cudaStream_t)stream);
        if (rc != MFSR_E_INVALID) return rc;
    }
    dim3 block(32, 8), grid(cdiv(geom->out_w, 32), cdiv(geom->out_h, 8));
    switch (geom->scale) {
        case 1: merge_generic_kernel<1><<<grid, block, 0, (cudaStream_t)stream>>>(A); break;
        case 2: merge_generic_kernel<2><<<grid, block, 0, (cudaStream_t)stream>>>(A); break;
        case 3: merge_generic_kernel<3><<<grid, block, 0, (cudaStream_t)stream>>>(A); break;
        case 4: merge_generic_kernel<4><<<grid, block, 0, (cudaStream_t)stream>>>(A); break;
        default: merge_generic_kernel<0><<<grid, block, 0, (cudaStream_t)stream>>>(A); break;
    }
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}
