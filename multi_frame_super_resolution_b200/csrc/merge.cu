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
    // S == 0: the scale comes from the geometry and may be rational (MFSR_SCALE_RATIONAL(num, den)): "/ s" is "* den / num"
    const int s = S ? S : MFSR_SCALE_NUM(g.scale), den = S ? 1 : MFSR_SCALE_DEN(g.scale);
    const float sf = S ? (float)S : __fdiv_rn((float)s, (float)den);
    // the reference skips the 1-px border of the output window (DeBayerKernels.cu:391)
    const bool interior = !(x < 1 || y < 1 || x >= g.out_w - 1 || y >= g.out_h - 1);
    if (interior) {
        const int X = x + g.org_x, Y = y + g.org_y;
        const float u = S ? __fdiv_rn((float)X + 0.5f, (float)s) : __fdiv_rn(__fmul_rn((float)X + 0.5f, (float)den), (float)s);
        const float v = S ? __fdiv_rn((float)Y + 0.5f, (float)s) : __fdiv_rn(__fmul_rn((float)Y + 0.5f, (float)den), (float)s);
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
            const int sx = (int)roundf(__fmul_rn(tex_mix(s00.x, s10.x, s01.x, s11.x, tx.a, ty.a), sf));
            const int sy = (int)roundf(__fmul_rn(tex_mix(s00.y, s10.y, s01.y, s11.y, tx.a, ty.a), sf));
#pragma unroll
            for (int py = -2; py <= 2; py++) {
                const int ppsy = min(max((Y + py + sy) * den / s, g.clamp_y0), g.clamp_y1);
                const int ppy = min(max((Y + py) * den / s, g.clamp_y0), g.clamp_y1);
                const uint16_t* rrow = row_ptr(raw, A.raw_pitch, ppsy);
                const float4* mrow = row_ptr(mask, A.mask_pitch, ppy / 2);
#pragma unroll
                for (int px = -2; px <= 2; px++) {
                    const int ppsx = min(max((X + px + sx) * den / s, g.clamp_x0), g.clamp_x1);
                    const int ppx = min(max((X + px) * den / s, g.clamp_x0), g.clamp_x1);
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

// ---- lean tap loop for the scales without a slot kernel (S >= 3; config 5 runs 3x) -------------------------------------------------
// Same taps, same clamps, same certainty / normalisation per tap as merge_generic_kernel, organised so that a tap costs ~25 instead of
// ~70 instructions (the generic loop's per-tap colour branches, IEEE division with its slow path, parameter-space indexing and index
// arithmetic were 1763 instructions per pixel and frame at S = 3):
//  * everything that depends on one axis only is computed once per axis and frame (5 + 5 clamped raw coordinates, their parities, row
//    pointers) or once per pixel (mask cells of the 5 + 5 unshifted taps: they do not depend on the frame);
//  * the 5 taps of a row read at most 3 consecutive raw samples and 2 mask cells: 3 + 2 loads per row instead of 5 + 5, normalised once
//    per sample, selected per tap by predicates that are computed once per column;
//  * the colour of a tap is the CFA entry of the ABSOLUTE parity of its raw sample, so value and weight are accumulated per parity class
//    (row parity x column parity) with predicated adds and mapped to colours once, after the frame loop — no per-tap colour branches;
//  * normalisation is (r - black) * (1 / white) with the reciprocal taken once per colour (<= 1 ulp from the division).
// Green's two parity classes are summed separately and added at the end: the result differs from the reference's summation order by
// fp32 round-off (the parity tests' bar is 1e-3).
__device__ __forceinline__ void padd(float& a, float v, bool p)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q add.f32 %0, %0, %1;\n\t}" : "+f"(a) : "f"(v), "r"((unsigned)p));
}

#ifndef LEAN_OCC
#define LEAN_OCC 2          // resident CTAs per SM the register allocation aims at (226 registers unconstrained = one CTA = 8 warps per SM)
#endif
template <int S>
__global__ void __launch_bounds__(256, LEAN_OCC)
merge_lean_kernel(const __grid_constant__ MergeArgs A)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const mfsr_merge_geom& g = A.g;
    if (x >= g.out_w || y >= g.out_h) return;
    float acc[3] = {0.f, 0.f, 0.f}, wacc[3] = {0.f, 0.f, 0.f};
    const bool interior = !(x < 1 || y < 1 || x >= g.out_w - 1 || y >= g.out_h - 1);
    if (interior) {
        const int X = x + g.org_x, Y = y + g.org_y;
        const float u = __fdiv_rn((float)X + 0.5f, (float)S), v = __fdiv_rn((float)Y + 0.5f, (float)S);
        const TexAxis tx = tex_axis(u, g.raw_w), ty = tex_axis(v, g.raw_h);
        float kx, ky, kz;
        {
            const float4 k00 = row_ptr(A.kern, A.kern_pitch, ty.i0)[tx.i0], k10 = row_ptr(A.kern, A.kern_pitch, ty.i0)[tx.i1];
            const float4 k01 = row_ptr(A.kern, A.kern_pitch, ty.i1)[tx.i0], k11 = row_ptr(A.kern, A.kern_pitch, ty.i1)[tx.i1];
            kx = tex_mix(k00.x, k10.x, k01.x, k11.x, tx.a, ty.a);
            ky = tex_mix(k00.y, k10.y, k01.y, k11.y, tx.a, ty.a);
            kz = tex_mix(k00.z, k10.z, k01.z, k11.z, tx.a, ty.a);
        }
        // 13 distinct weights: the quadratic form is even, w(py, px) == w(-py, -px); tap t = 5 py + px uses w[min(t, 24 - t)]
        float w[13];
#pragma unroll
        for (int t = 0; t < 13; t++) {
            const int py = t / 5 - 2, px = t % 5 - 2;
            float q = (float)(px * px) * kx + (float)(2 * px * py) * kz + (float)(py * py) * ky;
            float e = expf(-0.5f * q);
            if (!isfinite(e)) e = (px * py == 0) ? 1.0f : 0.0f;
            w[t] = e;
        }
        // frame independent: the mask cells of the 5 + 5 unshifted taps.  5 consecutive HR positions cover at most 3 raw pixels, hence
        // at most 2 half-resolution cells per axis: cell 0 = the first tap's, cell 1 = the last tap's, cxb / cyb say which one a tap uses
        int mx0, mx1, my0, my1; bool cxb[5], cyb[5];
        {
            int cx[5], cy[5];
#pragma unroll
            for (int t = 0; t < 5; t++) {
                cx[t] = min(max((X + t - 2) / S, g.clamp_x0), g.clamp_x1) / 2;
                cy[t] = min(max((Y + t - 2) / S, g.clamp_y0), g.clamp_y1) / 2;
            }
            mx0 = cx[0] * 16; mx1 = cx[4] * 16; my0 = cy[0]; my1 = cy[4];
#pragma unroll
            for (int t = 0; t < 5; t++) { cxb[t] = cx[t] != cx[0]; cyb[t] = cy[t] != cy[0]; }
        }
        // CFA constants per absolute parity class q = 2 * (row & 1) + (col & 1)
        int colq[4]; float blq[4], ivq[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = A.cfa.c[q];
            colq[q] = c;
            blq[q] = c == 0 ? A.black[0] : (c == 1 ? A.black[1] : A.black[2]);
            ivq[q] = 1.0f / (c == 0 ? A.white[0] : (c == 1 ? A.white[1] : A.white[2]));
        }
        float a4[4] = {0.f, 0.f, 0.f, 0.f}, w4[4] = {0.f, 0.f, 0.f, 0.f};      // value / weight sums per parity class
        for (int f = 0; f < A.n_frames; f++) {
            const float2* flow = (const float2*)((const char*)A.flow + A.flow_fs * f);
            const char* raw = (const char*)A.raw + A.raw_fs * f;
            const char* mask = (const char*)A.mask + A.mask_fs * f;
            const float2 s00 = row_ptr(flow, A.flow_pitch, ty.i0)[tx.i0], s10 = row_ptr(flow, A.flow_pitch, ty.i0)[tx.i1];
            const float2 s01 = row_ptr(flow, A.flow_pitch, ty.i1)[tx.i0], s11 = row_ptr(flow, A.flow_pitch, ty.i1)[tx.i1];
            const int sx = (int)roundf(__fmul_rn(tex_mix(s00.x, s10.x, s01.x, s11.x, tx.a, ty.a), (float)S));
            const int sy = (int)roundf(__fmul_rn(tex_mix(s00.y, s10.y, s01.y, s11.y, tx.a, ty.a), (float)S));
            // x axis: the 5 taps read at most 3 consecutive raw columns xb .. xb + 2 (xb kept inside the clamp range so that all three
            // loads are in bounds; a column beyond the last tap's is loaded and never selected)
            int rx0 = 0; bool j0[5], j1[5], oddx[5];
#pragma unroll
            for (int t = 0; t < 5; t++) {
                const int c = min(max((X + t - 2 + sx) / S, g.clamp_x0), g.clamp_x1);
                if (t == 0) rx0 = min(c, g.clamp_x1 - 2);
                j0[t] = c == rx0; j1[t] = c == rx0 + 1; oddx[t] = c & 1;
            }
            const bool pb = rx0 & 1;            // parity of column xb (columns xb + 1 / xb + 2: !pb / pb)
#pragma unroll
            for (int py = 0; py < 5; py++) {
                const int ry = min(max((Y + py - 2 + sy) / S, g.clamp_y0), g.clamp_y1);
                const bool oddy = ry & 1;
                const uint16_t* rrow = (const uint16_t*)(raw + A.raw_pitch * ry) + rx0;
                const char* mrow = mask + A.mask_pitch * (cyb[py] ? my1 : my0);
                const float r0 = (float)__ldg(rrow), r1 = (float)__ldg(rrow + 1), r2 = (float)__ldg(rrow + 2);
                const float4 M0 = __ldg((const float4*)(mrow + mx0)), M1 = __ldg((const float4*)(mrow + mx1));
                // constants of this row's even / odd columns, then of columns xb (A) and xb + 1 (B)
                const int colE = oddy ? colq[2] : colq[0], colO = oddy ? colq[3] : colq[1];
                const float blE = oddy ? blq[2] : blq[0], blO = oddy ? blq[3] : blq[1];
                const float ivE = oddy ? ivq[2] : ivq[0], ivO = oddy ? ivq[3] : ivq[1];
                const float blA = pb ? blO : blE, blB = pb ? blE : blO, ivA = pb ? ivO : ivE, ivB = pb ? ivE : ivO;
                const float n0 = (r0 - blA) * ivA, n1 = (r1 - blB) * ivB, n2 = (r2 - blA) * ivA;
                // certainty of (cell, column parity): the channel is the colour of the tap's raw sample
                const bool e0 = colE == 0, e1 = colE == 1, o0 = colO == 0, o1 = colO == 1;
                float cE0 = e0 ? M0.x : (e1 ? M0.y : M0.z), cO0 = o0 ? M0.x : (o1 ? M0.y : M0.z);
                float cE1 = e0 ? M1.x : (e1 ? M1.y : M1.z), cO1 = o0 ? M1.x : (o1 ? M1.y : M1.z);
                if (!isfinite(cE0)) cE0 = 0.0f;
                if (!isfinite(cO0)) cO0 = 0.0f;
                if (!isfinite(cE1)) cE1 = 0.0f;
                if (!isfinite(cO1)) cO1 = 0.0f;
                float aE = 0.f, aO = 0.f, wE = 0.f, wO = 0.f;
#pragma unroll
                for (int px = 0; px < 5; px++) {
                    const bool o = oddx[px];
                    const float rn = j0[px] ? n0 : (j1[px] ? n1 : n2);
                    const float cert = cxb[px] ? (o ? cO1 : cE1) : (o ? cO0 : cE0);
                    const float wv = w[(py * 5 + px) < 13 ? (py * 5 + px) : 24 - (py * 5 + px)] * cert, vv = rn * wv;
                    padd(aO, vv, o); padd(aE, vv, !o);
                    padd(wO, wv, o); padd(wE, wv, !o);
                }
                padd(a4[2], aE, oddy); padd(a4[0], aE, !oddy); padd(a4[3], aO, oddy); padd(a4[1], aO, !oddy);
                padd(w4[2], wE, oddy); padd(w4[0], wE, !oddy); padd(w4[3], wO, oddy); padd(w4[1], wO, !oddy);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (colq[q] == c) { acc[c] += a4[q]; wacc[c] += w4[q]; }
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
    if (n_frames < 0 || geom->scale < 1 || MFSR_SCALE_NUM(geom->scale) < 1 || geom->out_w <= 0 || geom->out_h <= 0) return MFSR_E_INVALID;
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
    // scales 3 and 4: the lean tap loop (same taps; per-class accumulation); MFSR_MERGE_GENERIC=1 keeps the plain reference loop
    if ((geom->scale == 3 || geom->scale == 4) && !force_generic && geom->clamp_x1 - geom->clamp_x0 >= 2) {
        if (geom->scale == 3) merge_lean_kernel<3><<<grid, block, 0, (cudaStream_t)stream>>>(A);
        else merge_lean_kernel<4><<<grid, block, 0, (cudaStream_t)stream>>>(A);
        MFSR_LAUNCH_CHECK();
        return MFSR_OK;
    }
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
