// robustness.cu — per-pixel, per-channel robustness (certainty) mask on half-resolution RGB.
//
// Replaces ComputeRobustnessMask (RobustnessModell.cu:29-157) and the absent host's min filter.
// The reference issues 26 flow-texture fetches per pixel of which 24 are dead (its 5x5 loop
// overwrites max/min with fmaxf(s, shiftf) each iteration, :67-70, so only the centre and the
// last sample (+2,+2) survive) and parks 9 float3 per thread in dynamic shared memory; here
// exactly the two live fetches are made and the 3x3 reference patch stays in registers.
#include "common.cuh"
#include "internal.h"

namespace mfsr {

struct Flow2 { float x, y; };

// MUFU-based square root / reciprocal / exp (<= 2 ulp): the certainty is a tolerance-checked quantity that enters the merge as a
// smooth weight; the IEEE division / sqrtf / expf sequences were ~250 of the kernel's ~680 instructions per pixel (profiles/r1t).
__device__ __forceinline__ float rb_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rb_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Flow "texture" fetch at half-resolution pixel centre (px, py): normalised coordinate (px + .5) / w on a texture of
// width 2w lands at texel coordinate 2 px + 1 (+- one ulp of the division, far from any rounding boundary), i.e. exactly
// between texels 2px and 2px+1 with the 1.8 fixed-point fraction 128/256: the bilinear fetch is tex_mix(..., .5, .5) of the
// four (clamped) texels.  Bit-identical to the generic tex_coord/tex_axis path it replaces, without its two IEEE divisions.
__device__ __forceinline__ Flow2 flow_fetch_half(const float2* __restrict__ flow, int64_t pitch, int fw, int fh, int px, int py)
{
    const int x0 = min(2 * px, fw - 1), x1 = min(2 * px + 1, fw - 1), y0 = min(2 * py, fh - 1), y1 = min(2 * py + 1, fh - 1);
    const float2* r0 = row_ptr(flow, pitch, y0);
    const float2* r1 = row_ptr(flow, pitch, y1);
    const float2 t00 = __ldg(r0 + x0), t10 = __ldg(r0 + x1), t01 = __ldg(r1 + x0), t11 = __ldg(r1 + x1);
    Flow2 r;
    r.x = tex_mix(t00.x, t10.x, t01.x, t11.x, 0.5f, 0.5f);
    r.y = tex_mix(t00.y, t10.y, t01.y, t11.y, 0.5f, 0.5f);
    return r;
}

// The reference-frame part of the certainty: mean and standard deviation of the 3 x 3 reference patch per channel.  It does not depend on
// the moved frame: mfsr_run computes it once per burst (ref_stats_kernel) and every frame's certainty reads 6 floats instead of loading the
// 27 reference samples and redoing ~170 of the ~430 instructions per pixel.  Same expressions in the same order either way: same bits.
struct RefStats { float mean[3], sd[3]; };
__device__ __forceinline__ RefStats ref_stats(const float* __restrict__ ref3, int64_t rgb_pitch, int px, int py)
{
    float pix[9][3];
    RefStats r;
    r.mean[0] = r.mean[1] = r.mean[2] = 0.f;
    // reference 3x3 patch: rows py-1..py+1, 9 contiguous floats each
#pragma unroll
    for (int y = -1; y <= 1; y++) {
        const float* p = row_ptr(ref3, rgb_pitch, py + y) + 3 * (px - 1);
#pragma unroll
        for (int x = 0; x < 3; x++) {
            const int i = (y + 1) * 3 + x;
            pix[i][0] = __ldg(p + 3 * x); pix[i][1] = __ldg(p + 3 * x + 1); pix[i][2] = __ldg(p + 3 * x + 2);
        }
    }
#pragma unroll
    for (int i = 0; i < 9; i++) { r.mean[0] += pix[i][0]; r.mean[1] += pix[i][1]; r.mean[2] += pix[i][2]; }
#pragma unroll
    for (int c = 0; c < 3; c++) r.mean[c] *= (1.0f / 9.0f);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float sd = 0.f;
#pragma unroll
        for (int i = 0; i < 9; i++) sd += (pix[i][c] - r.mean[c]) * (pix[i][c] - r.mean[c]);
        r.sd[c] = rb_sqrt(sd * (1.0f / 9.0f));
    }
    return r;
}

// Certainty of half-resolution pixel (px, py); 0 on the 1-pixel border the reference leaves unwritten (:48).
// stats: NULL, or the reference statistics image (6 floats per pixel, rows of w pixels) written by ref_stats_kernel.
__device__ __forceinline__ float4 robust_px(const float* __restrict__ ref3, const float* __restrict__ mov3, int64_t rgb_pitch,
                                            const float2* __restrict__ flow, int64_t flow_pitch, int fw, int fh,
                                            int w, int h, int px, int py, float alpha, float beta, float thresholdM,
                                            const float* __restrict__ stats)
{
    if (px >= w - 1 || py >= h - 1 || px < 1 || py < 1) return make_float4(0.f, 0.f, 0.f, 0.f);
    const Flow2 sf = flow_fetch_half(flow, flow_pitch, fw, fh, px, py);
    const Flow2 sl = flow_fetch_half(flow, flow_pitch, fw, fh, px + 2, py + 2);
    float maxx = fmaxf(sl.x, sf.x), maxy = fmaxf(sl.y, sf.y), minx = fminf(sl.x, sf.x), miny = fminf(sl.y, sf.y);
    const int shx = (int)roundf(__fmul_rn(sf.x, 0.5f)), shy = (int)roundf(__fmul_rn(sf.y, 0.5f));
    RefStats R;
    if (stats) {
        const float2* sp = (const float2*)(stats + ((size_t)py * w + px) * 6);
        const float2 a = __ldg(sp), b = __ldg(sp + 1), c = __ldg(sp + 2);
        R.mean[0] = a.x; R.mean[1] = a.y; R.mean[2] = b.x; R.sd[0] = b.y; R.sd[1] = c.x; R.sd[2] = c.y;
    } else {
        R = ref_stats(ref3, rgb_pitch, px, py);
    }
    float meanMov[3] = {0.f, 0.f, 0.f};
    // moved 3x3 patch at the rounded half-resolution shift, clamp addressing (:93-101); same summation order as the reference
    const int mx = px + shx, my = py + shy;
    const bool xin = mx >= 1 && mx <= w - 2;
#pragma unroll
    for (int y = -1; y <= 1; y++) {
        const float* q = row_ptr(mov3, rgb_pitch, min(max(my + y, 0), h - 1));
#pragma unroll
        for (int x = -1; x <= 1; x++) {
            const float* qq = q + 3 * (xin ? mx + x : min(max(mx + x, 0), w - 1));
            meanMov[0] += __ldg(qq); meanMov[1] += __ldg(qq + 1); meanMov[2] += __ldg(qq + 2);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) meanMov[c] *= (1.0f / 9.0f);
    const float* meanRef = R.mean;
    float meandist = fabsf(meanRef[0] - meanMov[0]) + fabsf(meanRef[1] - meanMov[1]) + fabsf(meanRef[2] - meanMov[2]);
    meandist *= (1.0f / 3.0f);
    maxx *= 0.5f * meandist; maxy *= 0.5f * meandist; minx *= 0.5f * meandist; miny *= 0.5f * meandist;
    const float Mv = rb_sqrt((maxx - minx) * (maxx - minx) + (maxy - miny) * (maxy - miny));
    float s = 1.5f;
    if (Mv > thresholdM) s = 0.f;
    const float tt = 0.12f;
    float mk[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float sd = R.sd[c];
        float sigmaMD = rb_sqrt(alpha * meanRef[c] + beta);
        if (c == 1) sigmaMD = sigmaMD * 0.70710678118654752440f;          // / sqrt(2): two greens averaged (:131)
        float dist = fabsf(meanRef[c] - meanMov[c]);
        const float sigma = fmaxf(sigmaMD, sd);
        dist = dist * (sd * sd * rb_rcp(sd * sd + sigmaMD * sigmaMD));
        mk[c] = fmaxf(fminf(s * __expf(-dist * dist * rb_rcp(sigma * sigma)) - tt, 1.0f), 0.0f);
    }
    return make_float4(mk[0], mk[1], mk[2], Mv);
}

__global__ void __launch_bounds__(256)
robustness_kernel(const float* __restrict__ ref3, const float* __restrict__ mov3, int64_t rgb_pitch,
                  const float2* __restrict__ flow, int64_t flow_pitch, int fw, int fh,
                  float4* __restrict__ mask, int64_t mask_pitch, int w, int h, float alpha, float beta, float thresholdM, FrameStrides fs,
                  const float* __restrict__ stats)
{
    mov3 = frame_ptr(mov3, fs.s[0], blockIdx.z);      // blockIdx.z = frame: moved image, its flow and its mask
    flow = frame_ptr(flow, fs.s[1], blockIdx.z);
    mask = frame_ptr(mask, fs.s[2], blockIdx.z);
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= w || py >= h) return;
    row_ptr(mask, mask_pitch, py)[px] = robust_px(ref3, mov3, rgb_pitch, flow, flow_pitch, fw, fh, w, h, px, py, alpha, beta, thresholdM, stats);
}

// Reference statistics image: 6 floats (mean.rgb, sd.rgb) per half-resolution pixel, rows of w pixels; border pixels are never read.
__global__ void __launch_bounds__(256)
ref_stats_kernel(const float* __restrict__ ref3, int64_t rgb_pitch, float* __restrict__ stats, int w, int h)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= w - 1 || py >= h - 1 || px < 1 || py < 1) return;
    const RefStats R = ref_stats(ref3, rgb_pitch, px, py);
    float2* sp = (float2*)(stats + ((size_t)py * w + px) * 6);
    sp[0] = make_float2(R.mean[0], R.mean[1]); sp[1] = make_float2(R.mean[2], R.sd[0]); sp[2] = make_float2(R.sd[1], R.sd[2]);
}

// Robustness + (2r+1)^2 min filter in ONE launch (round 2): the certainties of a 48 x 32 region (output tile (48 - 2r) x (32 - 2r) plus its
// halo, clamp border) are computed straight into shared memory, row minima, then column minima — the raw certainties never go to HBM
// (the two-kernel form wrote and re-read a 48 MB float4 mask per 12 MP frame) for 1.25 x the certainty arithmetic at r = 2.
// Same per-pixel function and the same min order as robustness_kernel + erode_kernel: bit-identical masks.
constexpr int RE_W = 48, RE_H = 32;
__global__ void __launch_bounds__(256)
robust_erode_kernel(const float* __restrict__ ref3, const float* __restrict__ mov3, int64_t rgb_pitch,
                    const float2* __restrict__ flow, int64_t flow_pitch, int fw, int fh,
                    float4* __restrict__ mask, int64_t mask_pitch, int w, int h, float alpha, float beta, float thresholdM, int r, FrameStrides fs,
                    const float* __restrict__ stats)
{
    extern __shared__ float4 s_re[];                 // [RE_H][RE_W] certainties, then [RE_H][RE_W - 2r] row minima
    mov3 = frame_ptr(mov3, fs.s[0], blockIdx.z);
    flow = frame_ptr(flow, fs.s[1], blockIdx.z);
    mask = frame_ptr(mask, fs.s[2], blockIdx.z);
    const int iw = RE_W - 2 * r, ih = RE_H - 2 * r;
    float4* s_in = s_re;
    float4* s_row = s_re + RE_W * RE_H;
    const int x0 = blockIdx.x * iw, y0 = blockIdx.y * ih, tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int i = tid; i < RE_W * RE_H; i += 256) {
        const int ly = i / RE_W, lx = i - ly * RE_W;
        s_in[i] = robust_px(ref3, mov3, rgb_pitch, flow, flow_pitch, fw, fh, w, h, clampi(x0 + lx - r, 0, w - 1), clampi(y0 + ly - r, 0, h - 1),
                            alpha, beta, thresholdM, stats);
    }
    __syncthreads();
    for (int i = tid; i < iw * RE_H; i += 256) {
        const int ly = i / iw, lx = i - ly * iw;
        float4 m = s_in[ly * RE_W + lx + r];
        for (int d = 0; d <= 2 * r; d++) {
            const float4 v = s_in[ly * RE_W + lx + d];
            m.x = fminf(m.x, v.x); m.y = fminf(m.y, v.y); m.z = fminf(m.z, v.z);
        }
        s_row[i] = m;
    }
    __syncthreads();
    for (int i = tid; i < iw * ih; i += 256) {
        const int ly = i / iw, lx = i - ly * iw;
        const int x = x0 + lx, y = y0 + ly;
        if (x >= w || y >= h) continue;
        float4 m = s_row[(ly + r) * iw + lx];
        for (int d = 0; d <= 2 * r; d++) {
            const float4 v = s_row[(ly + d) * iw + lx];
            m.x = fminf(m.x, v.x); m.y = fminf(m.y, v.y); m.z = fminf(m.z, v.z);
        }
        row_ptr(mask, mask_pitch, y)[x] = m;
    }
}

// (2r+1)^2 min filter on .xyz, clamp border; .w copied.  One launch: a 32 x 8 output tile with its halo is staged in shared
// memory, row minima in place, then column minima (the two-launch form moved every mask twice through HBM and was L1-bound:
// 2r+1 overlapping float4 loads per thread and pass).
constexpr int ER_MAX = 8, ER_TW = 32, ER_TH = 8;
__global__ void __launch_bounds__(256)
erode_kernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t pitch, int w, int h, int r, int64_t in_fs, int64_t out_fs)
{
    in = frame_ptr(in, in_fs, blockIdx.z);
    out = frame_ptr(out, out_fs, blockIdx.z);
    extern __shared__ float4 s_e[];                  // [ER_TH + 2r][ER_TW + 2r] input, then [ER_TH + 2r][ER_TW] row minima
    const int sw = ER_TW + 2 * r, sh = ER_TH + 2 * r;
    float4* s_in = s_e;
    float4* s_row = s_e + sw * sh;
    const int x0 = blockIdx.x * ER_TW, y0 = blockIdx.y * ER_TH, tid = threadIdx.y * ER_TW + threadIdx.x;
    for (int i = tid; i < sw * sh; i += ER_TW * ER_TH) {
        const int ly = i / sw, lx = i - ly * sw;
        s_in[i] = __ldg(row_ptr(in, pitch, clampi(y0 + ly - r, 0, h - 1)) + clampi(x0 + lx - r, 0, w - 1));
    }
    __syncthreads();
    for (int i = tid; i < ER_TW * sh; i += ER_TW * ER_TH) {
        const int ly = i / ER_TW, lx = i - ly * ER_TW;
        float4 m = s_in[ly * sw + lx + r];
        for (int d = 0; d <= 2 * r; d++) {
            const float4 v = s_in[ly * sw + lx + d];
            m.x = fminf(m.x, v.x); m.y = fminf(m.y, v.y); m.z = fminf(m.z, v.z);
        }
        s_row[i] = m;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    float4 m = s_row[(threadIdx.y + r) * ER_TW + threadIdx.x];
    for (int d = 0; d <= 2 * r; d++) {
        const float4 v = s_row[(threadIdx.y + d) * ER_TW + threadIdx.x];
        m.x = fminf(m.x, v.x); m.y = fminf(m.y, v.y); m.z = fminf(m.z, v.z);
    }
    row_ptr(out, pitch, y)[x] = m;
}

}  // namespace mfsr

using namespace mfsr;

// `frames` masks per launch (frame f: rgb_mov + f * rgb_fs, flow + f * flow_fs, mask + f * mask_fs, scratch + f * scratch_fs).
int mfsr::launch_robustness(const float* rgb_ref, const float* rgb_mov, int64_t rgb_pitch, int64_t rgb_fs, const float* flow, int64_t flow_pitch,
                            int64_t flow_fs, float* mask, int64_t mask_pitch, int64_t mask_fs, float* scratch, int64_t scratch_fs, int frames,
                            int w, int h, float alpha, float beta, float thresholdM, int erode_radius, cudaStream_t st, float* stats)
{
    if (!rgb_ref || !rgb_mov || !flow || !mask || w < 3 || h < 3 || frames < 1 || erode_radius < 0 || erode_radius > 8) return MFSR_E_INVALID;
    dim3 b(32, 8), g(cdiv(w, 32), cdiv(h, 8), frames);
    if (stats) {
        // stats: w * h * 6 floats of workspace: the reference part of the certainty, once for all frames of the launch
        ref_stats_kernel<<<dim3(cdiv(w, 32), cdiv(h, 8)), b, 0, st>>>(rgb_ref, rgb_pitch, stats, w, h);
        MFSR_LAUNCH_CHECK();
    }
    if (erode_radius > 0 && !scratch) {
        // fused form (what mfsr_run uses): no scratch image
        const int r = erode_radius;
        FrameStrides ff;
        ff.s[0] = rgb_fs; ff.s[1] = flow_fs; ff.s[2] = mask_fs;
        const size_t smem = (size_t)(RE_W * RE_H + (RE_W - 2 * r) * RE_H) * sizeof(float4);
        dim3 gf(cdiv(w, RE_W - 2 * r), cdiv(h, RE_H - 2 * r), frames);
        robust_erode_kernel<<<gf, b, smem, st>>>(rgb_ref, rgb_mov, rgb_pitch, (const float2*)flow, flow_pitch, 2 * w, 2 * h,
                                                 (float4*)mask, mask_pitch, w, h, alpha, beta, thresholdM, r, ff, stats);
        MFSR_LAUNCH_CHECK();
        return MFSR_OK;
    }
    // with a min filter the raw certainties go to `scratch` and the filter writes `mask`: every mask crosses HBM once per kernel
    float4* raw_out = erode_radius > 0 ? (float4*)scratch : (float4*)mask;
    FrameStrides fs;
    fs.s[0] = rgb_fs; fs.s[1] = flow_fs; fs.s[2] = erode_radius > 0 ? scratch_fs : mask_fs;
    robustness_kernel<<<g, b, 0, st>>>(rgb_ref, rgb_mov, rgb_pitch, (const float2*)flow, flow_pitch, 2 * w, 2 * h,
                                       raw_out, mask_pitch, w, h, alpha, beta, thresholdM, fs, stats);
    MFSR_LAUNCH_CHECK();
    if (erode_radius > 0) {
        const int r = erode_radius;
        const size_t smem = (size_t)((ER_TW + 2 * r) * (ER_TH + 2 * r) + ER_TW * (ER_TH + 2 * r)) * sizeof(float4);
        erode_kernel<<<g, b, smem, st>>>((const float4*)scratch, (float4*)mask, mask_pitch, w, h, r, scratch_fs, mask_fs);
        MFSR_LAUNCH_CHECK();
    }
    return MFSR_OK;
}

extern "C" int mfsr_stage_robustness(const float* rgb_ref, const float* rgb_mov, int64_t rgb_pitch, const float* flow, int64_t flow_pitch,
                                     float* mask, int64_t mask_pitch, float* scratch, int w, int h, float alpha, float beta,
                                     float thresholdM, int erode_radius, void* stream)
{
    return launch_robustness(rgb_ref, rgb_mov, rgb_pitch, 0, flow, flow_pitch, 0, mask, mask_pitch, 0, scratch, 0, 1, w, h, alpha, beta, thresholdM,
                             erode_radius, (cudaStream_t)stream, nullptr);
}
