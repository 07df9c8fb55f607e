// robustness.cu — per-pixel, per-channel robustness (certainty) mask on half-resolution RGB.
//
// Replaces ComputeRobustnessMask (RobustnessModell.cu:29-157) and the absent host's min filter.
// The reference issues 26 flow-texture fetches per pixel of which 24 are dead (its 5x5 loop
// overwrites max/min with fmaxf(s, shiftf) each iteration, :67-70, so only the centre and the
// last sample (+2,+2) survive) and parks 9 float3 per thread in dynamic shared memory; here
// exactly the two live fetches are made and the 3x3 reference patch stays in registers.
#include "common.cuh"

namespace mfsr {

struct Flow2 { float x, y; };

__device__ __forceinline__ Flow2 flow_fetch(const float2* __restrict__ flow, int64_t pitch, int fw, int fh, float px_half, float py_half, int w, int h)
{
    const TexAxis ax = tex_axis(tex_coord(px_half, w, fw), fw), ay = tex_axis(tex_coord(py_half, h, fh), fh);
    const float2 t00 = row_ptr(flow, pitch, ay.i0)[ax.i0], t10 = row_ptr(flow, pitch, ay.i0)[ax.i1];
    const float2 t01 = row_ptr(flow, pitch, ay.i1)[ax.i0], t11 = row_ptr(flow, pitch, ay.i1)[ax.i1];
    Flow2 r;
    r.x = tex_mix(t00.x, t10.x, t01.x, t11.x, ax.a, ay.a);
    r.y = tex_mix(t00.y, t10.y, t01.y, t11.y, ax.a, ay.a);
    return r;
}

__global__ void __launch_bounds__(256)
robustness_kernel(const float* __restrict__ ref3, const float* __restrict__ mov3, int64_t rgb_pitch,
                  const float2* __restrict__ flow, int64_t flow_pitch, int fw, int fh,
                  float4* __restrict__ mask, int64_t mask_pitch, int w, int h, float alpha, float beta, float thresholdM)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= w || py >= h) return;
    if (px >= w - 1 || py >= h - 1 || px < 1 || py < 1) {      // unwritten in the reference (:48): defined as 0
        row_ptr(mask, mask_pitch, py)[px] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const Flow2 sf = flow_fetch(flow, flow_pitch, fw, fh, (float)px + 0.5f, (float)py + 0.5f, w, h);
    const Flow2 sl = flow_fetch(flow, flow_pitch, fw, fh, (float)px + 2 + 0.5f, (float)py + 2 + 0.5f, w, h);
    float maxx = fmaxf(sl.x, sf.x), maxy = fmaxf(sl.y, sf.y), minx = fminf(sl.x, sf.x), miny = fminf(sl.y, sf.y);
    const int shx = (int)roundf(__fmul_rn(sf.x, 0.5f)), shy = (int)roundf(__fmul_rn(sf.y, 0.5f));
    float pix[9][3], meanRef[3] = {0.f, 0.f, 0.f}, meanMov[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int y = -1; y <= 1; y++)
#pragma unroll
        for (int x = -1; x <= 1; x++) {
            const float* p = row_ptr(ref3, rgb_pitch, py + y) + 3 * (px + x);
            const int i = (y + 1) * 3 + (x + 1);
            pix[i][0] = p[0]; pix[i][1] = p[1]; pix[i][2] = p[2];
            meanRef[0] += p[0]; meanRef[1] += p[1]; meanRef[2] += p[2];
            const int ppy = min(max(py + shy + y, 0), h - 1), ppx = min(max(px + shx + x, 0), w - 1);
            const float* q = row_ptr(mov3, rgb_pitch, ppy) + 3 * ppx;
            meanMov[0] += q[0]; meanMov[1] += q[1]; meanMov[2] += q[2];
        }
#pragma unroll
    for (int c = 0; c < 3; c++) { meanRef[c] /= 9.0f; meanMov[c] /= 9.0f; }
    float meandist = fabsf(meanRef[0] - meanMov[0]) + fabsf(meanRef[1] - meanMov[1]) + fabsf(meanRef[2] - meanMov[2]);
    meandist /= 3.0f;
    maxx *= 0.5f * meandist; maxy *= 0.5f * meandist; minx *= 0.5f * meandist; miny *= 0.5f * meandist;
    const float Mv = sqrtf((maxx - minx) * (maxx - minx) + (maxy - miny) * (maxy - miny));
    float s = 1.5f;
    if (Mv > thresholdM) s = 0.f;
    const float tt = 0.12f;
    float mk[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float sd = 0.f;
#pragma unroll
        for (int i = 0; i < 9; i++) sd += (pix[i][c] - meanRef[c]) * (pix[i][c] - meanRef[c]);
        sd = sqrtf(sd / 9.0f);
        float sigmaMD = sqrtf(alpha * meanRef[c] + beta);
        if (c == 1) sigmaMD = sigmaMD / sqrtf(2.0f);          // two greens averaged (:131)
        float dist = fabsf(meanRef[c] - meanMov[c]);
        const float sigma = fmaxf(sigmaMD, sd);
        dist = dist * (sd * sd / (sd * sd + sigmaMD * sigmaMD));
        mk[c] = fmaxf(fminf(s * expf(-dist * dist / (sigma * sigma)) - tt, 1.0f), 0.0f);
    }
    row_ptr(mask, mask_pitch, py)[px] = make_float4(mk[0], mk[1], mk[2], Mv);
}

// (2r+1)^2 min filter on .xyz, clamp border; .w copied.  Separable: rows then columns.
__global__ void __launch_bounds__(256)
erode_rows_kernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t pitch, int w, int h, int r)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float4* row = row_ptr(in, pitch, y);
    float4 m = row[x];
    for (int d = -r; d <= r; d++) {
        const float4 v = row[clampi(x + d, 0, w - 1)];
        m.x = fminf(m.x, v.x); m.y = fminf(m.y, v.y); m.z = fminf(m.z, v.z);
    }
    row_ptr(out, pitch, y)[x] = m;
}
__global__ void __launch_bounds__(256)
erode_cols_kernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t pitch, int w, int h, int r)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float4 m = row_ptr(in, pitch, y)[x];
    for (int d = -r; d <= r; d++) {
        const float4 v = row_ptr(in, pitch, clampi(y + d, 0, h - 1))[x];
        m.x = fminf(m.x, v.x); m.y = fminf(m.y, v.y); m.z = fminf(m.z, v.z);
    }
    row_ptr(out, pitch, y)[x] = m;
}

}  // namespace mfsr

using namespace mfsr;

extern "C" int mfsr_stage_robustness(const float* rgb_ref, const float* rgb_mov, int64_t rgb_pitch, const float* flow, int64_t flow_pitch,
                                     float* mask, int64_t mask_pitch, float* scratch, int w, int h, float alpha, float beta,
                                     float thresholdM, int erode_radius, void* stream)
{
    if (!rgb_ref || !rgb_mov || !flow || !mask || w < 3 || h < 3 || erode_radius < 0 || erode_radius > 8) return MFSR_E_INVALID;
    if (erode_radius > 0 && !scratch) return MFSR_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 b(32, 8), g(cdiv(w, 32), cdiv(h, 8));
    robustness_kernel<<<g, b, 0, st>>>(rgb_ref, rgb_mov, rgb_pitch, (const float2*)flow, flow_pitch, 2 * w, 2 * h,
                                       (float4*)mask, mask_pitch, w, h, alpha, beta, thresholdM);
    MFSR_LAUNCH_CHECK();
    if (erode_radius > 0) {
        erode_rows_kernel<<<g, b, 0, st>>>((const float4*)mask, (float4*)scratch, mask_pitch, w, h, erode_radius);
        MFSR_LAUNCH_CHECK();
        erode_cols_kernel<<<g, b, 0, st>>>((const float4*)scratch, (float4*)mask, mask_pitch, w, h, erode_radius);
        MFSR_LAUNCH_CHECK();
    }
    return MFSR_OK;
}
