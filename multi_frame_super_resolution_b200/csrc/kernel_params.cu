// kernel_params.cu — anisotropic merge-kernel estimation on the reference frame.
//
// One launch replaces ComputeDerivatives2Kernel (opticalFlow.cu:151), ComputeStructureTensor
// (kernel.cu:691), the absent host's box smoothing of the tensor (NPP FilterBox) and
// ComputeKernelParam (kernel.cu:718): gray tile + halo staged in shared memory, tensor kept in
// shared memory, only the float4 inverse covariance is written.  HBM: 4 B in, 16 B out per pixel
// (the reference chain moves 4 + 8 + 8 + 12 + 12 + 12 + 12 + 12 B).
#include "common.cuh"

namespace mfsr {

constexpr int KTW = 32, KTH = 16, KR_MAX = 3;
constexpr int KGW = KTW + 2 * (KR_MAX + 2), KGH = KTH + 2 * (KR_MAX + 2);
constexpr int KSW = KTW + 2 * KR_MAX, KSH = KTH + 2 * KR_MAX;

// MUFU square root / reciprocal (<= 2 ulp): the kernel parameters are a tolerance-checked quantity (99.9 % within 1e-3 relative)
__device__ __forceinline__ float kp_sqrt(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float kp_rcp(float x) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// R: box radius of the tensor smoothing, a template parameter so that the window loops unroll.  A thread produces two vertically
// adjacent pixels: their (2R+1)^2 windows share 2R of 2R+1 rows, so the row sums are formed once (2R+2 rows) and combined.
template <int R>
__global__ void __launch_bounds__(256)
kernel_params_kernel(const float* __restrict__ gray, int64_t gray_pitch, float4* __restrict__ out, int64_t out_pitch,
                     int w, int h, float Dth, float Dtr, float kDetail, float kDenoise, float kStretch, float kShrink)
{
    constexpr int r = R;
    __shared__ float s_g[KGH][KGW];
    __shared__ float s_t[3][KSH][KSW];
    const int x0 = blockIdx.x * KTW, y0 = blockIdx.y * KTH;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const int GW = KTW + 2 * (r + 2), GH = KTH + 2 * (r + 2), SW = KTW + 2 * r, SH = KTH + 2 * r;
    const int ox = x0 - r - 2, oy = y0 - r - 2;
    for (int i = tid; i < GW * GH; i += nthr) {
        const int ly = i / GW, lx = i - ly * GW;
        s_g[ly][lx] = row_ptr(gray, gray_pitch, clampi(oy + ly, 0, h - 1))[clampi(ox + lx, 0, w - 1)];
    }
    __syncthreads();
    // derivative + structure tensor at clamped coordinates (box filter uses clamp border)
    for (int i = tid; i < SW * SH; i += nthr) {
        const int ry = i / SW, rx = i - ry * SW;
        const int gx = clampi(x0 - r + rx, 0, w - 1), gy = clampi(y0 - r + ry, 0, h - 1);
        const int cx = gx - ox, cy = gy - oy;
        const int xm2 = clampi(gx - 2, 0, w - 1) - ox, xm1 = clampi(gx - 1, 0, w - 1) - ox;
        const int xp1 = clampi(gx + 1, 0, w - 1) - ox, xp2 = clampi(gx + 2, 0, w - 1) - ox;
        const int ym2 = clampi(gy - 2, 0, h - 1) - oy, ym1 = clampi(gy - 1, 0, h - 1) - oy;
        const int yp1 = clampi(gy + 1, 0, h - 1) - oy, yp2 = clampi(gy + 2, 0, h - 1) - oy;
        float dx = s_g[cy][xp2]; dx -= s_g[cy][xp1] * 8.0f; dx += s_g[cy][xm1] * 8.0f; dx -= s_g[cy][xm2]; dx *= (1.0f / 12.0f);
        float dy = s_g[yp2][cx]; dy -= s_g[yp1][cx] * 8.0f; dy += s_g[ym1][cx] * 8.0f; dy -= s_g[ym2][cx]; dy *= (1.0f / 12.0f);
        s_t[0][ry][rx] = dx * dx; s_t[1][ry][rx] = dy * dy; s_t[2][ry][rx] = dx * dy;
    }
    __syncthreads();
    const float invn = 1.0f / (float)((2 * r + 1) * (2 * r + 1));
    const float rDtr = 1.0f / Dtr, kShr = kDetail / kShrink;
    const int lx = threadIdx.x, ly0 = 2 * threadIdx.y, gx = x0 + lx;         // KTH == 2 * blockDim.y
    if (gx >= w) return;
    float rs[2 * R + 2][3];
#pragma unroll
    for (int dy = 0; dy < 2 * R + 2; dy++) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int dx = 0; dx <= 2 * R; dx++) { t0 += s_t[0][ly0 + dy][lx + dx]; t1 += s_t[1][ly0 + dy][lx + dx]; t2 += s_t[2][ly0 + dy][lx + dx]; }
        rs[dy][0] = t0; rs[dy][1] = t1; rs[dy][2] = t2;
    }
#pragma unroll
    for (int p = 0; p < 2; p++) {
        const int gy = y0 + ly0 + p;
        if (gy >= h) break;
        float a11 = 0.f, a22 = 0.f, a12 = 0.f;
#pragma unroll
        for (int dy = 0; dy <= 2 * R; dy++) { a11 += rs[p + dy][0]; a22 += rs[p + dy][1]; a12 += rs[p + dy][2]; }
        a11 *= invn; a22 *= invn; a12 *= invn;
        // ComputeKernelParam (kernel.cu:736-789)
        const float help = kp_sqrt((a22 - a11) * (a22 - a11) + 4.0f * a12 * a12);
        float c = 2.0f * a12, s = a22 - a11 + help;
        const float norm = kp_sqrt(c * c + s * s);
        if (norm > 0) { const float in = kp_rcp(norm); c *= in; s *= in; } else { c = 1; s = 0; }
        const float lam1 = (a11 + a22 + help) * 0.5f, lam2 = (a11 + a22 - help) * 0.5f;
        const float A = 1 + kp_sqrt((lam1 - lam2) * (lam1 - lam2) * kp_rcp((lam1 + lam2) * (lam1 + lam2)));
        float D = 1 - kp_sqrt(lam1) * rDtr + Dth;
        D = fmaxf(fminf(1.0f, D), 0.0f);
        const float k1h = kDetail * kStretch * A, k2h = kShr * A;
        float k1 = ((1.0f - D) * k1h + D * kDetail * kDenoise);
        float k2 = ((1.0f - D) * k2h + D * kDetail * kDenoise);
        k1 *= k1; k2 *= k2;
        const float x2 = c, y2 = s, x1 = s, y1 = -c;
        const float b11 = k1 * x1 * x1 + x2 * x2 * k2;
        const float b12 = k1 * x1 * y1 + x2 * y2 * k2;
        const float b22 = k1 * y1 * y1 + y2 * y2 * k2;
        const float idet = kp_rcp(b11 * b22 - b12 * b12 + 0.0000000001f);
        row_ptr(out, out_pitch, gy)[gx] = make_float4(b22 * idet, b11 * idet, -b12 * idet, 0.0f);
    }
}

}  // namespace mfsr

using namespace mfsr;

extern "C" int mfsr_stage_kernel_params(const float* gray, int64_t gray_pitch, float* kernel4, int64_t kernel_pitch,
                                        int width, int height, int box_radius, float Dth, float Dtr, float kDetail,
                                        float kDenoise, float kStretch, float kShrink, void* stream)
{
    if (!gray || !kernel4 || width < 1 || height < 1 || box_radius < 0 || box_radius > KR_MAX) return MFSR_E_INVALID;
    dim3 b(KTW, 8), g(cdiv(width, KTW), cdiv(height, KTH));
    static_assert(KTH == 16, "two rows per thread of an 8-row block");
    cudaStream_t st = (cudaStream_t)stream;
    float4* o = (float4*)kernel4;
    switch (box_radius) {
        case 0: kernel_params_kernel<0><<<g, b, 0, st>>>(gray, gray_pitch, o, kernel_pitch, width, height, Dth, Dtr, kDetail, kDenoise, kStretch, kShrink); break;
        case 1: kernel_params_kernel<1><<<g, b, 0, st>>>(gray, gray_pitch, o, kernel_pitch, width, height, Dth, Dtr, kDetail, kDenoise, kStretch, kShrink); break;
        case 2: kernel_params_kernel<2><<<g, b, 0, st>>>(gray, gray_pitch, o, kernel_pitch, width, height, Dth, Dtr, kDetail, kDenoise, kStretch, kShrink); break;
        default: kernel_params_kernel<3><<<g, b, 0, st>>>(gray, gray_pitch, o, kernel_pitch, width, height, Dth, Dtr, kDetail, kDenoise, kStretch, kShrink); break;
    }
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}
