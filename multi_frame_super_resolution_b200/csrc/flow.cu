// flow.cu — dense flow from the tile-shift grid and Lucas-Kanade refinement.
//
// lk_iteration_kernel fuses WarpingKernel (opticalFlow.cu:28), ComputeDerivativesKernel (:97)
// and lucasKanadeOptim (:190) into one launch per sweep: the warped image and the three
// derivative planes (4 x 4 B/px of HBM write + (2h+1)^2 x 3 uncoalesced re-reads per pixel in
// the reference) never leave shared memory; the window sums are separable (row sums, then
// column sums) instead of two full (2h+1)^2 passes per pixel.
// HBM traffic per sweep: ref gray 4 B + moved gray (gather, ~4 B) + flow 8 B in, 8 B out.
#include "common.cuh"
#include "internal.h"
#include <type_traits>

namespace mfsr {

// CreateFlowFieldFromTiles (opticalFlow.cu:48-93).  Band form: see internal.h (gh == h, gy0 == 0, gty == tilesY, trow0 == 0
// for a whole frame).
constexpr int FFT_ROWS = 8;      // rows per thread: the column-only part (texture column, weights, rotation terms) is computed once
__global__ void __launch_bounds__(256)
flow_from_tiles_kernel(const float2* __restrict__ tiles, int64_t tile_pitch, int tilesX, int tilesY,
                       float2* __restrict__ flow, int64_t flow_pitch, int w, int h, float bsx, float bsy, float cr, float sr,
                       int gh, int gy0, int gty, int trow0, const float* __restrict__ frame_pose, int64_t tiles_fs, int64_t flow_fs)
{
    tiles = frame_ptr(tiles, tiles_fs, blockIdx.z);      // blockIdx.z = frame
    flow = frame_ptr(flow, flow_fs, blockIdx.z);
    if (frame_pose) { frame_pose += 4 * blockIdx.z; bsx = frame_pose[0]; bsy = frame_pose[1]; cr = frame_pose[2]; sr = frame_pose[3]; }      // prealign.cu
    // the row part of the bilinear tile fetch (texture row, fraction, clamped tile rows) is the same for every pixel of a row: the
    // block's 64 rows are worked out once, by 64 threads, instead of once per pixel (an IEEE division + tex_axis + clamps, ~30 of the
    // ~70 instructions per pixel); the per-pixel arithmetic is unchanged, the flow field bit-identical
    __shared__ int s_i0[8 * FFT_ROWS], s_i1[8 * FFT_ROWS];
    __shared__ float s_a[8 * FFT_ROWS];
    {
        const int t = threadIdx.y * blockDim.x + threadIdx.x;
        if (t < 8 * FFT_ROWS) {
            const int y = blockIdx.y * 8 * FFT_ROWS + t;
            TexAxis ay = tex_axis(tex_coord((float)(min(y, h - 1) + gy0) + 0.5f, gh, gty), gty);
            s_i0[t] = clampi(ay.i0 - trow0, 0, tilesY - 1); s_i1[t] = clampi(ay.i1 - trow0, 0, tilesY - 1); s_a[t] = ay.a;
        }
    }
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, yb = (blockIdx.y * blockDim.y + threadIdx.y) * FFT_ROWS;
    if (x >= w || yb >= h) return;
    const float bx = cr * -bsx - sr * -bsy, by = sr * -bsx + cr * -bsy;
    const float pcx = (float)(x - w / 2);
    const TexAxis ax = tex_axis(tex_coord((float)x + 0.5f, w, tilesX), tilesX);
#pragma unroll 2
    for (int r = 0; r < FFT_ROWS; r++) {
        const int y = yb + r;
        if (y >= h) break;
        float sx = bx, sy = by;
        const float pcy = (float)(y + gy0 - gh / 2);
        sx += cr * pcx - sr * pcy - pcx;
        sy += sr * pcx + cr * pcy - pcy;
        const int rl = threadIdx.y * FFT_ROWS + r;
        const int i0 = s_i0[rl], i1 = s_i1[rl];
        const float aya = s_a[rl];
        const float2 t00 = row_ptr(tiles, tile_pitch, i0)[ax.i0], t10 = row_ptr(tiles, tile_pitch, i0)[ax.i1];
        const float2 t01 = row_ptr(tiles, tile_pitch, i1)[ax.i0], t11 = row_ptr(tiles, tile_pitch, i1)[ax.i1];
        sx += tex_mix(t00.x, t10.x, t01.x, t11.x, ax.a, aya);
        sy += tex_mix(t00.y, t10.y, t01.y, t11.y, ax.a, aya);
        row_ptr(flow, flow_pitch, y)[x] = make_float2(sx, sy);
    }
}

constexpr int LTW = 32, LTH = 32, LHW_MAX = 4;          // output tile of one CTA (256 threads)

// MUFU-based square root / reciprocal (<= 2 ulp) for the per-pixel pseudo-inverse and the derivative stencil.  The LK
// update is a tolerance-checked quantity (tests: 99.9 % of the flow within 2e-3 px of the oracle); the IEEE sqrtf / division
// sequences with their slow-path branches were ~110 of ~640 instructions per pixel of an issue-bound kernel.
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// cos / sin of theta = 0.5 * atan2(y, x) without trigonometry (half-angle identities, cancellation-free branch):
// theta in [-pi/2, pi/2], cos(theta) >= 0, sign(sin(theta)) = sign(y).  Agrees with cosf/sinf(0.5f * atan2f(y, x)) to a few ulp.
__device__ __forceinline__ void half_angle(float y, float x, float& c, float& s)
{
    const float r2 = x * x + y * y;
    if (!(r2 > 0.0f)) { c = 1.0f; s = 0.0f; if (r2 != r2) { c = r2; s = r2; } return; }     // atan2(0, 0) = 0; NaN propagates
    const float ir = rsqrtf(r2);
    const float cx = x * ir;                                   // cos(2 theta)
    if (x >= 0.0f) { c = fast_sqrt(0.5f * (1.0f + cx)); s = (0.5f * y * ir) * fast_rcp(c); }
    else           { const float sa = fast_sqrt(0.5f * (1.0f - cx)); s = copysignf(sa, y); c = (0.5f * fabsf(y) * ir) * fast_rcp(sa); }
}

// closed-form 2x2 SVD pseudo-inverse of the SYMMETRIC window matrix [[a,b],[b,d]] (opticalFlow.cu:236-292 with c == b),
// including the fminf(sigma1, sigma1) quirk (:255).  Returns false when the reference returns early.
// For c == b both rotation angles of the reference coincide (theta == eps = 0.5 atan2(2b(a+d), a^2 - d^2)), and their
// cos/sin come from half_angle() instead of atan2f + cosf + sinf (the kernel was bound by those, ~140 of ~500
// instructions per pixel).  Singular values, reciprocals, sign fix-ups and the NaN behaviour are kept verbatim.
__device__ __forceinline__ bool lk_pinv(float a, float b, float c, float d, float minDet, float inv[4])
{
    float ct, st;
    half_angle(2.0f * a * c + 2.0f * b * d, a * a + b * b - c * c - d * d, ct, st);
    const float UT0 = ct, UT2 = -st, UT1 = st, UT3 = ct;
    const float S1 = a * a + b * b + c * c + d * d;
    const float S2 = fast_sqrt((a * a + b * b - c * c - d * d) * (a * a + b * b - c * c - d * d) + 4 * (a * c + b * d) * (a * c + b * d));
    float sigma1 = fast_sqrt((S1 + S2) / 2), sigma2 = fast_sqrt((S1 - S2) / 2);
    const float smin = fminf(sigma1, sigma1);
    if (smin < minDet) return false;
    sigma1 = sigma1 != 0 ? fast_rcp(sigma1) : 0;
    sigma2 = sigma2 != 0 ? fast_rcp(sigma2) : 0;
    const float ce = ct, se = st;
    float s11 = (a * ct + c * st) * ce + (b * ct + d * st) * se;
    float s22 = (a * st - c * ct) * se + (-b * st + d * ct) * ce;
    s11 = s11 > 0.0f ? 1.0f : s11 < 0 ? -1.0f : 0.0f;
    s22 = s22 > 0.0f ? 1.0f : s22 < 0 ? -1.0f : 0.0f;
    const float V0 = s11 * ce, V1 = -s22 * se, V2 = s11 * se, V3 = s22 * ce;
    const float m0 = sigma1 * UT0 + 0.0f * UT2, m1 = sigma1 * UT1 + 0.0f * UT3;
    const float m2 = 0.0f * UT0 + sigma2 * UT2, m3 = 0.0f * UT1 + sigma2 * UT3;
    inv[0] = V0 * m0 + V1 * m2; inv[1] = V0 * m1 + V1 * m3;
    inv[2] = V2 * m0 + V3 * m2; inv[3] = V2 * m1 + V3 * m3;
    return true;
}

// One LK sweep on a 32x32 tile.  HW (half window) is a template parameter so that every region size, division and
// loop below is a compile-time constant: the first version (runtime hw, per-element clamped stencils, one output per
// thread in the window sums) executed 1170 instructions per pixel at 82 % issue utilisation (profiles/r1j_lk_ncu.txt).
//   W region (source + warped): tile + HW + 2 on each side      R region (derivatives): tile + HW
// MODE 0: the bilinear fetch of the warp step is evaluated in ALU with the 1.8 fixed-point texture model of common.cuh (bit-identical
//         to the oracle's restatement); MODE 1: the same for a row band (texture rows of the FULL frame);
// MODE 2: the fetch goes through the texture unit, exactly as the reference's WarpingKernel does it (opticalFlow.cu:36-41: linear
//         filter, clamp addressing, normalised coordinates) — one TEX instead of four loads and ~30 ALU instructions per element of
//         the haloed region, and the warped image is the reference kernel's bit for bit.
template <int HW, int MODE>
__global__ void __launch_bounds__(256, 4)
lk_iteration_kernel(const float* __restrict__ ref, const float* __restrict__ mov, int64_t img_pitch,
                    const float2* __restrict__ flow_in, float2* __restrict__ flow_out, int64_t flow_pitch,
                    int w, int h, float minDet, int gh, int gy0, cudaTextureObject_t movtex, int64_t mov_fs, int64_t flow_fs, int ref_frame)
{
    constexpr bool BAND = MODE == 1;
    // blockIdx.z = frame.  The reference frame against itself has Iz == 0, hence UV == 0: its flow is copied.
    mov = frame_ptr(mov, mov_fs, blockIdx.z);
    flow_in = frame_ptr(flow_in, flow_fs, blockIdx.z);
    flow_out = frame_ptr(flow_out, flow_fs, blockIdx.z);
    if ((int)blockIdx.z == ref_frame) {
        const int gx = blockIdx.x * LTW + threadIdx.x;
        if (gx < w)
            for (int gy = blockIdx.y * LTH + threadIdx.y; gy < min((int)(blockIdx.y + 1) * LTH, h); gy += blockDim.y)
                row_ptr(flow_out, flow_pitch, gy)[gx] = __ldg(row_ptr(flow_in, flow_pitch, gy) + gx);
        return;
    }
    // gh / gy0: height of the full frame and global row of local row 0 (row-band mode; gh == h, gy0 == 0 otherwise): the warp's
    // texture coordinates are normalised by the FULL frame so that a band reproduces the full-frame arithmetic bit for bit
    constexpr int RW = LTW + 2 * HW, RH = LTH + 2 * HW;       // derivative region
    constexpr int WW = RW + 4, WH = RH + 4;                   // source / warped region
    constexpr int NT = 256;
    // the row sums (step 3) reuse the storage of the source / warped regions, which are dead after step 2
    constexpr int A_FLOATS = (2 * WH * WW > 5 * RH * LTW) ? 2 * WH * WW : 5 * RH * LTW;
    __shared__ __align__(16) float s_a[A_FLOATS];
    __shared__ __align__(16) float s_ix[RH][RW], s_iy[RH][RW], s_it[RH][RW];
    float (*s_src)[WW] = reinterpret_cast<float (*)[WW]>(s_a);
    float (*s_wrp)[WW] = reinterpret_cast<float (*)[WW]>(s_a + WH * WW);
    float (*s_h)[RH][LTW] = reinterpret_cast<float (*)[RH][LTW]>(s_a);
    const int x0 = blockIdx.x * LTW, y0 = blockIdx.y * LTH;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int ox = x0 - HW - 2, oy = y0 - HW - 2;             // global coordinate of s_src[0][0]

    // 1. source + warped moved image (WarpingKernel, opticalFlow.cu:28) on the haloed region, clamp addressing.
    // tex_coord's (p / n) * n uses a hoisted correctly rounded reciprocal + one FMA correction step (Markstein): the same
    // correctly rounded quotient as __fdiv_rn without re-deriving the reciprocal and the FCHK slow-path test per element.
    const float fw = (float)w, fh = (float)gh, rw = 1.0f / fw, rh = 1.0f / fh;
    const int pe = (int)(img_pitch >> 2), pf = (int)(flow_pitch >> 3);          // pitches in elements (checked by the launcher)
    // A thread owns one column of the region and walks it in steps of RG rows, CNT rows per round with all loads of one kind
    // issued before any is consumed: element-at-a-time, this step was a chain of two dependent global latencies per element,
    // seven times per thread (64 % of the kernel's stall samples on 41 % of its instructions, profiles/r1q_lk).
    {
        constexpr int RG = NT / WW;                      // row groups
        constexpr int ROWS = (WH + RG - 1) / RG;         // rows per thread
        const int lx = tid % WW, rg = tid / WW;
        if (rg < RG) {
            const int gx = clampi(ox + lx, 0, w - 1);
            const float fgx = (float)gx + 0.5f;
            auto round = [&](auto cnt_tag, int k0) {
                constexpr int CNT = decltype(cnt_tag)::value;
                int ly[CNT], gy[CNT]; bool ok[CNT];
                float srcv[CNT]; float2 fl[CNT];
#pragma unroll
                for (int k = 0; k < CNT; k++) {
                    ly[k] = rg + RG * (k0 + k); ok[k] = ly[k] < WH;
                    gy[k] = clampi(oy + ly[k], 0, h - 1);
                    if (ok[k]) { srcv[k] = __ldg(ref + (unsigned)(gy[k] * pe + gx)); fl[k] = __ldg(flow_in + (unsigned)(gy[k] * pf + gx)); }
                    else { srcv[k] = 0.f; fl[k] = make_float2(0.f, 0.f); }
                }
                if constexpr (MODE == 2) {
                    float wv[CNT];
#pragma unroll
                    for (int k = 0; k < CNT; k++) {
                        const float px = fgx + fl[k].x, py = (float)gy[k] + 0.5f + fl[k].y;
                        float qx = px * rw, qy = py * rh;                    // correctly rounded px / w, py / h (see above)
                        qx = __fmaf_rn(__fmaf_rn(-fw, qx, px), rw, qx);
                        qy = __fmaf_rn(__fmaf_rn(-fh, qy, py), rh, qy);
                        wv[k] = ok[k] ? tex2D<float>(movtex, qx, qy) : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < CNT; k++)
                        if (ok[k]) { s_src[ly[k]][lx] = srcv[k] + wv[k]; s_wrp[ly[k]][lx] = wv[k] - srcv[k]; }
                } else {
                float t00[CNT], t10[CNT], t01[CNT], t11[CNT], fa[CNT], fb[CNT];
#pragma unroll
                for (int k = 0; k < CNT; k++) {
                    const float px = fgx + fl[k].x, py = (float)(gy[k] + gy0) + 0.5f + fl[k].y;
                    float qx = px * rw, qy = py * rh;
                    qx = __fmaf_rn(__fmaf_rn(-fw, qx, px), rw, qx);
                    qy = __fmaf_rn(__fmaf_rn(-fh, qy, py), rh, qy);
                    const TexAxis ax = tex_axis(__fmul_rn(qx, fw), w);
                    TexAxis ay = tex_axis(__fmul_rn(qy, fh), gh);
                    if (BAND) { ay.i0 = clampi(ay.i0 - gy0, 0, h - 1); ay.i1 = clampi(ay.i1 - gy0, 0, h - 1); }   // whole frame: already in [0, h)
                    // unsigned 32-bit element indices (the launcher checks pitch * height < 2^31): one IMAD.WIDE per load instead of
                    // a sign-extended 64-bit add chain (~5 instructions per load)
                    const unsigned r0 = (unsigned)(ay.i0 * pe), r1 = (unsigned)(ay.i1 * pe);
                    if (ok[k]) { t00[k] = __ldg(mov + (r0 + (unsigned)ax.i0)); t10[k] = __ldg(mov + (r0 + (unsigned)ax.i1));
                                 t01[k] = __ldg(mov + (r1 + (unsigned)ax.i0)); t11[k] = __ldg(mov + (r1 + (unsigned)ax.i1)); }
                    else { t00[k] = t10[k] = t01[k] = t11[k] = 0.f; }
                    fa[k] = ax.a; fb[k] = ay.a;
                }
#pragma unroll
                for (int k = 0; k < CNT; k++)
                    if (ok[k]) {
                        // the derivative stencil is linear: ((D src) / 12 + (D warped) / 12) / 2 == D(src + warped) / 24, and
                        // Iz = warped - src, so the sum and the difference are what the next step needs
                        const float wv = tex_mix(t00[k], t10[k], t01[k], t11[k], fa[k], fb[k]);
                        s_src[ly[k]][lx] = srcv[k] + wv;
                        s_wrp[ly[k]][lx] = wv - srcv[k];
                    }
                }
            };
            constexpr int WB = 4;
#pragma unroll
            for (int k0 = 0; k0 + WB <= ROWS; k0 += WB) round(std::integral_constant<int, WB>{}, k0);
            if constexpr (ROWS % WB != 0) round(std::integral_constant<int, ROWS % WB>{}, ROWS - ROWS % WB);
        }
    }
    __syncthreads();
    // 2. derivatives (ComputeDerivativesKernel, :97): 5-tap (1,-8,0,8,-1)/12 on source and warped, averaged == the same stencil
    // on (source + warped) / 24 (s_src holds the sum, s_wrp the difference; re-associated, tolerance-checked).  Region elements
    // were loaded at CLAMPED image coordinates, so plain local neighbours equal the reference's clamp addressing for every
    // element inside the image; elements outside the image are never read by a valid output (:205-207 skips the border).
    // Two horizontally adjacent outputs per thread (aligned float2 loads: WW, RW and the region origin are even).
    static_assert(WW % 2 == 0 && RW % 2 == 0, "float2 access");
    for (int i = tid; i < (RW / 2) * RH; i += NT) {
        const int ry = i / (RW / 2), rx = (i - ry * (RW / 2)) * 2;
        const int cy = ry + 2, cx = rx + 2;
        const float2 a = *(const float2*)&s_src[cy][cx - 2], b = *(const float2*)&s_src[cy][cx], c = *(const float2*)&s_src[cy][cx + 2];
        float2 ix;
        // texSource = warped, texTarget = reference: the stencil below is MINUS the derivative, so Iz = warped - ref is
        // the sign that makes `shift += UV` descend (restated host)
        ix.x = (c.x - b.y * 8.0f + a.y * 8.0f - a.x) * (1.0f / 24.0f);
        ix.y = (c.y - c.x * 8.0f + b.x * 8.0f - a.y) * (1.0f / 24.0f);
        *(float2*)&s_ix[ry][rx] = ix;
        *(float2*)&s_it[ry][rx] = *(const float2*)&s_wrp[cy][cx];
        const float2 u2 = *(const float2*)&s_src[cy - 2][cx], u1 = *(const float2*)&s_src[cy - 1][cx];
        const float2 d1 = *(const float2*)&s_src[cy + 1][cx], d2 = *(const float2*)&s_src[cy + 2][cx];
        float2 iy;
        iy.x = (d2.x - d1.x * 8.0f + u1.x * 8.0f - u2.x) * (1.0f / 24.0f);
        iy.y = (d2.y - d1.y * 8.0f + u1.y * 8.0f - u2.y) * (1.0f / 24.0f);
        *(float2*)&s_iy[ry][rx] = iy;
    }
    __syncthreads();
    // 3. row sums of the five products over [-HW, HW]: four adjacent outputs per thread from 2*HW+4 taps (LDS.64); the first is a
    // full sum, the next three slide the window (+ entering tap, - leaving tap: 15 instead of 35 operations per output)
    for (int i = tid; i < RH * (LTW / 4); i += NT) {
        const int ry = i / (LTW / 4), lx = (i - ry * (LTW / 4)) * 4;
        float dx[2 * HW + 4], dy[2 * HW + 4], dt[2 * HW + 4];
#pragma unroll
        for (int k = 0; k < HW + 2; k++) {
            const float2 a = *(const float2*)&s_ix[ry][lx + 2 * k], b = *(const float2*)&s_iy[ry][lx + 2 * k], c = *(const float2*)&s_it[ry][lx + 2 * k];
            dx[2 * k] = a.x; dx[2 * k + 1] = a.y; dy[2 * k] = b.x; dy[2 * k + 1] = b.y; dt[2 * k] = c.x; dt[2 * k + 1] = c.y;
        }
        float sxx = 0.f, sxy = 0.f, syy = 0.f, sxt = 0.f, syt = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * HW; k++) {
            sxx += dx[k] * dx[k]; sxy += dx[k] * dy[k]; syy += dy[k] * dy[k]; sxt += dx[k] * dt[k]; syt += dy[k] * dt[k];
        }
        s_h[0][ry][lx] = sxx; s_h[1][ry][lx] = sxy; s_h[2][ry][lx] = syy; s_h[3][ry][lx] = sxt; s_h[4][ry][lx] = syt;
#pragma unroll
        for (int o = 1; o < 4; o++) {
            const int kin = o + 2 * HW, kout = o - 1;
            sxx += dx[kin] * dx[kin] - dx[kout] * dx[kout]; sxy += dx[kin] * dy[kin] - dx[kout] * dy[kout];
            syy += dy[kin] * dy[kin] - dy[kout] * dy[kout]; sxt += dx[kin] * dt[kin] - dx[kout] * dt[kout];
            syt += dy[kin] * dt[kin] - dy[kout] * dt[kout];
            s_h[0][ry][lx + o] = sxx; s_h[1][ry][lx + o] = sxy; s_h[2][ry][lx + o] = syy; s_h[3][ry][lx + o] = sxt; s_h[4][ry][lx + o] = syt;
        }
    }
    __syncthreads();
    // 4. column sums, pseudo-inverse, update (lucasKanadeOptim, :190): 4 vertically adjacent pixels per thread
    {
        const int lx = threadIdx.x, ly0 = threadIdx.y * 4, gx = x0 + lx;
        // the four windows [p, p + 2 HW] share the rows 3 .. 2 HW and pairs of their edge rows: 13 additions per quantity
        // instead of 24
        float acc[4][5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            float v[2 * HW + 4];
#pragma unroll
            for (int r = 0; r < 2 * HW + 4; r++) v[r] = s_h[q][ly0 + r][lx];
            float core = 0.f;
#pragma unroll
            for (int r = 3; r <= 2 * HW; r++) core = (r == 3) ? v[r] : core + v[r];
            const float lo = v[1] + v[2], hi = v[2 * HW + 1] + v[2 * HW + 2];
            acc[0][q] = (v[0] + lo) + core;
            acc[1][q] = (lo + v[2 * HW + 1]) + core;
            acc[2][q] = (v[2] + hi) + core;
            acc[3][q] = (hi + v[2 * HW + 3]) + core;
        }
        if (gx < w) {
            float2 fin[4];                     // the four flow values are requested before the first pseudo-inverse
#pragma unroll
            for (int p = 0; p < 4; p++) fin[p] = __ldg(flow_in + (unsigned)(min(y0 + ly0 + p, h - 1) * pf + gx));
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const int gy = y0 + ly0 + p;
                if (gy >= h) break;
                float2 f = fin[p];
                if (!(gx < HW || gx >= w - HW || gy < HW || gy >= h - HW)) {
                    float inv[4];
                    if (lk_pinv(acc[p][0], acc[p][1], acc[p][1], acc[p][2], minDet, inv)) {
                        float u = inv[0] * acc[p][3] + inv[1] * acc[p][4];
                        float v = inv[2] * acc[p][3] + inv[3] * acc[p][4];
                        u = isnan(u) ? 0.f : u;
                        v = isnan(v) ? 0.f : v;
                        f.x += u; f.y += v;
                    }
                }
                row_ptr(flow_out, flow_pitch, gy)[gx] = f;
            }
        }
    }
}

}  // namespace mfsr

using namespace mfsr;

// frames > 1: frame f reads tiles + f * tiles_fs (bytes) and frame_pose + 4 f, writes flow + f * flow_fs
int mfsr::launch_flow_from_tiles(const float2* tiles, int64_t tile_pitch, int tilesX, int tilesY, float2* flow, int64_t flow_pitch, int w, int h,
                                 float bsx, float bsy, float rot, int gh, int gy0, int gty, int tile_row0, cudaStream_t st, const float* frame_pose,
                                 int frames, int64_t tiles_fs, int64_t flow_fs)
{
    if (!tiles || !flow || tilesX < 1 || tilesY < 1 || w < 1 || h < 1 || frames < 1) return MFSR_E_INVALID;
    if (gh <= 0) { gh = h; gy0 = 0; gty = tilesY; tile_row0 = 0; }
    dim3 b(32, 8), g(cdiv(w, 32), cdiv(h, 8 * FFT_ROWS), frames);
    flow_from_tiles_kernel<<<g, b, 0, st>>>(tiles, tile_pitch, tilesX, tilesY, flow, flow_pitch, w, h, bsx, bsy, cosf(rot), sinf(rot), gh, gy0, gty, tile_row0, frame_pose,
                                            tiles_fs, flow_fs);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

extern "C" int mfsr_stage_flow_from_tiles(const float* tile_shift, int64_t tile_pitch, int tilesX, int tilesY, int tile_size,
                                          float* flow, int64_t flow_pitch, int width, int height,
                                          float base_shift_x, float base_shift_y, float base_rotation, void* stream)
{
    (void)tile_size;
    return launch_flow_from_tiles((const float2*)tile_shift, tile_pitch, tilesX, tilesY, (float2*)flow, flow_pitch, width, height,
                                  base_shift_x, base_shift_y, base_rotation, 0, 0, 0, 0, (cudaStream_t)stream);
}

// frames > 1: one sweep for every frame of the burst (frame f: mov + f * mov_fs, flow + f * flow_fs; frame `ref_frame`, if >= 0, is the
// reference frame itself and only has its flow copied).  The texture form is per frame (one texture object per moved image).
int mfsr::launch_lk_iteration(const float* ref, const float* mov, int64_t img_pitch, const float2* flow_in, float2* flow_out, int64_t flow_pitch,
                              int width, int height, int half_window, float min_det, int gh, int gy0, cudaStream_t st, cudaTextureObject_t movtex,
                              int frames, int64_t mov_fs, int64_t flow_fs, int ref_frame)
{
    if (!ref || !mov || !flow_in || !flow_out || flow_in == flow_out || width < 1 || height < 1 || frames < 1) return MFSR_E_INVALID;
    if (movtex && frames != 1) return MFSR_E_INVALID;
    if (half_window < 1 || half_window > LHW_MAX) return MFSR_E_INVALID;
    // 32-bit element indexing inside the kernel
    if ((img_pitch & 3) || (flow_pitch & 7) || (int64_t)(img_pitch >> 2) * height >= (1ll << 31) || (int64_t)(flow_pitch >> 3) * height >= (1ll << 31)) return MFSR_E_INVALID;
    if (gh <= 0) { gh = height; gy0 = 0; }
    dim3 b(LTW, 8), g(cdiv(width, LTW), cdiv(height, LTH), frames);
    const bool band = !(gh == height && gy0 == 0);
    if (band) movtex = 0;             // a band reproduces the full frame bit for bit only with the ALU model (texture rows of the full frame)
#define MFSR_LK_ARGS ref, mov, img_pitch, flow_in, flow_out, flow_pitch, width, height, min_det, gh, gy0
#define MFSR_LK(HW_) do { if (band) lk_iteration_kernel<HW_, 1><<<g, b, 0, st>>>(MFSR_LK_ARGS, 0, mov_fs, flow_fs, ref_frame); \
                          else if (movtex) lk_iteration_kernel<HW_, 2><<<g, b, 0, st>>>(MFSR_LK_ARGS, movtex, mov_fs, flow_fs, ref_frame); \
                          else lk_iteration_kernel<HW_, 0><<<g, b, 0, st>>>(MFSR_LK_ARGS, 0, mov_fs, flow_fs, ref_frame); } while (0)
    switch (half_window) {
        case 1: MFSR_LK(1); break;
        case 2: MFSR_LK(2); break;
        case 3: MFSR_LK(3); break;
        default: MFSR_LK(4); break;
    }
#undef MFSR_LK
#undef MFSR_LK_ARGS
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

// linear-filter / clamp / normalised-coordinate texture over a pitched float image (what the reference binds, opticalFlow.cu:36-41)
int mfsr::make_gray_texture(const float* img, int64_t pitch, int width, int height, cudaTextureObject_t* out)
{
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = (void*)img; rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
    rd.res.pitch2D.width = (size_t)width; rd.res.pitch2D.height = (size_t)height; rd.res.pitch2D.pitchInBytes = (size_t)pitch;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
    MFSR_CUDA_TRY(cudaCreateTextureObject(out, &rd, &td, nullptr));
    return MFSR_OK;
}

extern "C" int mfsr_stage_lk_iteration(const float* ref, const float* mov, int64_t img_pitch, const float* flow_in, float* flow_out,
                                       int64_t flow_pitch, int width, int height, int half_window, float min_det, void* stream)
{
    return launch_lk_iteration(ref, mov, img_pitch, (const float2*)flow_in, (float2*)flow_out, flow_pitch, width, height, half_window, min_det,
                               0, 0, (cudaStream_t)stream, 0);
}

// The same sweep with the warp's bilinear fetch on the texture unit (what mfsr_run uses outside row-band mode).  `mov` must satisfy the
// device's texture alignment (base 512 B, pitch 32 B).  The texture object lives for the duration of the call (the stream is
// synchronised before it is destroyed): a test entry point, not a hot path.
extern "C" int mfsr_stage_lk_iteration_tex(const float* ref, const float* mov, int64_t img_pitch, const float* flow_in, float* flow_out,
                                           int64_t flow_pitch, int width, int height, int half_window, float min_det, void* stream)
{
    if (!mov || ((uintptr_t)mov & 511) || (img_pitch & 31)) return MFSR_E_INVALID;
    cudaTextureObject_t t = 0;
    int rc = make_gray_texture(mov, img_pitch, width, height, &t);
    if (rc) return rc;
    rc = launch_lk_iteration(ref, mov, img_pitch, (const float2*)flow_in, (float2*)flow_out, flow_pitch, width, height, half_window, min_det,
                             0, 0, (cudaStream_t)stream, t);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaDestroyTextureObject(t);
    return rc;
}
