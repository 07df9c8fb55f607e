// frontend.cu — per-frame front end: Bayer sub-sampling, demosaic, tracking image, pyramid.
//
// Compiled with -fmad=false (see build.py): the 7-bit tracking image feeds the
// integer block matcher, so its fp32 arithmetic must round exactly like a strict
// IEEE evaluation of the same formulas on a CPU — then the integer
// tile shifts are bit-exact by construction.  These kernels are HBM-bound
// (2 B in, 4..12 B out per pixel); the lost FMA contraction costs nothing.
//
// Replaces deBayersSubSample3 (DeBayerKernels.cu:244), deBayerGreenKernel (:55) +
// deBayerRedBlueKernel (:153) — fused: the green plane lives in shared memory
// instead of a second launch reading the first one's HBM output — and the absent
// host's B/W + Gaussian (gaussin_filter_1D, main.cpp:370) + resize steps.
#include "common.cuh"
#include "internal.h"

namespace mfsr {

// ---------------------------------------------------------------- subsample3
__global__ void __launch_bounds__(256)
subsample3_kernel(const uint16_t* __restrict__ raw, int64_t raw_pitch, float* __restrict__ out, int64_t out_pitch,
                  float factor, int dimX, int dimY, Cfa cfa, int64_t raw_fs, int64_t out_fs)
{
    raw = frame_ptr(raw, raw_fs, blockIdx.z);       // one launch per burst: blockIdx.z = frame
    out = frame_ptr(out, out_fs, blockIdx.z);
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dimX || y >= dimY) return;
    // one 32-bit load per Bayer row: two neighbouring u16 samples
    const uint32_t r0 = *(const uint32_t*)(row_ptr(raw, raw_pitch, 2 * y) + 2 * x);
    const uint32_t r1 = *(const uint32_t*)(row_ptr(raw, raw_pitch, 2 * y + 1) + 2 * x);
    const float v[4] = {(float)(r0 & 0xffffu), (float)(r0 >> 16), (float)(r1 & 0xffffu), (float)(r1 >> 16)};
    float px[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int ix = 0; ix < 2; ix++)
#pragma unroll
        for (int iy = 0; iy < 2; iy++) {
            const int col = cfa.c[iy * 2 + ix];
            const float r = v[iy * 2 + ix];
            if (col == MFSR_GREEN) px[1] += r * factor * 0.5f;      // two greens per quad (:267)
            else if (col == MFSR_RED) px[0] = r * factor;
            else if (col == MFSR_BLUE) px[2] = r * factor;
        }
    float* o = row_ptr(out, out_pitch, y) + 3 * x;
    o[0] = px[0]; o[1] = px[1]; o[2] = px[2];
}

// ---------------------------------------------------------------- demosaic tile
// Shared-memory staged demosaic of a TW x TH tile grown by HR pixels on each side.
//   raw plane : region grown by HR+3 (green needs +-2 of its own +-1 halo)
//   green     : region grown by HR+1
// Pixels in the 2-px image border (never written by the reference) are 0.
template <int TW, int TH, int HR>
struct DemosaicTile {
    static constexpr int RW = TW + 2 * (HR + 3), RH = TH + 2 * (HR + 3);
    static constexpr int GW = TW + 2 * (HR + 1), GH = TH + 2 * (HR + 1);
    float raw[RH][RW];
    float grn[GH][GW];
};

__device__ __forceinline__ float rawc(float v, int c, const F3& black, const F3& scale) { return (v - black.v[c]) * scale.v[c]; }

// PN ("pre-normalised"): the raw plane holds (v - black[c]) * scale[c] with c = the CFA colour of the sample's own
// position, computed once per sample instead of once per use (a sample is used ~8 times by the two demosaic steps).
// Identical values whenever every colour the reference passes to its RAW macro equals the colour at that position,
// i.e. for the four Bayer patterns and for an all-green (monochrome) CFA; the host selects PN only then.
template <bool PN>
__device__ __forceinline__ float rawv(float v, int c, const F3& black, const F3& scale) { return PN ? v : rawc(v, c, black, scale); }

template <int TW, int TH, int HR, bool PN>
__device__ void demosaic_stage(DemosaicTile<TW, TH, HR>& S, const uint16_t* __restrict__ raw, int64_t raw_pitch,
                               int w, int h, int x0, int y0, const Cfa& cfa, const F3& black, const F3& scale)
{
    using DT = DemosaicTile<TW, TH, HR>;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    for (int i = tid; i < DT::RW * DT::RH; i += nthr) {
        const int ry = i / DT::RW, rx = i - ry * DT::RW;
        const int nx = x0 - (HR + 3) + rx, ny = y0 - (HR + 3) + ry;
        const int gx = clampi(nx, 0, w - 1), gy = clampi(ny, 0, h - 1);
        const float v = (float)row_ptr(raw, raw_pitch, gy)[gx];
        // out-of-image samples are only read by pixels the reference never writes (2-px border): their value is irrelevant
        S.raw[ry][rx] = PN ? rawc(v, cfa.c[(ny & 1) * 2 + (nx & 1)], black, scale) : v;
    }
    __syncthreads();
    // green (deBayerGreenKernel, DeBayerKernels.cu:55-149)
    for (int i = tid; i < DT::GW * DT::GH; i += nthr) {
        const int qy = i / DT::GW, qx = i - qy * DT::GW;
        const int gx = x0 - (HR + 1) + qx, gy = y0 - (HR + 1) + qy;
        float g = 0.f;
        if (gx >= 2 && gx < w - 2 && gy >= 2 && gy < h - 2) {
            const int ry = qy + 2, rx = qx + 2;
            const int col = cfa.c[(gy & 1) * 2 + (gx & 1)];
            if (col == MFSR_GREEN) g = rawv<PN>(S.raw[ry][rx], 1, black, scale);
            else if (col == MFSR_RED || col == MFSR_BLUE) {
                const float p = rawv<PN>(S.raw[ry][rx], col, black, scale);
                const float xm2 = rawv<PN>(S.raw[ry][rx - 2], col, black, scale), xm1 = rawv<PN>(S.raw[ry][rx - 1], 1, black, scale);
                const float xp1 = rawv<PN>(S.raw[ry][rx + 1], 1, black, scale), xp2 = rawv<PN>(S.raw[ry][rx + 2], col, black, scale);
                const float ym2 = rawv<PN>(S.raw[ry - 2][rx], col, black, scale), ym1 = rawv<PN>(S.raw[ry - 1][rx], 1, black, scale);
                const float yp1 = rawv<PN>(S.raw[ry + 1][rx], 1, black, scale), yp2 = rawv<PN>(S.raw[ry + 2][rx], col, black, scale);
                const float gradX = 0.5f * fabsf(xp1 - xm1), gradY = 0.5f * fabsf(yp1 - ym1);
                const float lapX = 0.25f * fabsf(2.0f * p - xm2 - xp2), lapY = 0.25f * fabsf(2.0f * p - ym2 - yp2);
                const float ipX = 0.125f * (-xm2 + 4.0f * xm1 + 2.0f * p + 4.0f * xp1 - xp2);
                const float ipY = 0.125f * (-ym2 + 4.0f * ym1 + 2.0f * p + 4.0f * yp1 - yp2);
                const float wgt = (gradY + lapY) / (gradX + gradY + lapX + lapY + 0.000000001f);
                g = wgt * ipX + (1.0f - wgt) * ipY;
            }
        }
        S.grn[qy][qx] = g;
    }
    __syncthreads();
}

// red/blue of one pixel (deBayerRedBlueKernel, :153-231); (lx,ly) relative to the tile origin, in [-HR, T+HR)
template <int TW, int TH, int HR, bool PN>
__device__ __forceinline__ void demosaic_px(const DemosaicTile<TW, TH, HR>& S, int lx, int ly, int gx, int gy, int w, int h,
                                            const Cfa& cfa, const F3& black, const F3& scale, float& r, float& g, float& b)
{
    r = g = b = 0.f;
    if (gx < 2 || gx >= w - 2 || gy < 2 || gy >= h - 2) return;
    const int qx = lx + HR + 1, qy = ly + HR + 1, rx = lx + HR + 3, ry = ly + HR + 3;
#define RAWX(dx, dy, c) rawv<PN>(S.raw[ry + (dy)][rx + (dx)], c, black, scale)
#define GRNX(dx, dy) S.grn[qy + (dy)][qx + (dx)]
    const int col = cfa.c[(gy & 1) * 2 + (gx & 1)];
    const int row = cfa.c[(gy & 1) * 2 + ((gx + 1) & 1)];
    g = GRNX(0, 0);
    if (col == MFSR_GREEN) {
        if (row == MFSR_RED) {
            r = g + 0.5f * ((RAWX(-1, 0, 0) - GRNX(-1, 0)) + (RAWX(1, 0, 0) - GRNX(1, 0)));
            b = g + 0.5f * ((RAWX(0, -1, 2) - GRNX(0, -1)) + (RAWX(0, 1, 2) - GRNX(0, 1)));
        } else {
            b = g + 0.5f * ((RAWX(-1, 0, 2) - GRNX(-1, 0)) + (RAWX(1, 0, 2) - GRNX(1, 0)));
            r = g + 0.5f * ((RAWX(0, -1, 0) - GRNX(0, -1)) + (RAWX(0, 1, 0) - GRNX(0, 1)));
        }
    } else if (col == MFSR_RED) {
        r = RAWX(0, 0, 0);
        b = g + 0.25f * ((RAWX(-1, -1, 2) - GRNX(-1, -1)) + (RAWX(1, -1, 2) - GRNX(1, -1)) + (RAWX(1, 1, 2) - GRNX(1, 1)) + (RAWX(-1, 1, 2) - GRNX(-1, 1)));
    } else if (col == MFSR_BLUE) {
        b = RAWX(0, 0, 2);
        r = g + 0.25f * ((RAWX(-1, -1, 0) - GRNX(-1, -1)) + (RAWX(1, -1, 0) - GRNX(1, -1)) + (RAWX(1, 1, 0) - GRNX(1, 1)) + (RAWX(-1, 1, 0) - GRNX(-1, 1)));
    }
#undef RAWX
#undef GRNX
}

constexpr int DTW = 32, DTH = 16;

__global__ void __launch_bounds__(256)
demosaic_kernel(const uint16_t* __restrict__ raw, int64_t raw_pitch, float* __restrict__ rgb, int64_t rgb_pitch,
                int w, int h, Cfa cfa, F3 black, F3 scale)
{
    __shared__ DemosaicTile<DTW, DTH, 0> S;
    const int x0 = blockIdx.x * DTW, y0 = blockIdx.y * DTH;
    demosaic_stage<DTW, DTH, 0, false>(S, raw, raw_pitch, w, h, x0, y0, cfa, black, scale);
    for (int ly = threadIdx.y; ly < DTH; ly += blockDim.y) {
        const int gx = x0 + threadIdx.x, gy = y0 + ly;
        if (gx >= w || gy >= h) continue;
        float r, g, b;
        demosaic_px<DTW, DTH, 0, false>(S, (int)threadIdx.x, ly, gx, gy, w, h, cfa, black, scale, r, g, b);
        float* o = row_ptr(rgb, rgb_pitch, gy) + 3 * gx;
        o[0] = r; o[1] = g; o[2] = b;
    }
}

// ---------------------------------------------------------------- tracking image
constexpr int MAX_TAPS = 9;
struct Taps { float t[MAX_TAPS]; int n; };

// 64-wide tiles (TH = 32 rows for R <= 2, 16 for larger blurs: static shared memory <= 48 KB).  The first version used
// 32x16 tiles and re-normalised every raw sample at every use: 670 instructions per pixel (profiles/r1k).
template <int R, bool PN>
__global__ void __launch_bounds__(256)
tracking_kernel(const uint16_t* __restrict__ raw, int64_t raw_pitch, float* __restrict__ gray, int64_t gray_pitch,
                uint8_t* __restrict__ gray_q, int64_t gq_pitch, int w, int h, Cfa cfa, F3 black, F3 scale, Taps taps, float qmax, FrameStrides fs)
{
    raw = frame_ptr(raw, fs.s[0], blockIdx.z);
    if (gray) gray = frame_ptr(gray, fs.s[1], blockIdx.z);
    if (gray_q) gray_q = frame_ptr(gray_q, fs.s[2], blockIdx.z);
    constexpr int TW = 64, TH = (R <= 2) ? 32 : 16;
    constexpr int LW = TW + 2 * R, LH = TH + 2 * R;
    __shared__ DemosaicTile<TW, TH, R> S;
    __shared__ float lum[LH][LW];
    __shared__ float hb[LH][TW];
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    demosaic_stage<TW, TH, R, PN>(S, raw, raw_pitch, w, h, x0, y0, cfa, black, scale);
    for (int i = tid; i < LW * LH; i += nthr) {
        const int ly = i / LW - R, lx = i % LW - R;
        float r, g, b;
        demosaic_px<TW, TH, R, PN>(S, lx, ly, x0 + lx, y0 + ly, w, h, cfa, black, scale, r, g, b);
        // out-of-image pixels are never read (clamped indices below)
        lum[ly + R][lx + R] = 0.25f * r + 0.5f * g + 0.25f * b;
    }
    __syncthreads();
    // horizontal pass, clamp border: index = clamp(x+k-c, 0, w-1)
    for (int i = tid; i < TW * LH; i += nthr) {
        const int ry = i / TW, lx = i - ry * TW;
        const int gx = x0 + lx;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * R + 1; k++) {
            const int sx = clampi(gx + k - R, 0, w - 1) - x0 + R;
            acc += taps.t[k] * lum[ry][clampi(sx, 0, LW - 1)];
        }
        hb[ry][lx] = acc;
    }
    __syncthreads();
    for (int i = tid; i < TW * TH; i += nthr) {
        const int ly = i / TW, lx = i - ly * TW;
        const int gx = x0 + lx, gy = y0 + ly;
        if (gx >= w || gy >= h) continue;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * R + 1; k++) {
            const int sy = clampi(gy + k - R, 0, h - 1) - y0 + R;
            acc += taps.t[k] * hb[clampi(sy, 0, LH - 1)][lx];
        }
        if (gray) row_ptr(gray, gray_pitch, gy)[gx] = acc;
        if (gray_q) {
            float q = floorf(acc * qmax + 0.5f);
            q = fminf(fmaxf(q, 0.0f), qmax);
            row_ptr(gray_q, gq_pitch, gy)[gx] = (uint8_t)q;
        }
    }
}

// ---------------------------------------------------------------- tracking image, quad form
// Bayer / monochrome fast path.  Tiles are 64x32 with an even origin, so the CFA colour of every sample of a 2x2 quad
// is a COMPILE-TIME function of the pattern PAT (c00 | c10 << 2 | c01 << 4 | c11 << 6, row-major like c_cfaPattern) and
// of the region's origin parity: a thread processes whole quads with straight-line code — no per-sample colour lookups
// in the kernel-parameter bank (78 LDC + 105 ISETP per pixel in the element-wise version, profiles/r1m).
// Arithmetic (operation order, -fmad=false) is identical to deBayerGreenKernel / deBayerRedBlueKernel
// (DeBayerKernels.cu:55-231) on pre-normalised samples (see rawv above), followed by the restated luma + Gaussian.
template <int PAT> __device__ __forceinline__ constexpr int pat_col(int x, int y) { return (PAT >> (2 * ((y & 1) * 2 + (x & 1)))) & 3; }

template <int COL, int ROW, int RP, int GP>
__device__ __forceinline__ float px_luma(const float* __restrict__ rw, const float* __restrict__ gr)
{
    // rw / gr point at the pixel inside the normalised raw plane (pitch RP) and the green plane (pitch GP)
    const float g = gr[0];
    float r, b;
#define RW_(dx, dy) rw[(dy) * RP + (dx)]
#define GR_(dx, dy) gr[(dy) * GP + (dx)]
    if (COL == MFSR_GREEN) {
        const float hz = g + 0.5f * ((RW_(-1, 0) - GR_(-1, 0)) + (RW_(1, 0) - GR_(1, 0)));
        const float vt = g + 0.5f * ((RW_(0, -1) - GR_(0, -1)) + (RW_(0, 1) - GR_(0, 1)));
        if (ROW == MFSR_RED) { r = hz; b = vt; } else { b = hz; r = vt; }
    } else {
        const float dg = g + 0.25f * ((RW_(-1, -1) - GR_(-1, -1)) + (RW_(1, -1) - GR_(1, -1)) + (RW_(1, 1) - GR_(1, 1)) + (RW_(-1, 1) - GR_(-1, 1)));
        if (COL == MFSR_RED) { r = RW_(0, 0); b = dg; } else { b = RW_(0, 0); r = dg; }
    }
#undef RW_
#undef GR_
    return 0.25f * r + 0.5f * g + 0.25f * b;
}

template <int COL, int RP>
__device__ __forceinline__ float px_green(const float* __restrict__ rw)
{
    if (COL == MFSR_GREEN) return rw[0];
    const float p = rw[0];
    const float xm2 = rw[-2], xm1 = rw[-1], xp1 = rw[1], xp2 = rw[2];
    const float ym2 = rw[-2 * RP], ym1 = rw[-RP], yp1 = rw[RP], yp2 = rw[2 * RP];
    const float gradX = 0.5f * fabsf(xp1 - xm1), gradY = 0.5f * fabsf(yp1 - ym1);
    const float lapX = 0.25f * fabsf(2.0f * p - xm2 - xp2), lapY = 0.25f * fabsf(2.0f * p - ym2 - yp2);
    const float ipX = 0.125f * (-xm2 + 4.0f * xm1 + 2.0f * p + 4.0f * xp1 - xp2);
    const float ipY = 0.125f * (-ym2 + 4.0f * ym1 + 2.0f * p + 4.0f * yp1 - yp2);
    const float wgt = (gradY + lapY) / (gradX + gradY + lapX + lapY + 0.000000001f);
    return wgt * ipX + (1.0f - wgt) * ipY;
}

template <int R, int PAT>
__global__ void __launch_bounds__(256)
tracking_quad_kernel(const uint16_t* __restrict__ raw, int64_t raw_pitch, float* __restrict__ gray, int64_t gray_pitch,
                     uint8_t* __restrict__ gray_q, int64_t gq_pitch, int w, int h, F3 black, F3 scale, Taps taps, float qmax, FrameStrides fs)
{
    raw = frame_ptr(raw, fs.s[0], blockIdx.z);
    if (gray) gray = frame_ptr(gray, fs.s[1], blockIdx.z);
    if (gray_q) gray_q = frame_ptr(gray_q, fs.s[2], blockIdx.z);
    constexpr int TW = 64, TH = 32;
    constexpr int RW = TW + 2 * (R + 3), RH = TH + 2 * (R + 3);      // normalised raw plane, origin (x0 - R - 3, y0 - R - 3)
    constexpr int GW = TW + 2 * (R + 1), GH = TH + 2 * (R + 1);      // green plane,          origin (x0 - R - 1, y0 - R - 1)
    constexpr int LW = TW + 2 * R, LH = TH + 2 * R;                  // luma,                 origin (x0 - R, y0 - R)
    constexpr int LWP = LW + (LW & 1), LHP = LH + (LH & 1);          // luma computed on whole quads
    constexpr int PR = (R + 3) & 1, PG = (R + 1) & 1, PL = R & 1;    // origin parities (tile origins are even)
    __shared__ __align__(8) float s_raw[RH][RW];
    __shared__ float s_grn[GH][GW];
    __shared__ float s_lum[LHP][LWP];
    __shared__ float s_hb[LH][TW];
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    constexpr int NT = 256;

    // ---- normalised raw plane, one 2x2 quad per iteration
    {
        const int ox = x0 - (R + 3), oy = y0 - (R + 3);
        const bool inside = ox >= 0 && oy >= 0 && ox + RW <= w && oy + RH <= h && !(raw_pitch & 3) && !((uintptr_t)raw & 3) && !(ox & 1);
        for (int q = tid; q < (RW / 2) * (RH / 2); q += NT) {
            const int qy = q / (RW / 2), qx = q - qy * (RW / 2);
            const int rx = 2 * qx, ry = 2 * qy;
            float v[2][2];
            if (inside) {
#pragma unroll
                for (int dy = 0; dy < 2; dy++) {
                    const unsigned u = *(const unsigned*)(row_ptr(raw, raw_pitch, oy + ry + dy) + ox + rx);
                    v[dy][0] = (float)(u & 0xffffu); v[dy][1] = (float)(u >> 16);
                }
            } else {
#pragma unroll
                for (int dy = 0; dy < 2; dy++)
#pragma unroll
                    for (int dx = 0; dx < 2; dx++)
                        v[dy][dx] = (float)row_ptr(raw, raw_pitch, clampi(oy + ry + dy, 0, h - 1))[clampi(ox + rx + dx, 0, w - 1)];
            }
#pragma unroll
            for (int dy = 0; dy < 2; dy++) {
                const int c0 = pat_col<PAT>(PR, dy + PR), c1 = pat_col<PAT>(1 + PR, dy + PR);
                *(float2*)&s_raw[ry + dy][rx] = make_float2((v[dy][0] - black.v[c0]) * scale.v[c0], (v[dy][1] - black.v[c1]) * scale.v[c1]);
            }
        }
    }
    __syncthreads();
    // ---- green plane (deBayerGreenKernel): zero outside [2, dim-2)
    {
        const int ox = x0 - (R + 1), oy = y0 - (R + 1);
        const bool interior = ox >= 2 && oy >= 2 && ox + GW <= w - 2 && oy + GH <= h - 2;
        for (int q = tid; q < (GW / 2) * (GH / 2); q += NT) {
            const int qy = q / (GW / 2), qx = q - qy * (GW / 2);
            const int gx0 = 2 * qx, gy0 = 2 * qy;
            const float* rw = &s_raw[gy0 + 2][gx0 + 2];
            float g[2][2];
            g[0][0] = px_green<pat_col<PAT>(PG, PG), RW>(rw);
            g[0][1] = px_green<pat_col<PAT>(1 + PG, PG), RW>(rw + 1);
            g[1][0] = px_green<pat_col<PAT>(PG, 1 + PG), RW>(rw + RW);
            g[1][1] = px_green<pat_col<PAT>(1 + PG, 1 + PG), RW>(rw + RW + 1);
            if (!interior) {
#pragma unroll
                for (int dy = 0; dy < 2; dy++)
#pragma unroll
                    for (int dx = 0; dx < 2; dx++) {
                        const int ax = ox + gx0 + dx, ay = oy + gy0 + dy;
                        if (ax < 2 || ax >= w - 2 || ay < 2 || ay >= h - 2) g[dy][dx] = 0.f;
                    }
            }
            s_grn[gy0][gx0] = g[0][0]; s_grn[gy0][gx0 + 1] = g[0][1]; s_grn[gy0 + 1][gx0] = g[1][0]; s_grn[gy0 + 1][gx0 + 1] = g[1][1];
        }
    }
    __syncthreads();
    // ---- red / blue (deBayerRedBlueKernel) + luma 0.25 R + 0.5 G + 0.25 B on the blur's support
    {
        const int ox = x0 - R, oy = y0 - R;
        const bool interior = ox >= 2 && oy >= 2 && ox + LWP <= w - 2 && oy + LHP <= h - 2;
        for (int q = tid; q < (LWP / 2) * (LHP / 2); q += NT) {
            const int qy = q / (LWP / 2), qx = q - qy * (LWP / 2);
            const int lx0 = 2 * qx, ly0 = 2 * qy;
            const float* rw = &s_raw[ly0 + 3][lx0 + 3];
            const float* gr = &s_grn[ly0 + 1][lx0 + 1];
            float l[2][2];
            l[0][0] = px_luma<pat_col<PAT>(PL, PL), pat_col<PAT>(1 + PL, PL), RW, GW>(rw, gr);
            l[0][1] = px_luma<pat_col<PAT>(1 + PL, PL), pat_col<PAT>(PL, PL), RW, GW>(rw + 1, gr + 1);
            l[1][0] = px_luma<pat_col<PAT>(PL, 1 + PL), pat_col<PAT>(1 + PL, 1 + PL), RW, GW>(rw + RW, gr + GW);
            l[1][1] = px_luma<pat_col<PAT>(1 + PL, 1 + PL), pat_col<PAT>(PL, 1 + PL), RW, GW>(rw + RW + 1, gr + GW + 1);
            if (!interior) {
#pragma unroll
                for (int dy = 0; dy < 2; dy++)
#pragma unroll
                    for (int dx = 0; dx < 2; dx++) {
                        const int ax = ox + lx0 + dx, ay = oy + ly0 + dy;
                        if (ax < 2 || ax >= w - 2 || ay < 2 || ay >= h - 2) l[dy][dx] = 0.f;     // r = g = b = 0 (unwritten border)
                    }
            }
            s_lum[ly0][lx0] = l[0][0]; s_lum[ly0][lx0 + 1] = l[0][1]; s_lum[ly0 + 1][lx0] = l[1][0]; s_lum[ly0 + 1][lx0 + 1] = l[1][1];
        }
    }
    __syncthreads();
    // ---- separable Gaussian, clamp border (index = clamp(x + k - c, 0, w - 1)), quantisation
    const bool edge = x0 - R < 0 || y0 - R < 0 || x0 + TW + R > w || y0 + TH + R > h;
    for (int i = tid; i < TW * LH; i += NT) {
        const int ry = i / TW, lx = i - ry * TW;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * R + 1; k++) {
            const int sx = edge ? clampi(clampi(x0 + lx + k - R, 0, w - 1) - x0 + R, 0, LW - 1) : lx + k;
            acc += taps.t[k] * s_lum[ry][sx];
        }
        s_hb[ry][lx] = acc;
    }
    __syncthreads();
    for (int i = tid; i < TW * TH; i += NT) {
        const int ly = i / TW, lx = i - ly * TW;
        const int gx = x0 + lx, gy = y0 + ly;
        if (gx >= w || gy >= h) continue;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * R + 1; k++) {
            const int sy = edge ? clampi(clampi(gy + k - R, 0, h - 1) - y0 + R, 0, LH - 1) : ly + k;
            acc += taps.t[k] * s_hb[sy][lx];
        }
        if (gray) row_ptr(gray, gray_pitch, gy)[gx] = acc;
        if (gray_q) {
            float q = floorf(acc * qmax + 0.5f);
            q = fminf(fmaxf(q, 0.0f), qmax);
            row_ptr(gray_q, gq_pitch, gy)[gx] = (uint8_t)q;
        }
    }
}

// ---------------------------------------------------------------- pyramid
__global__ void __launch_bounds__(256)
pyramid_down_kernel(const uint8_t* __restrict__ in, int64_t in_pitch, uint8_t* __restrict__ out, int64_t out_pitch, int ow, int oh,
                    int64_t in_fs, int64_t out_fs)
{
    in = frame_ptr(in, in_fs, blockIdx.z);
    out = frame_ptr(out, out_fs, blockIdx.z);
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= ow || y >= oh) return;
    const uint8_t* r0 = row_ptr(in, in_pitch, 2 * y) + 2 * x;
    const uint8_t* r1 = row_ptr(in, in_pitch, 2 * y + 1) + 2 * x;
    const int s = (int)r0[0] + (int)r0[1] + (int)r1[0] + (int)r1[1];
    row_ptr(out, out_pitch, y)[x] = (uint8_t)((s + 2) >> 2);
}

// ---------------------------------------------------------------- fallback upsample
constexpr int FBU_ROWS = 8;      // output rows per thread: the column part (texture column, fraction, source offsets) is computed once
__global__ void __launch_bounds__(256)
fallback_upsample_kernel(const float* __restrict__ rgb, int64_t rgb_pitch, int w, int h,
                         float* __restrict__ out, int64_t out_pitch, mfsr_merge_geom g)
{
    // row part of the bilinear fetch once per block row (64 rows per block), as in flow_from_tiles_kernel; per-pixel arithmetic unchanged
    __shared__ int s_i0[8 * FBU_ROWS], s_i1[8 * FBU_ROWS];
    __shared__ float s_a[8 * FBU_ROWS];
    const float fs = (float)MFSR_SCALE_NUM(g.scale), fd = (float)MFSR_SCALE_DEN(g.scale);      // ((X + 0.5) * den) / num; den == 1: exact factor
    {
        const int t = threadIdx.y * blockDim.x + threadIdx.x;
        if (t < 8 * FBU_ROWS) {
            const int y = min((int)blockIdx.y * 8 * FBU_ROWS + t, g.out_h - 1);
            const TexAxis ty = tex_axis(__fdiv_rn(__fmul_rn((float)(y + g.org_y) + 0.5f, fd), fs), h);
            s_i0[t] = ty.i0; s_i1[t] = ty.i1; s_a[t] = ty.a;
        }
    }
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, yb = (blockIdx.y * blockDim.y + threadIdx.y) * FBU_ROWS;
    if (x >= g.out_w || yb >= g.out_h) return;
    const TexAxis tx = tex_axis(__fdiv_rn(__fmul_rn((float)(x + g.org_x) + 0.5f, fd), fs), w);
    const int o0 = 3 * tx.i0, o1 = 3 * tx.i1;
#pragma unroll 2
    for (int r = 0; r < FBU_ROWS; r++) {
        const int y = yb + r;
        if (y >= g.out_h) break;
        const int rl = threadIdx.y * FBU_ROWS + r;
        const float* r0 = row_ptr(rgb, rgb_pitch, s_i0[rl]);
        const float* r1 = row_ptr(rgb, rgb_pitch, s_i1[rl]);
        const float tya = s_a[rl];
        float* o = row_ptr(out, out_pitch, y) + 3 * x;
#pragma unroll
        for (int c = 0; c < 3; c++)
            o[c] = tex_mix(__ldg(r0 + o0 + c), __ldg(r0 + o1 + c), __ldg(r1 + o0 + c), __ldg(r1 + o1 + c), tx.a, tya);
    }
}

// host mirror of gaussin_filter_1D (main.cpp:370-391)
static int gauss_taps(float sigma, float* taps)
{
    if (sigma <= 0) { for (int i = 0; i < 9; i++) taps[i] = (i == 4) ? 1.0f : 0.0f; return 9; }
    int size = (int)(sigma / 0.6f - 0.4f) * 2 + 1 + 2;
    if (size > 99) size = 99;
    if (size > MAX_TAPS) return -1;
    const int center = size / 2;
    for (int i = 0; i < size; i++) { const int x = i - center; taps[i] = (float)(exp(-(x * x) / (2 * sigma * sigma))); }
    float sum = 0;
    for (int i = 0; i < size; i++) sum += taps[i];
    for (int i = 0; i < size; i++) taps[i] /= sum;
    return size;
}

}  // namespace mfsr

using namespace mfsr;

static Cfa mk_cfa(const int cfa[4]) { Cfa c; for (int i = 0; i < 4; i++) c.c[i] = cfa[i]; return c; }
static F3 mk_f3(const float v[3]) { F3 f; for (int i = 0; i < 3; i++) f.v[i] = v[i]; return f; }

int mfsr::launch_subsample3(const uint16_t* raw, int64_t raw_pitch, int64_t raw_fs, float* rgb_half, int64_t rgb_pitch, int64_t rgb_fs, int frames,
                      float maxVal, int dimX, int dimY, const int cfa[4], cudaStream_t stream)
{
    if (!raw || !rgb_half || !cfa || dimX <= 0 || dimY <= 0 || frames < 1 || ((raw_pitch | raw_fs) & 3) || ((uintptr_t)raw & 3)) return MFSR_E_INVALID;
    dim3 b(32, 8), g(cdiv(dimX, 32), cdiv(dimY, 8), frames);
    subsample3_kernel<<<g, b, 0, stream>>>(raw, raw_pitch, rgb_half, rgb_pitch, 1.0f / maxVal, dimX, dimY, mk_cfa(cfa), raw_fs, rgb_fs);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

extern "C" int mfsr_stage_subsample3(const uint16_t* raw, int64_t raw_pitch, float* rgb_half, int64_t rgb_pitch,
                                     float maxVal, int dimX, int dimY, const int cfa[4], void* stream)
{
    return launch_subsample3(raw, raw_pitch, 0, rgb_half, rgb_pitch, 0, 1, maxVal, dimX, dimY, cfa, (cudaStream_t)stream);
}

extern "C" int mfsr_stage_demosaic(const uint16_t* raw, int64_t raw_pitch, float* rgb, int64_t rgb_pitch,
                                   int width, int height, const int cfa[4], const float black[3], const float scale[3], void* stream)
{
    if (!raw || !rgb || !cfa || !black || !scale || width <= 0 || height <= 0) return MFSR_E_INVALID;
    dim3 b(DTW, 8), g(cdiv(width, DTW), cdiv(height, DTH));
    demosaic_kernel<<<g, b, 0, (cudaStream_t)stream>>>(raw, raw_pitch, rgb, rgb_pitch, width, height, mk_cfa(cfa), mk_f3(black), mk_f3(scale));
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

int mfsr::launch_tracking_image(const uint16_t* raw, int64_t raw_pitch, int64_t raw_fs, float* gray, int64_t gray_pitch, int64_t gray_fs,
                          uint8_t* gray_q, int64_t gray_q_pitch, int64_t gray_q_fs, int frames, int width, int height, const int cfa[4],
                          const float black[3], const float scale[3], float sigma, int track_bits, cudaStream_t stream)
{
    if (!raw || !cfa || !black || !scale || width <= 0 || height <= 0 || frames < 1 || track_bits < 1 || track_bits > 8) return MFSR_E_INVALID;
    FrameStrides fs;
    fs.s[0] = raw_fs; fs.s[1] = gray_fs; fs.s[2] = gray_q_fs;
    Taps t;
    t.n = gauss_taps(sigma, t.t);
    if (t.n < 0) return MFSR_E_INVALID;
    const float qmax = (float)((1 << track_bits) - 1);
    const Cfa c = mk_cfa(cfa); const F3 bl = mk_f3(black), sc = mk_f3(scale);
    cudaStream_t st = (cudaStream_t)stream;
    // pre-normalised raw plane: Bayer patterns (greens on one diagonal, red/blue on the other) or monochrome
    const bool mono = cfa[0] == 1 && cfa[1] == 1 && cfa[2] == 1 && cfa[3] == 1;
    const bool bayer = (cfa[1] == 1 && cfa[2] == 1 && cfa[0] != 1 && cfa[3] != 1 && cfa[0] != cfa[3]) ||
                       (cfa[0] == 1 && cfa[3] == 1 && cfa[1] != 1 && cfa[2] != 1 && cfa[1] != cfa[2]);
    const bool pn = mono || bayer;
    const int R = t.n / 2;
    if (R < 1 || R > 4) return MFSR_E_INVALID;
    if (pn && R <= 2) {
        const int pat = cfa[0] | (cfa[1] << 2) | (cfa[2] << 4) | (cfa[3] << 6);
        dim3 bq(64, 4), gq(cdiv(width, 64), cdiv(height, 32), frames);
#define MFSR_TQ(RR, PP) tracking_quad_kernel<RR, PP><<<gq, bq, 0, st>>>(raw, raw_pitch, gray, gray_pitch, gray_q, gray_q_pitch, width, height, bl, sc, t, qmax, fs)
#define MFSR_TQP(PP) if (pat == (PP)) { if (R == 1) MFSR_TQ(1, PP); else MFSR_TQ(2, PP); MFSR_LAUNCH_CHECK(); return MFSR_OK; }
        MFSR_TQP(0 | (1 << 2) | (1 << 4) | (2 << 6))      // RGGB
        MFSR_TQP(2 | (1 << 2) | (1 << 4) | (0 << 6))      // BGGR
        MFSR_TQP(1 | (0 << 2) | (2 << 4) | (1 << 6))      // GRBG
        MFSR_TQP(1 | (2 << 2) | (0 << 4) | (1 << 6))      // GBRG
        MFSR_TQP(1 | (1 << 2) | (1 << 4) | (1 << 6))      // monochrome
#undef MFSR_TQP
#undef MFSR_TQ
    }
    dim3 b(64, 4), g(cdiv(width, 64), cdiv(height, R <= 2 ? 32 : 16), frames);
#define MFSR_TRK(RR, PP) tracking_kernel<RR, PP><<<g, b, 0, st>>>(raw, raw_pitch, gray, gray_pitch, gray_q, gray_q_pitch, width, height, c, bl, sc, t, qmax, fs)
    switch (R * 2 + (pn ? 1 : 0)) {
        case 2: MFSR_TRK(1, false); break;  case 3: MFSR_TRK(1, true); break;
        case 4: MFSR_TRK(2, false); break;  case 5: MFSR_TRK(2, true); break;
        case 6: MFSR_TRK(3, false); break;  case 7: MFSR_TRK(3, true); break;
        case 8: MFSR_TRK(4, false); break;  default: MFSR_TRK(4, true); break;
    }
#undef MFSR_TRK
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

extern "C" int mfsr_stage_tracking_image(const uint16_t* raw, int64_t raw_pitch, float* gray, int64_t gray_pitch,
                                         uint8_t* gray_q, int64_t gray_q_pitch, int width, int height, const int cfa[4],
                                         const float black[3], const float scale[3], float sigma, int track_bits, void* stream)
{
    return launch_tracking_image(raw, raw_pitch, 0, gray, gray_pitch, 0, gray_q, gray_q_pitch, 0, 1, width, height, cfa, black, scale, sigma, track_bits,
                                 (cudaStream_t)stream);
}

int mfsr::launch_pyramid_down(const uint8_t* in, int64_t in_pitch, int64_t in_fs, int in_w, int in_h, uint8_t* out, int64_t out_pitch, int64_t out_fs, int frames,
                        cudaStream_t stream)
{
    if (!in || !out || in_w < 2 || in_h < 2 || frames < 1) return MFSR_E_INVALID;
    const int ow = in_w / 2, oh = in_h / 2;
    dim3 b(32, 8), g(cdiv(ow, 32), cdiv(oh, 8), frames);
    pyramid_down_kernel<<<g, b, 0, stream>>>(in, in_pitch, out, out_pitch, ow, oh, in_fs, out_fs);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

extern "C" int mfsr_stage_pyramid_down(const uint8_t* in, int64_t in_pitch, int in_w, int in_h, uint8_t* out, int64_t out_pitch, void* stream)
{
    return launch_pyramid_down(in, in_pitch, 0, in_w, in_h, out, out_pitch, 0, 1, (cudaStream_t)stream);
}

extern "C" int mfsr_stage_fallback_upsample(const float* rgb, int64_t rgb_pitch, int width, int height,
                                            float* out, int64_t out_pitch, const mfsr_merge_geom* geom, void* stream)
{
    if (!rgb || !out || !geom || geom->scale < 1) return MFSR_E_INVALID;
    dim3 b(32, 8), g(cdiv(geom->out_w, 32), cdiv(geom->out_h, 8 * FBU_ROWS));
    fallback_upsample_kernel<<<g, b, 0, (cudaStream_t)stream>>>(rgb, rgb_pitch, width, height, out, out_pitch, *geom);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}
