// align.cu — Gaussian-pyramid tile block matching and per-tile shift consolidation.
//
// Compiled with -fmad=false (discrete decisions hang on these fp32 values).
//
// tile_align_kernel replaces, in one launch and with zero HBM intermediates,
//   convertToTilesOverlapBorder / ...PreShift (kernel.cu:265,324)  -> u8 patches gathered to smem
//   cuFFT R2C x2, conjugateComplexMulKernel (:485), cuFFT C2R        -> direct integer correlation (dp4a)
//   squaredSum (:119), boxFilterWithBorderX/Y (:149,:186)            -> integer sums of squares (dp4a)
//   normalizedCC (:227)                                              -> SSD = sumT2 + sumI2 - 2 CC, exact in int32
//   findMinimum (:512)                                               -> warp (value,index) arg-min + serial 3x3 fit
// The reference's tile stacks alone are 110 MB each at 12 MP; here the
// only HBM traffic is the u8 images (1 B/px) and 8 B per tile of output.
//
// Exactness: with track_bits <= 7 and T <= 16 every term of the reference's fp32 formula
// (sumT2, box(I2), CC <= 256*127^2 = 4.13e6 and their combination <= 8.3e6 < 2^24) is an
// exactly representable integer, so float(int SSD) == the reference's fp32 SSD bit for
// bit, and so is the arg-min (ties -> lowest linear index, kernel.cu:536).
#include "common.cuh"
#include "internal.h"

namespace mfsr {

__constant__ float c_FA11[9] = {1.0f / 4.0f, -2.0f / 4.0f, 1.0f / 4.0f, 2.0f / 4.0f, -4.0f / 4.0f, 2.0f / 4.0f, 1.0f / 4.0f, -2.0f / 4.0f, 1.0f / 4.0f};
__constant__ float c_FA22[9] = {1.0f / 4.0f, 2.0f / 4.0f, 1.0f / 4.0f, -2.0f / 4.0f, -4.0f / 4.0f, -2.0f / 4.0f, 1.0f / 4.0f, 2.0f / 4.0f, 1.0f / 4.0f};
__constant__ float c_FA12[9] = {1.0f / 4.0f, 0.0f, -1.0f / 4.0f, 0.0f, 0.0f, 0.0f, -1.0f / 4.0f, 0.0f, 1.0f / 4.0f};
__constant__ float c_Fb1[9] = {-1.0f / 8.0f, 0.0f, 1.0f / 8.0f, -2.0f / 8.0f, 0.0f, 2.0f / 8.0f, -1.0f / 8.0f, 0.0f, 1.0f / 8.0f};
__constant__ float c_Fb2[9] = {-1.0f / 8.0f, -2.0f / 8.0f, -1.0f / 8.0f, 0.0f, 0.0f, 0.0f, 1.0f / 8.0f, 2.0f / 8.0f, 1.0f / 8.0f};

struct TileAlignArgs {
    TileAlignBatch b;
    float sf, cf;
};

// integer displacement of a tile (kernel.cu:299-311 / :358-370).
// Restated-host decision (round 2): the base shift / rotation is the pose of the MOVED image against the reference image, so
// only the moved patches are displaced by it (convertToTilesOverlapPreShift, kernel.cu:324); the reference tiles
// (convertToTilesOverlapBorder, :265) are taken with an identity base.  The reported tile shift is the RESIDUAL against the base
// pose — coord + round(base + pre) - round(base) — which is what CreateFlowFieldFromTiles (opticalFlow.cu:72-86) adds to the
// base displacement field.  (Round 1 displaced both patches, which cancels a base shift; with a zero base both forms coincide.)
__device__ __forceinline__ void tile_disp(const TileAlignArgs& A, int pair, int tix, int tiy, float prex, float prey, int& dx, int& dy)
{
    float bsx = A.b.bsx, bsy = A.b.bsy, cf = A.cf, sf = A.sf;
    if (A.b.pair_pose) {               // per-pair global pre-alignment (prealign.cu), shifts in full-resolution pixels
        const float4 pp = *(const float4*)(A.b.pair_pose + 4 * pair);
        bsx = pp.x * A.b.pose_scale; bsy = pp.y * A.b.pose_scale; cf = pp.z; sf = pp.w;
    }
    float sx = prex, sy = prey;
    sx += cf * -bsx - sf * -bsy;
    sy += sf * -bsx + cf * -bsy;
    const float pcx = (float)(tix * A.b.T + A.b.T / 2 - A.b.w / 2);
    const float pcy = (float)(tiy * A.b.T + A.b.T / 2 - A.b.h / 2);
    sx += cf * pcx - sf * pcy - pcx;
    sy += sf * pcx + cf * pcy - pcy;
    dx = (int)roundf(sx); dy = (int)roundf(sy);
}

// Generic tile sizes: one block per (tile, pair).  Dynamic smem: mov patch P rows x PW words, ref patch T rows x T/4 words, S*S floats.
__global__ void __launch_bounds__(96)
tile_align_generic_kernel(const __grid_constant__ TileAlignArgs AA)
{
    extern __shared__ uint32_t smem[];
    const TileAlignArgs& A = AA;
    const TileAlignBatch& B = A.b;
    const int pair = blockIdx.y;
    const uint8_t* img_ref = B.img + B.frame_stride * (int64_t)B.pt.from[pair];
    const uint8_t* img_mov = B.img + B.frame_stride * (int64_t)B.pt.to[pair];
    const float2* pre = B.pre ? (const float2*)((const char*)B.pre + B.pre_pair_stride * pair) : nullptr;
    float2* outp = (float2*)((char*)B.out + B.out_pair_stride * pair);
    int2* argminp = B.argmin ? (int2*)((char*)B.argmin + B.argmin_pair_stride * pair) : nullptr;
    float* ssdp = B.ssd ? (float*)((char*)B.ssd + B.ssd_pair_stride * pair) : nullptr;
    const int T = B.T, M = B.M, P = T + 2 * M, S = 2 * M + 1, nlag = S * S;
    const int PW = (P + 3) / 4 + 1;                 // words per mov row (+1: unaligned window reads one word past)
    const int TW = T / 4;
    uint32_t* s_mov = smem;                         // [P][PW]
    uint32_t* s_ref = s_mov + P * PW;               // [T][TW]
    float* s_ssd = (float*)(s_ref + T * TW);        // [nlag]
    __shared__ unsigned s_sumT2;
    const int t = blockIdx.x;
    const int tiy = t / B.tx, tix = t - tiy * B.tx;
    const int tid = threadIdx.x, nthr = blockDim.x;

    float prex = 0.f, prey = 0.f;
    if (pre) { const float2 p = row_ptr(pre, B.pre_pitch, tiy)[tix]; prex = p.x; prey = p.y; }
    int dxm, dym, dxr, dyr;
    tile_disp(A, pair, tix, tiy, prex, prey, dxm, dym);
    tile_disp(A, pair, tix, tiy, 0.f, 0.f, dxr, dyr);

    // gather patches (clamped like kernel.cu:312-313 / :371-372), pack 4 px per word
    for (int i = tid; i < P * PW; i += nthr) {
        const int py = i / PW, pw = i - py * PW;
        const int gy = clampi(tiy * T + py + dym, 0, B.h - 1);
        const uint8_t* row = row_ptr(img_mov, B.pitch, gy);
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int px = pw * 4 + b;
            const int gx = clampi(tix * T + px + dxm, 0, B.w - 1);
            v |= (px < P ? (uint32_t)row[gx] : 0u) << (8 * b);
        }
        s_mov[i] = v;
    }
    for (int i = tid; i < T * TW; i += nthr) {
        const int py = i / TW, pw = i - py * TW;
        const int gy = clampi(tiy * T + M + py, 0, B.h - 1);           // reference tiles are taken undisplaced (see tile_disp)
        const uint8_t* row = row_ptr(img_ref, B.pitch, gy);
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int gx = clampi(tix * T + M + pw * 4 + b, 0, B.w - 1);
            v |= (uint32_t)row[gx] << (8 * b);
        }
        s_ref[i] = v;
    }
    if (tid == 0) s_sumT2 = 0u;
    __syncthreads();
    // sum of squares of the template (squaredSum, kernel.cu:119)
    {
        unsigned part = 0;
        for (int i = tid; i < T * TW; i += nthr) part = __dp4a(s_ref[i], s_ref[i], part);
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((tid & 31) == 0) atomicAdd(&s_sumT2, part);
    }
    __syncthreads();
    const unsigned sumT2 = s_sumT2;
    // one lag per thread: SSD(s) = sumT2 + sum_window I^2 - 2 sum T*I
    for (int lag = tid; lag < nlag; lag += nthr) {
        const int ly = lag / S, lx = lag - ly * S;   // shift = (lx - M, ly - M); window origin (lx, ly) in the mov patch
        const int w0 = lx >> 2, sh = (lx & 3) * 8;
        unsigned cc = 0, i2 = 0;
        for (int y = 0; y < T; y++) {
            const uint32_t* mr = s_mov + (y + ly) * PW + w0;
            const uint32_t* rr = s_ref + y * TW;
            uint32_t lo = mr[0];
            for (int xw = 0; xw < TW; xw++) {
                const uint32_t hi = mr[xw + 1];
                const uint32_t iv = __funnelshift_r(lo, hi, sh);
                cc = __dp4a(rr[xw], iv, cc);
                i2 = __dp4a(iv, iv, i2);
                lo = hi;
            }
        }
        const float v = (float)((int)(sumT2 + i2) - 2 * (int)cc);
        s_ssd[lag] = v;
        if (ssdp) ssdp[(size_t)t * nlag + lag] = v;
    }
    __syncthreads();
    // findMinimum (kernel.cu:512-636): warp 0, (value, index) arg-min with strict '<' => lowest index wins ties
    if (tid < 32) {
        float minVal = FLT_MAX, maxVal = -FLT_MAX; int minIdx = -1;
        for (int i = tid; i < nlag; i += 32) {
            const float v = s_ssd[i];
            maxVal = fmaxf(maxVal, v);
            if (v < minVal) { minVal = v; minIdx = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, minVal, o);
            const int oi = __shfl_xor_sync(0xffffffffu, minIdx, o);
            const float om = __shfl_xor_sync(0xffffffffu, maxVal, o);
            maxVal = fmaxf(maxVal, om);
            if (oi >= 0 && (ov < minVal || (ov == minVal && (minIdx < 0 || oi < minIdx)))) { minVal = ov; minIdx = oi; }
        }
        if (tid == 0) {
            float cy = (float)(minIdx / S);
            float cx = (float)minIdx - cy * (float)S;
            if (argminp) argminp[t] = make_int2((int)cx - M, (int)cy - M);
            if (cx < 1 || cy < 1 || cx >= 2 * M || cy >= 2 * M) { cx = 0; cy = 0; }
            else {
                float A11 = 0, A22 = 0, A12 = 0, b1 = 0, b2 = 0;
#pragma unroll
                for (int i = 0; i < 9; i++) {
                    const int off = (i < 3) ? (i - 1 - S) : (i < 6 ? i - 4 : i - 7 + S);
                    const float g = s_ssd[minIdx + off];
                    A11 += c_FA11[i] * g; A22 += c_FA22[i] * g; A12 += c_FA12[i] * g; b1 += c_Fb1[i] * g; b2 += c_Fb2[i] * g;
                }
                A11 = fmaxf(A11, 0.0f); A22 = fmaxf(A22, 0.0f);
                float det = A11 * A22 - A12 * A12;
                if (det < 0) { A12 = 0; det = A11 * A22; }
                if (det != 0) {
                    float muX = (A22 * b1 - A12 * b2) / det;
                    float muY = (A11 * b2 - A12 * b1) / det;
                    if (fabsf(muX) > 1) muX = 0;
                    if (fabsf(muY) > 1) muY = 0;
                    cx -= muX; cy -= muY;
                }
                cx -= M; cy -= M;
            }
            if (B.threshold + minVal > maxVal) { cx = 0; cy = 0; }
            row_ptr(outp, B.out_pitch, tiy)[tix] = make_float2(cx + (float)(dxm - dxr), cy + (float)(dym - dyr));
        }
    }
}

// findMinimum (kernel.cu:512-636) by one warp on an S x S SSD map in shared memory: (value, index) arg-min with strict
// '<' so that the lowest linear index wins ties, 3x3 quadratic sub-pixel fit by lane 0.  Writes the tile's shift.
__device__ __forceinline__ void find_minimum_warp(const float* s_ssd, int S, int M, int lane, float threshold, int2* argminp, int t,
                                                  float2* outp, int64_t out_pitch, int tix, int tiy, float addx, float addy)
{
    const int nlag = S * S;
    float minVal = FLT_MAX, maxVal = -FLT_MAX; int minIdx = -1;
    for (int i = lane; i < nlag; i += 32) {
        const float v = s_ssd[i];
        maxVal = fmaxf(maxVal, v);
        if (v < minVal) { minVal = v; minIdx = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, minVal, o);
        const int oi = __shfl_xor_sync(0xffffffffu, minIdx, o);
        const float om = __shfl_xor_sync(0xffffffffu, maxVal, o);
        maxVal = fmaxf(maxVal, om);
        if (oi >= 0 && (ov < minVal || (ov == minVal && (minIdx < 0 || oi < minIdx)))) { minVal = ov; minIdx = oi; }
    }
    if (lane == 0) {
        float cy = (float)(minIdx / S);
        float cx = (float)minIdx - cy * (float)S;
        if (argminp) argminp[t] = make_int2((int)cx - M, (int)cy - M);
        if (cx < 1 || cy < 1 || cx >= 2 * M || cy >= 2 * M) { cx = 0; cy = 0; }
        else {
            float A11 = 0, A22 = 0, A12 = 0, b1 = 0, b2 = 0;
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const int off = (i < 3) ? (i - 1 - S) : (i < 6 ? i - 4 : i - 7 + S);
                const float g = s_ssd[minIdx + off];
                A11 += c_FA11[i] * g; A22 += c_FA22[i] * g; A12 += c_FA12[i] * g; b1 += c_Fb1[i] * g; b2 += c_Fb2[i] * g;
            }
            A11 = fmaxf(A11, 0.0f); A22 = fmaxf(A22, 0.0f);
            float det = A11 * A22 - A12 * A12;
            if (det < 0) { A12 = 0; det = A11 * A22; }
            if (det != 0) {
                float muX = (A22 * b1 - A12 * b2) / det;
                float muY = (A11 * b2 - A12 * b1) / det;
                if (fabsf(muX) > 1) muX = 0;
                if (fabsf(muY) > 1) muY = 0;
                cx -= muX; cy -= muY;
            }
            cx -= M; cy -= M;
        }
        if (threshold + minVal > maxVal) { cx = 0; cy = 0; }
        row_ptr(outp, out_pitch, tiy)[tix] = make_float2(cx + addx, cy + addy);
    }
}

// Correlation of one lane's group of up to 4 horizontally adjacent lags (lx = 4g .. 4g+3, one ly) for T = 16.
// AL = byte alignment of the moved patch inside its aligned words (warp-uniform): window word offsets and funnel
// shifts are compile-time, the 6 moved words of a row serve all 4 lags.
template <int AL>
__device__ __forceinline__ void cc_group16(const uint32_t* __restrict__ mov, int pwa, const uint32_t* __restrict__ ref, unsigned (&cc)[4])
{
#pragma unroll 4
    for (int y = 0; y < 16; y++) {
        uint32_t W[6], R[4];
#pragma unroll
        for (int k = 0; k < 6; k++) W[k] = mov[y * pwa + k];
#pragma unroll
        for (int k = 0; k < 4; k++) R[k] = ref[y * 4 + k];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int o = (AL + j) >> 2, sh = ((AL + j) & 3) * 8;
#pragma unroll
            for (int xw = 0; xw < 4; xw++) {
                const uint32_t iv = sh ? __funnelshift_r(W[xw + o], W[xw + o + 1], sh) : W[xw + o];
                cc[j] = __dp4a(R[xw], iv, cc[j]);
            }
        }
    }
}

// T = 16 fast path: ONE WARP per (tile, pair), four per block, no block barriers.
//  * patches are fetched as aligned 32-bit words (the byte-gather version issued ~930 u8 loads per tile and pair),
//    the moved patch keeps its alignment and is funnel-shifted at use;
//  * the window energy sum I^2 comes from row sums + a column prefix (exact integers, any order) instead of a second
//    dp4a per lag and word;
//  * a lane computes 4 adjacent lags from one set of loaded words.
// MT: compile-time search radius (0: read it from the batch).  With M known every i / PWA, i / S, it / NG of the staging loops is
// a multiply-shift instead of a ~20-instruction integer division: the correlation loop was only a third of the 2200 instructions
// per tile and pair (gpurun_out/r1r_misc.ncu-rep).
template <int MT>
__global__ void __launch_bounds__(128)
tile_align16_kernel(const __grid_constant__ TileAlignArgs AA)
{
    extern __shared__ uint32_t smem[];
    const TileAlignArgs& A = AA;
    const TileAlignBatch& B = A.b;
    constexpr int T = 16, TW = 4;
    const int M = MT ? MT : B.M, P = T + 2 * M, S = 2 * M + 1, nlag = S * S;
    const int PWA = (P + 6) / 4 + 1;
    const int per_warp = P * PWA + T * TW + (P + 1) * S + nlag;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 4 + warp, pair = blockIdx.y;
    if (t >= B.tx * B.ty) return;
    uint32_t* s_mov = smem + warp * per_warp;             // [P][PWA] aligned words
    uint32_t* s_ref = s_mov + P * PWA;                    // [T][TW] exact patch words
    int* s_cp = (int*)(s_ref + T * TW);                   // [P + 1][S] column prefix of the row window energies
    float* s_ssd = (float*)(s_cp + (P + 1) * S);          // [nlag]
    const uint8_t* img_ref = B.img + B.frame_stride * (int64_t)B.pt.from[pair];
    const uint8_t* img_mov = B.img + B.frame_stride * (int64_t)B.pt.to[pair];
    const float2* pre = B.pre ? (const float2*)((const char*)B.pre + B.pre_pair_stride * pair) : nullptr;
    float2* outp = (float2*)((char*)B.out + B.out_pair_stride * pair);
    int2* argminp = B.argmin ? (int2*)((char*)B.argmin + B.argmin_pair_stride * pair) : nullptr;
    float* ssdp = B.ssd ? (float*)((char*)B.ssd + B.ssd_pair_stride * pair) : nullptr;
    const int tiy = t / B.tx, tix = t - tiy * B.tx;

    float prex = 0.f, prey = 0.f;
    if (pre) { const float2 p = row_ptr(pre, B.pre_pitch, tiy)[tix]; prex = p.x; prey = p.y; }
    int dxm, dym, dxr, dyr;
    tile_disp(A, pair, tix, tiy, prex, prey, dxm, dym);
    tile_disp(A, pair, tix, tiy, 0.f, 0.f, dxr, dyr);

    // ---- moved patch (clamped like kernel.cu:371-372)
    const int gxs = tix * T + dxm, gys = tiy * T + dym;
    const bool mfast = gxs >= 0 && gxs + P <= B.w && (int64_t)((gxs & ~3) + 4 * PWA) <= B.pitch;
    const int al = mfast ? (gxs & 3) : 0;
    if (mfast) {
        const int xa = gxs & ~3;
        for (int i = lane; i < P * PWA; i += 32) {
            const int py = i / PWA, k = i - py * PWA;
            s_mov[i] = __ldg((const uint32_t*)(row_ptr(img_mov, B.pitch, clampi(gys + py, 0, B.h - 1)) + xa) + k);
        }
    } else {
        for (int i = lane; i < P * PWA; i += 32) {
            const int py = i / PWA, k = i - py * PWA;
            const uint8_t* row = row_ptr(img_mov, B.pitch, clampi(gys + py, 0, B.h - 1));
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int px = k * 4 + b;
                v |= (px < P ? (uint32_t)row[clampi(gxs + px, 0, B.w - 1)] : 0u) << (8 * b);
            }
            s_mov[i] = v;
        }
    }
    // ---- template patch (kernel.cu:312-313), exact alignment
    const int gxr = tix * T + M, gyr = tiy * T + M;                        // reference tiles are taken undisplaced (see tile_disp)
    const bool rfast = gxr >= 0 && gxr + T <= B.w && (int64_t)((gxr & ~3) + T + 4) <= B.pitch;
    for (int i = lane; i < T * TW; i += 32) {
        const int py = i >> 2, xw = i & 3;
        const uint8_t* row = row_ptr(img_ref, B.pitch, clampi(gyr + py, 0, B.h - 1));
        uint32_t v;
        if (rfast) {
            const uint32_t* wp = (const uint32_t*)(row + (gxr & ~3)) + xw;
            const int sh = (gxr & 3) * 8;
            const uint32_t w0 = __ldg(wp);
            v = sh ? __funnelshift_r(w0, __ldg(wp + 1), sh) : w0;
        } else {
            v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) v |= (uint32_t)row[clampi(gxr + xw * 4 + b, 0, B.w - 1)] << (8 * b);
        }
        s_ref[i] = v;
    }
    __syncwarp();
    // ---- sum of squares of the template (squaredSum, kernel.cu:119)
    unsigned sumT2 = 0;
    for (int i = lane; i < T * TW; i += 32) sumT2 = __dp4a(s_ref[i], s_ref[i], sumT2);
    for (int o = 16; o > 0; o >>= 1) sumT2 += __shfl_xor_sync(0xffffffffu, sumT2, o);
    // ---- window energies (boxFilterWithBorderX/Y of I^2, kernel.cu:149,:186): row sums, then column prefix.
    // Compile-time radius with P <= 32: lane y owns row y — the window at lag 0 is four dp4a, every further lag slides it by one
    // sample (- leaving^2 + entering^2; exact integers, any order), instead of four funnel shifts + four dp4a per (row, lag).
    if (MT && P <= 32) {
        if (lane < P) {
            const uint32_t* mr = s_mov + lane * PWA;
            uint32_t V[(T + 2 * (MT ? MT : 1)) / 4 + 1];                       // the row's bytes al .. al + P + 3, realigned
#pragma unroll
            for (int k = 0; k < (T + 2 * (MT ? MT : 1)) / 4 + 1; k++) V[k] = al ? __funnelshift_r(mr[k], mr[k + 1], 8 * al) : mr[k];
            unsigned e = 0;
#pragma unroll
            for (int xw = 0; xw < TW; xw++) e = __dp4a(V[xw], V[xw], e);
            s_cp[(lane + 1) * S] = (int)e;
#pragma unroll
            for (int lx = 1; lx < 2 * (MT ? MT : 1) + 1; lx++) {
                const unsigned out = (V[(lx - 1) >> 2] >> (8 * ((lx - 1) & 3))) & 0xffu, in = (V[(lx + 15) >> 2] >> (8 * ((lx + 15) & 3))) & 0xffu;
                e = e - out * out + in * in;
                s_cp[(lane + 1) * S + lx] = (int)e;
            }
        }
    } else
    for (int i = lane; i < P * S; i += 32) {
        const int y = i / S, lx = i - y * S;
        const int b = al + lx, wq = b >> 2, sh = (b & 3) * 8;
        const uint32_t* mr = s_mov + y * PWA + wq;
        unsigned e = 0;
#pragma unroll
        for (int xw = 0; xw < TW; xw++) {
            const uint32_t iv = __funnelshift_r(mr[xw], mr[xw + 1], sh);
            e = __dp4a(iv, iv, e);
        }
        s_cp[(y + 1) * S + lx] = (int)e;
    }
    __syncwarp();
    if (lane < S) {
        int run = 0;
        s_cp[lane] = 0;
        for (int y = 1; y <= P; y++) { run += s_cp[y * S + lane]; s_cp[y * S + lane] = run; }
    }
    __syncwarp();
    // ---- correlation + SSD map (normalizedCC, kernel.cu:227): SSD = sumT2 + sumI2 - 2 CC, exact in int32
    const int NG = (S + 3) / 4;
    for (int it = lane; it < S * NG; it += 32) {
        const int ly = it / NG, g = it - ly * NG;
        unsigned cc[4] = {0u, 0u, 0u, 0u};
        const uint32_t* mv = s_mov + ly * PWA + g;
        switch (al) {
            case 0: cc_group16<0>(mv, PWA, s_ref, cc); break;
            case 1: cc_group16<1>(mv, PWA, s_ref, cc); break;
            case 2: cc_group16<2>(mv, PWA, s_ref, cc); break;
            default: cc_group16<3>(mv, PWA, s_ref, cc); break;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int lx = 4 * g + j;
            if (lx < S) {
                const int i2 = s_cp[(ly + T) * S + lx] - s_cp[ly * S + lx];
                const float v = (float)((int)(sumT2 + (unsigned)i2) - 2 * (int)cc[j]);
                s_ssd[ly * S + lx] = v;
                if (ssdp) ssdp[(size_t)t * nlag + ly * S + lx] = v;
            }
        }
    }
    __syncwarp();
    find_minimum_warp(s_ssd, S, M, lane, B.threshold, argminp, t, outp, B.out_pitch, tix, tiy, (float)(dxm - dxr), (float)(dym - dyr));
}

// UpSampleShifts (kernel.cu:642-688)
__global__ void __launch_bounds__(256)
upsample_shifts_kernel(const __grid_constant__ UpsampleBatch U)
{
    const int nx = blockIdx.x * blockDim.x + threadIdx.x, ny = blockIdx.y * blockDim.y + threadIdx.y;
    const int oldLevel = U.oldLevel, newLevel = U.newLevel, oldCX = U.oldCX, oldCY = U.oldCY, newCX = U.newCX, newCY = U.newCY;
    const int oldT = U.oldT, newT = U.newT;
    const int64_t in_pitch = U.in_pitch, out_pitch = U.out_pitch;
    const float2* in = (const float2*)((const char*)U.in + U.in_pair_stride * blockIdx.z);
    float2* out = (float2*)((char*)U.out + U.out_pair_stride * blockIdx.z);
    if (nx >= newCX || ny >= newCY) return;
    const float factor = (float)oldLevel * oldT / (float)(newLevel * newT);
    const float oldX = nx / factor, oldY = ny / factor;
    int xmin = (int)floorf(oldX), xmax = (int)ceilf(oldX), ymin = (int)floorf(oldY), ymax = (int)ceilf(oldY);
    xmin = min(xmin, oldCX - 1); xmax = min(xmax, oldCX - 1); ymin = min(ymin, oldCY - 1); ymax = min(ymax, oldCY - 1);
    const float2 mm = row_ptr(in, in_pitch, ymin)[xmin], Mm = row_ptr(in, in_pitch, ymin)[xmax];
    const float2 mM = row_ptr(in, in_pitch, ymax)[xmin], MM = row_ptr(in, in_pitch, ymax)[xmax];
    const float fx = 1.0f - (xmax - oldX), fy = 1.0f - (ymax - oldY);
    float t1 = mm.x + (Mm.x - mm.x) * fx, t2 = mM.x + (MM.x - mM.x) * fx;
    float2 o;
    o.x = t1 + (t2 - t1) * fy;
    t1 = mm.y + (Mm.y - mm.y) * fx; t2 = mM.y + (MM.y - mM.y) * fx;
    o.y = t1 + (t2 - t1) * fy;
    o.x *= oldLevel / (float)newLevel;
    o.y *= oldLevel / (float)newLevel;
    row_ptr(out, out_pitch, ny)[nx] = o;
}

// ------------------------------------------------------------------ consolidation
// One warp per tile.  The design matrix A (m x n1, ShiftMinimizerKernels.cu:137 layout) is never
// materialised: row k is the indicator of [from_k, to_k) and a 64-bit mask holds the rows
// still active.  AtA / inverse live in shared memory; Gauss-Jordan with partial pivoting,
// lane r owns row r.  Replaces concatenateShifts/copyShiftMatrix/setPointers/transposeShifts/
// checkForOutliers/getOptimalShifts/separateShifts + 4 cuBLAS batched calls per sweep.

// AtA of the active measurement rows and its inverse by Gauss-Jordan with partial pivoting (one warp, lane r owns row r).
// Returns false when a pivot vanishes.  Shared by the per-tile kernel and by the one-warp kernel that inverts the FULL system once
// per burst: with every pair active the matrix is the same for all tiles, so tiles without outliers (most) only load that inverse.
__device__ __forceinline__ bool ata_inverse(float* a, float* inv, int n1, int m, unsigned long long active, const PairTable& pt, int lane)
{
    // AtA (small exact integers) and identity
    for (int e = lane; e < n1 * n1; e += 32) {
        const int i = e / n1, j = e - i * n1, lo = min(i, j), hi = max(i, j);
        int cnt = 0;
        for (int k = 0; k < m; k++) cnt += ((active >> k) & 1ull) && pt.from[k] <= lo && hi < pt.to[k];
        a[e] = (float)cnt; inv[e] = (i == j) ? 1.0f : 0.0f;
    }
    __syncwarp();
    bool singular = false;
    for (int c = 0; c < n1; c++) {
        // partial pivot: first row with the largest |a[r][c]|, r >= c
        int piv = c; float best = fabsf(a[c * n1 + c]);
        for (int r = c + 1; r < n1; r++) { const float v = fabsf(a[r * n1 + c]); if (v > best) { best = v; piv = r; } }
        if (best < 1e-6f) { singular = true; break; }
        if (piv != c) {
            for (int j = lane; j < n1; j += 32) {
                float tmp = a[c * n1 + j]; a[c * n1 + j] = a[piv * n1 + j]; a[piv * n1 + j] = tmp;
                tmp = inv[c * n1 + j]; inv[c * n1 + j] = inv[piv * n1 + j]; inv[piv * n1 + j] = tmp;
            }
            __syncwarp();
        }
        const float d = 1.0f / a[c * n1 + c];
        __syncwarp();
        for (int j = lane; j < n1; j += 32) { a[c * n1 + j] *= d; inv[c * n1 + j] *= d; }
        __syncwarp();
        if (lane < n1 && lane != c) {
            const int r = lane;
            const float f = a[r * n1 + c];
            if (f != 0.0f)
                for (int j = 0; j < n1; j++) { a[r * n1 + j] -= f * a[c * n1 + j]; inv[r * n1 + j] -= f * inv[c * n1 + j]; }
        }
        __syncwarp();
    }
    return !singular;
}

__global__ void __launch_bounds__(32)
consolidate_inverse_kernel(PairTable pt, int m, int imageCount, float* __restrict__ inv0)
{
    extern __shared__ float cs0[];
    const int n1 = imageCount - 1, lane = threadIdx.x;
    float* a = cs0;
    float* inv = cs0 + n1 * n1;
    const unsigned long long active = (m >= 64) ? ~0ull : ((1ull << m) - 1ull);
    const bool ok = ata_inverse(a, inv, n1, m, active, pt, lane);
    __syncwarp();
    for (int e = lane; e < n1 * n1; e += 32) inv0[e] = inv[e];
    if (lane == 0) inv0[n1 * n1] = ok ? 1.0f : 0.0f;
}

__global__ void __launch_bounds__(128)
consolidate_kernel(const float2* __restrict__ measured, int64_t tile_stride, int64_t pair_stride, PairTable pt, int m, int imageCount, int nTiles, int referenceImage,
                   float2* __restrict__ one_to_one, float2* __restrict__ frame_shift, int* __restrict__ status, const float* __restrict__ inv0,
                   const unsigned long long* __restrict__ active0)
{
    extern __shared__ float cs[];
    const int n1 = imageCount - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + warp;
    const int per_warp = 2 * n1 * n1 + 2 * m + 4 * n1;
    float* a = cs + warp * per_warp;          // [n1][n1]
    float* inv = a + n1 * n1;                 // [n1][n1]
    float* b = inv + n1 * n1;                 // [m][2]
    float* atb = b + 2 * m;                   // [n1][2]
    float* x = atb + 2 * n1;                  // [n1][2]
    if (t >= nTiles) return;
    for (int k = lane; k < m; k += 32) { const float2 v = measured[(int64_t)t * tile_stride + (int64_t)k * pair_stride]; b[2 * k] = v.x; b[2 * k + 1] = v.y; }
    const unsigned long long full = (m >= 64) ? ~0ull : ((1ull << m) - 1ull);
    // pairs the global pre-alignment ruled out (relative rotation too large for translation-only tile matching, prealign.cu) start removed
    unsigned long long active = active0 ? (*active0 & full) : full;
    __syncwarp();
    if (active != full)
        for (int k = lane; k < m; k += 32) if (!((active >> k) & 1ull)) { b[2 * k] = 0.f; b[2 * k + 1] = 0.f; }
    int removed = 0, st = 0;
    __syncwarp();
    for (;;) {
        bool singular;
        if (removed == 0 && active == full && inv0 && inv0[n1 * n1] != 0.0f) {       // the full system: inverted once per burst
            for (int e = lane; e < n1 * n1; e += 32) inv[e] = inv0[e];
            singular = false;
            __syncwarp();
        } else {
            singular = !ata_inverse(a, inv, n1, m, active, pt, lane);
        }
        if (singular) { for (int i = lane; i < 2 * n1; i += 32) x[i] = 0.0f; st = -1; __syncwarp(); break; }
        // Atb (ascending k), x = inv * Atb (ascending j)
        if (lane < n1) {
            float sx = 0.f, sy = 0.f;
            for (int k = 0; k < m; k++)
                if (((active >> k) & 1ull) && pt.from[k] <= lane && lane < pt.to[k]) { sx += b[2 * k]; sy += b[2 * k + 1]; }
            atb[2 * lane] = sx; atb[2 * lane + 1] = sy;
        }
        __syncwarp();
        if (lane < n1) {
            float sx = 0.f, sy = 0.f;
            for (int j = 0; j < n1; j++) { sx += inv[lane * n1 + j] * atb[2 * j]; sy += inv[lane * n1 + j] * atb[2 * j + 1]; }
            x[2 * lane] = sx; x[2 * lane + 1] = sy;
        }
        __syncwarp();
        // checkForOutliers (ShiftMinimizerKernels.cu:81-139): largest residual above 1 px^2
        float mx = 1.0f; int idx = -1;
        for (int k = 0; k < m; k++) {
            float ox = 0.f, oy = 0.f;
            if ((active >> k) & 1ull) for (int c = pt.from[k]; c < pt.to[k]; c++) { ox += x[2 * c]; oy += x[2 * c + 1]; }
            const float dx = b[2 * k] - ox, dy = b[2 * k + 1] - oy, dd = dx * dx + dy * dy;
            if (dd > mx) { idx = k; mx = dd; }
        }
        if (idx == -1) break;
        __syncwarp();
        if (lane == 0) { b[2 * idx] = 0.f; b[2 * idx + 1] = 0.f; }
        active &= ~(1ull << idx);
        removed++;
        __syncwarp();
    }
    if (status && lane == 0) status[t] = st < 0 ? -1 : removed;
    if (one_to_one) for (int i = lane; i < n1; i += 32) one_to_one[(size_t)t * n1 + i] = make_float2(x[2 * i], x[2 * i + 1]);
    // getOptimalShifts (:179-218) for every frame
    for (int f = lane; f < imageCount; f += 32) {
        float tsx = 0.f, tsy = 0.f;
        if (referenceImage < f) for (int i = referenceImage; i < f; i++) { tsx += x[2 * i]; tsy += x[2 * i + 1]; }
        else if (f < referenceImage) for (int i = f; i < referenceImage; i++) { tsx -= x[2 * i]; tsy -= x[2 * i + 1]; }
        frame_shift[(size_t)f * nTiles + t] = make_float2(tsx, tsy);
    }
}

}  // namespace mfsr

using namespace mfsr;

int mfsr::launch_tile_align(const TileAlignBatch& b, cudaStream_t st)
{
    if (!b.img || !b.out || b.T < 4 || (b.T & 3) || b.M < 1 || b.tx < 1 || b.ty < 1 || b.n_pairs < 1 || b.n_pairs > CONS_MAX_M) return MFSR_E_INVALID;
    if (b.tx * b.T + 2 * b.M > b.w || b.ty * b.T + 2 * b.M > b.h) return MFSR_E_INVALID;
    // int32 headroom of sumT2 + sumI2 (u8 samples)
    if ((int64_t)b.T * b.T * 255 * 255 * 2 >= (1ll << 31)) return MFSR_E_INVALID;
    TileAlignArgs A;
    A.b = b; A.sf = sinf(b.rot); A.cf = cosf(b.rot);
    const int P = b.T + 2 * b.M, S = 2 * b.M + 1, PW = (P + 3) / 4 + 1;
    // T = 16 with word-aligned images: warp-per-tile kernel
    if (b.T == 16 && !(((uintptr_t)b.img | (uintptr_t)b.pitch | (uintptr_t)b.frame_stride) & 3)) {
        const int PWA = (P + 6) / 4 + 1;
        const size_t smem16 = (size_t)4 * (P * PWA + 16 * 4 + (P + 1) * S + S * S) * 4;
        if (smem16 <= 48 * 1024) {
            const dim3 g16(cdiv(b.tx * b.ty, 4), b.n_pairs);
            switch (b.M) {
                case 2: tile_align16_kernel<2><<<g16, 128, smem16, st>>>(A); break;
                case 4: tile_align16_kernel<4><<<g16, 128, smem16, st>>>(A); break;
                case 8: tile_align16_kernel<8><<<g16, 128, smem16, st>>>(A); break;
                default: tile_align16_kernel<0><<<g16, 128, smem16, st>>>(A); break;
            }
            MFSR_LAUNCH_CHECK();
            return MFSR_OK;
        }
    }
    const size_t smem = (size_t)(P * PW + b.T * (b.T / 4)) * 4 + (size_t)S * S * 4;
    if (smem > 48 * 1024) return MFSR_E_INVALID;
    tile_align_generic_kernel<<<dim3(b.tx * b.ty, b.n_pairs), 96, smem, st>>>(A);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

int mfsr::launch_upsample_shifts(const UpsampleBatch& u, cudaStream_t st)
{
    if (!u.in || !u.out || u.oldLevel < 1 || u.newLevel < 1 || u.oldCX < 1 || u.oldCY < 1 || u.newCX < 1 || u.newCY < 1 || u.n_pairs < 1) return MFSR_E_INVALID;
    dim3 b(32, 8), g(cdiv(u.newCX, 32), cdiv(u.newCY, 8), u.n_pairs);
    upsample_shifts_kernel<<<g, b, 0, st>>>(u);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

int mfsr::launch_consolidate(const float2* measured, int64_t tile_stride, int64_t pair_stride, const PairTable& pt, int m,
                             int imageCount, int nTiles, int referenceImage, float2* one_to_one, float2* frame_shift,
                             int* status, float* inv0_scratch, cudaStream_t st, const unsigned long long* active0)
{
    const int n1 = imageCount - 1, warps = 4;
    const size_t smem = (size_t)warps * (2 * n1 * n1 + 2 * m + 4 * n1) * sizeof(float);
    if (smem > 48 * 1024) return MFSR_E_INVALID;
    // inv0_scratch (n1 * n1 + 1 floats, may be null): the inverse of the full system, computed once and shared by all tiles
    if (inv0_scratch) {
        consolidate_inverse_kernel<<<1, 32, (size_t)2 * n1 * n1 * sizeof(float), st>>>(pt, m, imageCount, inv0_scratch);
        MFSR_LAUNCH_CHECK();
    }
    consolidate_kernel<<<cdiv(nTiles, warps), warps * 32, smem, st>>>(measured, tile_stride, pair_stride, pt, m, imageCount, nTiles,
                                                                    referenceImage, one_to_one, frame_shift, status, inv0_scratch, active0);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

extern "C" int mfsr_stage_tile_align(const uint8_t* ref, const uint8_t* mov, int64_t img_pitch, int width, int height,
                                     const float* pre_shift, int64_t pre_shift_pitch, float* out_shift, int64_t out_shift_pitch,
                                     int32_t* out_argmin, float* out_ssd, int tile_size, int max_shift, int tilesX, int tilesY,
                                     float base_shift_x, float base_shift_y, float base_rotation, float threshold, void* stream)
{
    if (!ref || !mov) return MFSR_E_INVALID;
    TileAlignBatch b = {};
    b.img = ref; b.pitch = img_pitch; b.frame_stride = (int64_t)(mov - ref); b.w = width; b.h = height;
    b.pre = (const float2*)pre_shift; b.pre_pitch = pre_shift_pitch;
    b.out = (float2*)out_shift; b.out_pitch = out_shift_pitch;
    b.argmin = (int2*)out_argmin; b.ssd = out_ssd;
    b.pt.from[0] = 0; b.pt.to[0] = 1; b.n_pairs = 1;
    b.T = tile_size; b.M = max_shift; b.tx = tilesX; b.ty = tilesY;
    b.bsx = base_shift_x; b.bsy = base_shift_y; b.rot = base_rotation; b.threshold = threshold;
    return launch_tile_align(b, (cudaStream_t)stream);
}

extern "C" int mfsr_stage_upsample_shifts(const float* in_shift, int64_t in_pitch, float* out_shift, int64_t out_pitch,
                                          int oldLevel, int newLevel, int oldCountX, int oldCountY, int newCountX, int newCountY,
                                          int oldTileSize, int newTileSize, void* stream)
{
    UpsampleBatch u = {};
    u.in = (const float2*)in_shift; u.in_pitch = in_pitch; u.out = (float2*)out_shift; u.out_pitch = out_pitch; u.n_pairs = 1;
    u.oldLevel = oldLevel; u.newLevel = newLevel; u.oldCX = oldCountX; u.oldCY = oldCountY; u.newCX = newCountX; u.newCY = newCountY;
    u.oldT = oldTileSize; u.newT = newTileSize;
    return launch_upsample_shifts(u, (cudaStream_t)stream);
}

extern "C" int mfsr_stage_consolidate_shifts(const float* measured, const int* pair_from, const int* pair_to, int m, int imageCount,
                                             int tilesX, int tilesY, int referenceImage, float* one_to_one, float* frame_shift,
                                             int32_t* status, void* stream)
{
    if (!measured || !pair_from || !pair_to || !frame_shift || m < 1 || m > CONS_MAX_M || imageCount < 2 || imageCount - 1 > CONS_MAX_N) return MFSR_E_INVALID;
    if (referenceImage < 0 || referenceImage >= imageCount) return MFSR_E_INVALID;
    PairTable pt;
    for (int k = 0; k < m; k++) {
        if (pair_from[k] < 0 || pair_to[k] <= pair_from[k] || pair_to[k] >= imageCount) return MFSR_E_INVALID;
        pt.from[k] = (int8_t)pair_from[k]; pt.to[k] = (int8_t)pair_to[k];
    }
    return launch_consolidate((const float2*)measured, m, 1, pt, m, imageCount, tilesX * tilesY, referenceImage,
                              (float2*)one_to_one, (float2*)frame_shift, status, nullptr, (cudaStream_t)stream);
}
