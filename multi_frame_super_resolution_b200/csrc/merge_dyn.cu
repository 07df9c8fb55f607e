// merge_dyn.cu — scale-2 kernel-regression merge for PER-PIXEL shifts (the production scale-2 kernel).
//
// Same arithmetic as merge_generic_kernel (N x accumulateImagesSuperRes DeBayerKernels.cu:379-468 +
// ApplyWeighting kernel.cu:426 + GammasRGB kernel.cu:393).  Measured on real pipeline flows the integer
// HR shift round(2*flow) differs between horizontally adjacent output pixels in ~26 % of 4-pixel groups
// (Lucas-Kanade noise), so a warp never sees one shift parity: the static (RHO, EY) specialisation of
// merge_fast.cu diverges.  Here nothing branches on the shift:
//
//  * warp w owns row w of the tile and runs four passes J = 0..3; in a pass lane l owns the output pixel at
//    absolute HR column 4B + J, so (J = X%4, YM = Y%4) — which certainty sample each of the
//    25 taps reads — are compile-time, every warp scheduler (warp % 4 == YM) executes ONE code variant
//    and at most four variants are live per SM (the first version, with 8 variants interleaved on an SM,
//    stalled 23 of 24 issue slots on instruction fetch: profiles/r1b_merge_dyn_icache_ncu.txt);
//  * which raw sample a tap reads depends on the parity of X+sx / Y+sy only for the taps at +-1: those
//    taps are issued once per candidate destination under a lane predicate (49 predicated FMAs instead
//    of 25) and fold into a 3x3 per-raw-sample weight G, then 18 FMAs accumulate value and weight;
//  * the CFA phase of the window centre only permutes certainty channels and accumulators: certainty is
//    staged in four channel-permuted planes and fetched with a phase-dependent ADDRESS, the four class
//    sums are routed to absolute-phase accumulators by a 2-level select.
//  * raw windows are staged once per tile and frame as normalised float, de-interleaved by column
//    parity so that a warp's stride-2 window reads are bank-conflict free.
//  * the staged raw window is centred on the tile's MEAN shift; a pixel whose own window falls outside
//    it (alignment outliers) fetches its 3x3 raw samples from global memory instead, same code after.
// Taps that hit the clamp range (:414-419) and shifts beyond +-127 take the per-pixel generic path of
// merge_s2_common.cuh.
// Tile = 128 x TH output pixels, TH warps: 20 rows when all frames fit the 227 KB of shared memory (up to 9 frames), 16 rows for a
// 10th frame; longer bursts are merged in chunks of frames by the 20-row kernel when sum / weight images are available (the
// pipeline provides them), else by the 8- / 4-row variants (launch_merge_s2 at the end of this file has the measurements).
#include "merge_s2_common.cuh"

// frame-loop unroll factor (1: one code variant of ~2.5 KB per warp scheduler; tools/ab_build.sh for A/B builds)
#ifndef MFSR_LOOP_UNROLL
#define MFSR_LOOP_UNROLL 1
#endif
#define MFSR_STR(x) #x
#define MFSR_UNROLL(n) _Pragma(MFSR_STR(unroll n))

#ifdef MFSR_MERGE_TIMING
// debug builds (tools/merge_phases.py): cycles per phase, summed over the CTAs (thread 0 of every CTA)
__device__ unsigned long long g_merge_phase_cycles[8];
extern "C" int mfsr_debug_merge_phase_cycles(unsigned long long* host8, int reset)
{
    if (host8 && cudaMemcpyFromSymbol(host8, g_merge_phase_cycles, sizeof(g_merge_phase_cycles)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[8] = {0}; if (cudaMemcpyToSymbol(g_merge_phase_cycles, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}
#define MFSR_TICK(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_merge_phase_cycles[i], (unsigned long long)(t_ - t_prev_)); t_prev_ = t_; } } while (0)
#else
#define MFSR_TICK(i)
#endif

namespace mfsr {

namespace {

using namespace s2;

constexpr int RHALF = RWS / 2;   // odd raw columns live RHALF floats after the even ones of the same row

template <int TH> struct DCfg {
    static constexpr int NW_ = TH;                   // warps: one per tile row
    static constexpr int NT = 32 * TH;               // threads
    static constexpr int RHS = TH / 2 + 15;          // staged raw rows (TH/2 + taps 3 + shift slack +-6 raw rows)
    static constexpr int MHS = TH / 4 + 2;           // staged certainty rows
    static constexpr int PLANE = MHS * MWS;          // float2 elements per certainty plane
    static constexpr int SHIFT_BYTES = TH * TW * 2;
    static constexpr int RAW_BYTES = RHS * RWS * 4;
    static constexpr int MASK_BYTES = 4 * PLANE * 8;
    static constexpr int FRAME_BYTES = SHIFT_BYTES + RAW_BYTES + MASK_BYTES;
    static constexpr int KWS = TW / 2 + 2, KHS = TH / 2 + 2;      // staged kernel-parameter window (raw resolution + 1 each side)
    static constexpr int KERN_BYTES = KWS * KHS * 16;
};

MFSR_CX int g_e(int e, int p) { return mt::fl2(e + p); }      // raw sample of tap p relative to the window centre

// g = fma(w, c, g) under a lane predicate (p != 0).  Inline PTX keeps ptxas from turning the 40 conditional
// taps into FSEL + FFMA pairs (it did: 48 FSEL per pixel and frame in profiles/r1c).
__device__ __forceinline__ void pfma(float& g, float w, float c, int p)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q fma.rn.f32 %0, %1, %2, %0;\n\t}" : "+f"(g) : "f"(w), "f"(c), "r"(p));
}
__device__ __forceinline__ int bfe_s8(unsigned v, int pos)
{
    int r;
    asm("bfe.s32 %0, %1, %2, 8;" : "=r"(r) : "r"(v), "r"(pos));
    return r;
}

// One pixel, one frame, everything in shared memory.  J = X % 4 and YM = Y % 4 are static; the parities of
// X+sx / Y+sy (ex, ey) are lane predicates; the CFA phase of the window centre only moves addresses (pe/po,
// q0/q1) and the final routing (phx, phy).
//   R  : normalised raw samples R[gy+1][gx+1] around (k, ky)
//   q0 / q1 : certainty planes of y class 0 / 1 at the thread's first mask pixel
template <int J, int YM>
__device__ __forceinline__ void pixel_fast(const float (&w)[mt::NW], const float (&R)[3][3],
                                           const float2* __restrict__ q0, const float2* __restrict__ q1,
                                           int ex, int ey, int phx, int phy, float (&acc)[4], float (&wacc)[4])
{
    float Q[2][2][2][2];     // [mask row][mask col][y class][x class]
#pragma unroll
    for (int mr = 0; mr < 2; mr++)
#pragma unroll
        for (int mc = 0; mc < 2; mc++) {
            const float2 v0 = q0[mr * MWS + mc], v1 = q1[mr * MWS + mc];
            Q[mr][mc][0][0] = v0.x; Q[mr][mc][0][1] = v0.y; Q[mr][mc][1][0] = v1.x; Q[mr][mc][1][1] = v1.y;
        }
    const int fx[2] = {ex ^ 1, ex}, fy[2] = {ey ^ 1, ey};
    const int fxy[2][2] = {{fx[0] & fy[0], fx[0] & fy[1]}, {fx[1] & fy[0], fx[1] & fy[1]}};   // [x candidate][y candidate]
    // G[gy+1][gx+1] = sum over the taps that read raw sample (gx, gy) of weight * certainty.
    // The 9 taps with px, py in {-2, 0, 2} have one destination each and initialise G; the 16 taps at +-1 are
    // issued once per candidate destination under the matching predicate (40 predicated FMAs).
    float G[3][3];
#pragma unroll
    for (int py = -2; py <= 2; py += 2)
#pragma unroll
        for (int px = -2; px <= 2; px += 2) {
            const int gy = g_e(0, py), gx = g_e(0, px);
            G[gy + 1][gx + 1] = w[mt::widx(px, py)] * Q[mt::y_mrow(YM, py)][mt::x_mcol(J, px) - (J >= 2 ? 1 : 0)][gy & 1][gx & 1];
        }
#pragma unroll
    for (int py = -2; py <= 2; py++) {
#pragma unroll
        for (int px = -2; px <= 2; px++) {
            const bool ambx = (px & 1) != 0, amby = (py & 1) != 0;
            if (!ambx && !amby) continue;
            const int mr = mt::y_mrow(YM, py), mc = mt::x_mcol(J, px) - (J >= 2 ? 1 : 0);
            const float wt = w[mt::widx(px, py)];
#pragma unroll
            for (int eyc = 0; eyc < (amby ? 2 : 1); eyc++)
#pragma unroll
                for (int exc = 0; exc < (ambx ? 2 : 1); exc++) {
                    const int gy = g_e(eyc, py), gx = g_e(exc, px);
                    const int flag = (ambx && amby) ? fxy[exc][eyc] : (ambx ? fx[exc] : fy[eyc]);
                    pfma(G[gy + 1][gx + 1], wt, Q[mr][mc][gy & 1][gx & 1], flag);
                }
        }
    }
    // class sums (relative to the window centre)
    float t[4], u[4];
    t[0] = G[1][1] * R[1][1];                                   u[0] = G[1][1];
    t[1] = fmaf(G[1][2], R[1][2], G[1][0] * R[1][0]);           u[1] = G[1][0] + G[1][2];
    t[2] = fmaf(G[2][1], R[2][1], G[0][1] * R[0][1]);           u[2] = G[0][1] + G[2][1];
    t[3] = fmaf(G[2][2], R[2][2], fmaf(G[2][0], R[2][0], fmaf(G[0][2], R[0][2], G[0][0] * R[0][0])));
    u[3] = (G[0][0] + G[0][2]) + (G[2][0] + G[2][2]);
    // route to absolute CFA phase: absolute = relative ^ (phy, phx)
    {
        const float t0 = phx ? t[1] : t[0], t1 = phx ? t[0] : t[1], t2 = phx ? t[3] : t[2], t3 = phx ? t[2] : t[3];
        const float u0 = phx ? u[1] : u[0], u1 = phx ? u[0] : u[1], u2 = phx ? u[3] : u[2], u3 = phx ? u[2] : u[3];
        acc[0] += phy ? t2 : t0; acc[1] += phy ? t3 : t1; acc[2] += phy ? t0 : t2; acc[3] += phy ? t1 : t3;
        wacc[0] += phy ? u2 : u0; wacc[1] += phy ? u3 : u1; wacc[2] += phy ? u0 : u2; wacc[3] += phy ? u1 : u3;
    }
}

// ---- cold code shared by all 16 (J, YM) variants.  Kept out of line ON PURPOSE: inlined into every variant it is
// ---- ~350 once-executed instructions per pass (90 KB in total) that stream through the instruction caches and
// ---- evict the frame loops (profiles/r1e: 23 % of the stall samples on 14 % of the instructions).

// 13 regression weights of absolute HR pixel (X, Y) (:401, :427-430) from the staged kernel-parameter window.
// kwin: float4 [KHS][KWS], origin (kx0, ky0) in raw coordinates, clamp addressing already applied while staging.
__device__ __forceinline__ float4 lds_f4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// kwin_s: 32-bit SHARED address of the window (a generic pointer would turn the four loads into generic LDs)
static __device__ __noinline__ void compute_weights(unsigned kwin_s, int kws, int kx0, int ky0, int X, int Y, float* __restrict__ wl)
{
    const int fx = ((X - 1) >> 1) - kx0, fy = ((Y - 1) >> 1) - ky0;
    const unsigned a00 = kwin_s + (unsigned)(fy * kws + fx) * 16u, a01 = a00 + (unsigned)kws * 16u;
    const float4 K00 = lds_f4(a00), K10 = lds_f4(a00 + 16u), K01 = lds_f4(a01), K11 = lds_f4(a01 + 16u);
    const float ta = (X & 1) ? 0.25f : 0.75f, tb = (Y & 1) ? 0.25f : 0.75f;
    const float kx = tex_mix(K00.x, K10.x, K01.x, K11.x, ta, tb);
    const float ky = tex_mix(K00.y, K10.y, K01.y, K11.y, ta, tb);
    const float kz = tex_mix(K00.z, K10.z, K01.z, K11.z, ta, tb);
#pragma unroll
    for (int py = 0; py <= 2; py++)
#pragma unroll
        for (int px = -2; px <= 2; px++) {
            if (py == 0 && px < 0) continue;
            const float q = (float)(px * px) * kx + (float)(2 * px * py) * kz + (float)(py * py) * ky;
            float e;                                     // exp(-q/2) = 2^(-q/2 * log2 e), 2-ulp MUFU
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q * -0.72134752044448170368f));
            if (!(fabsf(e) < INFINITY)) e = (px * py == 0) ? 1.0f : 0.0f;      // :429-430
            wl[mt::widx(px, py)] = e;
        }
}

// CFA phase -> colour, ApplyWeighting (kernel.cu:426), GammasRGB (:393), one write of one pixel.
// Everything arrives BY VALUE: an out-of-line function can reach the kernel parameters only through a generic pointer
// (LD through the constant window instead of LDC), which put ~10 dependent generic loads in front of every pixel's store.
// cfa4: the four CFA colours packed 2 bits each.
// (so / si and wo / wi may alias: the chunked merge read-modify-writes the partial sums in place)
static __device__ __noinline__ void epilogue_px(float* __restrict__ orow, float* so, float* wo,
                                                const float* si, const float* wi, unsigned cfa4, float threshold, int flags,
                                                float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3, float f0, float f1, float f2)
{
    const float acc[4] = {a0, a1, a2, a3}, wacc[4] = {b0, b1, b2, b3}, fb3[3] = {f0, f1, f2};
    float s3[3] = {0.f, 0.f, 0.f}, w3[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const unsigned col = (cfa4 >> (2 * q)) & 3u;
#pragma unroll
        for (int c = 0; c < 3; c++)
            if (col == (unsigned)c) { s3[c] += acc[q]; w3[c] += wacc[q]; }
    }
    if (si) {                              // frame-chunked merge: sums of the earlier chunks
#pragma unroll
        for (int c = 0; c < 3; c++) { s3[c] = si[c] + s3[c]; w3[c] = wi[c] + w3[c]; }
    }
    if (so) {
        so[0] = s3[0]; so[1] = s3[1]; so[2] = s3[2];
        wo[0] = w3[0]; wo[1] = w3[1]; wo[2] = w3[2];
    }
    if (flags & MFSR_MERGE_PARTIAL_INTERNAL) return;
#pragma unroll
    for (int c = 0; c < 3; c++) orow[c] = finish_px(apply_weighting(s3[c], w3[c], fb3[c], threshold), flags);
}

// 3x3 normalised raw samples around (k, ky) of an alignment outlier whose window is not staged, from global memory.
// No clamping: the caller has checked that every tap stays inside the clamp range.
// rawf: frame base, norm_s: 32-bit SHARED address of {black[4], 1/white[4]} per CFA phase (by value / shared for the same
// reason as epilogue_px: no generic loads of kernel parameters on this path).
static __device__ __noinline__ void fetch_raw_global(const uint16_t* __restrict__ rawf, int64_t raw_pitch, unsigned norm_s, int k, int ky, float* __restrict__ out9)
{
    unsigned v[9];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const uint16_t* rrow = row_ptr(rawf, raw_pitch, ky - 1 + r) + (k - 1);
#pragma unroll
        for (int c = 0; c < 3; c++) v[r * 3 + c] = __ldg(rrow + c);
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int ph = ((ky - 1 + r) & 1) * 2 + ((k - 1 + c) & 1);
            float bl, iv;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bl) : "r"(norm_s + 4u * ph));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(iv) : "r"(norm_s + 16u + 4u * ph));
            out9[r * 3 + c] = ((float)v[r * 3 + c] - bl) * iv;
        }
}

// clamped taps and outsized shifts (sentinel sx == -128): the reference loop
static __device__ __noinline__ void slow_pixel(const FastArgs& F, int f, int X, int Y, int sx, int sy, const float* wl, float* ab)
{
    if (sx == -128) { const int2 s2 = shift_global(F.a, f, X, Y); sx = s2.x; sy = s2.y; }
    generic_pixel(F, f, X, Y, sx, sy, wl, ab);
}

// One pass of one warp: tile row `row` (absolute Y % 4 == YM), the 32 pixels X = X0abs + 4*lane + J.
template <int TH, int YM, int J>
__device__ __forceinline__ void run_row(const FastArgs& F, const unsigned char* smem, const int2* fbase, unsigned norm_s, int row, int x0, int y0, int X0abs, int Y0abs)
{
    using C = DCfg<TH>;
    const MergeArgs& A = F.a;
    const mfsr_merge_geom& g = A.g;
    const int lane = threadIdx.x & 31;
    const int y = y0 + row, Y = Y0abs + row;                  // window / absolute row
    const int x = x0 + 4 * lane + J, X = X0abs + 4 * lane + J;
    const int N = A.n_frames;
    if (x < 0 || x >= g.out_w || y < 0 || y >= g.out_h) return;
    const bool pix_on = x >= 1 && x < g.out_w - 1 && y >= 1 && y < g.out_h - 1;      // the reference skips the window border (:391)

    float acc[4] = {0.f, 0.f, 0.f, 0.f}, wacc[4] = {0.f, 0.f, 0.f, 0.f};
    float fb3[3] = {0.f, 0.f, 0.f};             // ApplyWeighting's inOutImg value, fetched early: its latency hides behind the frame loop
    if (A.fallback) {
        const float* p = row_ptr(A.fallback, A.fb_pitch, y) + 3 * x;
        fb3[0] = __ldg(p); fb3[1] = __ldg(p + 1); fb3[2] = __ldg(p + 2);
    }
    if (pix_on) {
        // ---- regression weights, frame independent (shared out-of-line code; wl doubles as the slow path's copy)
        float wl[mt::NW], W[mt::NW];
        compute_weights((unsigned)__cvta_generic_to_shared(smem + (size_t)N * C::FRAME_BYTES), C::KWS, (X0abs >> 1) - 1, (Y0abs >> 1) - 1, X, Y, wl);
#pragma unroll
        for (int i = 0; i < mt::NW; i++) W[i] = wl[i];
        // a pixel-frame runs the staged path when none of its taps (shifted or not) touches the clamp range (:414-419)
        const int lox = 2 * g.clamp_x0 + 2, spanx = 2 * g.clamp_x1 - 1 - lox, loy = 2 * g.clamp_y0 + 2, spany = 2 * g.clamp_y1 - 1 - loy;
        const bool pix_stat = spanx >= 0 && spany >= 0 && (unsigned)(Y - loy) <= (unsigned)spany && (unsigned)(X - lox) <= (unsigned)spanx;
        const int mx0 = (X0abs >> 2) - 1, my0 = (Y0abs >> 2) - 1;
        // byte offset of the thread's first certainty pixel inside a plane, and of its shift
        const int mask_off = ((((Y >> 2) + (YM < 2 ? -1 : 0)) - my0) * MWS + (((X >> 2) + (J < 2 ? -1 : 0)) - mx0)) * 8;
        const unsigned char* shp = smem + (row * TW + 4 * lane + J) * 2;
        const unsigned char* rawS = smem + C::SHIFT_BYTES;
        const unsigned char* maskS = smem + C::SHIFT_BYTES + C::RAW_BYTES + mask_off;

        MFSR_UNROLL(MFSR_LOOP_UNROLL)
        for (int f = 0; f < N; f++) {
            const int2 fb = fbase[f];               // .x = rx0, .y = ry0 of the staged raw window
            const unsigned sw = *(const unsigned short*)(shp + f * C::FRAME_BYTES);
            const int sx = bfe_s8(sw, 0), sy = bfe_s8(sw, 8);
            const int Xs = X + sx, Ys = Y + sy;
            const int k = Xs >> 1, ky = Ys >> 1;
            const int cc = k - 1 - fb.x, r0 = ky - 1 - fb.y;              // window column / row of sample (k-1, ky-1)
            // sentinel sx == -128 marks |shift| > 127 / NaN flow
            const bool noclamp = pix_stat && sx != -128 && (unsigned)(Xs - lox) <= (unsigned)spanx && (unsigned)(Ys - loy) <= (unsigned)spany;
            if (noclamp) {
                float R[3][3];
                if ((unsigned)cc <= (unsigned)(RWS - 3) && (unsigned)r0 <= (unsigned)(C::RHS - 3)) {
                    const int o = cc & 1;
                    const float* pe = (const float*)(rawS + f * C::FRAME_BYTES + r0 * (RWS * 4) + (cc >> 1) * 4 + o * (RHALF * 4));
                    const float* po = pe + (o ? 1 - RHALF : RHALF);
#pragma unroll
                    for (int r = 0; r < 3; r++) { R[r][0] = pe[r * RWS]; R[r][1] = po[r * RWS]; R[r][2] = pe[r * RWS + 1]; }
                } else {                                   // alignment outlier: its window is not staged
                    float tmp[9];
                    fetch_raw_global((const uint16_t*)((const char*)A.raw + A.raw_fs * f), A.raw_pitch, norm_s, k, ky, tmp);
#pragma unroll
                    for (int r = 0; r < 3; r++) { R[r][0] = tmp[3 * r]; R[r][1] = tmp[3 * r + 1]; R[r][2] = tmp[3 * r + 2]; }
                }
                const int pl = (k & 1) * 2 + (ky & 1);
                const unsigned char* q0 = maskS + f * C::FRAME_BYTES + pl * (C::PLANE * 8);
                const unsigned char* q1 = maskS + f * C::FRAME_BYTES + (pl ^ 1) * (C::PLANE * 8);
                pixel_fast<J, YM>(W, R, (const float2*)q0, (const float2*)q1, Xs & 1, Ys & 1, k & 1, ky & 1, acc, wacc);
            } else {
                float ab[8];
                slow_pixel(F, f, X, Y, sx, sy, wl, ab);
#pragma unroll
                for (int q = 0; q < 4; q++) { acc[q] += ab[q]; wacc[q] += ab[4 + q]; }
            }
        }
    }

    const unsigned cfa4 = (unsigned)A.cfa.c[0] | ((unsigned)A.cfa.c[1] << 2) | ((unsigned)A.cfa.c[2] << 4) | ((unsigned)A.cfa.c[3] << 6);
    epilogue_px(row_ptr(A.out, A.out_pitch, y) + 3 * x, A.sum_out ? row_ptr(A.sum_out, A.acc_pitch, y) + 3 * x : nullptr,
                A.sum_out ? row_ptr(A.weight_out, A.acc_pitch, y) + 3 * x : nullptr,
                A.sum_in ? row_ptr(A.sum_in, A.acc_pitch, y) + 3 * x : nullptr, A.sum_in ? row_ptr(A.weight_in, A.acc_pitch, y) + 3 * x : nullptr,
                cfa4, A.threshold, A.flags,
                acc[0], acc[1], acc[2], acc[3], wacc[0], wacc[1], wacc[2], wacc[3], fb3[0], fb3[1], fb3[2]);
}

template <int TH, int YM>
__device__ __forceinline__ void run_rows(const FastArgs& F, const unsigned char* smem, const int2* fbase, unsigned norm_s, int row, int x0, int y0, int X0abs, int Y0abs)
{
    // (A block barrier between the passes, to keep all warps on the same code variant, changes nothing: 6.58 vs 6.54 ms.
    // Unrolling the frame loop by 2 / 4 costs 1.34x / 2.4x: the 2.5 KB loop body must stay resident in the scheduler's L0.)
    run_row<TH, YM, 0>(F, smem, fbase, norm_s, row, x0, y0, X0abs, Y0abs);
    run_row<TH, YM, 1>(F, smem, fbase, norm_s, row, x0, y0, X0abs, Y0abs);
    run_row<TH, YM, 2>(F, smem, fbase, norm_s, row, x0, y0, X0abs, Y0abs);
    run_row<TH, YM, 3>(F, smem, fbase, norm_s, row, x0, y0, X0abs, Y0abs);
}

template <int TH>
__global__ void __launch_bounds__(DCfg<TH>::NT, DCfg<TH>::NT <= 512 ? 512 / DCfg<TH>::NT : 1)
merge_s2_dyn_kernel(const __grid_constant__ FastArgs F)
{
    using C = DCfg<TH>;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int2 fbase[MAXF];           // origin (rx0, ry0) of the staged raw window per frame
    __shared__ int s_part[MAXF * (TH / 2)][3];   // per (frame, row pair) item: sum sx, sum sy, count of sampled (non-outsized) shifts
    __shared__ float s_norm[8];            // black level [4] and reciprocal white level [4] per CFA phase, for the out-of-line outlier fetch
    if (threadIdx.x < 8) s_norm[threadIdx.x] = threadIdx.x < 4 ? F.black_ph[threadIdx.x] : F.inv_ph[threadIdx.x - 4];
    const unsigned norm_s = (unsigned)__cvta_generic_to_shared(s_norm);
    const MergeArgs& A = F.a;
    const mfsr_merge_geom& g = A.g;
    const int N = A.n_frames;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = (int)blockIdx.x * TW - F.x_off, y0 = (int)blockIdx.y * TH - F.y_off;   // window coords of the tile origin
    const int X0abs = x0 + g.org_x, Y0abs = y0 + g.org_y;                                // multiples of 4

#ifdef MFSR_MERGE_TIMING
    long long t_prev_ = clock64();
#endif
#ifndef MFSR_NO_L2_PREFETCH
    // L2 prefetch of what phase 1 will stage (certainty rows, kernel-parameter rows, fallback rows and the UNSHIFTED raw window):
    // the addresses are known now, so their DRAM latency overlaps phase 0 instead of following it.  Rows / columns are clamped
    // into the images; a line the staging does not need after all is harmless.
    {
        auto pf = [](const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
        const int mw = g.raw_w / 2, mh = g.raw_h / 2;
        const int mx0 = clampi((X0abs >> 2) - 1, 0, mw - 1), my0 = (Y0abs >> 2) - 1;
        constexpr int ML = (MWS * 16 + 127) / 128 + 1;                       // 128-byte lines per certainty row of the window
        for (int i = tid; i < N * C::MHS * ML; i += C::NT) {
            const int f = i / (C::MHS * ML), j = i - f * (C::MHS * ML), r = j / ML, l = j - r * ML;
            const int cx = min(mx0 + l * 8, mw - 1);
            pf((const char*)A.mask + A.mask_fs * f + A.mask_pitch * clampi(my0 + r, 0, mh - 1) + 16 * cx);
        }
        const int kx0 = clampi((X0abs >> 1) - 1, 0, g.raw_w - 1), ky0 = (Y0abs >> 1) - 1;
        constexpr int KL = (C::KWS * 16 + 127) / 128 + 1;
        for (int i = tid; i < C::KHS * KL; i += C::NT) {
            const int r = i / KL, l = i - r * KL;
            pf(row_ptr(A.kern, A.kern_pitch, clampi(ky0 + r, 0, g.raw_h - 1)) + min(kx0 + l * 8, g.raw_w - 1));
        }
        constexpr int RL = (TW / 2 + 32) * 2 / 128 + 2;                       // raw window: tile footprint + 16 columns each side
        const int rx0 = clampi((X0abs >> 1) - 16, 0, g.raw_w - 1), ry0 = (Y0abs >> 1) - 4;
        for (int i = tid; i < N * (TH / 2 + 8) * RL; i += C::NT) {
            const int f = i / ((TH / 2 + 8) * RL), j = i - f * ((TH / 2 + 8) * RL), r = j / RL, l = j - r * RL;
            pf((const char*)A.raw + A.raw_fs * f + A.raw_pitch * clampi(ry0 + r, 0, g.raw_h - 1) + 2 * min(rx0 + l * 64, g.raw_w - 1));
        }
        if (A.fallback) {
            constexpr int FL = TW * 12 / 128 + 1;
            const int fx0 = clampi(x0, 0, g.out_w - 1);
            for (int i = tid; i < TH * FL; i += C::NT) {
                const int r = i / FL, l = i - r * FL;
                pf(row_ptr(A.fallback, A.fb_pitch, clampi(y0 + r, 0, g.out_h - 1)) + 3 * min(fx0 + l * 10, g.out_w - 1));
            }
        }
    }
#endif

    // ---------------- phase 0: integer HR shifts of every tile pixel and frame -> char2 in shared memory.
    // Work item = (frame, row pair): 8 pixels per lane from a 3 x 4 flow window.  Two items are in flight per warp
    // (the next window is requested before the current one is consumed) — this phase is pure load latency otherwise.
    {
        const int Bq = (X0abs >> 2) + lane;
        const int fxb = 2 * Bq - 1;
        int cx[4];
#pragma unroll
        for (int c = 0; c < 4; c++) cx[c] = clampi(fxb + c, 0, g.raw_w - 1);
        const bool xin = fxb >= 0 && fxb + 3 < g.raw_w;           // the 4 flow columns are contiguous (no clamping)
        const int total = (TH / 2) * N;
        auto load = [&](int item, float2 (&Fl)[3][4]) {
            const int f = item / (TH / 2), rp = item - f * (TH / 2);
            const int a = (Y0abs >> 1) + rp;
            const float2* flow = (const float2*)((const char*)A.flow + A.flow_fs * f);
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const float2* fr = row_ptr(flow, A.flow_pitch, clampi(a - 1 + r, 0, g.raw_h - 1));
                if (xin) {
#pragma unroll
                    for (int c = 0; c < 4; c++) Fl[r][c] = __ldg(fr + fxb + c);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; c++) Fl[r][c] = __ldg(fr + cx[c]);
                }
            }
        };
        auto emit = [&](int item, const float2 (&Fl)[3][4]) {
            const int f = item / (TH / 2), rp = item - f * (TH / 2);
            int sum_x = 0, sum_y = 0, cnt = 0;
            unsigned packed[2][2];
#pragma unroll
            for (int yy = 0; yy < 2; yy++) {
                // even row 2a: flow rows (a-1, a) frac .75 ; odd row 2a+1: rows (a, a+1) frac .25
                const int rt = yy, rb = yy + 1;
                const bool ay = (yy == 0);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c0 = (j + 1) >> 1;
                    const bool ax = !(j & 1);
                    const float vx = mix25(mix25(Fl[rt][c0].x, Fl[rt][c0 + 1].x, ax), mix25(Fl[rb][c0].x, Fl[rb][c0 + 1].x, ax), ay);
                    const float vy = mix25(mix25(Fl[rt][c0].y, Fl[rt][c0 + 1].y, ax), mix25(Fl[rb][c0].y, Fl[rb][c0 + 1].y, ax), ay);
                    const float rx = roundf(__fmul_rn(vx, 2.0f)), ryf = roundf(__fmul_rn(vy, 2.0f));
                    const bool big = !(fabsf(rx) <= 127.0f) || !(fabsf(ryf) <= 127.0f);     // NaN / huge: sentinel -128
                    const int sx = big ? -128 : (int)rx, sy = big ? 0 : (int)ryf;
                    if (j == 0 && !big) { sum_x += sx; sum_y += sy; cnt++; }                // mean from a 1-in-4 subsample
                    const unsigned v = (unsigned)(sx & 0xff) | ((unsigned)(sy & 0xff) << 8);
                    if (j & 1) packed[yy][j >> 1] |= v << 16; else packed[yy][j >> 1] = v;
                }
            }
            unsigned char* sh = smem + (size_t)f * C::FRAME_BYTES;
            *(uint2*)(sh + ((2 * rp) * TW + 4 * lane) * 2) = make_uint2(packed[0][0], packed[0][1]);
            *(uint2*)(sh + ((2 * rp + 1) * TW + 4 * lane) * 2) = make_uint2(packed[1][0], packed[1][1]);
            sum_x = __reduce_add_sync(0xffffffffu, sum_x); sum_y = __reduce_add_sync(0xffffffffu, sum_y);
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0) { s_part[item][0] = sum_x; s_part[item][1] = sum_y; s_part[item][2] = cnt; }
        };
        float2 FA[3][4], FB[3][4];
        int item = warp;
        if (item < total) load(item, FA);
        while (item < total) {
            int next = item + C::NW_;
            if (next < total) load(next, FB);
            emit(item, FA);
            item = next;
            if (item >= total) break;
            next = item + C::NW_;
            if (next < total) load(next, FA);
            emit(item, FB);
            item = next;
        }
    }
    MFSR_TICK(0);
    // ---------------- phase 1a: certainty planes and kernel-parameter window (independent of the shifts)
    {
        const int mw = g.raw_w / 2, mh = g.raw_h / 2;
        const int mx0 = (X0abs >> 2) - 1, my0 = (Y0abs >> 2) - 1;
        // (frame, mask pixel) items flattened over the block (balanced: the barrier below waits for the slowest warp),
        // 4 loads in flight per thread
        const int nmask = N * C::PLANE;
        for (int i0 = tid; i0 < nmask; i0 += 4 * C::NT) {
            float4 m4[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = i0 + k * C::NT;
                if (i < nmask) {
                    const int f = i / C::PLANE, ii = i - f * C::PLANE;
                    const int r = ii / MWS, c = ii - r * MWS;
                    m4[k] = __ldg((const float4*)((const char*)A.mask + A.mask_fs * f + A.mask_pitch * clampi(my0 + r, 0, mh - 1) + 16 * clampi(mx0 + c, 0, mw - 1)));
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = i0 + k * C::NT;
                if (i < nmask) {
                    const int f = i / C::PLANE, ii = i - f * C::PLANE;
                    const float4 m = m4[k];
                    const float ch[3] = {isfinite(m.x) ? m.x : 0.f, isfinite(m.y) ? m.y : 0.f, isfinite(m.z) ? m.z : 0.f};   // :438-439
                    float q4[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) { const int col = A.cfa.c[q]; q4[q] = col == 0 ? ch[0] : (col == 1 ? ch[1] : ch[2]); }
                    // plane = xswap * 2 + absolute y phase; element = (x class 0, x class 1)
                    float2* m2 = (float2*)(smem + (size_t)f * C::FRAME_BYTES + C::SHIFT_BYTES + C::RAW_BYTES) + ii;
                    m2[0 * C::PLANE] = make_float2(q4[0], q4[1]);
                    m2[1 * C::PLANE] = make_float2(q4[2], q4[3]);
                    m2[2 * C::PLANE] = make_float2(q4[1], q4[0]);
                    m2[3 * C::PLANE] = make_float2(q4[3], q4[2]);
                }
            }
        }
        // kernel parameters (texture clamp addressing applied here)
        float4* ks = (float4*)(smem + (size_t)N * C::FRAME_BYTES);
        const int kx0 = (X0abs >> 1) - 1, ky0 = (Y0abs >> 1) - 1;
        for (int i = tid; i < C::KWS * C::KHS; i += C::NT) {
            const int r = i / C::KWS, c = i - r * C::KWS;
            ks[i] = __ldg(row_ptr(A.kern, A.kern_pitch, clampi(ky0 + r, 0, g.raw_h - 1)) + clampi(kx0 + c, 0, g.raw_w - 1));
        }
    }
    MFSR_TICK(1);
    __syncthreads();
    MFSR_TICK(2);
    // staged raw window: the tile's own footprint displaced by the MEAN shift, spare rows/columns split evenly.  Every warp
    // derives the (identical) origins itself — a benign same-value race instead of a second block barrier.
    for (int f = lane; f < N; f += 32) {
        int sx = 0, sy = 0, cnt = 0;
        for (int k = 0; k < TH / 2; k++) { sx += s_part[f * (TH / 2) + k][0]; sy += s_part[f * (TH / 2) + k][1]; cnt += s_part[f * (TH / 2) + k][2]; }
        cnt = max(cnt, 1);
        const int mx = (int)floorf((float)sx / (float)cnt + 0.5f), my = (int)floorf((float)sy / (float)cnt + 0.5f);
        fbase[f] = make_int2((((X0abs + mx - 2) >> 1) - (RWS - (TW / 2 + 3)) / 2) & ~3,
                             ((Y0abs + my - 2) >> 1) - (C::RHS - (TH / 2 + 3)) / 2);
    }
    __syncwarp();

    // ---------------- phase 1b: normalised raw windows (de-interleaved by column parity).  Thread t owns the 4-column chunk
    // position t of the window (row r, chunk c4: computed once, 32-bit offsets inside a frame) and walks the frames four at a
    // time with the loads in flight; the CH - NT positions beyond the block size are flattened over (position, frame) so that
    // no thread runs a second walk.  (The fully flattened form spent ~25 instructions per sample on index arithmetic.)
    {
        constexpr int CH = C::RHS * (RWS / 4);
        constexpr int FSTRIDE = C::FRAME_BYTES / 4;                 // floats between the windows of consecutive frames
        const unsigned rpitch = (unsigned)A.raw_pitch;
        const float bk0 = F.black_ph[0], bk1 = F.black_ph[1], bk2 = F.black_ph[2], bk3 = F.black_ph[3];
        const float iv0 = F.inv_ph[0], iv1 = F.inv_ph[1], iv2 = F.inv_ph[2], iv3 = F.inv_ph[3];
        auto fetch = [&](int f, int r, int c4, uint2& pv, int& odd) {
            const int2 fi = fbase[f];
            const int yy = clampi(fi.y + r, 0, g.raw_h - 1), xx = fi.x + 4 * c4;        // xx is a multiple of 4
            const char* fb = (const char*)A.raw + A.raw_fs * f;
            odd = yy & 1;
            if ((unsigned)xx <= (unsigned)(g.raw_w - 4)) pv = __ldg((const uint2*)(fb + ((unsigned)yy * rpitch + 2u * (unsigned)xx)));
            else {
                const uint16_t* rrow = (const uint16_t*)(fb + (unsigned)yy * rpitch);
                unsigned v[4];
#pragma unroll
                for (int q = 0; q < 4; q++) v[q] = __ldg(rrow + clampi(xx + q, 0, g.raw_w - 1));
                pv = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
            }
        };
        auto store = [&](float* rs, uint2 pv, int odd) {
            const float be = odd ? bk2 : bk0, bo = odd ? bk3 : bk1, ie = odd ? iv2 : iv0, io = odd ? iv3 : iv1;
            *(float2*)rs = make_float2(((float)(pv.x & 0xffffu) - be) * ie, ((float)(pv.y & 0xffffu) - be) * ie);
            *(float2*)(rs + RHALF) = make_float2(((float)(pv.x >> 16) - bo) * io, ((float)(pv.y >> 16) - bo) * io);
        };
        float* win0 = (float*)(smem + C::SHIFT_BYTES);
        if (tid < CH) {
            const int r = tid / (RWS / 4), c4 = tid - r * (RWS / 4);
            float* rs = win0 + r * RWS + 2 * c4;
            for (int f0 = 0; f0 < N; f0 += 4) {
                uint2 p4[4]; int o4[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (f0 + k < N) fetch(f0 + k, r, c4, p4[k], o4[k]);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (f0 + k < N) store(rs + (f0 + k) * FSTRIDE, p4[k], o4[k]);
            }
        }
        if (CH > C::NT) {
            constexpr int REST = CH > C::NT ? CH - C::NT : 1;
            for (int i = tid; i < REST * N; i += C::NT) {
                const int f = i / REST, pp = C::NT + (i - f * REST);
                const int r = pp / (RWS / 4), c4 = pp - r * (RWS / 4);
                uint2 pv; int od;
                fetch(f, r, c4, pv, od);
                store(win0 + f * FSTRIDE + r * RWS + 2 * c4, pv, od);
            }
        }
    }
    MFSR_TICK(3);
    __syncthreads();
    MFSR_TICK(4);

    // ---------------- phase 2: warp w owns tile row w (Y % 4 == w % 4 == its scheduler): four passes J = 0..3,
    // each a small loop over the frames (one code variant per warp scheduler at any time)
    switch (warp & 3) {
        case 0: run_rows<TH, 0>(F, smem, fbase, norm_s, warp, x0, y0, X0abs, Y0abs); break;
        case 1: run_rows<TH, 1>(F, smem, fbase, norm_s, warp, x0, y0, X0abs, Y0abs); break;
        case 2: run_rows<TH, 2>(F, smem, fbase, norm_s, warp, x0, y0, X0abs, Y0abs); break;
        default: run_rows<TH, 3>(F, smem, fbase, norm_s, warp, x0, y0, X0abs, Y0abs); break;
    }
    MFSR_TICK(5);
#ifdef MFSR_MERGE_TIMING
    __syncthreads();
    MFSR_TICK(6);
#endif
}

template <int TH>
int launch_th(const FastArgs& F, cudaStream_t st)
{
    using C = DCfg<TH>;
    const mfsr_merge_geom& g = F.a.g;
    const size_t smem = (size_t)C::FRAME_BYTES * F.a.n_frames + C::KERN_BYTES;
    // opt-in shared memory: 227 KB per block on sm_100 minus this instantiation's static tables
    // per device: the attribute belongs to the current device's context (a process may hold handles on several GPUs)
    static size_t max_dyn_dev[64] = {0};
    int dev = 0;
    MFSR_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return MFSR_E_INVALID;
    if (!max_dyn_dev[dev]) {
        cudaFuncAttributes at;
        MFSR_CUDA_TRY(cudaFuncGetAttributes(&at, merge_s2_dyn_kernel<TH>));
        const size_t lim = (size_t)227 * 1024 - at.sharedSizeBytes;
        MFSR_CUDA_TRY(cudaFuncSetAttribute(merge_s2_dyn_kernel<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim));
        max_dyn_dev[dev] = lim;
    }
    const size_t max_dyn = max_dyn_dev[dev];
    if (smem > max_dyn) return MFSR_E_INVALID;
    dim3 grid(cdiv(g.out_w + F.x_off, TW), cdiv(g.out_h + F.y_off, TH));
    merge_s2_dyn_kernel<TH><<<grid, C::NT, smem, st>>>(F);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

}  // namespace

int launch_merge_s2(const MergeArgs& A, cudaStream_t st)
{
    const mfsr_merge_geom& g = A.g;
    if (g.scale != 2 || A.n_frames < 1 || A.n_frames > MAXF) return MFSR_E_INVALID;
    if (g.org_x < 0 || g.org_y < 0 || (g.raw_w & 1) || (g.raw_h & 1) || g.raw_w < 8 || g.raw_h < 8) return MFSR_E_INVALID;
    // vector loads of the staging phase
    if (((uintptr_t)A.raw & 7) || (A.raw_pitch & 7) || (A.raw_fs & 7)) return MFSR_E_INVALID;
    if (A.raw_pitch <= 0 || A.raw_pitch * (int64_t)g.raw_h >= (1ll << 32)) return MFSR_E_INVALID;      // 32-bit offsets inside a frame (staging)
    if (((uintptr_t)A.mask & 15) || (A.mask_pitch & 15) || (A.mask_fs & 15) || ((uintptr_t)A.kern & 15) || (A.kern_pitch & 15)) return MFSR_E_INVALID;
    if (((uintptr_t)A.flow & 7) || (A.flow_pitch & 7) || (A.flow_fs & 7)) return MFSR_E_INVALID;
    FastArgs F;
    F.a = A;
    for (int c = 0; c < 3; c++) F.inv_white[c] = 1.0f / A.white[c];
    for (int q = 0; q < 4; q++) { F.black_ph[q] = A.black[A.cfa.c[q]]; F.inv_ph[q] = F.inv_white[A.cfa.c[q]]; }
    F.x_off = g.org_x & 3; F.y_off = g.org_y & 3;
    F.ph2c = 0;
    for (int q = 0; q < 4; q++) {
        if (A.cfa.c[q] >= 0 && A.cfa.c[q] < 3) F.ph2c |= 1u << (3 * q + A.cfa.c[q]);
        for (int c = 0; c < 3; c++) F.cfa_sel[q][c] = (A.cfa.c[q] == c || (c == 2 && (A.cfa.c[q] < 0 || A.cfa.c[q] > 2))) ? 1.0f : 0.0f;
        F.nbi_ph[q] = -F.black_ph[q] * F.inv_ph[q];
    }
    // Round 2: the predicate-free slot kernel (merge_pf.cu) takes every burst it can keep resident (10 frames), longer bursts in
    // balanced chunks of frames when sum / weight images are available; the kernels below remain for bursts without them.
    static const char* oldenv = getenv("MFSR_MERGE_OLD");
    if (!(oldenv && atoi(oldenv))) {
        const int cap = merge_pf_capacity(), n = A.n_frames;
        if (n <= cap) return launch_merge_pf(F, st);
        if (A.sum_out && A.weight_out && A.acc_pitch >= (int64_t)g.out_w * 12) {
            const int chunks = (n + cap - 1) / cap, per = (n + chunks - 1) / chunks;
            for (int c = 0, f0 = 0; c < chunks; c++, f0 += per) {
                FastArgs Fc = F;
                Fc.a.raw = (const uint16_t*)((const char*)A.raw + A.raw_fs * f0);
                Fc.a.mask = (const float4*)((const char*)A.mask + A.mask_fs * f0);
                Fc.a.flow = (const float2*)((const char*)A.flow + A.flow_fs * f0);
                Fc.a.n_frames = n - f0 < per ? n - f0 : per;
                Fc.a.sum_in = c ? A.sum_out : nullptr; Fc.a.weight_in = c ? A.weight_out : nullptr;
                if (c < chunks - 1) { Fc.a.flags |= MFSR_MERGE_PARTIAL_INTERNAL; Fc.a.fallback = nullptr; }
                const int rcc = launch_merge_pf(Fc, st);
                if (rcc != MFSR_OK) return rcc;
            }
            return MFSR_OK;
        }
    }
    static const char* thenv = getenv("MFSR_MERGE_TH");
    const int want = thenv ? atoi(thenv) : 0;
    const size_t n = (size_t)A.n_frames, budget1 = 227 * 1024 - 6144;     // dynamic part; launch_th re-checks against the exact limit
    // More frames than the 20- / 16-row tiles hold in shared memory (9 / 10 at the default geometry): with sum / weight images available
    // the burst is merged in balanced chunks of frames with the 20-row kernel (the chunk's partial sums are read-modify-written
    // by the same thread, the last chunk normalises) — the 8- and 4-row variants that would hold all frames at once cost 1.6x
    // per pixel and frame (10.1 vs 6.4 ms at config 2), the extra 96 B per pixel and chunk boundary is cheap against that.
    // Tile height = warps per SM.  20 rows (640 threads x 96 registers, 191 KB at 8 frames) beat 16 rows by 6 % (6.10 vs 6.49 ms,
    // same box); 24 rows spill (80 registers), 8 rows are 1.6x slower.  16 rows take one more frame (10) than 20 rows (9).
    auto fits = [&](size_t frame_bytes, size_t kern_bytes) { return n * frame_bytes + kern_bytes <= budget1; };
    if (!want && fits(DCfg<20>::FRAME_BYTES, DCfg<20>::KERN_BYTES)) { const int rc20 = launch_th<20>(F, st); if (rc20 != MFSR_E_INVALID) return rc20; }
    if (!want && fits(DCfg<16>::FRAME_BYTES, DCfg<16>::KERN_BYTES)) { const int rc16 = launch_th<16>(F, st); if (rc16 != MFSR_E_INVALID) return rc16; }
    const size_t cap20 = (budget1 - DCfg<20>::KERN_BYTES) / DCfg<20>::FRAME_BYTES;
    if (!want && n > cap20 && A.sum_out && A.weight_out && A.acc_pitch >= (int64_t)g.out_w * 12) {
        const int chunks = (int)((n + cap20 - 1) / cap20), per = (int)((n + chunks - 1) / chunks);
        for (int c = 0, f0 = 0; c < chunks; c++, f0 += per) {
            FastArgs Fc = F;
            Fc.a.raw = (const uint16_t*)((const char*)A.raw + A.raw_fs * f0);
            Fc.a.mask = (const float4*)((const char*)A.mask + A.mask_fs * f0);
            Fc.a.flow = (const float2*)((const char*)A.flow + A.flow_fs * f0);
            Fc.a.n_frames = (int)n - f0 < per ? (int)n - f0 : per;
            Fc.a.sum_in = c ? A.sum_out : nullptr; Fc.a.weight_in = c ? A.weight_out : nullptr;
            if (c < chunks - 1) { Fc.a.flags |= MFSR_MERGE_PARTIAL_INTERNAL; Fc.a.fallback = nullptr; }
            const int rcc = launch_th<20>(Fc, st);
            if (rcc != MFSR_OK) return rcc;
        }
        return MFSR_OK;
    }
    int rc = MFSR_E_INVALID;
    if (want == 24 && fits(DCfg<24>::FRAME_BYTES, DCfg<24>::KERN_BYTES)) rc = launch_th<24>(F, st);
    if (rc == MFSR_E_INVALID && want == 20 && fits(DCfg<20>::FRAME_BYTES, DCfg<20>::KERN_BYTES)) rc = launch_th<20>(F, st);
    if (rc == MFSR_E_INVALID && want == 8 && fits(DCfg<8>::FRAME_BYTES, DCfg<8>::KERN_BYTES)) rc = launch_th<8>(F, st);
    if (rc == MFSR_E_INVALID && fits(DCfg<16>::FRAME_BYTES, DCfg<16>::KERN_BYTES)) rc = launch_th<16>(F, st);
    if (rc == MFSR_E_INVALID && fits(DCfg<8>::FRAME_BYTES, DCfg<8>::KERN_BYTES)) rc = launch_th<8>(F, st);
    if (rc == MFSR_E_INVALID && fits(DCfg<4>::FRAME_BYTES, DCfg<4>::KERN_BYTES)) rc = launch_th<4>(F, st);
    return rc;
}

}  // namespace mfsr
