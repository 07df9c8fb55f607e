// merge_pf.cu — scale-2 kernel-regression merge, predicate-free slot formulation (the production scale-2 kernel of round 2).
//
// Same arithmetic as merge_generic_kernel (N x accumulateImagesSuperRes DeBayerKernels.cu:379-468 + ApplyWeighting
// kernel.cu:426 + GammasRGB kernel.cu:393), organised around what the B200 issues cheaply.  Measured issue cost per warp
// instruction and scheduler (tools/microbench, profiles/r2a_issue_rates_*.txt): FFMA / FMUL / FADD / IADD 1 cycle,
// LOP3 / SHF / SEL / IMAD / ISETP / PRMT 2 cycles, FFMA2 3 cycles, F2I 8+, shared-memory loads on their own pipe.  Hence:
//
//  * the per-pixel, per-frame selection "which raw sample does tap +-1 read" is arithmetic, not predicates or selects: the
//    parity of X+sx enters as the floats fx / nx = 1-fx (merge_slots.h), 20 + 16 multiply-adds fold the 25 weights into the
//    4 x 4 (destination, certainty cell) slots, 16 more apply the certainty, 14 accumulate value and weight per CFA class;
//  * everything an output pixel needs to know about a frame is ONE 16-bit descriptor written by phase 0:
//    bits 2..13 byte offset of its 3x3 raw window in the staged (column-parity de-interleaved) frame window, bit 0 / 1 the
//    parities of X+sx / Y+sy, bit 14 / 15 the column / row parity of the window (they select the certainty plane and route
//    the class sums to absolute CFA phase).  Phase 0 builds it from two small lookup tables (column part + row part):
//    no integer division, shift-and-mask chain or float->int conversion per pixel (rounding by magic-number adds);
//  * descriptors >= SPECIAL mark pixel-frames whose window is not staged (alignment outliers), whose taps touch the clamp
//    range (:414-419) or whose shift is not finite: an out-of-line path recomputes them from global memory.
//
// Tile = 128 x TH output pixels, one warp per row, four passes J = X % 4 (compile time, like YM = Y % 4); all frames of the
// tile resident in shared memory (descriptors 2 B per pixel, raw window 96 x (TH/2+15) floats, four certainty planes of
// 2 KB per frame); bursts beyond the capacity are merged in chunks of frames (partial sums read-modify-written in place).
#include "merge_s2_common.cuh"
#include "merge_slots.h"

namespace mfsr {

namespace {

using namespace s2;

constexpr int RHALF = RWS / 2;               // odd raw columns live RHALF floats after the even ones of the same row
constexpr unsigned SPECIAL = 0x3000u;        // descriptor byte offsets from here on: not a staged pixel-frame
constexpr unsigned POISON = 0x10000u;        // table entry of an index outside the staged window (saturates the sum)
constexpr int PLANE_BYTES = 0x800;           // one certainty plane; plane (o, rho) of a frame at (o + 2 rho) * PLANE_BYTES
constexpr int MASK_FRAME_BYTES = 4 * PLANE_BYTES;

template <int TH> struct PCfg {
    static constexpr int NW_ = TH;
    static constexpr int NT = 32 * TH;
    static constexpr int RHS = TH / 2 + 15;          // staged raw rows (TH/2 + taps 3 + shift slack +-6 raw rows)
    static constexpr int MHS = TH / 4 + 2;           // staged certainty rows
    static constexpr int DESC_BYTES = TH * TW * 2;
    static constexpr int RAW_BYTES = RHS * RWS * 4;
    static constexpr int KWS = TW / 2 + 2, KHS = TH / 2 + 2;
    static constexpr int KERN_BYTES = KWS * KHS * 16;
    static constexpr int VN = 2 * (RWS - 2), UN = 2 * (RHS - 2);      // valid table indices; entry [VN] / [UN] is the poison
    static constexpr int TAB_BYTES = ((VN + 1 + UN + 1) * 4 + 127) & ~127;
    static constexpr int OT_ROW_BYTES = TW * 12;                  // one output row of the tile (float3)
    static constexpr int OT_BYTES = TH * OT_ROW_BYTES;            // fallback-in / result-out tile moved by TMA bulk copies
    static_assert(MHS * MWS * 8 <= PLANE_BYTES, "certainty plane does not fit its slot");
    static_assert((RHS - 3) * RWS * 4 + (RHALF + RHALF) * 4 < (int)SPECIAL, "raw window offsets collide with the special range");
    static size_t smem_bytes(int n, bool tma) { return (size_t)n * (DESC_BYTES + RAW_BYTES + MASK_FRAME_BYTES) + KERN_BYTES + TAB_BYTES + (tma ? OT_BYTES : 0); }
};

__device__ __forceinline__ float4 lds_f4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// ---- TMA bulk copies (cp.async.bulk, Hopper+ / sm_100a) for the tile's fallback rows (in) and result rows (out): pure copies of
// 1536-byte row segments that the per-pixel code otherwise moved as 3 + 3 scalar accesses with a 48-byte lane stride (32 sectors per
// warp instruction).  One elected thread issues them; completion of the loads is tracked by an mbarrier, of the stores by a bulk group.
__device__ __forceinline__ void mbar_init(unsigned mbar_s, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_s), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar_s, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar_s, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(mbar_s), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_row(unsigned dst_s, const void* src, unsigned bytes, unsigned mbar_s)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_s), "l"(src), "r"(bytes), "r"(mbar_s) : "memory");
}
__device__ __forceinline__ void tma_store_row(void* dst, unsigned src_s, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 13 regression weights of absolute HR pixel (X, Y) (:401, :427-430) from the staged kernel-parameter window
// (float4 [KHS][KWS], origin (kx0, ky0) in raw coordinates, clamp addressing applied while staging).
// kwin_s is a 32-bit SHARED address.  Inlined ONCE into the pass loop (the pass index J is a run-time loop there): the weights
// stay in registers (round 2, ncu: the out-of-line version returned them through local memory and every pass paid that
// round trip while the other three warps of its scheduler were the only cover).
__device__ __forceinline__ void compute_weights(unsigned kwin_s, int kws, int kx0, int ky0, int X, int Y, float (&wl)[mt::NW])
{
    const int fx = ((X - 1) >> 1) - kx0, fy = ((Y - 1) >> 1) - ky0;
    const unsigned a00 = kwin_s + (unsigned)(fy * kws + fx) * 16u, a01 = a00 + (unsigned)kws * 16u;
    const float4 K00 = lds_f4(a00), K10 = lds_f4(a00 + 16u), K01 = lds_f4(a01), K11 = lds_f4(a01 + 16u);
    const float ta = (X & 1) ? 0.25f : 0.75f, tb = (Y & 1) ? 0.25f : 0.75f;
    // exp(-q/2) = 2^(q * -log2(e)/2): the scale is applied to the three parameters once
    const float kx = tex_mix(K00.x, K10.x, K01.x, K11.x, ta, tb) * -0.72134752044448170368f;
    const float ky = tex_mix(K00.y, K10.y, K01.y, K11.y, ta, tb) * -0.72134752044448170368f;
    const float kz = tex_mix(K00.z, K10.z, K01.z, K11.z, ta, tb) * -0.72134752044448170368f;
    bool all_finite = true;
#pragma unroll
    for (int py = 0; py <= 2; py++)
#pragma unroll
        for (int px = -2; px <= 2; px++) {
            if (py == 0 && px < 0) continue;
            const float q = (float)(px * px) * kx + (float)(2 * px * py) * kz + (float)(py * py) * ky;
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
            all_finite = all_finite && (fabsf(e) < INFINITY);
            wl[mt::widx(px, py)] = e;
        }
    if (!all_finite) {                       // :429-430, rare: a non-finite weight becomes 1 on the centre cross, 0 elsewhere
#pragma unroll
        for (int py = 0; py <= 2; py++)
#pragma unroll
            for (int px = -2; px <= 2; px++) {
                if (py == 0 && px < 0) continue;
                const float e = wl[mt::widx(px, py)];
                if (!(fabsf(e) < INFINITY)) wl[mt::widx(px, py)] = (px * py == 0) ? 1.0f : 0.0f;
            }
    }
}

// relative class sums (cy*2+cx around the window centre) -> absolute CFA phase: absolute = relative ^ (phy, phx)
__device__ __forceinline__ void route_add(const float (&t)[4], const float (&u)[4], bool phx, bool phy, float (&acc)[4], float (&wacc)[4])
{
    const float t0 = phx ? t[1] : t[0], t1 = phx ? t[0] : t[1], t2 = phx ? t[3] : t[2], t3 = phx ? t[2] : t[3];
    const float u0 = phx ? u[1] : u[0], u1 = phx ? u[0] : u[1], u2 = phx ? u[3] : u[2], u3 = phx ? u[2] : u[3];
    acc[0] += phy ? t2 : t0; acc[1] += phy ? t3 : t1; acc[2] += phy ? t0 : t2; acc[3] += phy ? t1 : t3;
    wacc[0] += phy ? u2 : u0; wacc[1] += phy ? u3 : u1; wacc[2] += phy ? u0 : u2; wacc[3] += phy ? u1 : u3;
}

template <int J, int YM>
__device__ __forceinline__ void fold_rt_case(const float (&w)[mt::NW], float fx, float fy, const float (&Q)[2][2][2][2], const float (&R)[3][3], float (&t)[4], float (&u)[4])
{
    ms::pixel_fold<J, YM>(w, fx, 1.0f - fx, fy, 1.0f - fy, Q, R, t, u);
}

// A pixel-frame that phase 0 could not describe: recomputed from global memory.  Alignment outliers (window not staged) whose
// taps stay inside the clamp range run the slot fold on 3x3 raw samples and 2x2 certainty cells fetched from global memory;
// clamped taps and non-finite / outsized shifts run the reference loop (generic_pixel).  ab[0..3] / ab[4..7]: value / weight sums
// per ABSOLUTE CFA phase.  The weights are recomputed here (rare path) so that the common path never spills them.
static __device__ __noinline__ void special_pixel(const FastArgs& F, unsigned kwin_s, int kws, int kx0, int ky0, int f, int X, int Y, float* __restrict__ ab)
{
    const MergeArgs& A = F.a;
    const mfsr_merge_geom& g = A.g;
    float wl[mt::NW];
    compute_weights(kwin_s, kws, kx0, ky0, X, Y, wl);
    const int2 s = shift_global(A, f, X, Y);
    const int Xs = X + s.x, Ys = Y + s.y;
    const int lox = 2 * g.clamp_x0 + 2, hix = 2 * g.clamp_x1 - 1, loy = 2 * g.clamp_y0 + 2, hiy = 2 * g.clamp_y1 - 1;
    const bool fin = abs(s.x) < (1 << 20) && abs(s.y) < (1 << 20);
    if (!(fin && X >= lox && X <= hix && Y >= loy && Y <= hiy && Xs >= lox && Xs <= hix && Ys >= loy && Ys <= hiy)) {
        generic_pixel(F, f, X, Y, s.x, s.y, wl, ab);
        return;
    }
    // unclamped: the taps read raw samples (k-1..k+1, ky-1..ky+1) and certainty cells ((X-2)>>2 .. +1, (Y-2)>>2 .. +1)
    const int k = Xs >> 1, ky = Ys >> 1, phx = k & 1, phy = ky & 1;
    const uint16_t* rawf = (const uint16_t*)((const char*)A.raw + A.raw_fs * f);
    const float4* maskf = (const float4*)((const char*)A.mask + A.mask_fs * f);
    unsigned rv[3][3];
    float4 mv[2][2];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const uint16_t* rrow = row_ptr(rawf, A.raw_pitch, ky - 1 + r) + (k - 1);
#pragma unroll
        for (int c = 0; c < 3; c++) rv[r][c] = __ldg(rrow + c);
    }
#pragma unroll
    for (int mr = 0; mr < 2; mr++) {
        const float4* mrow = row_ptr(maskf, A.mask_pitch, ((Y - 2) >> 2) + mr) + ((X - 2) >> 2);
#pragma unroll
        for (int mc = 0; mc < 2; mc++) mv[mr][mc] = __ldg(mrow + mc);
    }
    float R[3][3], Q[2][2][2][2];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int ph = ((phy ^ ((r + 1) & 1)) * 2) + (phx ^ ((c + 1) & 1));        // CFA phase of sample (k-1+c, ky-1+r)
            R[r][c] = fmaf((float)rv[r][c], F.inv_ph[ph], F.nbi_ph[ph]);
        }
#pragma unroll
    for (int mr = 0; mr < 2; mr++)
#pragma unroll
        for (int mc = 0; mc < 2; mc++) {
            const float4 m = mv[mr][mc];
            const float c0 = isfinite(m.x) ? m.x : 0.f, c1 = isfinite(m.y) ? m.y : 0.f, c2 = isfinite(m.z) ? m.z : 0.f;
#pragma unroll
            for (int cy = 0; cy < 2; cy++)
#pragma unroll
                for (int cx = 0; cx < 2; cx++) {
                    const int q = (cy ^ phy) * 2 + (cx ^ phx);
                    Q[mr][mc][cy][cx] = c0 * F.cfa_sel[q][0] + c1 * F.cfa_sel[q][1] + c2 * F.cfa_sel[q][2];
                }
        }
    const float fx = (float)(Xs & 1), fy = (float)(Ys & 1);
    float t[4], u[4];
    switch ((Y & 3) * 4 + (X & 3)) {
#define MFSR_CASE(YMv, Jv) case YMv * 4 + Jv: fold_rt_case<Jv, YMv>(wl, fx, fy, Q, R, t, u); break;
        MFSR_CASE(0, 0) MFSR_CASE(0, 1) MFSR_CASE(0, 2) MFSR_CASE(0, 3) MFSR_CASE(1, 0) MFSR_CASE(1, 1) MFSR_CASE(1, 2) MFSR_CASE(1, 3)
        MFSR_CASE(2, 0) MFSR_CASE(2, 1) MFSR_CASE(2, 2) MFSR_CASE(2, 3) MFSR_CASE(3, 0) MFSR_CASE(3, 1) MFSR_CASE(3, 2) default: fold_rt_case<3, 3>(wl, fx, fy, Q, R, t, u); break;
#undef MFSR_CASE
    }
    float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
    route_add(t, u, phx != 0, phy != 0, a4, b4);
#pragma unroll
    for (int q = 0; q < 4; q++) { ab[q] = a4[q]; ab[4 + q] = b4[q]; }
}

// A pixel-frame whose 3x3 raw window is not staged (an alignment outlier: mis-matched tiles whose shift is far from the tile's
// mean; on some burst seeds 5 % of the pixel-frames, concentrated in a fifth of the tiles): the shift is recomputed from the flow,
// the nine raw samples come from global memory, normalised, and the frame loop continues with them exactly as with staged ones.
// Returns the descriptor flag bits (bit 0 / 1 parities of X+sx / Y+sy, bit 14 = !phx, bit 15 = phy) or 0xFFFFFFFF when a tap
// would touch the clamp range / the shift is not finite (the caller then runs the reference loop).  Out of line: ~80 instructions
// that most warps never execute.
static __device__ __noinline__ unsigned outlier_fetch(const FastArgs& F, int f, int X, int Y, float* __restrict__ R9)
{
    const MergeArgs& A = F.a;
    const mfsr_merge_geom& g = A.g;
    const int2 s = shift_global(A, f, X, Y);
    const int Xs = X + s.x, Ys = Y + s.y;
    const int lox = 2 * g.clamp_x0 + 2, hix = 2 * g.clamp_x1 - 1, loy = 2 * g.clamp_y0 + 2, hiy = 2 * g.clamp_y1 - 1;
    if (!(abs(s.x) < (1 << 20) && abs(s.y) < (1 << 20) && Xs >= lox && Xs <= hix && Ys >= loy && Ys <= hiy)) return 0xFFFFFFFFu;
    const int k = Xs >> 1, ky = Ys >> 1, phx = k & 1, phy = ky & 1;
    const uint16_t* rawf = (const uint16_t*)((const char*)A.raw + A.raw_fs * f);
    unsigned rv[9];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const uint16_t* rrow = row_ptr(rawf, A.raw_pitch, ky - 1 + r) + (k - 1);
#pragma unroll
        for (int c = 0; c < 3; c++) rv[3 * r + c] = __ldg(rrow + c);
    }
    // CFA phase of sample (k-1+c, ky-1+r): (phy ^ !(r&1)) * 2 + (phx ^ !(c&1))
    const float iv_cc = F.inv_ph[phy * 2 + phx], nb_cc = F.nbi_ph[phy * 2 + phx];                  // centre
    const float iv_ch = F.inv_ph[phy * 2 + (phx ^ 1)], nb_ch = F.nbi_ph[phy * 2 + (phx ^ 1)];      // left / right
    const float iv_cv = F.inv_ph[(phy ^ 1) * 2 + phx], nb_cv = F.nbi_ph[(phy ^ 1) * 2 + phx];      // up / down
    const float iv_cd = F.inv_ph[(phy ^ 1) * 2 + (phx ^ 1)], nb_cd = F.nbi_ph[(phy ^ 1) * 2 + (phx ^ 1)];   // corners
    R9[0] = fmaf((float)rv[0], iv_cd, nb_cd); R9[1] = fmaf((float)rv[1], iv_cv, nb_cv); R9[2] = fmaf((float)rv[2], iv_cd, nb_cd);
    R9[3] = fmaf((float)rv[3], iv_ch, nb_ch); R9[4] = fmaf((float)rv[4], iv_cc, nb_cc); R9[5] = fmaf((float)rv[5], iv_ch, nb_ch);
    R9[6] = fmaf((float)rv[6], iv_cd, nb_cd); R9[7] = fmaf((float)rv[7], iv_cv, nb_cv); R9[8] = fmaf((float)rv[8], iv_cd, nb_cd);
    return (unsigned)(Xs & 1) | ((unsigned)(Ys & 1) << 1) | ((unsigned)(phx ^ 1) << 14) | ((unsigned)phy << 15);
}

// The frame loop of one output pixel: J = X % 4 and YM = Y % 4 are compile time (they fix which certainty cell every tap
// reads and therefore the slot tables); everything frame dependent comes from the 16-bit descriptor.
template <int TH, int J, int YM>
__device__ __forceinline__ void frame_loop(const FastArgs& F, int N, const unsigned char* dp, const unsigned char* rp, const unsigned char* mp, unsigned celloff,
                                           const float (&W)[mt::NW], unsigned kwin_s, int kx0, int ky0, int X, int Y, float (&acc)[4], float (&wacc)[4])
{
    using C = PCfg<TH>;
#pragma unroll 1
    for (int f = 0; f < N; f++, dp += C::DESC_BYTES, rp += C::RAW_BYTES, mp += MASK_FRAME_BYTES) {
        unsigned d = *(const unsigned short*)dp;
        const unsigned po = d & 0x3FFCu;
        float R[3][3];
        if (po < SPECIAL) {
            const float* pe = (const float*)(rp + po);
            const float* pc = pe + ((d & 0x4000u) ? 1 - RHALF : RHALF);
#pragma unroll
            for (int r = 0; r < 3; r++) { R[r][0] = pe[r * RWS]; R[r][1] = pc[r * RWS]; R[r][2] = pe[r * RWS + 1]; }
        } else {
            float R9[9];
            d = outlier_fetch(F, f, X, Y, R9);
            if (d == 0xFFFFFFFFu) {                 // clamped taps / non-finite shift: the reference loop
                float ab[8];
                special_pixel(F, kwin_s, C::KWS, kx0, ky0, f, X, Y, ab);
#pragma unroll
                for (int q = 0; q < 4; q++) { acc[q] += ab[q]; wacc[q] += ab[4 + q]; }
                continue;
            }
#pragma unroll
            for (int r = 0; r < 3; r++) { R[r][0] = R9[3 * r]; R[r][1] = R9[3 * r + 1]; R[r][2] = R9[3 * r + 2]; }
        }
        const unsigned rq0 = ((d >> 3) & 0x1800u) | celloff, rq1 = rq0 ^ 0x1000u;
        const float2* q0 = (const float2*)(mp + rq0);
        const float2* q1 = (const float2*)(mp + rq1);
        float Q[2][2][2][2];
#pragma unroll
        for (int mr = 0; mr < 2; mr++)
#pragma unroll
            for (int mc = 0; mc < 2; mc++) {
                const float2 v0 = q0[mr * MWS + mc], v1 = q1[mr * MWS + mc];
                Q[mr][mc][0][0] = v0.x; Q[mr][mc][0][1] = v0.y; Q[mr][mc][1][0] = v1.x; Q[mr][mc][1][1] = v1.y;
            }
        const float fx = (d & 1u) ? 1.0f : 0.0f, fy = (d & 2u) ? 1.0f : 0.0f;
        float t[4], u[4];
        ms::pixel_fold<J, YM>(W, fx, 1.0f - fx, fy, 1.0f - fy, Q, R, t, u);
        route_add(t, u, !(d & 0x4000u), (d & 0x8000u) != 0, acc, wacc);      // the window column parity bit is o = !phx
    }
}

// Phase 2 of one warp: tile row `row`, four passes J = 0..3 over the 32 pixels X = X0abs + 4*lane + J.  The per-pixel prologue
// (weights) and epilogue (CFA phase -> colour, ApplyWeighting kernel.cu:426, GammasRGB :393, one write) exist once; only the
// frame loop is specialised (16 copies selected by a switch).
template <int TH>
__device__ __forceinline__ void run_row(const FastArgs& F, const unsigned char* smem, int row, int x0, int y0, int X0abs, int Y0abs, float* ot_row)
{
    using C = PCfg<TH>;
    const MergeArgs& A = F.a;
    const mfsr_merge_geom& g = A.g;
    const int lane = threadIdx.x & 31;
    const int N = A.n_frames;
    const int y = y0 + row, Y = Y0abs + row;                  // window / absolute row
    if (y < F.in_y0 || y >= F.in_y1) return;                  // outside the window or in the clamp band
    const unsigned char* descS = smem;
    const unsigned char* rawS = smem + (size_t)N * C::DESC_BYTES;
    const unsigned char* maskS = rawS + (size_t)N * C::RAW_BYTES;
    const unsigned kwin_s = (unsigned)__cvta_generic_to_shared(maskS + (size_t)N * MASK_FRAME_BYTES);
    const int kx0 = (X0abs >> 1) - 1, ky0 = (Y0abs >> 1) - 1;
    const int mx0 = (X0abs >> 2) - 1, my0 = (Y0abs >> 2) - 1;
    const int YM = row & 3;                                   // Y0abs is a multiple of 4
    const unsigned rowcell = (unsigned)((((Y - 2) >> 2) - my0) * MWS);
#ifdef MFSR_PF_SKEW
    // two of the four warps of a scheduler start half a pass late: their prologues / epilogues (latency bound) then meet the other
    // two warps' frame loops (issue bound) instead of each other
    if ((row >> 2) & 1) __nanosleep(MFSR_PF_SKEW);
#endif

#pragma unroll 1
    for (int J = 0; J < 4; J++) {
        const int x = x0 + 4 * lane + J, X = X0abs + 4 * lane + J;
        if (x < F.in_x0 || x >= F.in_x1) continue;            // outside the window or in the clamp band (merge_band_kernel's pixels)
        float fb3[3] = {0.f, 0.f, 0.f};         // ApplyWeighting's inOutImg value, fetched early: its latency hides behind the frame loop
        // ot_row != null: the tile's fallback rows were brought to shared memory by TMA (and the result leaves the same way)
        if (!ot_row && A.fallback) { const float* fbp = row_ptr(A.fallback, A.fb_pitch, y) + 3 * x; fb3[0] = __ldg(fbp); fb3[1] = __ldg(fbp + 1); fb3[2] = __ldg(fbp + 2); }
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, wacc[4] = {0.f, 0.f, 0.f, 0.f};
        {
            float W[mt::NW];
            compute_weights(kwin_s, C::KWS, kx0, ky0, X, Y, W);
            const unsigned celloff = (rowcell + (unsigned)(((X - 2) >> 2) - mx0)) * 8u;     // certainty cell of tap -2, byte offset inside a plane
            const unsigned char* dp = descS + (row * TW + 4 * lane + J) * 2;
            switch (YM * 4 + J) {
#define MFSR_CASE(YMv, Jv) case YMv * 4 + Jv: frame_loop<TH, Jv, YMv>(F, N, dp, rawS, maskS, celloff, W, kwin_s, kx0, ky0, X, Y, acc, wacc); break;
                MFSR_CASE(0, 0) MFSR_CASE(0, 1) MFSR_CASE(0, 2) MFSR_CASE(0, 3) MFSR_CASE(1, 0) MFSR_CASE(1, 1) MFSR_CASE(1, 2) MFSR_CASE(1, 3)
                MFSR_CASE(2, 0) MFSR_CASE(2, 1) MFSR_CASE(2, 2) MFSR_CASE(2, 3) MFSR_CASE(3, 0) MFSR_CASE(3, 1) MFSR_CASE(3, 2)
                default: frame_loop<TH, 3, 3>(F, N, dp, rawS, maskS, celloff, W, kwin_s, kx0, ky0, X, Y, acc, wacc); break;
#undef MFSR_CASE
            }
        }
        // ---- epilogue: CFA phase -> colour by 0 / 1 blends, partial sums of a frame-chunked merge, normalisation, one write
        float s3[3], w3[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            s3[c] = acc[0] * F.cfa_sel[0][c] + acc[1] * F.cfa_sel[1][c] + acc[2] * F.cfa_sel[2][c] + acc[3] * F.cfa_sel[3][c];
            w3[c] = wacc[0] * F.cfa_sel[0][c] + wacc[1] * F.cfa_sel[1][c] + wacc[2] * F.cfa_sel[2][c] + wacc[3] * F.cfa_sel[3][c];
        }
        // (row pointers are formed here, after the frame loop: held across it they were spilled and reloaded from local memory)
        if (A.sum_in) {
            const float* si = row_ptr(A.sum_in, A.acc_pitch, y) + 3 * x; const float* wi = row_ptr(A.weight_in, A.acc_pitch, y) + 3 * x;
#pragma unroll
            for (int c = 0; c < 3; c++) { s3[c] = si[c] + s3[c]; w3[c] = wi[c] + w3[c]; }
        }
        if (A.sum_out) {
            float* so = row_ptr(A.sum_out, A.acc_pitch, y) + 3 * x; float* wo = row_ptr(A.weight_out, A.acc_pitch, y) + 3 * x;
#pragma unroll
            for (int c = 0; c < 3; c++) { so[c] = s3[c]; wo[c] = w3[c]; }
        }
        if (!(A.flags & MFSR_MERGE_PARTIAL_INTERNAL)) {
            if (ot_row) {
                float* o = ot_row + 3 * (4 * lane + J);
#pragma unroll
                for (int c = 0; c < 3; c++) o[c] = finish_px(apply_weighting(s3[c], w3[c], o[c], A.threshold), A.flags);
            } else {
                float* orow = row_ptr(A.out, A.out_pitch, y);
#pragma unroll
                for (int c = 0; c < 3; c++) orow[3 * x + c] = finish_px(apply_weighting(s3[c], w3[c], fb3[c], A.threshold), A.flags);
            }
        }
    }
}

// round-half-away-from-zero of |v| as an integer, without F2I (8 cycles per warp on the conversion pipe): two round-toward-zero
// adds leave floor(|v| + 0.5) in the mantissa of 2^23 + n.  |v| >= 2^22, NaN and Inf give a large positive number (the caller's
// table index saturates to the poison entry).
__device__ __forceinline__ int round_mag(float v)
{
    float a, b;
    asm("add.rz.f32 %0, %1, 0f3F000000;" : "=f"(a) : "f"(fabsf(v)));
    asm("add.rz.f32 %0, %1, 0f4B000000;" : "=f"(b) : "f"(a));
    return __float_as_int(b) - 0x4B000000;
}

__device__ __forceinline__ unsigned lds_u32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Phases 0 (descriptors), 1a (certainty planes, kernel-parameter window) and 1b (raw windows) of one tile.
// BORDER = false: the tile's taps and every staged window lie inside the clamp range (hence inside the image): no index clamping
// anywhere, row / column parities are frame independent.  BORDER = true: the general code.
template <int TH, bool BORDER>
__device__ __forceinline__ void stage_tile(const FastArgs& F, unsigned char* smem, const int2* fbase, int x0, int y0, int X0abs, int Y0abs)
{
    using C = PCfg<TH>;
    const MergeArgs& A = F.a;
    const mfsr_merge_geom& g = A.g;
    const int N = A.n_frames;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* descS = smem;
    unsigned char* rawS = smem + (size_t)N * C::DESC_BYTES;
    unsigned char* maskS = rawS + (size_t)N * C::RAW_BYTES;
    unsigned char* kernS = maskS + (size_t)N * MASK_FRAME_BYTES;
    const unsigned colA = (unsigned)__cvta_generic_to_shared(kernS + C::KERN_BYTES), rowA = colA + 4u * (C::VN + 1);
    const int raw_w = g.raw_w, raw_h = g.raw_h;

    // ---------------- phase 0: one 16-bit descriptor per tile pixel and frame.
    // Work item = (frame, row pair): 8 pixels per lane from a 3 x 4 flow window; two items in flight per warp.
    {
        const int fxb = 2 * ((X0abs >> 2) + lane) - 1;
        int cx[4];
#pragma unroll
        for (int c = 0; c < 4; c++) cx[c] = BORDER ? clampi(fxb + c, 0, raw_w - 1) : fxb + c;
        const int total = (TH / 2) * N;
        const int lox = 2 * g.clamp_x0 + 2, spanx = 2 * g.clamp_x1 - 1 - lox, loy = 2 * g.clamp_y0 + 2, spany = 2 * g.clamp_y1 - 1 - loy;
        const int64_t fpitch = A.flow_pitch, ffs = A.flow_fs;
        const char* flow0 = (const char*)A.flow;
        const int Xl = X0abs + 4 * lane;
        auto load = [&](int item, float2 (&Fl)[3][4]) {
            const int f = item / (TH / 2), rp = item - f * (TH / 2);
            const int a = (Y0abs >> 1) + rp;
            const char* fl = flow0 + ffs * f;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const int yy = BORDER ? clampi(a - 1 + r, 0, raw_h - 1) : a - 1 + r;
                const float2* fr = (const float2*)(fl + fpitch * yy);
#pragma unroll
                for (int c = 0; c < 4; c++) Fl[r][c] = __ldg(fr + cx[c]);
            }
        };
        auto emit = [&](int item, const float2 (&Fl)[3][4]) {
            const int f = item / (TH / 2), rp = item - f * (TH / 2);
            const int2 fb = fbase[f];
            const int Yl = Y0abs + 2 * rp;
            const int cX = Xl - 2 - 2 * fb.x, cY = Yl - 2 - 2 * fb.y;
            // horizontal mixes once per flow row: even column 2k uses (k-1, k) frac .75, odd column (k, k+1) frac .25
            float2 H[3][4];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c0 = (j + 1) >> 1;
                    const bool ax = !(j & 1);
                    H[r][j].x = mix25(Fl[r][c0].x, Fl[r][c0 + 1].x, ax);
                    H[r][j].y = mix25(Fl[r][c0].y, Fl[r][c0 + 1].y, ax);
                }
            unsigned packed[2][2];
#pragma unroll
            for (int yy = 0; yy < 2; yy++) {
                const bool ay = (yy == 0);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // round(2 * shift): the factor 2 is exact, the rounding is half away from zero like roundf (:403)
                    const float vx = __fmul_rn(mix25(H[yy][j].x, H[yy + 1][j].x, ay), 2.0f);
                    const float vy = __fmul_rn(mix25(H[yy][j].y, H[yy + 1][j].y, ay), 2.0f);
                    const int nx = round_mag(vx), ny = round_mag(vy);
                    const int v = (vx < 0.0f) ? cX + j - nx : cX + j + nx;          // X - 2 - 2 fbx + sx
                    const int u = (vy < 0.0f) ? cY + yy - ny : cY + yy + ny;
                    unsigned dsc = lds_u32(colA + 4u * min((unsigned)v, (unsigned)C::VN)) + lds_u32(rowA + 4u * min((unsigned)u, (unsigned)C::UN));
                    if (BORDER) {
                        const int Xs = v + 2 + 2 * fb.x, Ys = u + 2 + 2 * fb.y;
                        const bool noclamp = spanx >= 0 && spany >= 0 && (unsigned)(Xl + j - lox) <= (unsigned)spanx && (unsigned)(Yl + yy - loy) <= (unsigned)spany &&
                                             (unsigned)(Xs - lox) <= (unsigned)spanx && (unsigned)(Ys - loy) <= (unsigned)spany;
                        if (!noclamp) dsc = POISON;
                    }
                    dsc = min(dsc, 0xFFFFu);
                    if (j & 1) packed[yy][j >> 1] |= dsc << 16; else packed[yy][j >> 1] = dsc;
                }
            }
            unsigned char* sh = descS + (size_t)f * C::DESC_BYTES;
            *(uint2*)(sh + ((2 * rp) * TW + 4 * lane) * 2) = make_uint2(packed[0][0], packed[0][1]);
            *(uint2*)(sh + ((2 * rp + 1) * TW + 4 * lane) * 2) = make_uint2(packed[1][0], packed[1][1]);
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
    // ---------------- phase 1a: certainty planes.  A thread owns one cell of the window and walks every other frame
    // (two thread groups), four loads in flight.
    {
        const int mw = raw_w / 2, mh = raw_h / 2;
        const int mx0 = (X0abs >> 2) - 1, my0 = (Y0abs >> 2) - 1;
        constexpr int PL = C::MHS * MWS;
        static_assert(2 * PL <= C::NT, "two thread groups per certainty window");
        if (tid < 2 * PL) {
            const int grp = tid >= PL ? 1 : 0, cell = tid - grp * PL;
            const int r = cell / MWS, c = cell - r * MWS;
            const int my = BORDER ? clampi(my0 + r, 0, mh - 1) : my0 + r, mx = BORDER ? clampi(mx0 + c, 0, mw - 1) : mx0 + c;
            const char* src = (const char*)A.mask + A.mask_pitch * my + 16 * mx;
            const int64_t mfs = A.mask_fs;
            unsigned char* dst = maskS + cell * 8;
            const float s00 = F.cfa_sel[0][0], s01 = F.cfa_sel[0][1], s02 = F.cfa_sel[0][2], s10 = F.cfa_sel[1][0], s11 = F.cfa_sel[1][1], s12 = F.cfa_sel[1][2];
            const float s20 = F.cfa_sel[2][0], s21 = F.cfa_sel[2][1], s22 = F.cfa_sel[2][2], s30 = F.cfa_sel[3][0], s31 = F.cfa_sel[3][1], s32 = F.cfa_sel[3][2];
            for (int f0 = grp; f0 < N; f0 += 8) {
                float4 m4[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (f0 + 2 * k < N) m4[k] = __ldg((const float4*)(src + mfs * (f0 + 2 * k)));
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (f0 + 2 * k < N) {
                        const float4 m = m4[k];
                        const float c0 = isfinite(m.x) ? m.x : 0.f, c1 = isfinite(m.y) ? m.y : 0.f, c2 = isfinite(m.z) ? m.z : 0.f;   // :438-439
                        // certainty per CFA phase q: the channel of colour cfa[q], picked by a 0 / 1 blend (no compare chains)
                        const float q0 = c0 * s00 + c1 * s01 + c2 * s02, q1 = c0 * s10 + c1 * s11 + c2 * s12;
                        const float q2 = c0 * s20 + c1 * s21 + c2 * s22, q3 = c0 * s30 + c1 * s31 + c2 * s32;
                        // plane (o, rho) serves window centres of column parity phx = !o and row parity phy = rho; element = (x class 0, x class 1);
                        // the plane of the other row parity (y class 1) is 2 planes away (rq ^ 0x1000 in the frame loop)
                        unsigned char* mb = dst + (size_t)(f0 + 2 * k) * MASK_FRAME_BYTES;
                        *(float2*)(mb + 0 * PLANE_BYTES) = make_float2(q1, q0);      // o = 0, rho = 0: phase (y 0, x 1)
                        *(float2*)(mb + 1 * PLANE_BYTES) = make_float2(q0, q1);      // o = 1, rho = 0: phase (0, 0)
                        *(float2*)(mb + 2 * PLANE_BYTES) = make_float2(q3, q2);      // o = 0, rho = 1: phase (1, 1)
                        *(float2*)(mb + 3 * PLANE_BYTES) = make_float2(q2, q3);      // o = 1, rho = 1: phase (1, 0)
                    }
            }
        }
        // kernel parameters (texture clamp addressing applied here)
        float4* ks = (float4*)kernS;
        const int kx0 = (X0abs >> 1) - 1, ky0 = (Y0abs >> 1) - 1;
        for (int i = tid; i < C::KWS * C::KHS; i += C::NT) {
            const int r = i / C::KWS, c = i - r * C::KWS;
            ks[i] = __ldg(row_ptr(A.kern, A.kern_pitch, clampi(ky0 + r, 0, raw_h - 1)) + clampi(kx0 + c, 0, raw_w - 1));
        }
    }
    // ---------------- phase 1b: normalised raw windows, de-interleaved by column parity.  Thread t owns the 4-column chunk position t
    // of the window and walks the frames four at a time with the loads in flight.  u16 -> float without the conversion pipe:
    // PRMT builds 2^23 + v, one FADD removes 2^23 (exact), one FFMA normalises.
    {
        constexpr int CH = C::RHS * (RWS / 4);
        constexpr int FSTRIDE = C::RAW_BYTES / 4;
        const unsigned rpitch = (unsigned)A.raw_pitch;
        const int64_t rfs = A.raw_fs;
        const char* raw0 = (const char*)A.raw;
        auto cvt4 = [](uint2 pv, float be, float bo, float ie, float io, float* rs) {
            // (v - black) * inv as v * inv - black * inv: be / bo hold -black * inv of the even / odd column's colour
            const float e0 = __uint_as_float(__byte_perm(pv.x, 0x4B000000u, 0x7410)) - 8388608.0f, o0 = __uint_as_float(__byte_perm(pv.x, 0x4B000000u, 0x7432)) - 8388608.0f;
            const float e1 = __uint_as_float(__byte_perm(pv.y, 0x4B000000u, 0x7410)) - 8388608.0f, o1 = __uint_as_float(__byte_perm(pv.y, 0x4B000000u, 0x7432)) - 8388608.0f;
            *(float2*)rs = make_float2(fmaf(e0, ie, be), fmaf(e1, ie, be));
            *(float2*)(rs + RHALF) = make_float2(fmaf(o0, io, bo), fmaf(o1, io, bo));
        };
        float* win0 = (float*)rawS;
        const float nb0 = F.nbi_ph[0], nb1 = F.nbi_ph[1], nb2 = F.nbi_ph[2], nb3 = F.nbi_ph[3];
        const float iv0 = F.inv_ph[0], iv1 = F.inv_ph[1], iv2 = F.inv_ph[2], iv3 = F.inv_ph[3];
        auto walk = [&](int r, int c4, int f_lo, int f_step) {
            float* rs = win0 + r * RWS + 2 * c4;
            // interior tiles: window rows have the parity of fbase.y + r (fbase.y is odd for every frame) and nothing is clamped, so the
            // normalisation constants of the row are chosen once
            const int odd_static = (r + 1) & 1;
            float be = odd_static ? nb2 : nb0, bo = odd_static ? nb3 : nb1, ie = odd_static ? iv2 : iv0, io = odd_static ? iv3 : iv1;
            for (int f0 = f_lo; f0 < N; f0 += 4 * f_step) {
                uint2 p4[4]; int o4[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int f = f0 + k * f_step;
                    if (f < N) {
                        const int2 fi = fbase[f];
                        const char* fbp = raw0 + rfs * f;
                        if (!BORDER) {
                            p4[k] = __ldg((const uint2*)(fbp + ((unsigned)(fi.y + r) * rpitch + 2u * (unsigned)(fi.x + 4 * c4))));
                        } else {
                            const int yy = clampi(fi.y + r, 0, raw_h - 1), xx = fi.x + 4 * c4;        // xx is a multiple of 4
                            o4[k] = yy & 1;
                            if ((unsigned)xx <= (unsigned)(raw_w - 4)) p4[k] = __ldg((const uint2*)(fbp + ((unsigned)yy * rpitch + 2u * (unsigned)xx)));
                            else {
                                const uint16_t* rrow = (const uint16_t*)(fbp + (unsigned)yy * rpitch);
                                unsigned v[4];
#pragma unroll
                                for (int q = 0; q < 4; q++) v[q] = __ldg(rrow + clampi(xx + q, 0, raw_w - 1));
                                p4[k] = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int f = f0 + k * f_step;
                    if (f < N) {
                        if (BORDER) { const int odd = o4[k]; be = odd ? nb2 : nb0; bo = odd ? nb3 : nb1; ie = odd ? iv2 : iv0; io = odd ? iv3 : iv1; }
                        cvt4(p4[k], be, bo, ie, io, rs + f * FSTRIDE);
                    }
                }
            }
        };
        if (tid < CH) walk(tid / (RWS / 4), tid % (RWS / 4), 0, 1);
        if (CH > C::NT) {
            // the CH - NT positions beyond the block size: thread t takes position NT + t % REST and frames t / REST, + NT / REST, ...
            constexpr int REST = CH > C::NT ? CH - C::NT : 1;
            constexpr int GR = C::NT / REST;
            if (tid < GR * REST) {
                const int pp = C::NT + tid % REST;
                walk(pp / (RWS / 4), pp % (RWS / 4), tid / REST, GR);
            }
        }
    }
}

// L2 prefetch of everything the staging phases of tile (x0, y0) will read (certainty rows, kernel-parameter rows, fallback rows, the
// UNSHIFTED raw window and the flow window): issued while the previous tile of this persistent CTA is in its frame loops, so the
// staging loads of the next tile meet L2 instead of DRAM.  Rows / columns are clamped into the images; a line too many is harmless.
template <int TH>
__device__ __forceinline__ void prefetch_tile(const FastArgs& F, int x0, int y0)
{
    using C = PCfg<TH>;
    const MergeArgs& A = F.a;
    const mfsr_merge_geom& g = A.g;
    const int N = A.n_frames, tid = threadIdx.x;
    const int X0abs = x0 + g.org_x, Y0abs = y0 + g.org_y;
    auto pf = [](const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
    const int mw = g.raw_w / 2, mh = g.raw_h / 2;
    const int mx0 = clampi((X0abs >> 2) - 1, 0, mw - 1), my0 = (Y0abs >> 2) - 1;
    constexpr int ML = (MWS * 16 + 127) / 128 + 1;
    for (int i = tid; i < N * C::MHS * ML; i += C::NT) {
        const int f = i / (C::MHS * ML), j = i - f * (C::MHS * ML), r = j / ML, l = j - r * ML;
        pf((const char*)A.mask + A.mask_fs * f + A.mask_pitch * clampi(my0 + r, 0, mh - 1) + 16 * min(mx0 + l * 8, mw - 1));
    }
    const int kx0 = clampi((X0abs >> 1) - 1, 0, g.raw_w - 1), ky0 = (Y0abs >> 1) - 1;
    constexpr int KL = (C::KWS * 16 + 127) / 128 + 1;
    for (int i = tid; i < C::KHS * KL; i += C::NT) {
        const int r = i / KL, l = i - r * KL;
        pf(row_ptr(A.kern, A.kern_pitch, clampi(ky0 + r, 0, g.raw_h - 1)) + min(kx0 + l * 8, g.raw_w - 1));
    }
    constexpr int FLL = (C::KWS * 8 + 127) / 128 + 1;                        // flow window: same footprint as the kernel-parameter window
    for (int i = tid; i < N * C::KHS * FLL; i += C::NT) {
        const int f = i / (C::KHS * FLL), j = i - f * (C::KHS * FLL), r = j / FLL, l = j - r * FLL;
        pf((const char*)A.flow + A.flow_fs * f + A.flow_pitch * clampi(ky0 + r, 0, g.raw_h - 1) + 8 * min(kx0 + l * 16, g.raw_w - 1));
    }
    constexpr int RL = (TW / 2 + 32) * 2 / 128 + 2;
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

// Persistent kernel: one CTA per SM walks the tiles (tile index = blockIdx.x + k * gridDim.x, row-major: concurrent CTAs work on
// neighbouring tiles and share their halo lines in L2).
template <int TH>
__global__ void __launch_bounds__(PCfg<TH>::NT, 1)
merge_pf_kernel(const __grid_constant__ FastArgs F, int tiles_x, int n_tiles)
{
    using C = PCfg<TH>;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int2 fbase[MAXF];           // origin (rx0, ry0) of the staged raw window per frame
    __shared__ int s_border;               // 1: some pixel-frame of the tile may touch the clamp range
    __shared__ __align__(8) unsigned long long s_mbar[TH];  // completion of each row's TMA fallback load (a warp owns a row)
    const MergeArgs& A = F.a;
    const mfsr_merge_geom& g = A.g;
    const int N = A.n_frames;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned mbar_s = (unsigned)__cvta_generic_to_shared(&s_mbar[warp]);
    if (lane == 0) mbar_init(mbar_s, 1);
    __syncwarp();
    unsigned tma_phase = 0;
    unsigned char* kernS = smem + (size_t)N * (C::DESC_BYTES + C::RAW_BYTES + MASK_FRAME_BYTES);
    unsigned* colT = (unsigned*)(kernS + C::KERN_BYTES);
    unsigned* rowT = colT + C::VN + 1;
    float* otS = (float*)(kernS + C::KERN_BYTES + C::TAB_BYTES);

    // ---------------- lookup tables of phase 0 (same for every tile and frame: indices are relative to the window origin)
    for (int i = tid; i <= C::VN + C::UN + 1; i += C::NT) {
        if (i <= C::VN) {
            const int v = i, cc = v >> 1;
            colT[i] = (v == C::VN) ? POISON : (unsigned)(((cc >> 1) + (cc & 1) * RHALF) * 4) | (unsigned)(v & 1) | ((unsigned)(cc & 1) << 14);
        } else {
            const int u = i - C::VN - 1, r0 = u >> 1;
            rowT[u] = (u == C::UN) ? POISON : (unsigned)(r0 * RWS * 4) | ((unsigned)(u & 1) << 1) | ((unsigned)(r0 & 1) << 15);
        }
    }
    int tile = blockIdx.x;
    if (tile < n_tiles) prefetch_tile<TH>(F, (tile % tiles_x) * TW - F.x_off, (tile / tiles_x) * TH - F.y_off);

    for (; tile < n_tiles; tile += gridDim.x) {
        const int x0 = (tile % tiles_x) * TW - F.x_off, y0 = (tile / tiles_x) * TH - F.y_off;     // window coords of the tile origin
        const int X0abs = x0 + g.org_x, Y0abs = y0 + g.org_y;                                  // multiples of 4
        // ---------------- window origins: the tile's own footprint displaced by the mean shift of a 32-point sample per frame
        // (spare rows / columns split evenly; x a multiple of 4 for the 8-byte raw loads, y ODD so that the row parity of the
        // window equals the absolute row parity)
        if (tid == 0) s_border = 0;
        __syncthreads();                   // also: the previous tile's frame loops are done with the shared-memory windows
        // ---------------- TMA: the warp's fallback row on its way to shared memory (needed by the epilogues, long after).  Everything
        // about a row (load, mbarrier, result, store) belongs to the warp that owns it: no block-level synchronisation is added.
#ifdef MFSR_PF_NO_TMA
        const bool tma = false;
#else
        const bool tma = F.use_tma && x0 >= 0 && x0 + TW <= g.out_w && y0 + warp >= 0 && y0 + warp < g.out_h;     // a whole row segment inside the image
#endif
        float* ot_row = otS + warp * (TW * 3);
        if (tma && lane == 0) {
            mbar_expect_tx(mbar_s, (unsigned)C::OT_ROW_BYTES);
            tma_load_row((unsigned)__cvta_generic_to_shared(ot_row), row_ptr(A.fallback, A.fb_pitch, y0 + warp) + 3 * x0, C::OT_ROW_BYTES, mbar_s);
        }
        for (int f = warp; f < N; f += C::NW_) {
            const float2* flow = (const float2*)((const char*)A.flow + A.flow_fs * f);
            const int sxp = clampi((X0abs >> 1) + 4 + 8 * (lane & 7), 0, g.raw_w - 1);
            const int syp = clampi((Y0abs >> 1) + ((TH / 2) * (2 * (lane >> 3) + 1)) / 8, 0, g.raw_h - 1);
            const float2 v = __ldg(row_ptr(flow, A.flow_pitch, syp) + sxp);
            const bool ok = fabsf(v.x) < 1.0e4f && fabsf(v.y) < 1.0e4f;
            int sx = ok ? __float2int_rn(2.0f * v.x) : 0, sy = ok ? __float2int_rn(2.0f * v.y) : 0, cnt = ok ? 1 : 0;
            sx = __reduce_add_sync(0xffffffffu, sx); sy = __reduce_add_sync(0xffffffffu, sy); cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0) {
                cnt = max(cnt, 1);
                const int mx = (int)floorf((float)sx / (float)cnt + 0.5f), my = (int)floorf((float)sy / (float)cnt + 0.5f);
                const int2 fb = make_int2((((X0abs + mx - 2) >> 1) - (RWS - (TW / 2 + 3)) / 2) & ~3,
                                          (((Y0abs + my - 2) >> 1) - (C::RHS - (TH / 2 + 3)) / 2) | 1);
                fbase[f] = fb;
                // clamp range (:414-419): the tile is "interior" when neither its own taps nor any staged window can touch it
                const int lox = 2 * g.clamp_x0 + 2, hix = 2 * g.clamp_x1 - 1, loy = 2 * g.clamp_y0 + 2, hiy = 2 * g.clamp_y1 - 1;
                const bool inside = X0abs >= lox && X0abs + TW - 1 <= hix && Y0abs >= loy && Y0abs + TH - 1 <= hiy &&
                                    fb.x >= g.clamp_x0 && fb.x + RWS - 1 <= g.clamp_x1 && fb.y >= g.clamp_y0 && fb.y + C::RHS - 1 <= g.clamp_y1 &&
                                    (X0abs >> 1) - 2 >= 0 && (X0abs >> 1) + TW / 2 + 2 < g.raw_w && (Y0abs >> 1) - 2 >= 0 && (Y0abs >> 1) + TH / 2 + 2 < g.raw_h;
                if (!inside) s_border = 1;
            }
        }
        __syncthreads();
        // ---------------- phases 0 and 1 (independent of each other: the window origins are already known): interior tiles run the
        // variant without clamp handling
        if (s_border) stage_tile<TH, true>(F, smem, fbase, x0, y0, X0abs, Y0abs);
        else stage_tile<TH, false>(F, smem, fbase, x0, y0, X0abs, Y0abs);
        __syncthreads();
        // ---------------- the next tile's lines on their way to L2 while this one computes
        {
            const int nt = tile + gridDim.x;
            if (nt < n_tiles) prefetch_tile<TH>(F, (nt % tiles_x) * TW - F.x_off, (nt / tiles_x) * TH - F.y_off);
        }
        // ---------------- phase 2: warp w owns tile row w: four passes J = 0..3, each a loop over the frames
        if (tma) { mbar_wait(mbar_s, tma_phase); tma_phase ^= 1u; }
        run_row<TH>(F, smem, warp, x0, y0, X0abs, Y0abs, tma ? ot_row : nullptr);
        if (tma) {
            // ---------------- TMA: the finished row leaves as one bulk store (the warp's generic-proxy writes made visible to the async proxy)
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_row(row_ptr(A.out, A.out_pitch, y0 + warp) + 3 * x0, (unsigned)__cvta_generic_to_shared(ot_row), C::OT_ROW_BYTES);
                tma_store_commit_wait();
            }
        }
    }
}

// The clamp band: output pixels within BAND HR pixels of the clamp range (:414-419) — where a tap, shifted or not, can be clamped —
// and the 1-pixel border of the window that the reference skips (:391).  One thread per band pixel runs the reference loop
// (generic_pixel) for every frame.  In the tile kernel such a pixel would occupy a whole warp for ~1200 instructions per frame
// (round 2, ncu: 1.6 % of the warp iterations cost 10 % of the kernel's time); here 32 of them share a warp.
constexpr int BAND = 8;
__global__ void __launch_bounds__(128)
merge_band_kernel(const __grid_constant__ FastArgs F, int n_top, int n_mid_rows, int n_left, int n_right, int total)
{
    const MergeArgs& A = F.a;
    const mfsr_merge_geom& g = A.g;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    // strips: rows [0, in_y0) and [in_y1, out_h) over the full width, then the left / right parts of the rows in between
    int x, y;
    const int n_bot_start = n_top * g.out_w, n_bot = (g.out_h - F.in_y1) * g.out_w;
    if (i < n_bot_start) { y = i / g.out_w; x = i - y * g.out_w; }
    else if (i < n_bot_start + n_bot) { const int j = i - n_bot_start; y = j / g.out_w; x = j - y * g.out_w; y += F.in_y1; }
    else {
        const int j = i - n_bot_start - n_bot, per = n_left + n_right;
        y = j / per; x = j - y * per; y += F.in_y0;
        if (x >= n_left) x = F.in_x1 + (x - n_left);
    }
    (void)n_mid_rows;
    const int X = x + g.org_x, Y = y + g.org_y;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, wacc[4] = {0.f, 0.f, 0.f, 0.f};
    if (x >= 1 && x < g.out_w - 1 && y >= 1 && y < g.out_h - 1) {
        // 13 weights from global memory (texture model of :401, clamp addressing)
        float wl[mt::NW];
        {
            const int fx = (X - 1) >> 1, fy = (Y - 1) >> 1;
            const int xa = clampi(fx, 0, g.raw_w - 1), xb = clampi(fx + 1, 0, g.raw_w - 1), ya = clampi(fy, 0, g.raw_h - 1), yb = clampi(fy + 1, 0, g.raw_h - 1);
            const float4 K00 = __ldg(row_ptr(A.kern, A.kern_pitch, ya) + xa), K10 = __ldg(row_ptr(A.kern, A.kern_pitch, ya) + xb);
            const float4 K01 = __ldg(row_ptr(A.kern, A.kern_pitch, yb) + xa), K11 = __ldg(row_ptr(A.kern, A.kern_pitch, yb) + xb);
            const float ta = (X & 1) ? 0.25f : 0.75f, tb = (Y & 1) ? 0.25f : 0.75f;
            const float kx = tex_mix(K00.x, K10.x, K01.x, K11.x, ta, tb) * -0.72134752044448170368f;
            const float ky = tex_mix(K00.y, K10.y, K01.y, K11.y, ta, tb) * -0.72134752044448170368f;
            const float kz = tex_mix(K00.z, K10.z, K01.z, K11.z, ta, tb) * -0.72134752044448170368f;
#pragma unroll
            for (int py = 0; py <= 2; py++)
#pragma unroll
                for (int px = -2; px <= 2; px++) {
                    if (py == 0 && px < 0) continue;
                    const float q = (float)(px * px) * kx + (float)(2 * px * py) * kz + (float)(py * py) * ky;
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
                    if (!(fabsf(e) < INFINITY)) e = (px * py == 0) ? 1.0f : 0.0f;      // :429-430
                    wl[mt::widx(px, py)] = e;
                }
        }
        for (int f = 0; f < A.n_frames; f++) {
            const int2 s = shift_global(A, f, X, Y);
            float ab[8];
            generic_pixel(F, f, X, Y, s.x, s.y, wl, ab);
#pragma unroll
            for (int q = 0; q < 4; q++) { acc[q] += ab[q]; wacc[q] += ab[4 + q]; }
        }
    }
    float s3[3], w3[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        s3[c] = acc[0] * F.cfa_sel[0][c] + acc[1] * F.cfa_sel[1][c] + acc[2] * F.cfa_sel[2][c] + acc[3] * F.cfa_sel[3][c];
        w3[c] = wacc[0] * F.cfa_sel[0][c] + wacc[1] * F.cfa_sel[1][c] + wacc[2] * F.cfa_sel[2][c] + wacc[3] * F.cfa_sel[3][c];
    }
    if (A.sum_in) {
        const float* si = row_ptr(A.sum_in, A.acc_pitch, y) + 3 * x; const float* wi = row_ptr(A.weight_in, A.acc_pitch, y) + 3 * x;
#pragma unroll
        for (int c = 0; c < 3; c++) { s3[c] = si[c] + s3[c]; w3[c] = wi[c] + w3[c]; }
    }
    if (A.sum_out) {
        float* so = row_ptr(A.sum_out, A.acc_pitch, y) + 3 * x; float* wo = row_ptr(A.weight_out, A.acc_pitch, y) + 3 * x;
#pragma unroll
        for (int c = 0; c < 3; c++) { so[c] = s3[c]; wo[c] = w3[c]; }
    }
    if (!(A.flags & MFSR_MERGE_PARTIAL_INTERNAL)) {
        float fb3[3] = {0.f, 0.f, 0.f};
        if (A.fallback) { const float* fp = row_ptr(A.fallback, A.fb_pitch, y) + 3 * x; fb3[0] = fp[0]; fb3[1] = fp[1]; fb3[2] = fp[2]; }
        float* o = row_ptr(A.out, A.out_pitch, y) + 3 * x;
#pragma unroll
        for (int c = 0; c < 3; c++) o[c] = finish_px(apply_weighting(s3[c], w3[c], fb3[c], A.threshold), A.flags);
    }
}

template <int TH>
int launch_th(const FastArgs& Fin, cudaStream_t st)
{
    using C = PCfg<TH>;
    FastArgs F = Fin;
    const mfsr_merge_geom& g = F.a.g;
    // TMA bulk copies of the fallback / result rows (cp.async.bulk + mbarrier, one row per warp) need 16-byte aligned row segments:
    // window origin on the tile grid, pitches and bases multiples of 16, a fallback image, and an image (not only partial sums) to
    // write.  OFF by default: measured on one box (tools/ab_merge.sh, 12 MP x 8) the kernel takes 5.90 ms with them against 5.66 ms
    // without — the fallback loads were already hidden behind the frame loop, the bulk store adds a proxy fence and a wait per warp,
    // and the 24.5 KB tile buffer comes out of the L1 that the staging loads use.  MFSR_PF_TMA=1 switches them on.
    {
        static const char* te = getenv("MFSR_PF_TMA");
        const bool want = te && te[0] == '1';
        F.use_tma = want && F.x_off == 0 && F.a.fallback && !(F.a.flags & MFSR_MERGE_PARTIAL_INTERNAL) &&
                    !(((uintptr_t)F.a.fallback | (uintptr_t)F.a.out | (uintptr_t)F.a.fb_pitch | (uintptr_t)F.a.out_pitch) & 15);
    }
    if (F.use_tma && C::smem_bytes(F.a.n_frames, true) > (size_t)225 * 1024) F.use_tma = 0;      // the tile buffer does not fit beside this many frames
    const size_t smem = C::smem_bytes(F.a.n_frames, F.use_tma != 0);
    // opt-in shared memory (227 KB per block on sm_100 minus this instantiation's static tables); the attribute belongs to the
    // current device's context, so it is set per device (a function-static flag broke the second GPU of a process)
    static size_t max_dyn[64] = {0};
    int dev = 0;
    MFSR_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return MFSR_E_INVALID;
    if (!max_dyn[dev]) {
        cudaFuncAttributes at;
        MFSR_CUDA_TRY(cudaFuncGetAttributes(&at, merge_pf_kernel<TH>));
        const size_t lim = (size_t)227 * 1024 - at.sharedSizeBytes;
        MFSR_CUDA_TRY(cudaFuncSetAttribute(merge_pf_kernel<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim));
        max_dyn[dev] = lim;
    }
    if (smem > max_dyn[dev]) return MFSR_E_INVALID;
    // interior rectangle of the window (window coordinates): pixels at least BAND HR pixels inside the clamp range and off the
    // window's 1-pixel border; everything else is merge_band_kernel's
    {
        const int lox = 2 * g.clamp_x0 + 2, hix = 2 * g.clamp_x1 - 1, loy = 2 * g.clamp_y0 + 2, hiy = 2 * g.clamp_y1 - 1;
        int ax0 = lox + BAND - g.org_x, ax1 = hix - BAND - g.org_x + 1, ay0 = loy + BAND - g.org_y, ay1 = hiy - BAND - g.org_y + 1;
        ax0 = ax0 < 1 ? 1 : ax0; ay0 = ay0 < 1 ? 1 : ay0;
        ax1 = ax1 > g.out_w - 1 ? g.out_w - 1 : ax1; ay1 = ay1 > g.out_h - 1 ? g.out_h - 1 : ay1;
        if (ax1 <= ax0 || ay1 <= ay0) { ax0 = ax1 = 0; ay0 = 0; ay1 = 0; }          // no interior: the band kernel takes the whole window
        F.in_x0 = ax0; F.in_x1 = ax1; F.in_y0 = ay0; F.in_y1 = ay1;
    }
    const int tiles_x = cdiv(g.out_w + F.x_off, TW), tiles_y = cdiv(g.out_h + F.y_off, TH);
    static int n_sm[64] = {0};
    if (!n_sm[dev]) MFSR_CUDA_TRY(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev));
    const int n_tiles = tiles_x * tiles_y;
    if (F.in_x1 > F.in_x0) {
        // one CTA per tile by default: the persistent grid (MFSR_PF_GRID=persistent: one CTA per SM walking the tiles with the next
        // tile's lines prefetched to L2) measured 2.5 % slower — its per-tile barrier costs more than the prefetch saves
        static const char* ge = getenv("MFSR_PF_GRID");
        const bool full = !(ge && ge[0] == 'p');
        merge_pf_kernel<TH><<<(full || n_tiles < n_sm[dev]) ? n_tiles : n_sm[dev], C::NT, smem, st>>>(F, tiles_x, n_tiles);
        MFSR_LAUNCH_CHECK();
    }
    {
        const int n_top = F.in_y0, n_mid = F.in_y1 - F.in_y0, n_left = F.in_x0, n_right = g.out_w - F.in_x1;
        const long long total = (long long)(n_top + (g.out_h - F.in_y1)) * g.out_w + (long long)n_mid * (n_left + n_right);
        if (total >= (1ll << 31)) return MFSR_E_INVALID;
        if (total > 0) merge_band_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(F, n_top, n_mid, n_left, n_right, (int)total);
    }
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

}  // namespace

// frames the 16-row tile keeps resident (the pipeline and the tests ask; the bench's config 2 has 8)
namespace s2 {
int merge_pf_capacity() { return (int)((227 * 1024 - 2048 - PCfg<16>::KERN_BYTES - PCfg<16>::TAB_BYTES) / (PCfg<16>::DESC_BYTES + PCfg<16>::RAW_BYTES + MASK_FRAME_BYTES)); }

int launch_merge_pf(const FastArgs& F, cudaStream_t st)
{
    // 20-row tiles (640 threads) while the frames fit beside them, else 16-row tiles (512 threads): same-box A/B at 12 MP x 8 frames
    // 5.53 ms vs 5.85 ms (fewer staged halo rows per output row, 20 warps instead of 16 to hide the staging latencies).
    // MFSR_PF_TH=16 / 20 forces one.
    static const char* e = getenv("MFSR_PF_TH");
    const int th = e ? atoi(e) : 0;
    if (th == 20 || (th == 0 && PCfg<20>::smem_bytes(F.a.n_frames, false) <= (size_t)226 * 1024)) {
        const int rc = launch_th<20>(F, st);
        if (rc != MFSR_E_INVALID || th == 20) return rc;
    }
    return launch_th<16>(F, st);
}
}  // namespace s2

}  // namespace mfsr
