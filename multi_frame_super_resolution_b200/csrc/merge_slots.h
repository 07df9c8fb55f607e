// merge_slots.h — slot tables of the scale-2 merge (merge_pf.cu), shared with a plain g++ host check
// (tests/host/merge_slots_check.cpp), hence no CUDA types in here.
//
// One axis of accumulateImagesSuperRes (DeBayerKernels.cu:398-437), absolute HR coordinates: tap p in [-2,2] of output
// pixel X with integer HR shift s reads raw column (X+s+p)>>1 and certainty column (X+p)>>2.  With e = (X+s)&1 and
// c = X&3 (compile time: J for x, YM for y) the tap goes to
//     destination   g(e,p) = floor((e+p)/2)            in {-1,0,1}   (raw sample relative to the window centre (X+s)>>1)
//     certainty cell m(c,p) = floor((c+p)/4) - floor((c-2)/4)  in {0,1}   (relative to the cell of tap -2)
// A SLOT is a (destination, cell) pair that occurs for some (p, e); there are always four per axis.  The weight a slot
// collects is    const(even tap, if any)  +  nx * [tap -1 lands here when e=0] * w(-1)  +  fx * [.. when e=1] * w(-1)
//                                          +  nx * [tap +1 ..  e=0] * w(+1)            +  fx * [.. e=1] * w(+1)
// with fx = e, nx = 1-e as floats: four multiply-adds per row and NO predicate — which of the odd taps moves is data.
// Measured on B200 (profiles/r2a_issue_rates_*.txt): FFMA/FMUL/FADD issue every cycle, LOP3/SHF/SEL/IMAD/ISETP/PRMT every
// second cycle, FFMA2 every third — so selection logic is done in FP arithmetic and nothing is packed.
#pragma once
#include "merge_taps.h"

namespace mfsr {
namespace ms {

MFSR_CX int dest(int e, int p) { return mt::fl2(e + p); }
MFSR_CX int cell(int c, int p) { return mt::fl4(c + p) - mt::fl4(c - 2); }

// slot id = (g+1)*2 + m  in 0..5; present(c, id) says whether the slot occurs for axis class c
MFSR_CX int slot_id(int c, int e, int p) { return (dest(e, p) + 1) * 2 + cell(c, p); }
MFSR_CX bool present(int c, int id)
{
    for (int e = 0; e < 2; e++)
        for (int p = -2; p <= 2; p++)
            if (slot_id(c, e, p) == id) return true;
    return false;
}
// k-th present slot (k = 0..3) of axis class c, as slot id
MFSR_CX int nth_slot(int c, int k)
{
    int n = 0;
    for (int id = 0; id < 6; id++)
        if (present(c, id)) { if (n == k) return id; n++; }
    return -1;
}
MFSR_CX int slot_dest(int id) { return id / 2 - 1; }
MFSR_CX int slot_cell(int id) { return id & 1; }
// even tap (−2, 0, 2) that always lands in slot id, or 9 if none
MFSR_CX int slot_even_tap(int c, int id)
{
    for (int p = -2; p <= 2; p += 2)
        if (slot_id(c, 0, p) == id) return p;
    return 9;
}
// does odd tap p (−1 / +1) land in slot id when the parity is e?
MFSR_CX bool slot_has_odd(int c, int id, int e, int p) { return slot_id(c, e, p) == id; }

// One axis fold on the host or the device: out[k] = weight collected by the k-th slot.  w5[p+2] = the five tap weights.
template <int C>
MFSR_HD void fold_axis(const float (&w5)[5], float fx, float nx, float (&out)[4])
{
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) {
        const int id = nth_slot(C, k);
        const int pe = slot_even_tap(C, id);
        bool has = pe != 9;
        float v = has ? w5[(pe != 9 ? pe : 0) + 2] : 0.0f;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int p = -1; p <= 1; p += 2) {
            if (slot_has_odd(C, id, 0, p)) { v = has ? nx * w5[p + 2] + v : nx * w5[p + 2]; has = true; }
            if (slot_has_odd(C, id, 1, p)) { v = has ? fx * w5[p + 2] + v : fx * w5[p + 2]; has = true; }
        }
        out[k] = v;
    }
}


// One output pixel, one frame (the arithmetic of merge_pf.cu's frame loop; also run on the host by the slot check).
//   w    : the pixel's 13 regression weights (mt::widx)
//   fx,nx: (X+sx)&1 and its complement as floats; fy,ny likewise for y
//   Q    : certainty Q[mask row 0..1][mask col 0..1][y class][x class]  (classes relative to the window centre)
//   R    : raw samples R[gy+1][gx+1] around the window centre
//   t,u  : per class (cy*2+cx) RELATIVE to the window centre:  t = sum w*cert*raw,  u = sum w*cert   (overwritten)
template <int J, int YM>
MFSR_HD void pixel_fold(const float (&w)[mt::NW], float fx, float nx, float fy, float ny,
                        const float (&Q)[2][2][2][2], const float (&R)[3][3], float (&t)[4], float (&u)[4])
{
    // x stage: per tap row, the four x slots
    float X[5][4];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int py = -2; py <= 2; py++) {
        float w5[5];
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int px = -2; px <= 2; px++) w5[px + 2] = w[mt::widx(px, py)];
        fold_axis<J>(w5, fx, nx, X[py + 2]);
    }
    // y stage: per x slot, the four y slots
    float S[4][4];      // [y slot][x slot]
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int a = 0; a < 4; a++) {
        float c5[5], o4[4];
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int r = 0; r < 5; r++) c5[r] = X[r][a];
        fold_axis<YM>(c5, fy, ny, o4);
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int b = 0; b < 4; b++) S[b][a] = o4[b];
    }
    // certainty and fold onto the 3x3 raw samples
    float G[3][3];
    bool gset[3][3] = {{false, false, false}, {false, false, false}, {false, false, false}};
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int b = 0; b < 4; b++) {
        const int idy = nth_slot(YM, b), gy = slot_dest(idy), mr = slot_cell(idy);
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int a = 0; a < 4; a++) {
            const int idx = nth_slot(J, a), gx = slot_dest(idx), mc = slot_cell(idx);
            const float q = Q[mr][mc][gy & 1][gx & 1];
            G[gy + 1][gx + 1] = gset[gy + 1][gx + 1] ? S[b][a] * q + G[gy + 1][gx + 1] : S[b][a] * q;
            gset[gy + 1][gx + 1] = true;
        }
    }
    // class sums relative to the window centre: (0,0) centre, (0,1) left/right, (1,0) up/down, (1,1) corners
    t[0] = G[1][1] * R[1][1];                                              u[0] = G[1][1];
    t[1] = G[1][2] * R[1][2] + G[1][0] * R[1][0];                          u[1] = G[1][0] + G[1][2];
    t[2] = G[2][1] * R[2][1] + G[0][1] * R[0][1];                          u[2] = G[0][1] + G[2][1];
    t[3] = G[2][2] * R[2][2] + (G[2][0] * R[2][0] + (G[0][2] * R[0][2] + G[0][0] * R[0][0]));
    u[3] = (G[0][0] + G[0][2]) + (G[2][0] + G[2][2]);
}

}  // namespace ms
}  // namespace mfsr
