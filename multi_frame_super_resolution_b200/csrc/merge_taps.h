// merge_taps.h — static tap tables of the scale-2 kernel-regression merge.
//
// Shared by the CUDA kernel (merge_fast.cu) and by a plain g++ host check
// (tests/host/merge_taps_check.cpp), hence no CUDA types in here.
//
// Geometry restated from accumulateImagesSuperRes (DeBayerKernels.cu:398-437) for one axis,
// all coordinates ABSOLUTE high-resolution (X = x + dimX/2 in the reference):
//   tap p in [-2,2] of output pixel X whose integer HR shift is s reads
//     raw column   (X + s + p) >> 1          (:414  ppsx/2)
//     mask column  ((X + p) >> 1) >> 1       (:418,:437  ppx/2 then /2 again)
//   and the tap's CFA colour is the parity of the raw column (:423).
// A thread owns the 4 pixels X = 4B + J (J = 0..3) of one output row; with one shift s for
// all four, X' = X + s, RHO = (4B + s) & 3 and
//   raw column  = k0 + dk(J) + g(e, p),   k0 = (4B + s) >> 1,  e = (RHO + J) & 1
//   mask column = B + floor((J + p) / 4)
// so every index below depends only on compile-time (J, RHO, EY, YM) — the kernel switches on
// (RHO, EY) once per thread and frame and runs straight-line FMA code.
#pragma once

#if defined(__CUDACC__)
#define MFSR_HD __host__ __device__ __forceinline__
#define MFSR_CX __host__ __device__ __forceinline__ constexpr
#else
#define MFSR_HD inline
#define MFSR_CX constexpr
#endif

namespace mfsr {
namespace mt {

MFSR_CX int fl2(int v) { return (v + 8) / 2 - 4; }      // floor(v / 2) for v >= -8
MFSR_CX int fl4(int v) { return (v + 16) / 4 - 4; }     // floor(v / 4) for v >= -16

// The regression weight depends on (px*px, px*py, py*py) only: w(px,py) == w(-px,-py) -> 13 distinct values.
MFSR_CX int widx(int px, int py)
{
    const bool neg = py < 0 || (py == 0 && px < 0);
    const int ax = neg ? -px : px, ay = neg ? -py : py;
    return ay == 0 ? ax : (ay == 1 ? 5 + ax : 10 + ax);
}
constexpr int NW = 13;

// ---- x axis -------------------------------------------------------------------------------
// columns the thread loads: an even-aligned window of 6 raw columns starting at cbase = (k0-1) & ~1
MFSR_CX int x_off0(int rho) { return 1 - ((rho >> 1) & 1); }                    // (k0-1) - cbase
MFSR_CX int x_dk(int rho, int j) { return ((rho + j) >> 1) - (rho >> 1); }      // k_J - k0
MFSR_CX int x_e(int rho, int j) { return (rho + j) & 1; }
MFSR_CX int x_g(int rho, int j, int p) { return fl2(x_e(rho, j) + p); }         // raw column of tap p relative to k_J
MFSR_CX int x_rcol(int rho, int j, int g) { return x_off0(rho) + x_dk(rho, j) + g + 1; }   // index into the 6-column window
MFSR_CX int x_phase(int rho, int j, int g) { return (((rho + j) >> 1) + g + 8) & 1; }     // absolute CFA x parity of that column
MFSR_CX int x_mcol(int j, int p) { return fl4(j + p) + 1; }                     // index into mask columns (B-1, B, B+1)

// ---- y axis (the thread owns one row; EY = (Y+sy)&1 and the parity of (Y+sy)>>1 are runtime) ---
MFSR_CX int y_g(int ey, int p) { return fl2(ey + p); }                          // raw row of tap p relative to ky, in -1..1
MFSR_CX int y_cls(int g) { return g & 1; }                                       // parity class relative to ky
MFSR_CX int y_mrow(int ym, int p) { return fl4(ym + p) + (ym < 2 ? 1 : 0); }    // index into the 2 loaded mask rows

// One output pixel, one frame.
//   w   : the pixel's 13 regression weights
//   R   : raw window, R[gy+1][column], 3 rows x 6 columns, normalised samples
//   Q   : certainty, Q[mask row 0..1][mask col 0..2][ycls*2 + xphase]  (y relative to ky, x absolute)
//   t,u : per (ycls*2 + xphase) class:  t += sum w*cert*raw,  u += sum w*cert
template <int J, int RHO, int EY, int YM>
MFSR_HD void pixel_taps(const float (&w)[NW], const float (&R)[3][6], const float (&Q)[2][3][4], float (&t)[4], float (&u)[4])
{
    float G[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) G[a][b] = 0.0f;
#pragma unroll
    for (int py = -2; py <= 2; py++) {
        const int gy = y_g(EY, py), mr = y_mrow(YM, py), cy = y_cls(gy);
#pragma unroll
        for (int px = -2; px <= 2; px++) {
            const int gx = x_g(RHO, J, px), mc = x_mcol(J, px), qx = x_phase(RHO, J, gx);
            G[gy + 1][gx + 1] += w[widx(px, py)] * Q[mr][mc][cy * 2 + qx];
        }
    }
#pragma unroll
    for (int gy = -1; gy <= 1; gy++)
#pragma unroll
        for (int gx = -1; gx <= 1; gx++) {
            const int c = y_cls(gy) * 2 + x_phase(RHO, J, gx);
            t[c] += G[gy + 1][gx + 1] * R[gy + 1][x_rcol(RHO, J, gx)];
            u[c] += G[gy + 1][gx + 1];
        }
}

}  // namespace mt
}  // namespace mfsr
