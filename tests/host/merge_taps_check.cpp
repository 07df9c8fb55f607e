// Host check of csrc/merge_taps.h: the static (J, RHO, EY, YM) tap tables must reproduce the reference's
// per-tap index arithmetic (DeBayerKernels.cu:398-437 as restated in oracle/mfsr_oracle.c:orc_accumulate)
// for every residue of (X, Y, sx, sy).  Built and run by tests/test_merge_taps_cpu.py with plain g++.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../multi_frame_super_resolution_b200/csrc/merge_taps.h"

using namespace mfsr::mt;

static const int W = 96, H = 80;            // raw dims; mask dims W/2 x H/2
static std::vector<float> raw, mask;        // raw[W*H] normalised, mask[(H/2)*(W/2)*4] abs-phase certainty (4 phases)
static float frand() { return (float)rand() / (float)RAND_MAX; }

template <int J, int RHO, int EY, int YM>
static int check_one(int B, int By, int s, int sy, const float* w25, double* maxerr)
{
    // thread geometry
    const int X0 = 4 * B, Y = 4 * By + YM;
    const int X0p = X0 + s, Yp = Y + sy;
    if (((X0p % 4) + 4) % 4 != RHO || (((Yp % 2) + 2) % 2) != EY) return 0;
    const int k0 = X0p >> 1, ky = Yp >> 1, phiy = ky & 1;
    const int cbase = (k0 - 1) & ~1;
    float w[NW];
    for (int py = -2; py <= 2; py++)
        for (int px = -2; px <= 2; px++) w[widx(px, py)] = w25[(py + 2) * 5 + px + 2];
    float R[3][6], Q[2][3][4];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 6; c++) R[r][c] = raw[(ky - 1 + r) * W + cbase + c];
    const int mrow0 = By + (YM < 2 ? -1 : 0);
    for (int mr = 0; mr < 2; mr++)
        for (int mc = 0; mc < 3; mc++)
            for (int cy = 0; cy < 2; cy++)
                for (int qx = 0; qx < 2; qx++)
                    Q[mr][mc][cy * 2 + qx] = mask[((mrow0 + mr) * (W / 2) + (B - 1 + mc)) * 4 + ((cy ^ phiy) * 2 + qx)];
    float t[4] = {0, 0, 0, 0}, u[4] = {0, 0, 0, 0};
    pixel_taps<J, RHO, EY, YM>(w, R, Q, t, u);
    // direct evaluation, reference index arithmetic, absolute CFA phase classes
    double dt[4] = {0, 0, 0, 0}, du[4] = {0, 0, 0, 0};
    const int X = X0 + J;
    for (int py = -2; py <= 2; py++)
        for (int px = -2; px <= 2; px++) {
            const int ppsx = (X + px + s) / 2, ppsy = (Y + py + sy) / 2;
            const int ppx = (X + px) / 2, ppy = (Y + py) / 2;
            const int q = (ppsy % 2) * 2 + (ppsx % 2);
            const double wt = w25[(py + 2) * 5 + px + 2];
            const double cert = mask[((ppy / 2) * (W / 2) + (ppx / 2)) * 4 + q];
            dt[q] += wt * cert * raw[ppsy * W + ppsx];
            du[q] += wt * cert;
        }
    int bad = 0;
    for (int cy = 0; cy < 2; cy++)
        for (int qx = 0; qx < 2; qx++) {
            const int q = (cy ^ phiy) * 2 + qx;
            const double e1 = std::fabs(t[cy * 2 + qx] - dt[q]), e2 = std::fabs(u[cy * 2 + qx] - du[q]);
            if (e1 > *maxerr) *maxerr = e1;
            if (e2 > *maxerr) *maxerr = e2;
            if (e1 > 1e-4 || e2 > 1e-4) bad++;
        }
    return bad;
}

template <int J, int RHO, int EY, int YM>
static int sweep(double* maxerr, long* n)
{
    int bad = 0;
    float w25[25];
    for (int B = 4; B < 12; B++)
        for (int By = 4; By < 10; By++)
            for (int s = -9; s <= 9; s++)
                for (int sy = -7; sy <= 7; sy++) {
                    // symmetric weights: w(px,py) == w(-px,-py)
                    for (int py = -2; py <= 2; py++)
                        for (int px = -2; px <= 2; px++) {
                            const int a = widx(px, py);
                            srand(1000 * a + B * 31 + By * 7 + s * 3 + sy + 12345);
                            w25[(py + 2) * 5 + px + 2] = 0.05f + frand();
                        }
                    const int b = check_one<J, RHO, EY, YM>(B, By, s, sy, w25, maxerr);
                    bad += b;
                    (*n)++;
                }
    return bad;
}

template <int J, int RHO>
static int sweep_y(double* maxerr, long* n)
{
    return sweep<J, RHO, 0, 0>(maxerr, n) + sweep<J, RHO, 0, 1>(maxerr, n) + sweep<J, RHO, 0, 2>(maxerr, n) + sweep<J, RHO, 0, 3>(maxerr, n) +
           sweep<J, RHO, 1, 0>(maxerr, n) + sweep<J, RHO, 1, 1>(maxerr, n) + sweep<J, RHO, 1, 2>(maxerr, n) + sweep<J, RHO, 1, 3>(maxerr, n);
}
template <int J>
static int sweep_x(double* maxerr, long* n)
{
    return sweep_y<J, 0>(maxerr, n) + sweep_y<J, 1>(maxerr, n) + sweep_y<J, 2>(maxerr, n) + sweep_y<J, 3>(maxerr, n);
}

int main()
{
    srand(7);
    raw.resize(W * H);
    mask.resize((H / 2) * (W / 2) * 4);
    for (auto& v : raw) v = frand();
    for (auto& v : mask) v = frand();
    double maxerr = 0;
    long n = 0;
    const int bad = sweep_x<0>(&maxerr, &n) + sweep_x<1>(&maxerr, &n) + sweep_x<2>(&maxerr, &n) + sweep_x<3>(&maxerr, &n);
    std::printf("cases %ld bad %d maxerr %.3g\n", n, bad, maxerr);
    return bad ? 1 : 0;
}
