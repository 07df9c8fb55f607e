// Host check of csrc/merge_slots.h: the slot fold (x stage, y stage, certainty fold, class sums) must reproduce the reference's
// per-tap index arithmetic (DeBayerKernels.cu:398-437 as restated in oracle/mfsr_oracle.c:orc_accumulate) for every residue of
// (X, Y, sx, sy).  Built and run by tests/test_merge_taps_cpu.py with plain g++.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../multi_frame_super_resolution_b200/csrc/merge_slots.h"

using namespace mfsr;

static const int W = 96, H = 80;            // raw dims; mask dims W/2 x H/2
static std::vector<float> raw, mask;        // raw[W*H] normalised, mask[(H/2)*(W/2)*4] certainty per absolute CFA phase
static float frand() { return (float)rand() / (float)RAND_MAX; }

template <int J, int YM>
static int check_one(int B, int By, int sx, int sy, const float* w25, double* maxerr)
{
    const int X = 4 * B + J, Y = 4 * By + YM;
    const int Xs = X + sx, Ys = Y + sy;
    const int k = Xs >> 1, ky = Ys >> 1, ex = Xs & 1, ey = Ys & 1, phx = k & 1, phy = ky & 1;
    float w[mt::NW];
    for (int py = -2; py <= 2; py++)
        for (int px = -2; px <= 2; px++) w[mt::widx(px, py)] = w25[(py + 2) * 5 + px + 2];
    float R[3][3], Q[2][2][2][2];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[r][c] = raw[(ky - 1 + r) * W + (k - 1 + c)];
    const int mrow0 = (Y - 2) >> 2, mcol0 = (X - 2) >> 2;       // the cell of tap -2
    for (int mr = 0; mr < 2; mr++)
        for (int mc = 0; mc < 2; mc++)
            for (int cy = 0; cy < 2; cy++)
                for (int cx = 0; cx < 2; cx++)
                    Q[mr][mc][cy][cx] = mask[((mrow0 + mr) * (W / 2) + (mcol0 + mc)) * 4 + ((cy ^ phy) * 2 + (cx ^ phx))];
    float t[4], u[4];
    ms::pixel_fold<J, YM>(w, (float)ex, (float)(1 - ex), (float)ey, (float)(1 - ey), Q, R, t, u);
    double dt[4] = {0, 0, 0, 0}, du[4] = {0, 0, 0, 0};
    for (int py = -2; py <= 2; py++)
        for (int px = -2; px <= 2; px++) {
            const int ppsx = (X + px + sx) / 2, ppsy = (Y + py + sy) / 2;
            const int ppx = (X + px) / 2, ppy = (Y + py) / 2;
            const int q = (ppsy % 2) * 2 + (ppsx % 2);
            const double wt = w25[(py + 2) * 5 + px + 2];
            const double cert = mask[((ppy / 2) * (W / 2) + (ppx / 2)) * 4 + q];
            dt[q] += wt * cert * raw[ppsy * W + ppsx];
            du[q] += wt * cert;
        }
    int bad = 0;
    for (int cy = 0; cy < 2; cy++)
        for (int cx = 0; cx < 2; cx++) {
            const int q = (cy ^ phy) * 2 + (cx ^ phx);
            const double e1 = std::fabs(t[cy * 2 + cx] - dt[q]), e2 = std::fabs(u[cy * 2 + cx] - du[q]);
            if (e1 > *maxerr) *maxerr = e1;
            if (e2 > *maxerr) *maxerr = e2;
            if (e1 > 1e-4 || e2 > 1e-4) bad++;
        }
    return bad;
}

template <int J, int YM>
static int check_class(double* maxerr, long* n)
{
    int bad = 0;
    float w25[25];
    for (int rep = 0; rep < 40; rep++) {
        // symmetric weights w(px,py) == w(-px,-py), like the regression kernel
        for (int py = 0; py <= 2; py++)
            for (int px = -2; px <= 2; px++) {
                if (py == 0 && px < 0) continue;
                const float v = 0.05f + frand();
                w25[(py + 2) * 5 + px + 2] = v;
                w25[(-py + 2) * 5 + (-px) + 2] = v;
            }
        const int B = 8 + rand() % 6, By = 6 + rand() % 5;
        for (int sy = -7; sy <= 7; sy++)
            for (int sx = -7; sx <= 7; sx++) { bad += check_one<J, YM>(B, By, sx, sy, w25, maxerr); (*n)++; }
    }
    return bad;
}

int main()
{
    srand(1234);
    raw.resize(W * H); mask.resize((H / 2) * (W / 2) * 4);
    for (auto& v : raw) v = frand();
    for (auto& v : mask) v = frand();
    // slot structure: always four slots per axis class
    for (int c = 0; c < 4; c++) {
        int n = 0;
        for (int id = 0; id < 6; id++) n += ms::present(c, id);
        if (n != 4) { printf("class %d has %d slots\n", c, n); return 1; }
    }
    double maxerr = 0; long n = 0; int bad = 0;
#define ROW(YM) bad += check_class<0, YM>(&maxerr, &n) + check_class<1, YM>(&maxerr, &n) + check_class<2, YM>(&maxerr, &n) + check_class<3, YM>(&maxerr, &n);
    ROW(0) ROW(1) ROW(2) ROW(3)
    printf("checked %ld pixel-frames, max err %.3g, bad %d\n", n, maxerr, bad);
    return bad != 0;
}
