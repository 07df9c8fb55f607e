/*
 * ref_driver.cu — restated host driver for the REFERENCE's own hot-path kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference ships 33 host-less
 * `extern "C" __global__` kernels; nothing in /root/reference
 * launches them.  This file is the missing launcher: it #includes the reference
 * .cu files UNMODIFIED from where they lie (-I/root/reference/test_opencv, see
 * oracle/Makefile; no reference source is copied into this repo) and exposes a
 * C ABI so tests can run the reference kernels on the same device buffers as
 * the product kernels.  Output: oracle/_ref/libmfsr_ref.so (git-ignored).
 *
 * Host decisions the reference does not pin (Appendix C of SURVEY.md):
 *   textures      : cudaArray, normalizedCoords=1, linear filter, clamp address,
 *                   readMode=ElementType
 *   block shapes  : 16x16 for 2-D image kernels; box filters (P,1,1)/(1,P,1)
 *   IFFT scaling  : cc *= 1/P^2 after the unnormalised cuFFT C2R
 *   c_cfaPattern  : cudaMemcpyToSymbol of int[2][2]
 * All buffers are DENSE (pitch = width * sizeof(elem)).
 */
#include <cuda_runtime.h>
#include <cufft.h>
#include <cublas_v2.h>
#include <vector>
#include <stdint.h>
#include <stdio.h>

#include "kernel.cu"
#include "ShiftMinimizerKernels.cu"
#include "opticalFlow.cu"
#include "DeBayerKernels.cu"
#include "RobustnessModell.cu"

/* DeBayerKernels.cu:40-41 only DECLARES the CFA table (`extern "C" __device__ __constant__ ...;` —
 * upstream's host loads the kernels as a PTX module and the symbol is defined elsewhere there);
 * the restated host supplies the definition. */
extern "C" { __device__ __constant__ BayerColor c_cfaPattern[2][2] = {{Red, Green}, {Green, Blue}}; }

#define RTRY(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { fprintf(stderr, "ref_driver: %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); return (int)_e; } } while (0)
#define RSYNC() do { RTRY(cudaGetLastError()); RTRY(cudaDeviceSynchronize()); } while (0)

/* device time of the kernels of the last ref_* call (events around the launches only: texture / buffer set-up excluded) */
static cudaEvent_t g_e0, g_e1; static bool g_ev = false; static float g_last_ms = -1.0f;
#define TIC() do { if (!g_ev) { cudaEventCreate(&g_e0); cudaEventCreate(&g_e1); g_ev = true; } cudaEventRecord(g_e0); } while (0)
#define TOC() do { cudaEventRecord(g_e1); cudaEventSynchronize(g_e1); cudaEventElapsedTime(&g_last_ms, g_e0, g_e1); } while (0)

static inline dim3 grid2(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y, 1); }
static const dim3 B2(16, 16, 1);

/* ---- texture helper ---------------------------------------------------- */
struct RefTex { cudaArray_t arr; cudaTextureObject_t tex; };
static int make_tex(RefTex* t, const void* dev, int w, int h, int nch /*1,2,4 floats*/)
{
    cudaChannelFormatDesc d = nch == 1 ? cudaCreateChannelDesc<float>() : (nch == 2 ? cudaCreateChannelDesc<float2>() : cudaCreateChannelDesc<float4>());
    RTRY(cudaMallocArray(&t->arr, &d, w, h));
    RTRY(cudaMemcpy2DToArray(t->arr, 0, 0, dev, (size_t)w * nch * 4, (size_t)w * nch * 4, h, cudaMemcpyDeviceToDevice));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = t->arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
    RTRY(cudaCreateTextureObject(&t->tex, &rd, &td, nullptr));
    return 0;
}
static void free_tex(RefTex* t) { cudaDestroyTextureObject(t->tex); cudaFreeArray(t->arr); }

static int set_cfa(const int cfa[4])
{
    BayerColor p[2][2] = {{(BayerColor)cfa[0], (BayerColor)cfa[1]}, {(BayerColor)cfa[2], (BayerColor)cfa[3]}};
    RTRY(cudaMemcpyToSymbol(c_cfaPattern, p, sizeof(p)));
    return 0;
}

/* direct cross-correlation (exact-sum variant (ii) of SURVEY §7): same lags as the
 * FFT result, cc[s] = sum_p a[p] * b[(p+s) mod P], row-major serial sum per lag. */
__global__ void ref_direct_cc(const float* a, const float* b, float* cc, int P, int tiles)
{
    int sx = blockIdx.x * blockDim.x + threadIdx.x, sy = blockIdx.y * blockDim.y + threadIdx.y, t = blockIdx.z;
    if (sx >= P || sy >= P || t >= tiles) return;
    const float* A = a + (size_t)t * P * P; const float* Bm = b + (size_t)t * P * P;
    float acc = 0;
    for (int py = 0; py < P; py++) {
        int by = py + sy; if (by >= P) by -= P;
        for (int px = 0; px < P; px++) {
            float av = A[py * P + px];
            if (av == 0.0f) continue;
            int bx = px + sx; if (bx >= P) bx -= P;
            acc += av * Bm[by * P + bx];
        }
    }
    cc[(size_t)t * P * P + sy * P + sx] = acc;
}
__global__ void ref_scale(float* v, float f, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i < n) v[i] *= f; }

extern "C" {

int ref_version(void) { return 2; }
float ref_last_kernel_ms(void) { return g_last_ms; }

int ref_subsample3(const uint16_t* raw, float* rgb3, float maxVal, int dimX, int dimY, const int cfa[4])
{
    if (set_cfa(cfa)) return -1;
    TIC(); deBayersSubSample3<<<grid2(dimX, dimY, B2), B2>>>((unsigned short*)raw, (float3*)rgb3, maxVal, dimX, dimY, dimX * 12);
    TOC(); RSYNC(); return 0;
}

/* raw_f: float image; rgb3 must be zeroed by the caller (border stays unwritten) */
int ref_debayer(const float* raw_f, float* rgb3, int w, int h, const int cfa[4], const float black[3], const float scale[3])
{
    if (set_cfa(cfa)) return -1;
    float3 bp = make_float3(black[0], black[1], black[2]), sc = make_float3(scale[0], scale[1], scale[2]);
    TIC(); deBayerGreenKernel<<<grid2(w, h, B2), B2>>>(w, h, raw_f, w * 4, (float3*)rgb3, w * 12, bp, sc);
    deBayerRedBlueKernel<<<grid2(w, h, B2), B2>>>(w, h, raw_f, w * 4, (float3*)rgb3, w * 12, bp, sc);
    TOC(); RSYNC(); return 0;
}

/* Full tile-matching chain on float images.  pre2: dense float2 [ty][tx] or NULL.
 * coord2: findMinimum output (dense float2 [ty][tx]); ssd: [tiles][S*S]. use_fft: cuFFT CC or direct CC. */
int ref_tile_align(const float* ref_img, const float* mov_img, int w, int h, const float* pre2,
                   float* coord2, float* ssd, int T, int M, int tx, int ty,
                   float bsx, float bsy, float rot, float threshold, int use_fft)
{
    int P = T + 2 * M, S = 2 * M + 1, nt = tx * ty;
    size_t tile_elems = (size_t)nt * P * P;
    float *ta, *tb, *cc, *bx, *by, *sq, *pre = nullptr;
    RTRY(cudaMalloc(&ta, tile_elems * 4)); RTRY(cudaMalloc(&tb, tile_elems * 4)); RTRY(cudaMalloc(&cc, tile_elems * 4));
    RTRY(cudaMalloc(&bx, tile_elems * 4)); RTRY(cudaMalloc(&by, tile_elems * 4)); RTRY(cudaMalloc(&sq, (size_t)nt * 4));
    if (!pre2) { RTRY(cudaMalloc(&pre, (size_t)nt * 8)); RTRY(cudaMemset(pre, 0, (size_t)nt * 8)); }
    dim3 bt(8, 8, 4), gt((P + 7) / 8, (P + 7) / 8, (nt + 3) / 4);
    float2 bs = make_float2(bsx, bsy);
    float ms_total = 0; TIC(); convertToTilesOverlapBorder<<<gt, bt>>>(ref_img, ta, w, h, w * 4, M, T, tx, ty, bs, rot);
    convertToTilesOverlapPreShift<<<gt, bt>>>(mov_img, tb, (const float2*)(pre2 ? pre2 : pre), tx * 8, w, h, w * 4, M, T, tx, ty, bs, rot);
    TOC(); ms_total += g_last_ms; RSYNC();
    if (use_fft) {
        cufftHandle r2c, c2r; int n[2] = {P, P};
        size_t cplx = (size_t)nt * P * (P / 2 + 1);
        float2 *fa, *fb;
        RTRY(cudaMalloc(&fa, cplx * 8)); RTRY(cudaMalloc(&fb, cplx * 8));
        if (cufftPlanMany(&r2c, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, nt) != CUFFT_SUCCESS) return -2;
        if (cufftPlanMany(&c2r, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, nt) != CUFFT_SUCCESS) return -2;
        TIC(); if (cufftExecR2C(r2c, ta, (cufftComplex*)fa) != CUFFT_SUCCESS) return -3;
        if (cufftExecR2C(r2c, tb, (cufftComplex*)fb) != CUFFT_SUCCESS) return -3;
        conjugateComplexMulKernel<<<(unsigned)((cplx + 255) / 256), 256>>>(fa, fb, (int)cplx);
        if (cufftExecC2R(c2r, (cufftComplex*)fb, cc) != CUFFT_SUCCESS) return -3;
        ref_scale<<<(unsigned)((tile_elems + 255) / 256), 256>>>(cc, 1.0f / (float)(P * P), tile_elems);
        TOC(); ms_total += g_last_ms; RSYNC();
        cufftDestroy(r2c); cufftDestroy(c2r); cudaFree(fa); cudaFree(fb);
    } else {
        dim3 bc(8, 8, 1), gc((P + 7) / 8, (P + 7) / 8, nt);
        TIC(); ref_direct_cc<<<gc, bc>>>(ta, tb, cc, P, nt);
        TOC(); ms_total += g_last_ms; RSYNC();
    }
    TIC(); squaredSum<<<(nt + 127) / 128, 128>>>(ta, sq, M, T, nt);
    /* blockDim.x = P, one row per block, one tile per z (kernel.cu:145-147) */
    boxFilterWithBorderX<<<dim3(1, P, nt), dim3(P, 1, 1), P * 4>>>(tb, bx, M, T, nt);
    boxFilterWithBorderY<<<dim3(P, 1, nt), dim3(1, P, 1), P * 4>>>(bx, by, M, T, nt);
    dim3 bn(S, S, 1), gn(1, 1, nt);
    normalizedCC<<<gn, bn>>>(cc, sq, by, ssd, M, T, nt);
    findMinimum<<<(nt + 127) / 128, 128>>>(ssd, (float2*)coord2, tx * 8, M, nt, tx, threshold);
    TOC(); ms_total += g_last_ms; g_last_ms = ms_total; RSYNC();
    cudaFree(ta); cudaFree(tb); cudaFree(cc); cudaFree(bx); cudaFree(by); cudaFree(sq); if (pre) cudaFree(pre);
    return 0;
}

int ref_upsample_shifts(const float* in2, float* out2, int oldLevel, int newLevel, int oldCX, int oldCY, int newCX, int newCY, int oldT, int newT)
{
    UpSampleShifts<<<grid2(newCX, newCY, B2), B2>>>((const float2*)in2, (float2*)out2, oldCX * 8, newCX * 8, oldLevel, newLevel, oldCX, oldCY, newCX, newCY, oldT, newT);
    RSYNC(); return 0;
}

int ref_flow_from_tiles(const float* tile2, int tilesX, int tilesY, int T, float* flow2, int w, int h, float bsx, float bsy, float rot)
{
    RefTex t; if (make_tex(&t, tile2, tilesX, tilesY, 2)) return -1;
    TIC(); CreateFlowFieldFromTiles<<<grid2(w, h, B2), B2>>>((float2*)flow2, t.tex, T, tilesX, tilesY, w, h, w * 8, make_float2(bsx, bsy), rot);
    TOC(); RSYNC(); free_tex(&t); return 0;
}

int ref_warp(const float* flow2, const float* img, float* out, int w, int h)
{
    RefTex tf, ti; if (make_tex(&tf, flow2, w, h, 2) || make_tex(&ti, img, w, h, 1)) return -1;
    TIC(); WarpingKernel<<<grid2(w, h, B2), B2>>>(w, h, w * 4, tf.tex, out, ti.tex);
    TOC(); RSYNC(); free_tex(&tf); free_tex(&ti); return 0;
}

int ref_derivatives(const float* src, const float* tgt, float* Ix, float* Iy, float* Iz, int w, int h)
{
    RefTex ts, tt; if (make_tex(&ts, src, w, h, 1) || make_tex(&tt, tgt, w, h, 1)) return -1;
    TIC(); ComputeDerivativesKernel<<<grid2(w, h, B2), B2>>>(w, h, w * 4, Ix, Iy, Iz, ts.tex, tt.tex);
    TOC(); RSYNC(); free_tex(&ts); free_tex(&tt); return 0;
}

int ref_derivatives2(const float* img, float* Ix, float* Iy, int w, int h)
{
    RefTex t; if (make_tex(&t, img, w, h, 1)) return -1;
    TIC(); ComputeDerivatives2Kernel<<<grid2(w, h, B2), B2>>>(w, h, w * 4, Ix, Iy, t.tex);
    TOC(); RSYNC(); free_tex(&t); return 0;
}

int ref_lucas_kanade(float* flow2, const float* Ix, const float* Iy, const float* It, int w, int h, int halfWin, float minDet)
{
    TIC(); lucasKanadeOptim<<<grid2(w, h, B2), B2>>>((float2*)flow2, Ix, Iy, It, w * 8, w * 4, w, h, halfWin, minDet);
    TOC(); RSYNC(); return 0;
}

int ref_structure_tensor(const float* Ix, const float* Iy, float* t3, int w, int h)
{
    TIC(); ComputeStructureTensor<<<grid2(w, h, B2), B2>>>(Ix, Iy, (float3*)t3, w, h, w * 4, w * 12);
    TOC(); RSYNC(); return 0;
}

int ref_kernel_param(float* k3, int w, int h, float Dth, float Dtr, float kDetail, float kDenoise, float kStretch, float kShrink)
{
    TIC(); ComputeKernelParam<<<grid2(w, h, B2), B2>>>((float3*)k3, w, h, w * 12, Dth, Dtr, kDetail, kDenoise, kStretch, kShrink);
    TOC(); RSYNC(); return 0;
}

/* mask4 zeroed by the caller; flow2 is fw x fh */
int ref_robustness_mask(const float* ref3, const float* mov3, float* mask4, const float* flow2, int fw, int fh,
                        int w, int h, float alpha, float beta, float thresholdM)
{
    RefTex t; if (make_tex(&t, flow2, fw, fh, 2)) return -1;
    size_t smem = (size_t)B2.x * B2.y * 9 * sizeof(float3);
    TIC(); ComputeRobustnessMask<<<grid2(w, h, B2), B2, smem>>>((const float3*)ref3, (const float3*)mov3, (float4*)mask4, t.tex, w, h, w * 12, w * 16, alpha, beta, thresholdM);
    TOC(); RSYNC(); free_tex(&t); return 0;
}

/* one frame of accumulateImagesSuperRes (2x, output dims == raw dims); sum3/weight3 RMW */
int ref_accumulate_superres(const uint16_t* raw, float* sum3, float* weight3, const float* mask4, const float* kernel4,
                            const float* flow2, int dimX, int dimY, const int cfa[4], const float white[3], const float black[3])
{
    if (set_cfa(cfa)) return -1;
    RefTex tk, ts; if (make_tex(&tk, kernel4, dimX, dimY, 4) || make_tex(&ts, flow2, dimX, dimY, 2)) return -1;
    TIC(); accumulateImagesSuperRes<<<grid2(dimX, dimY, B2), B2>>>((unsigned short*)raw, (float3*)sum3, (float3*)weight3, (const float4*)mask4, tk.tex, ts.tex,
        make_float3(white[0], white[1], white[2]), make_float3(black[0], black[1], black[2]), dimX, dimY, dimX * 12, (dimX / 2) * 16, dimX * 16, dimX * 8);
    TOC(); RSYNC(); free_tex(&tk); free_tex(&ts); return 0;
}

/* one frame of accumulateImages (1x). kernel3: float3 image with pitch == output pitch (DeBayerKernels.cu:308) */
int ref_accumulate_1x(const uint16_t* raw, float* sum3, float* weight3, const float* mask4, const float* kernel3,
                      const float* flow2, int dimX, int dimY, const int cfa[4], const float white[3], const float black[3])
{
    if (set_cfa(cfa)) return -1;
    accumulateImages<<<grid2(dimX, dimY, B2), B2>>>((unsigned short*)raw, (float3*)sum3, (float3*)weight3, (const float4*)mask4, (const float3*)kernel3, (const float2*)flow2,
        make_float3(white[0], white[1], white[2]), make_float3(black[0], black[1], black[2]), dimX, dimY, dimX * 12, (dimX / 2) * 16, dimX * 8);
    RSYNC(); return 0;
}

int ref_apply_weighting(float* inout3, const float* final3, const float* weight3, int w, int h, float threshold)
{
    ApplyWeighting<<<grid2(w, h, B2), B2>>>((float3*)inout3, (const float3*)final3, (const float3*)weight3, w, h, w * 12, threshold);
    RSYNC(); return 0;
}

int ref_gamma(float* img3, int w, int h)
{
    GammasRGB<<<grid2(w, h, B2), B2>>>((float3*)img3, w, h, w * 12);
    RSYNC(); return 0;
}

/* Whole reference merge chain for timing: N RMW passes + normalise (+gamma); returns ms via *ms_out.
 * Frame f buffers are raw + f*dimX*dimY etc. (dense stacks). Textures are created outside the timed region. */
int ref_merge_chain_timed(const uint16_t* raw, const float* mask4, const float* kernel4, const float* flow2,
                          float* sum3, float* weight3, float* inout3, int n_frames, int dimX, int dimY,
                          const int cfa[4], const float white[3], const float black[3], float threshold, int gamma, float* ms_out)
{
    if (set_cfa(cfa)) return -1;
    RefTex tk; if (make_tex(&tk, kernel4, dimX, dimY, 4)) return -1;
    RefTex* ts = new RefTex[n_frames];
    for (int f = 0; f < n_frames; f++) if (make_tex(&ts[f], flow2 + (size_t)f * dimX * dimY * 2, dimX, dimY, 2)) return -1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    RTRY(cudaMemset(sum3, 0, (size_t)dimX * dimY * 12)); RTRY(cudaMemset(weight3, 0, (size_t)dimX * dimY * 12));
    RTRY(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int f = 0; f < n_frames; f++)
        accumulateImagesSuperRes<<<grid2(dimX, dimY, B2), B2>>>((unsigned short*)(raw + (size_t)f * dimX * dimY), (float3*)sum3, (float3*)weight3,
            (const float4*)(mask4 + (size_t)f * (dimX / 2) * (dimY / 2) * 4), tk.tex, ts[f].tex,
            make_float3(white[0], white[1], white[2]), make_float3(black[0], black[1], black[2]), dimX, dimY, dimX * 12, (dimX / 2) * 16, dimX * 16, dimX * 8);
    ApplyWeighting<<<grid2(dimX, dimY, B2), B2>>>((float3*)inout3, (const float3*)sum3, (const float3*)weight3, dimX, dimY, dimX * 12, threshold);
    if (gamma) GammasRGB<<<grid2(dimX, dimY, B2), B2>>>((float3*)inout3, dimX, dimY, dimX * 12);
    cudaEventRecord(e1);
    RSYNC();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (int f = 0; f < n_frames; f++) free_tex(&ts[f]);
    delete[] ts; free_tex(&tk);
    return 0;
}

/* ---- shift consolidation (ShiftMinimizerKernels.cu) ------------------------------------------------
 * Upstream's host runs, per sweep, cuBLAS batched AtA -> matinv -> inv*At -> solved*measured -> A*x and then
 * checkForOutliers (:81) until no tile removes a measurement.  The host is absent from the reference; this is the
 * restated one, built on the reference's own kernels copyShiftMatrix (:29), setPointers (:51), transposeShifts (:143),
 * checkForOutliers (:81), getOptimalShifts (:179) and on cuBLAS batched for the linear algebra.
 *   measured2   : device float2 [nt][m]   (tile-major, as concatenateShifts :223 produces)
 *   pair_from/to: host int [m]; row k of the design matrix has ones in columns from..to-1
 *   one_to_one2 : device float2 [nt][n1]      frame_shift2: device float2 [imageCount][ty][tx]
 *   status      : device int [nt] final reference status (-1 everywhere when converged)
 *   removed     : device int [nt] number of measurements checkForOutliers removed per tile (host-side count)
 */
__global__ void ref_to_transposed(const float2* m2, float* mT, int nt, int m)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= nt * m) return;
    int t = i / m, k = i - t * m;
    mT[2 * (size_t)t * m + k] = m2[i].x; mT[2 * (size_t)t * m + k + m] = m2[i].y;
}
__global__ void ref_count_removed(const int* status, int* removed, int nt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < nt && status[i] >= 0) removed[i]++;
}
#define BTRY(e) do { cublasStatus_t _s = (e); if (_s != CUBLAS_STATUS_SUCCESS) { fprintf(stderr, "ref_driver: cublas %d at %s:%d\n", (int)_s, __FILE__, __LINE__); return -4; } } while (0)

int ref_consolidate(const float* measured2, const int* pair_from, const int* pair_to, int m, int imageCount,
                    int tx, int ty, int referenceImage, float* one_to_one2, float* frame_shift2, int* status, int* removed)
{
    const int n1 = imageCount - 1, nt = tx * ty;
    if (n1 < 1 || n1 > 32 || m < 1) return -1;
    float *A, *Asafe, *Sq, *Inv, *Solved, *measT, *o2oT, *optT; float2 *meas, *o2o;
    float **pA, **pAsafe, **pSq, **pInv, **pSolved; float2 **pO2o, **pMeas, **pOpt; int* info;
    RTRY(cudaMalloc(&A, (size_t)nt * n1 * m * 4)); RTRY(cudaMalloc(&Asafe, (size_t)nt * n1 * m * 4));
    RTRY(cudaMalloc(&Sq, (size_t)nt * n1 * n1 * 4)); RTRY(cudaMalloc(&Inv, (size_t)nt * n1 * n1 * 4));
    RTRY(cudaMalloc(&Solved, (size_t)nt * n1 * m * 4));
    RTRY(cudaMalloc(&measT, (size_t)nt * m * 8)); RTRY(cudaMalloc(&meas, (size_t)nt * m * 8));
    RTRY(cudaMalloc(&o2oT, (size_t)nt * n1 * 8)); RTRY(cudaMalloc(&o2o, (size_t)nt * n1 * 8)); RTRY(cudaMalloc(&optT, (size_t)nt * m * 8));
    RTRY(cudaMalloc(&pA, nt * sizeof(void*))); RTRY(cudaMalloc(&pAsafe, nt * sizeof(void*))); RTRY(cudaMalloc(&pSq, nt * sizeof(void*)));
    RTRY(cudaMalloc(&pInv, nt * sizeof(void*))); RTRY(cudaMalloc(&pSolved, nt * sizeof(void*))); RTRY(cudaMalloc(&pO2o, nt * sizeof(void*)));
    RTRY(cudaMalloc(&pMeas, nt * sizeof(void*))); RTRY(cudaMalloc(&pOpt, nt * sizeof(void*))); RTRY(cudaMalloc(&info, nt * 4));
    /* design matrix of tile 0 (column-major m x n1, element idx + col*m as :137 addresses it), replicated by copyShiftMatrix */
    std::vector<float> a0((size_t)m * n1, 0.0f);
    for (int k = 0; k < m; k++) for (int c = pair_from[k]; c < pair_to[k]; c++) a0[k + (size_t)c * m] = 1.0f;
    RTRY(cudaMemcpy(A, a0.data(), a0.size() * 4, cudaMemcpyHostToDevice));
    const int TB = 128, GB = (nt + TB - 1) / TB;
    copyShiftMatrix<<<GB, TB>>>(A, nt, imageCount, m);
    setPointers<<<GB, TB>>>(pA, pAsafe, pSq, pInv, pSolved, pO2o, pMeas, pOpt, A, Asafe, Sq, Inv, Solved,
                            (float2*)o2oT, (float2*)measT, (float2*)optT, nt, imageCount, m);
    ref_to_transposed<<<(nt * m + 255) / 256, 256>>>((const float2*)measured2, measT, nt, m);
    dim3 bt(32, 8), gt((nt + 31) / 32, (m + 7) / 8);
    transposeShifts<<<gt, bt>>>(meas, measT, o2oT, o2o, nt, imageCount, m);      /* measured (float2) from measuredT */
    RTRY(cudaMemset(status, 0, nt * 4)); RTRY(cudaMemset(removed, 0, nt * 4));
    RSYNC();
    cublasHandle_t hb; BTRY(cublasCreate(&hb));
    BTRY(cublasSetMathMode(hb, CUBLAS_PEDANTIC_MATH));
    const float one = 1.0f, zero = 0.0f;
    std::vector<int> hst(nt);
    for (int sweep = 0; sweep <= m; sweep++) {
        BTRY(cublasSgemmBatched(hb, CUBLAS_OP_T, CUBLAS_OP_N, n1, n1, m, &one, (const float* const*)pA, m, (const float* const*)pA, m, &zero, pSq, n1, nt));
        BTRY(cublasSmatinvBatched(hb, n1, (const float* const*)pSq, n1, pInv, n1, info, nt));
        BTRY(cublasSgemmBatched(hb, CUBLAS_OP_N, CUBLAS_OP_T, n1, m, n1, &one, (const float* const*)pInv, n1, (const float* const*)pA, m, &zero, pSolved, n1, nt));
        BTRY(cublasSgemmBatched(hb, CUBLAS_OP_N, CUBLAS_OP_N, n1, 2, m, &one, (const float* const*)pSolved, n1, (const float* const*)pMeas, m, &zero, (float**)pO2o, n1, nt));
        BTRY(cublasSgemmBatched(hb, CUBLAS_OP_N, CUBLAS_OP_N, m, 2, n1, &one, (const float* const*)pA, m, (const float* const*)pO2o, n1, &zero, (float**)pOpt, m, nt));
        checkForOutliers<<<GB, TB>>>(meas, optT, A, status, info, nt, imageCount, m);
        ref_count_removed<<<GB, TB>>>(status, removed, nt);
        RSYNC();
        RTRY(cudaMemcpy(hst.data(), status, nt * 4, cudaMemcpyDeviceToHost));
        bool any = false; for (int t = 0; t < nt; t++) any |= hst[t] >= 0;
        if (!any) break;
    }
    cublasDestroy(hb);
    transposeShifts<<<gt, bt>>>(meas, measT, o2oT, o2o, nt, imageCount, m);      /* oneToOne (float2) from oneToOneT */
    RTRY(cudaMemcpy(one_to_one2, o2o, (size_t)nt * n1 * 8, cudaMemcpyDeviceToDevice));
    for (int f = 0; f < imageCount; f++)
        getOptimalShifts<<<grid2(tx, ty, B2), B2>>>((float2*)frame_shift2 + (size_t)f * nt, o2o, imageCount, tx, ty, tx * 8, referenceImage, f);
    RSYNC();
    cudaFree(A); cudaFree(Asafe); cudaFree(Sq); cudaFree(Inv); cudaFree(Solved); cudaFree(measT); cudaFree(meas); cudaFree(o2oT); cudaFree(o2o); cudaFree(optT);
    cudaFree(pA); cudaFree(pAsafe); cudaFree(pSq); cudaFree(pInv); cudaFree(pSolved); cudaFree(pO2o); cudaFree(pMeas); cudaFree(pOpt); cudaFree(info);
    return 0;
}

/* probe: what does the texture unit return for a linear-filtered fetch at unnormalised
 * coordinate u of a 1-row ramp texture? (used to pin the 1.8 fixed-point model) */
__global__ void ref_probe_kernel(cudaTextureObject_t t, const float* xn, float* out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = tex2D<float>(t, xn[i], 0.5f);
}
int ref_texture_probe(const float* texels, int w, const float* xn, float* out, int n)
{
    RefTex t; if (make_tex(&t, texels, w, 1, 1)) return -1;
    ref_probe_kernel<<<(n + 255) / 256, 256>>>(t.tex, xn, out, n);
    RSYNC(); free_tex(&t); return 0;
}

}  /* extern "C" */
