// prealign.cu — global pre-alignment: one shift + rotation per frame against the reference frame.
//
// Compiled with -fmad=false: the sample positions hang on strictly rounded fp32 values (the oracle evaluates the same
// expressions with -ffp-contract=off), the scores are integer sums, candidates are compared as exact fractions and ties go to
// the lowest candidate index — the estimate is bit-identical to the CPU restatement the tests hold beside it.
//
// The reference carries the stage only as a skeleton (class PreAlignment, boxFilterNPP.cpp:102-166: buffers for an FFT phase
// correlation of a rotated image, no code) plus the kernels that CONSUME its result: convertToTilesOverlapBorder / PreShift
// (kernel.cu:265,324) and CreateFlowFieldFromTiles (opticalFlow.cu:48) take `baseShift` / `baseRotation` and read the moved
// image at p + round(R(theta) (c - b) - c) for a pixel with centred coordinates c.  The restated estimator searches (theta, b)
// exhaustively with exactly that transform on two levels of the 7-bit tracking pyramid:
//   stage A  small level (longer side <= 192 px): theta in [-20, 20] deg step 1, b in [-8, 8]^2
//   stage B  two levels finer: theta +- 1 deg step 0.125 around A's, b +- 4 around A's (scaled), every 2nd pixel
// score = mean squared difference over the pixels whose sample lies inside the moved image (at least a quarter of them).
#include "common.cuh"
#include "internal.h"

namespace mfsr {

namespace {

// (idx0, cx, cy): first angle-table index and centre of the shift search of one frame
struct FrameSearch { int idx0, cx, cy, pad; };

__global__ void __launch_bounds__(128)
prealign_score_kernel(const uint8_t* __restrict__ img, int64_t pitch, int64_t frame_stride, int w, int h, int ref_idx,
                      const float* __restrict__ cs, const FrameSearch* __restrict__ fs, int step, int R, int sub,
                      unsigned long long* __restrict__ ssd, unsigned* __restrict__ cnt, int ncand)
{
    const int c = blockIdx.x, f = blockIdx.y;
    const int S = 2 * R + 1;
    const int a = c / (S * S), r = c - a * S * S, iy = r / S, ix = r - iy * S;
    const FrameSearch q = fs[f];
    const float cf = cs[2 * (q.idx0 + a * step)], sf = cs[2 * (q.idx0 + a * step) + 1];
    const int bx = q.cx + ix - R, by = q.cy + iy - R;
    const uint8_t* ref = img + frame_stride * ref_idx;
    const uint8_t* mov = img + frame_stride * f;
    const int nx = (w + sub - 1) / sub, ny = (h + sub - 1) / sub;
    unsigned long long s = 0; unsigned n = 0;
    for (int i = threadIdx.x; i < nx * ny; i += blockDim.x) {
        const int yy = i / nx, y = yy * sub, x = (i - yy * nx) * sub;
        const float pcx = (float)(x - w / 2), pcy = (float)(y - h / 2);
        const float ax = pcx - (float)bx, ay = pcy - (float)by;
        const float dxf = (cf * ax - sf * ay) - pcx, dyf = (sf * ax + cf * ay) - pcy;
        const int mx = x + (int)roundf(dxf), my = y + (int)roundf(dyf);
        if (mx < 0 || my < 0 || mx >= w || my >= h) continue;
        const int d = (int)ref[(int64_t)y * pitch + x] - (int)mov[(int64_t)my * pitch + mx];
        s += (unsigned long long)(d * d); n++;
    }
    __shared__ unsigned long long s_s[4]; __shared__ unsigned s_n[4];
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
    if ((threadIdx.x & 31) == 0) { s_s[threadIdx.x >> 5] = s; s_n[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        ssd[(size_t)f * ncand + c] = s_s[0] + s_s[1] + s_s[2] + s_s[3];
        cnt[(size_t)f * ncand + c] = s_n[0] + s_n[1] + s_n[2] + s_n[3];
    }
}

// is candidate (sa, na, ia) better than (sb, nb, ib)?  exact fraction compare, ties -> lower index
__device__ __forceinline__ bool better(unsigned long long sa, unsigned na, int ia, unsigned long long sb, unsigned nb, int ib)
{
    if (ib < 0) return true;
    if (ia < 0) return false;
    const unsigned long long l = sa * (unsigned long long)nb, r = sb * (unsigned long long)na;
    return l < r || (l == r && ia < ib);
}

// One block per frame: best candidate.  Writes result3 = (angle index a or -1, bx, by) and, when `next` is given, the next
// stage's search (angle table centre, shift centre scaled by `next_scale`); when `pose` is given, the frame's final pose
// (bx, by in full-resolution pixels = level pixels * pose_scale, cos, sin).
__global__ void __launch_bounds__(256)
prealign_pick_kernel(const unsigned long long* __restrict__ ssd, const unsigned* __restrict__ cnt, int ncand, int w, int h, int sub, int R,
                     const float* __restrict__ cs, const FrameSearch* __restrict__ fs, int step, int ref_idx, int zero_idx,
                     int* __restrict__ result3, FrameSearch* __restrict__ next, int next_scale, int next_half, int next_step,
                     float* __restrict__ pose, int pose_scale)
{
    const int f = blockIdx.x;
    const unsigned total = (unsigned)(((w + sub - 1) / sub) * ((h + sub - 1) / sub)), cnt_min = total / 4 > 0 ? total / 4 : 1;
    unsigned long long bs = 0; unsigned bn = 1; int bi = -1;
    for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
        const unsigned n = cnt[(size_t)f * ncand + c];
        if (n < cnt_min) continue;
        const unsigned long long s = ssd[(size_t)f * ncand + c];
        if (better(s, n, c, bs, bn, bi)) { bs = s; bn = n; bi = c; }
    }
    __shared__ unsigned long long s_s[256]; __shared__ unsigned s_n[256]; __shared__ int s_i[256];
    s_s[threadIdx.x] = bs; s_n[threadIdx.x] = bn; s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o && better(s_s[threadIdx.x + o], s_n[threadIdx.x + o], s_i[threadIdx.x + o], s_s[threadIdx.x], s_n[threadIdx.x], s_i[threadIdx.x])) {
            s_s[threadIdx.x] = s_s[threadIdx.x + o]; s_n[threadIdx.x] = s_n[threadIdx.x + o]; s_i[threadIdx.x] = s_i[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int S = 2 * R + 1, best = (f == ref_idx) ? -1 : s_i[0];
        const FrameSearch q = fs[f];
        int a = -1, bx = 0, by = 0, tidx = zero_idx;
        if (best >= 0) {
            a = best / (S * S);
            const int r = best - a * S * S, iy = r / S, ix = r - iy * S;
            bx = q.cx + ix - R; by = q.cy + iy - R; tidx = q.idx0 + a * step;
        }
        result3[3 * f] = a; result3[3 * f + 1] = bx; result3[3 * f + 2] = by;
        if (next) { FrameSearch nq; nq.idx0 = tidx - next_half * next_step; nq.cx = bx * next_scale; nq.cy = by * next_scale; nq.pad = 0; next[f] = nq; }
        if (pose) {
            pose[4 * f] = (float)(bx * pose_scale); pose[4 * f + 1] = (float)(by * pose_scale);
            pose[4 * f + 2] = cs[2 * tidx]; pose[4 * f + 3] = cs[2 * tidx + 1];
        }
    }
}

__global__ void prealign_init_kernel(FrameSearch* fs, int n, int idx0)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n) { FrameSearch q; q.idx0 = idx0; q.cx = 0; q.cy = 0; q.pad = 0; fs[f] = q; }
}

// pose of frame `to` relative to frame `from` (both given relative to the reference): rotation theta_to - theta_from,
// shift R(theta_from) (b_to - b_from).  A pair whose relative rotation exceeds PAIR_MAX_DEG is ruled out of the shift
// consolidation (bit k of *pair_valid cleared): the tile matcher displaces patches but does not rotate their content
// (kernel.cu:299-311), and 16 x 16 patches decorrelate beyond that.
constexpr float PAIR_COS_MIN = 0.96126169593831886f;        // cos(16 degrees)
__global__ void pair_pose_kernel(const float* __restrict__ pose, PairTable pt, int m, float* __restrict__ pair_pose, unsigned long long* __restrict__ pair_valid)
{
    const int k = threadIdx.x;
    bool ok = false;
    if (k < m) {
        const float* pi = pose + 4 * pt.from[k]; const float* pj = pose + 4 * pt.to[k];
        const float dbx = pj[0] - pi[0], dby = pj[1] - pi[1];
        const float ci = pi[2], si = pi[3], cj = pj[2], sj = pj[3];
        const float cr = cj * ci + sj * si;
        pair_pose[4 * k] = ci * dbx - si * dby;
        pair_pose[4 * k + 1] = si * dbx + ci * dby;
        pair_pose[4 * k + 2] = cr;
        pair_pose[4 * k + 3] = sj * ci - cj * si;
        ok = cr >= PAIR_COS_MIN;
    }
    const unsigned lo = __ballot_sync(0xffffffffu, ok);
    __shared__ unsigned s_w[2];
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = lo;
    __syncthreads();
    if (threadIdx.x == 0 && pair_valid) *pair_valid = (unsigned long long)s_w[0] | ((unsigned long long)s_w[1] << 32);
}

}  // namespace

int launch_prealign_stage(const uint8_t* img, int64_t pitch, int64_t frame_stride, int w, int h, int n_frames, int ref_idx,
                          const float* cs, int zero_idx, void* fs_in, int step, int n_ang, int R, int sub,
                          unsigned long long* ssd, unsigned* cnt, int* result3,
                          void* fs_next, int next_scale, int next_half, int next_step, float* pose, int pose_scale, cudaStream_t st)
{
    const int S = 2 * R + 1, ncand = n_ang * S * S;
    prealign_score_kernel<<<dim3(ncand, n_frames), 128, 0, st>>>(img, pitch, frame_stride, w, h, ref_idx, cs, (const FrameSearch*)fs_in, step, R, sub, ssd, cnt, ncand);
    MFSR_LAUNCH_CHECK();
    prealign_pick_kernel<<<n_frames, 256, 0, st>>>(ssd, cnt, ncand, w, h, sub, R, cs, (const FrameSearch*)fs_in, step, ref_idx, zero_idx,
                                                   result3, (FrameSearch*)fs_next, next_scale, next_half, next_step, pose, pose_scale);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

int launch_prealign_init(void* fs, int n, int idx0, cudaStream_t st)
{
    prealign_init_kernel<<<1, 64, 0, st>>>((FrameSearch*)fs, n, idx0);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

int launch_pair_pose(const float* pose, const PairTable& pt, int m, float* pair_pose, unsigned long long* pair_valid, cudaStream_t st)
{
    if (m < 1 || m > 64) return MFSR_OK;
    pair_pose_kernel<<<1, 64, 0, st>>>(pose, pt, m, pair_pose, pair_valid);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

}  // namespace mfsr

using namespace mfsr;

// One search stage on one image pair (test entry point; the pipeline batches all frames of a burst).
extern "C" int mfsr_stage_prealign_search(const uint8_t* ref, const uint8_t* mov, int64_t pitch, int w, int h,
                                          const float* cs_table, int n_table, int idx0, int step, int n_ang, int cx, int cy, int R, int sub,
                                          int* out3, void* stream)
{
    if (!ref || !mov || !cs_table || !out3 || w < 8 || h < 8 || pitch < w || n_ang < 1 || step < 1 || R < 0 || R > 16 || sub < 1) return MFSR_E_INVALID;
    if (idx0 < 0 || idx0 + (n_ang - 1) * step >= n_table) return MFSR_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int S = 2 * R + 1, ncand = n_ang * S * S;
    // a two-frame "stack": frame 0 = ref, frame 1 = mov (frame stride = their address difference)
    const int64_t fstride = (const char*)mov - (const char*)ref;
    char* scratch = nullptr;
    MFSR_CUDA_TRY(cudaMallocAsync(&scratch, (size_t)2 * ncand * 12 + 256, st));
    unsigned long long* ssd = (unsigned long long*)scratch;
    unsigned* cnt = (unsigned*)(scratch + (size_t)2 * ncand * 8);
    FrameSearch hfs[2] = {{idx0, cx, cy, 0}, {idx0, cx, cy, 0}};
    FrameSearch* dfs = (FrameSearch*)(scratch + (size_t)2 * ncand * 12);
    MFSR_CUDA_TRY(cudaMemcpyAsync(dfs, hfs, sizeof(hfs), cudaMemcpyHostToDevice, st));
    int* res = (int*)(dfs + 2);
    int rc = launch_prealign_stage(ref, pitch, fstride, w, h, 2, 0, cs_table, idx0, dfs, step, n_ang, R, sub, ssd, cnt, res,
                                   nullptr, 1, 0, 1, nullptr, 1, st);
    if (rc == MFSR_OK) {
        cudaError_t e = cudaMemcpyAsync(out3, res + 3, 12, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) rc = (int)e;
    }
    cudaFreeAsync(scratch, st);
    return rc;
}
