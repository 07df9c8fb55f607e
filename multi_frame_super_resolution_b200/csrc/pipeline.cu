// pipeline.cu — one-burst pipeline handle (SURVEY §3.2 A..I) and the misc C-ABI entry points.
//
// Mirrors the way finalProject/Project/multi_frame_sr.cpp:165-194 drives cv::superres
// (create -> configure -> give it the frames -> pull the result), with the reference's
// implicit conventions made explicit: one handle per GPU, an owned stream, one workspace
// allocation, status codes instead of silent failure.
#include "common.cuh"
#include "internal.h"
#include <cuda_fp16.h>
#include <new>
#include <string.h>
#include <vector>
#include <math.h>

namespace mfsr { namespace s2 { int merge_pf_capacity(); } }      // frames the scale-2 merge kernel keeps resident (merge_pf.cu)

using namespace mfsr;

namespace {

enum Stage { ST_UPLOAD, ST_FRONTEND, ST_ALIGN, ST_CONSOLIDATE, ST_FLOW, ST_KERNEL, ST_ROBUST, ST_FALLBACK, ST_MERGE, ST_DOWNLOAD, ST_COUNT };
const char* kStageNames[ST_COUNT] = {"upload", "frontend", "align", "consolidate", "flow", "kernel_params", "robustness", "fallback", "merge", "download"};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// global pre-alignment search (prealign.cu): angle table of 0.125 degree steps over +-21 degrees
constexpr int PA_FINE = 8, PA_MAXDEG = 21, PA_TABLE = 2 * PA_MAXDEG * PA_FINE + 1, PA_ZERO = PA_MAXDEG * PA_FINE;
constexpr int PA_A_DEG = 20, PA_A_ANG = 2 * PA_A_DEG + 1, PA_A_R = 8, PA_B_HALF = 8, PA_B_ANG = 2 * PA_B_HALF + 1, PA_B_R = 4, PA_B_SUB = 2;
constexpr int PA_SMALL = 192;

struct Level { int w, h, tx, ty; int64_t pitch, frame_stride; uint8_t* img; float2* shift; float2* pre; };

}  // namespace

struct mfsr_context {
    mfsr_params p;
    int device, max_w, max_h, max_frames;
    cudaStream_t stream;
    cudaEvent_t ev[ST_COUNT + 1];
    float stage_ms[ST_COUNT];
    bool timed;
    // workspace
    char* ws; size_t ws_bytes;
    // burst state
    int n, w, h, ref_idx, format; bool have_frames, ran;
    int launches;
    // buffers (all inside ws)
    uint16_t* raw; int64_t raw_pitch, raw_fs;          // workspace copy of the frames
    const uint16_t* rawp; int64_t rawp_pitch, rawp_fs; // the stack the kernels read: the copy, or the caller's frames in place
    float* rgb_half; int64_t rgbh_pitch, rgbh_fs;
    float* gray; int64_t gray_pitch, gray_fs;
    float* rgb_ref; int64_t rgb_pitch;
    float2* flowA; float2* flowB; int64_t flow_pitch, flow_fs;
    float4* mask; int64_t mask_pitch, mask_fs; float* rstats;
    float4* kern; int64_t kern_pitch;
    float* fallback; float* outbuf; int64_t out_pitch_own;
    float* part_sum; float* part_weight;               // partial sums of the frame-chunked merge (bursts of more than 10 frames), else null
    std::vector<Level> lv;
    PairTable pt; int m;
    int2* argmin; float2* one_to_one; float2* frame_shift; int* cons_status; float* cons_inv0;
    float2* flow_final;   // which of flowA/flowB holds the final flow
    // global pre-alignment (prealign.cu): extra pyramid levels below the matcher's, search scratch, poses
    std::vector<Level> pal;                             // levels lv.size() .. of the tracking pyramid (img only)
    int pa_la, pa_lb;                                   // search levels of stages A and B (global level numbers)
    float* pa_cs; char* pa_scratch; size_t pa_scratch_bytes;
    float* pose; float* pair_pose; int* pa_result;      // per frame (bx, by, cos, sin); per pair; per frame and stage (a, bx, by)
    unsigned long long* pair_valid;                     // bit k: pair k takes part in the shift consolidation
    // linear-filter textures over the tracking gray frames (the LK warp step samples the moved frame through the texture unit)
    cudaTextureObject_t gray_tex[CONS_MAX_N + 1]; int n_gray_tex; const float* gray_tex_base; int gray_tex_w, gray_tex_h;
    mfsr_merge_geom geom;
};

// float3 image -> the caller's output format (mfsr_run_format): two floats per thread over the dense image.
// F16: round-to-nearest-even halves.  U8: floor(v * 255 + 0.5) saturated, NaN -> 0 (what the reference program writes to its PNGs
// after GammasRGB, multi_frame_sr.cpp:207).
template <int FMT>
__global__ void __launch_bounds__(256)
convert_out_kernel(const float2* __restrict__ in, void* __restrict__ out, int64_t n2)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const float2 v = __ldg(in + i);
    if (FMT == MFSR_OUT_F16) ((__half2*)out)[i] = __floats2half2_rn(v.x, v.y);
    else {
        const float a = fminf(fmaxf(floorf(v.x * 255.0f + 0.5f), 0.0f), 255.0f), b = fminf(fmaxf(floorf(v.y * 255.0f + 0.5f), 0.0f), 255.0f);
        ((uchar2*)out)[i] = make_uchar2((unsigned char)(a == a ? a : 0.0f), (unsigned char)(b == b ? b : 0.0f));
    }
}

// Device-to-device copy of `rows` rows of `row_bytes` bytes between two pitched images (8-byte words when everything is 8-byte aligned,
// else 4-byte).  cudaMemcpy2DAsync moved the kept rows of a row band at ~210 GB/s (1.4 ms for the 297 MB of an 8-rank band of config 4,
// 11 ms for the whole 2.3 GB image); this runs at copy bandwidth.
template <typename T>
__global__ void __launch_bounds__(256)
copy_rows_kernel(const char* __restrict__ src, int64_t src_pitch, char* __restrict__ dst, int64_t dst_pitch, int words_per_row, int rows)
{
    const int y = blockIdx.y;
    const T* s = (const T*)(src + src_pitch * y);
    T* d = (T*)(dst + dst_pitch * y);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words_per_row; i += gridDim.x * blockDim.x) d[i] = __ldg(s + i);
}

static int copy_rows_d2d(const void* src, int64_t src_pitch, void* dst, int64_t dst_pitch, int64_t row_bytes, int rows, cudaStream_t st)
{
    if (rows <= 0 || row_bytes <= 0) return MFSR_OK;
    if (rows > 65535 || (row_bytes & 3)) {
        MFSR_CUDA_TRY(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, (size_t)row_bytes, rows, cudaMemcpyDeviceToDevice, st));
        return MFSR_OK;
    }
    const bool w8 = !(((uintptr_t)src | (uintptr_t)dst | (uintptr_t)src_pitch | (uintptr_t)dst_pitch | (uintptr_t)row_bytes) & 7);
    const int words = (int)(row_bytes / (w8 ? 8 : 4));
    const dim3 g((unsigned)std::min(8, cdiv(words, 256 * 4)), (unsigned)rows);
    if (w8) copy_rows_kernel<float2><<<g, 256, 0, st>>>((const char*)src, src_pitch, (char*)dst, dst_pitch, words, rows);
    else copy_rows_kernel<float><<<g, 256, 0, st>>>((const char*)src, src_pitch, (char*)dst, dst_pitch, words, rows);
    MFSR_LAUNCH_CHECK();
    return MFSR_OK;
}

static void make_geom(const mfsr_params& p, int w, int h, mfsr_merge_geom* g)
{
    g->raw_w = w; g->raw_h = h; g->scale = p.scale;
    if (p.band_global_h > 0) {
        // row band: only the kept rows are merged; taps of kept rows stay inside the band (halo), so the clamp range is the band
        const int keep0 = p.band_keep_row0, keepn = p.band_keep_rows > 0 ? p.band_keep_rows : h - p.band_keep_row0;
        g->out_w = w * p.scale; g->out_h = keepn * p.scale; g->org_x = 0; g->org_y = keep0 * p.scale;
        g->clamp_x0 = 0; g->clamp_x1 = w - 1; g->clamp_y0 = 0; g->clamp_y1 = h - 1;
    } else if (p.full_frame) {
        const int num = MFSR_SCALE_NUM(p.scale), den = MFSR_SCALE_DEN(p.scale);
        g->out_w = w * num / den; g->out_h = h * num / den; g->org_x = 0; g->org_y = 0;
        g->clamp_x0 = 0; g->clamp_x1 = w - 1; g->clamp_y0 = 0; g->clamp_y1 = h - 1;
    } else {
        // generalisation of DeBayerKernels.cu:398-423 (s = 2: org = dim/2, clamp = [dim/4, dim/4 + dim/2 - 1])
        const int s = p.scale;
        g->out_w = w; g->out_h = h;
        g->org_x = w * (s - 1) / 2; g->org_y = h * (s - 1) / 2;
        g->clamp_x0 = g->org_x / s; g->clamp_x1 = g->clamp_x0 + w / s - 1;
        g->clamp_y0 = g->org_y / s; g->clamp_y1 = g->clamp_y0 + h / s - 1;
    }
}

extern "C" int mfsr_abi_version(void) { return MFSR_ABI_VERSION; }

extern "C" const char* mfsr_error_string(int status)
{
    switch (status) {
        case MFSR_OK: return "ok";
        case MFSR_E_INVALID: return "invalid argument or unsupported configuration";
        case MFSR_E_STATE: return "call order violated";
        case MFSR_E_NOMEM: return "out of memory";
        case MFSR_E_NODEVICE: return "no compute-capability-10.x CUDA device";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown mfsr status";
    }
}

extern "C" int mfsr_default_params(mfsr_params* p)
{
    if (!p) return MFSR_E_INVALID;
    memset(p, 0, sizeof(*p));
    p->abi_version = MFSR_ABI_VERSION;
    p->scale = 2; p->full_frame = 1;
    p->cfa[0] = MFSR_RED; p->cfa[1] = MFSR_GREEN; p->cfa[2] = MFSR_GREEN; p->cfa[3] = MFSR_BLUE;
    for (int c = 0; c < 3; c++) { p->black_level[c] = 64.0f; p->white_level[c] = 1023.0f - 64.0f; }
    p->tile_size = 16; p->max_shift = 4; p->levels = 4; p->pair_span = 2; p->track_bits = 7; p->track_sigma = 0.5f;
    p->min_threshold = 1024.0f;     // flat SSD surface (max - min below ~2 grey levels rms over a 16x16 tile): zero shift instead of a noise arg-min (tools/threshold_sweep.py)
    p->lk_iterations = 3; p->lk_half_window = 3; p->lk_min_det = 1e-3f;
    p->Dth = 0.005f; p->Dtr = 0.012f; p->kDetail = 0.3f; p->kDenoise = 4.0f; p->kStretch = 4.0f; p->kShrink = 2.0f;
    p->tensor_box_radius = 2;
    p->alpha = 1e-3f; p->beta = 1e-5f; p->thresholdM = 0.8f; p->mask_erode_radius = 2;
    p->weight_threshold = 0.1f; p->merge_flags = 0;
    return MFSR_OK;
}

extern "C" int mfsr_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

static int validate_params(const mfsr_params* p)
{
    if (!p || p->abi_version != MFSR_ABI_VERSION) return MFSR_E_INVALID;
    {
        // integer scales 1..4, or a rational MFSR_SCALE_RATIONAL(num, den) with 1 <= num / den <= 4 on the full frame (not in a row band)
        const int num = MFSR_SCALE_NUM(p->scale), den = MFSR_SCALE_DEN(p->scale);
        if (p->scale < 1 || num < den || num > 4 * den || den > 16) return MFSR_E_INVALID;
        if (den > 1 && (!p->full_frame || p->band_global_h > 0)) return MFSR_E_INVALID;
    }
    if (p->tile_size < 4 || (p->tile_size & 3) || p->max_shift < 1 || p->levels < 1 || p->levels > 8) return MFSR_E_INVALID;
    if (p->pair_span < 1 || p->track_bits < 1 || p->track_bits > 8) return MFSR_E_INVALID;
    // exactness contract of the integer SSD (align.cu): 2 * T^2 * qmax^2 < 2^24
    const int64_t qmax = (1 << p->track_bits) - 1;
    if (2 * (int64_t)p->tile_size * p->tile_size * qmax * qmax >= (1ll << 24)) return MFSR_E_INVALID;
    if (p->lk_iterations < 0 || p->lk_half_window < 1 || p->lk_half_window > 4) return MFSR_E_INVALID;
    if (p->tensor_box_radius < 0 || p->tensor_box_radius > 3 || p->mask_erode_radius < 0 || p->mask_erode_radius > 8) return MFSR_E_INVALID;
    for (int i = 0; i < 4; i++) if (p->cfa[i] < 0 || p->cfa[i] > 2) return MFSR_E_INVALID;
    if ((p->prealign != 0 && p->prealign != 1) || (p->lk_texture != 0 && p->lk_texture != 1)) return MFSR_E_INVALID;
    if (p->band_global_h < 0 || p->band_row0 < 0 || p->band_keep_row0 < 0 || p->band_keep_rows < 0 || p->band_margin < 0) return MFSR_E_INVALID;
    if (p->band_global_h > 0) {
        const int grid = p->tile_size << (p->levels - 1);
        if (!p->full_frame || p->base_rotation != 0.0f || p->prealign || (p->band_row0 % grid) || p->band_row0 >= p->band_global_h) return MFSR_E_INVALID;
    }
    return MFSR_OK;
}

// measured pairs (i, j), i < j: every pair at most pair_span frames apart; with the global pre-alignment also every frame
// against the reference frame (distant frames may be rotated too far against their neighbours to be matched, prealign.cu).
// Lexicographic order.  Returns the count; fills from / to when given.
static int build_pairs(const mfsr_params& p, int n, int ref_idx, int8_t* from, int8_t* to, int cap)
{
    int m = 0;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            const bool take = (j - i <= p.pair_span) || (p.prealign && (i == ref_idx || j == ref_idx));
            if (!take) continue;
            if (from && m < cap) { from[m] = (int8_t)i; to[m] = (int8_t)j; }
            m++;
        }
    return m;
}

// carve the workspace; returns required bytes. If base == nullptr only measures.
static size_t carve(mfsr_context* c, char* base, int n, int w, int h)
{
    size_t off = 0;
    auto take = [&](size_t bytes) -> char* { char* p = base ? base + off : nullptr; off += align_up(bytes, 512); return p; };
    const mfsr_params& p = c->p;
    const int hw = w / 2, hh = h / 2;
    c->raw_pitch = align_up((size_t)w * 2, 16); c->raw_fs = c->raw_pitch * h;
    c->raw = (uint16_t*)take((size_t)c->raw_fs * n);
    c->rgbh_pitch = (int64_t)hw * 12; c->rgbh_fs = c->rgbh_pitch * hh;
    c->rgb_half = (float*)take((size_t)c->rgbh_fs * n);
    c->gray_pitch = align_up((size_t)w * 4, 32); c->gray_fs = align_up((size_t)c->gray_pitch * h, 512);      // texture alignment (pitch 32 B, base 512 B)
    c->gray = (float*)take((size_t)c->gray_fs * n);
    c->rgb_pitch = (int64_t)w * 12;
    c->rgb_ref = (float*)take((size_t)c->rgb_pitch * h);
    c->flow_pitch = align_up((size_t)w * 8, 16); c->flow_fs = c->flow_pitch * h;
    c->flowA = (float2*)take((size_t)c->flow_fs * n);
    c->flowB = (float2*)take((size_t)c->flow_fs * n);
    c->mask_pitch = (int64_t)hw * 16; c->mask_fs = c->mask_pitch * hh;
    c->mask = (float4*)take((size_t)c->mask_fs * n);
    c->rstats = (float*)take((size_t)hw * hh * 24);            // reference patch statistics of the robustness model (6 floats per half-res pixel)
    c->kern_pitch = (int64_t)w * 16;
    c->kern = (float4*)take((size_t)c->kern_pitch * h);
    mfsr_merge_geom g; make_geom(p, w, h, &g);
    c->out_pitch_own = (int64_t)g.out_w * 12;
    c->fallback = (float*)take((size_t)c->out_pitch_own * (g.out_h + 2));      // + 2: a band's merge window grows by one
    c->outbuf = (float*)take((size_t)c->out_pitch_own * (g.out_h + 2));        //      row at each interior seam
    c->part_sum = c->part_weight = nullptr;
    if (n > s2::merge_pf_capacity() && p.scale == 2) {      // more frames than the merge tile holds at once: merged in chunks of frames (partial sums)
        c->part_sum = (float*)take((size_t)c->out_pitch_own * (g.out_h + 2));
        c->part_weight = (float*)take((size_t)c->out_pitch_own * (g.out_h + 2));
    }
    // measured pairs (upper bound over the possible reference frames when pre-aligning)
    int m = build_pairs(p, n, 0, nullptr, nullptr, 0);
    if (p.prealign) for (int r = 1; r < n; r++) { const int mr = build_pairs(p, n, r, nullptr, nullptr, 0); if (mr > m) m = mr; }
    if (m < 1) m = 1;
    // pyramid
    c->lv.clear();
    int lw = w, lh = h;
    for (int l = 0; l < p.levels; l++) {
        Level L; L.w = lw; L.h = lh;
        L.tx = (lw - 2 * p.max_shift) / p.tile_size; L.ty = (lh - 2 * p.max_shift) / p.tile_size;
        if (L.tx < 1 || L.ty < 1) break;
        L.pitch = align_up((size_t)lw, 16); L.frame_stride = L.pitch * lh;
        L.img = (uint8_t*)take((size_t)L.frame_stride * n);
        L.shift = (float2*)take((size_t)L.tx * L.ty * 8 * m);
        L.pre = (float2*)take((size_t)L.tx * L.ty * 8 * m);
        c->lv.push_back(L);
        lw /= 2; lh /= 2;
    }
    // global pre-alignment: the tracking pyramid continues below the matcher's levels until the longer side is <= PA_SMALL
    c->pal.clear(); c->pa_la = c->pa_lb = 0; c->pa_cs = nullptr; c->pa_scratch = nullptr; c->pose = c->pair_pose = nullptr; c->pa_result = nullptr; c->pair_valid = nullptr;
    if (p.prealign && !c->lv.empty()) {
        int la = 0, aw = w, ah = h;
        while ((aw > PA_SMALL || ah > PA_SMALL) && aw / 2 >= 16 && ah / 2 >= 16) { aw /= 2; ah /= 2; la++; }
        c->pa_la = la; c->pa_lb = la >= 2 ? la - 2 : 0;
        int pw = c->lv.back().w, ph = c->lv.back().h;
        for (int l = (int)c->lv.size(); l <= la; l++) {
            pw /= 2; ph /= 2;
            Level L = {}; L.w = pw; L.h = ph; L.pitch = align_up((size_t)pw, 16); L.frame_stride = L.pitch * ph;
            L.img = (uint8_t*)take((size_t)L.frame_stride * n);
            c->pal.push_back(L);
        }
        c->pa_cs = (float*)take((size_t)PA_TABLE * 8);
        const size_t ncand = (size_t)PA_A_ANG * (2 * PA_A_R + 1) * (2 * PA_A_R + 1);
        c->pa_scratch_bytes = (size_t)n * ncand * 12 + (size_t)n * 2 * 16;
        c->pa_scratch = take(c->pa_scratch_bytes);
        c->pose = (float*)take((size_t)n * 16);
        c->pair_pose = (float*)take((size_t)m * 16);
        c->pa_result = (int*)take((size_t)n * 2 * 12);
        c->pair_valid = (unsigned long long*)take(8);
    }
    if (!c->lv.empty()) {
        const size_t nt = (size_t)c->lv[0].tx * c->lv[0].ty;
        c->argmin = (int2*)take(nt * 8 * m);
        c->one_to_one = (float2*)take(nt * 8 * (n > 1 ? n - 1 : 1));
        c->frame_shift = (float2*)take(nt * 8 * n);
        c->cons_status = (int*)take(nt * 4);
        c->cons_inv0 = (float*)take((size_t)(CONS_MAX_N * CONS_MAX_N + 1) * 4);
    }
    return off;
}

extern "C" int mfsr_create(const mfsr_params* params, int device, int max_w, int max_h, int max_frames, mfsr_handle* out)
{
    if (!out) return MFSR_E_INVALID;
    *out = nullptr;
    int rc = validate_params(params);
    if (rc) return rc;
    if (max_w < 64 || max_h < 64 || (max_w & 1) || (max_h & 1) || max_frames < 1 || max_frames > CONS_MAX_N + 1) return MFSR_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return MFSR_E_NODEVICE; }
    if (device < 0 || device >= ndev) return MFSR_E_INVALID;
    int major = 0;
    MFSR_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return MFSR_E_NODEVICE;      // the library is compiled for sm_100a only
    MFSR_CUDA_TRY(cudaSetDevice(device));
    mfsr_context* c = new (std::nothrow) mfsr_context();
    if (!c) return MFSR_E_NOMEM;
    c->p = *params; c->device = device; c->max_w = max_w; c->max_h = max_h; c->max_frames = max_frames;
    c->have_frames = false; c->ran = false; c->timed = true; c->ws = nullptr; c->launches = 0;
    c->n_gray_tex = 0; c->gray_tex_base = nullptr; c->gray_tex_w = c->gray_tex_h = 0;
    c->ws_bytes = carve(c, nullptr, max_frames, max_w, max_h);
    if (c->lv.empty()) { delete c; return MFSR_E_INVALID; }
    cudaError_t e = cudaMalloc(&c->ws, c->ws_bytes);
    if (e != cudaSuccess) { cudaGetLastError(); delete c; return MFSR_E_NOMEM; }
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { cudaFree(c->ws); delete c; return (int)e; }
    for (int i = 0; i <= ST_COUNT; i++) cudaEventCreate(&c->ev[i]);
    memset(c->stage_ms, 0, sizeof(c->stage_ms));
    *out = c;
    return MFSR_OK;
}

extern "C" int mfsr_destroy(mfsr_handle h)
{
    if (!h) return MFSR_E_INVALID;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int f = 0; f < h->n_gray_tex; f++) cudaDestroyTextureObject(h->gray_tex[f]);
    for (int i = 0; i <= ST_COUNT; i++) cudaEventDestroy(h->ev[i]);
    cudaStreamDestroy(h->stream);
    cudaFree(h->ws);
    delete h;
    return MFSR_OK;
}

extern "C" int mfsr_output_size(mfsr_handle h, int width, int height, int* out_w, int* out_h)
{
    if (!h || !out_w || !out_h) return MFSR_E_INVALID;
    mfsr_merge_geom g; make_geom(h->p, width, height, &g);
    *out_w = g.out_w; *out_h = g.out_h;
    return MFSR_OK;
}

extern "C" int64_t mfsr_workspace_bytes(mfsr_handle h) { return h ? (int64_t)h->ws_bytes : 0; }
extern "C" void* mfsr_stream(mfsr_handle h) { return h ? (void*)h->stream : nullptr; }
extern "C" int mfsr_synchronize(mfsr_handle h)
{
    if (!h) return MFSR_E_INVALID;
    MFSR_CUDA_TRY(cudaSetDevice(h->device));
    MFSR_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MFSR_OK;
}

extern "C" int mfsr_set_frames(mfsr_handle h, const void* const* frames, int n, int width, int height, int64_t pitch,
                               int format, int ref_idx, int on_host)
{
    if (!h || !frames || n < 1 || n > h->max_frames || width > h->max_w || height > h->max_h) return MFSR_E_INVALID;
    if (width < 64 || height < 64 || (width & 1) || (height & 1) || ref_idx < 0 || ref_idx >= n || pitch < (int64_t)width * 2) return MFSR_E_INVALID;
    if (format != MFSR_FMT_BAYER_U16 && format != MFSR_FMT_GRAY_U16) return MFSR_E_INVALID;
    // everything that can be rejected is checked BEFORE the handle's state changes (a failed call leaves no half-configured burst)
    for (int f = 0; f < n; f++)
        if (!frames[f]) return MFSR_E_INVALID;
    if (build_pairs(h->p, n, ref_idx, nullptr, nullptr, 0) > CONS_MAX_M) return MFSR_E_INVALID;
    h->have_frames = false; h->ran = false;
    MFSR_CUDA_TRY(cudaSetDevice(h->device));
    if (carve(h, h->ws, n, width, height) > h->ws_bytes) return MFSR_E_INVALID;      // cannot happen for n, width, height within the handle's maxima
    if (h->lv.empty()) return MFSR_E_INVALID;
    h->n = n; h->w = width; h->h = height; h->ref_idx = ref_idx; h->format = format;
    make_geom(h->p, width, height, &h->geom);
    // measured pairs (i, j), 0 < j - i <= pair_span
    h->m = build_pairs(h->p, n, ref_idx, h->pt.from, h->pt.to, CONS_MAX_M);
    // textures over the gray frames: re-created only when the burst geometry changes (the handle's stream must be idle then:
    // a texture object may not be destroyed under a running kernel)
    if (h->gray_tex_base != h->gray || h->gray_tex_w != width || h->gray_tex_h != height || h->n_gray_tex != n) {
        if (h->n_gray_tex) MFSR_CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (int f = 0; f < h->n_gray_tex; f++) cudaDestroyTextureObject(h->gray_tex[f]);
        h->n_gray_tex = 0;
        if (h->p.band_global_h == 0 && h->p.lk_texture) {
            for (int f = 0; f < n; f++) {
                const int rc = make_gray_texture((const float*)((const char*)h->gray + h->gray_fs * f), h->gray_pitch, width, height, &h->gray_tex[f]);
                if (rc) return rc;
                h->n_gray_tex = f + 1;
            }
        }
        h->gray_tex_base = h->gray; h->gray_tex_w = width; h->gray_tex_h = height;
    }
    if (h->pa_cs) {
        // (cos, sin) of the candidate angles, computed in double and rounded once: the oracle uses the same table
        static float table[2 * PA_TABLE];
        static bool filled = false;
        if (!filled) {
            for (int i = 0; i < PA_TABLE; i++) {
                const double th = (double)(i - PA_ZERO) * (0.125 * 3.14159265358979323846 / 180.0);
                table[2 * i] = (float)cos(th); table[2 * i + 1] = (float)sin(th);
            }
            filled = true;
        }
        MFSR_CUDA_TRY(cudaMemcpyAsync(h->pa_cs, table, sizeof(table), cudaMemcpyHostToDevice, h->stream));
    }
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_UPLOAD], h->stream));
    // Device frames that form an evenly spaced, vector-load-aligned stack are used IN PLACE (no staging copy);
    // the caller keeps them alive until the run has finished.  Anything else is copied into the workspace.
    bool in_place = !on_host && !((uintptr_t)frames[0] & 15) && !(pitch & 15);
    int64_t stride = n > 1 ? (const char*)frames[1] - (const char*)frames[0] : pitch * height;
    for (int f = 1; in_place && f < n; f++) in_place = ((const char*)frames[f] - (const char*)frames[0]) == stride * f;
    if (in_place && n > 1 && (stride < pitch * height || (stride & 15))) in_place = false;
    if (in_place) {
        h->rawp = (const uint16_t*)frames[0]; h->rawp_pitch = pitch; h->rawp_fs = stride;
    } else {
        // dense rows on both sides: one linear copy per frame (the 2-D form is not guaranteed to collapse to it)
        const bool dense = pitch == (int64_t)width * 2 && h->raw_pitch == pitch;
        for (int f = 0; f < n; f++) {
            const cudaMemcpyKind kind = on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
            if (dense) MFSR_CUDA_TRY(cudaMemcpyAsync((char*)h->raw + h->raw_fs * f, frames[f], (size_t)pitch * height, kind, h->stream));
            else MFSR_CUDA_TRY(cudaMemcpy2DAsync((char*)h->raw + h->raw_fs * f, h->raw_pitch, frames[f], pitch, (size_t)width * 2, height, kind, h->stream));
        }
        h->rawp = h->raw; h->rawp_pitch = h->raw_pitch; h->rawp_fs = h->raw_fs;
    }
    h->have_frames = true; h->ran = false;
    return MFSR_OK;
}

#define RUN(expr) do { int _rc = (expr); if (_rc) return _rc; h->launches++; } while (0)

static int run_impl(mfsr_handle h, void* out, int64_t out_pitch, int out_on_host, bool sync_host, int out_format);

extern "C" int mfsr_run(mfsr_handle h, float* out, int64_t out_pitch, int out_on_host) { return run_impl(h, out, out_pitch, out_on_host, true, MFSR_OUT_F32); }
extern "C" int mfsr_run_async(mfsr_handle h, float* out, int64_t out_pitch, int out_on_host) { return run_impl(h, out, out_pitch, out_on_host, false, MFSR_OUT_F32); }
extern "C" int mfsr_run_format(mfsr_handle h, void* out, int64_t out_pitch, int out_on_host, int out_format, int async)
{
    return run_impl(h, out, out_pitch, out_on_host, !async, out_format);
}

static int run_impl(mfsr_handle h, void* out_any, int64_t out_pitch, int out_on_host, bool sync_host, int out_format)
{
    float* out = (float*)out_any;
    if (!h || !out) return MFSR_E_INVALID;
    if (out_format != MFSR_OUT_F32 && out_format != MFSR_OUT_F16 && out_format != MFSR_OUT_U8) return MFSR_E_INVALID;
    if (!h->have_frames) return MFSR_E_STATE;
    MFSR_CUDA_TRY(cudaSetDevice(h->device));
    const mfsr_params& p = h->p;
    cudaStream_t st = h->stream;
    const int n = h->n, w = h->w, hh = h->h, hw2 = w / 2, hh2 = hh / 2;
    h->launches = 0;
    int cfa[4];
    for (int i = 0; i < 4; i++) cfa[i] = (h->format == MFSR_FMT_GRAY_U16) ? MFSR_GREEN : p.cfa[i];
    float scale[3];
    for (int c = 0; c < 3; c++) scale[c] = 1.0f / p.white_level[c];

    // ---- A. front end: half-res RGB, tracking gray (float + 7-bit), pyramid, demosaiced reference
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_FRONTEND], st));
    const float maxVal = p.white_level[1] + p.black_level[1];
    // one launch per burst for every per-frame kernel (grid.z = frame)
    RUN(launch_subsample3(h->rawp, h->rawp_pitch, h->rawp_fs, h->rgb_half, h->rgbh_pitch, h->rgbh_fs, n, maxVal, hw2, hh2, cfa, st));
    RUN(launch_tracking_image(h->rawp, h->rawp_pitch, h->rawp_fs, h->gray, h->gray_pitch, h->gray_fs, h->lv[0].img, h->lv[0].pitch, h->lv[0].frame_stride, n,
                              w, hh, cfa, p.black_level, scale, p.track_sigma, p.track_bits, st));
    for (size_t l = 1; l < h->lv.size(); l++)
        RUN(launch_pyramid_down(h->lv[l - 1].img, h->lv[l - 1].pitch, h->lv[l - 1].frame_stride, h->lv[l - 1].w, h->lv[l - 1].h,
                                h->lv[l].img, h->lv[l].pitch, h->lv[l].frame_stride, n, st));
    RUN(mfsr_stage_demosaic((const uint16_t*)((const char*)h->rawp + h->rawp_fs * h->ref_idx), h->rawp_pitch, h->rgb_ref, h->rgb_pitch,
                            w, hh, cfa, p.black_level, scale, st));

    // ---- C. pyramid tile matching, all measured pairs per launch, coarse -> fine
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_ALIGN], st));
    const int L = (int)h->lv.size();
    const bool pre = p.prealign && n > 1 && h->pose;
    if (pre) {
        // ---- B. global pre-alignment: tracking pyramid below the matcher's levels, two exhaustive search stages, poses per frame / pair
        auto level_img = [&](int l, const uint8_t*& img, int64_t& pitch, int64_t& fstride, int& lw, int& lh) {
            const Level& V = l < L ? h->lv[l] : h->pal[l - L];
            img = V.img; pitch = V.pitch; fstride = V.frame_stride; lw = V.w; lh = V.h;
        };
        for (size_t k = 0; k < h->pal.size(); k++) {
            const Level& src = k == 0 ? h->lv.back() : h->pal[k - 1];
            RUN(launch_pyramid_down(src.img, src.pitch, src.frame_stride, src.w, src.h, h->pal[k].img, h->pal[k].pitch, h->pal[k].frame_stride, n, st));
        }
        const size_t ncandA = (size_t)PA_A_ANG * (2 * PA_A_R + 1) * (2 * PA_A_R + 1);
        unsigned long long* ssd = (unsigned long long*)h->pa_scratch;
        unsigned* cnt = (unsigned*)(h->pa_scratch + (size_t)n * ncandA * 8);
        char* fsA = h->pa_scratch + (size_t)n * ncandA * 12; char* fsB = fsA + (size_t)n * 16;
        const uint8_t* img; int64_t pitch, fstride; int lw, lh;
        RUN(launch_prealign_init(fsA, n, PA_ZERO - PA_A_DEG * PA_FINE, st));
        level_img(h->pa_la, img, pitch, fstride, lw, lh);
        RUN(launch_prealign_stage(img, pitch, fstride, lw, lh, n, h->ref_idx, h->pa_cs, PA_ZERO, fsA, PA_FINE, PA_A_ANG, PA_A_R, 1, ssd, cnt, h->pa_result,
                                  fsB, 1 << (h->pa_la - h->pa_lb), PA_B_HALF, 1, nullptr, 1, st));
        h->launches++;
        level_img(h->pa_lb, img, pitch, fstride, lw, lh);
        RUN(launch_prealign_stage(img, pitch, fstride, lw, lh, n, h->ref_idx, h->pa_cs, PA_ZERO, fsB, 1, PA_B_ANG, PA_B_R, PA_B_SUB, ssd, cnt, h->pa_result + 3 * n,
                                  nullptr, 1, 0, 1, h->pose, 1 << h->pa_lb, st));
        h->launches++;
        RUN(launch_pair_pose(h->pose, h->pt, h->m, h->pair_pose, h->pair_valid, st));
    }
    if (n > 1) {
        for (int l = L - 1; l >= 0; l--) {
            Level& V = h->lv[l];
            const int64_t grid_bytes = (int64_t)V.tx * V.ty * 8;
            if (l < L - 1) {
                Level& C = h->lv[l + 1];
                UpsampleBatch u = {};
                u.in = C.shift; u.in_pitch = (int64_t)C.tx * 8; u.in_pair_stride = (int64_t)C.tx * C.ty * 8;
                u.out = V.pre; u.out_pitch = (int64_t)V.tx * 8; u.out_pair_stride = grid_bytes;
                u.n_pairs = h->m; u.oldLevel = 1 << (l + 1); u.newLevel = 1 << l;
                u.oldCX = C.tx; u.oldCY = C.ty; u.newCX = V.tx; u.newCY = V.ty; u.oldT = p.tile_size; u.newT = p.tile_size;
                RUN(launch_upsample_shifts(u, st));
            }
            TileAlignBatch b = {};
            b.img = V.img; b.pitch = V.pitch; b.frame_stride = V.frame_stride; b.w = V.w; b.h = V.h;
            b.pre = (l < L - 1) ? V.pre : nullptr; b.pre_pitch = (int64_t)V.tx * 8; b.pre_pair_stride = grid_bytes;
            b.out = V.shift; b.out_pitch = (int64_t)V.tx * 8; b.out_pair_stride = grid_bytes;
            b.argmin = (l == 0) ? h->argmin : nullptr; b.argmin_pair_stride = grid_bytes;
            b.ssd = nullptr; b.ssd_pair_stride = 0;
            b.pt = h->pt; b.n_pairs = h->m; b.T = p.tile_size; b.M = p.max_shift; b.tx = V.tx; b.ty = V.ty;
            // the global pre-alignment is expressed in full-resolution pixels
            b.bsx = p.base_shift[0] / (float)(1 << l); b.bsy = p.base_shift[1] / (float)(1 << l); b.rot = p.base_rotation;
            b.threshold = p.min_threshold;
            b.pair_pose = pre ? h->pair_pose : nullptr; b.pose_scale = 1.0f / (float)(1 << l);
            RUN(launch_tile_align(b, st));
        }
    }
    // ---- D. per-tile least squares -> reference->frame tile shifts
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_CONSOLIDATE], st));
    const int tx = h->lv[0].tx, ty = h->lv[0].ty, nt = tx * ty;
    if (n > 1) {
        RUN(launch_consolidate(h->lv[0].shift, 1, nt, h->pt, h->m, n, nt, h->ref_idx, h->one_to_one, h->frame_shift, h->cons_status, h->cons_inv0, st, pre ? h->pair_valid : nullptr));
    } else {
        MFSR_CUDA_TRY(cudaMemsetAsync(h->frame_shift, 0, (size_t)nt * 8, st));
    }
    // ---- E. dense flow + Lucas-Kanade refinement
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_FLOW], st));
    float2* cur = h->flowA; float2* nxt = h->flowB;
    const int gh = p.band_global_h, gty = gh > 0 ? (gh - 2 * p.max_shift) / p.tile_size : 0;
    // Row-band mode with band_margin > 0: the per-pixel stages (flow, kernel parameters, robustness) only run on the kept
    // rows + margin, rows [ra, rb) of the band; the wide halo is only needed by the pyramid tile matcher.  Every kernel
    // below is a stencil with clamp addressing whose footprint (LK: 5 rows per sweep + |flow|) stays inside the margin for
    // the kept rows, so the kept rows' values do not change (tests/test_rowband_gpu.py checks bit-identity).
    int ra = 0, rb = hh;
    if (gh > 0 && p.band_margin > 0) {
        const int keepn = p.band_keep_rows > 0 ? p.band_keep_rows : hh - p.band_keep_row0;
        // ra is a multiple of 32 (the LK tile height): a pixel then sits at the same place inside its CTA's row groups as in the
        // full-frame run, which keeps the re-associated column sums bit-identical
        ra = p.band_keep_row0 - p.band_margin; if (ra < 0) ra = 0; ra &= ~31;
        rb = (p.band_keep_row0 + keepn + p.band_margin + 1) & ~1; if (rb > hh) rb = hh;
    }
    const int rh = rb - ra, gy0 = p.band_row0 + ra;
    RUN(launch_flow_from_tiles(h->frame_shift, (int64_t)tx * 8, tx, ty, (float2*)((char*)cur + h->flow_pitch * ra), h->flow_pitch,
                               w, rh, p.base_shift[0], p.base_shift[1], p.base_rotation, gh, gy0, gty, p.band_row0 / p.tile_size, st,
                               pre ? h->pose : nullptr, n, (int64_t)nt * 8, h->flow_fs));
    const float* gray_ref = (const float*)((const char*)h->gray + h->gray_fs * h->ref_idx + h->gray_pitch * ra);
    for (int it = 0; it < p.lk_iterations; it++) {
        if (h->n_gray_tex == n) {
            // texture form: one texture object per moved image, one launch per frame
            for (int f = 0; f < n; f++) {
                if (f == h->ref_idx) {   // reference against itself: Iz == 0 -> UV == 0, flow unchanged
                    MFSR_CUDA_TRY(cudaMemcpyAsync((char*)nxt + h->flow_fs * f + h->flow_pitch * ra, (char*)cur + h->flow_fs * f + h->flow_pitch * ra,
                                                  (size_t)h->flow_pitch * rh, cudaMemcpyDeviceToDevice, st));
                    continue;
                }
                RUN(launch_lk_iteration(gray_ref, (const float*)((const char*)h->gray + h->gray_fs * f + h->gray_pitch * ra), h->gray_pitch,
                                        (const float2*)((const char*)cur + h->flow_fs * f + h->flow_pitch * ra),
                                        (float2*)((char*)nxt + h->flow_fs * f + h->flow_pitch * ra),
                                        h->flow_pitch, w, rh, p.lk_half_window, p.lk_min_det, gh, gy0, st, h->gray_tex[f]));
            }
        } else {
            RUN(launch_lk_iteration(gray_ref, (const float*)((const char*)h->gray + h->gray_pitch * ra), h->gray_pitch,
                                    (const float2*)((const char*)cur + h->flow_pitch * ra), (float2*)((char*)nxt + h->flow_pitch * ra),
                                    h->flow_pitch, w, rh, p.lk_half_window, p.lk_min_det, gh, gy0, st, 0, n, h->gray_fs, h->flow_fs, h->ref_idx));
        }
        float2* t = cur; cur = nxt; nxt = t;
    }
    h->flow_final = cur;
    // ---- F. merge kernel parameters from the reference frame
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_KERNEL], st));
    RUN(mfsr_stage_kernel_params((const float*)((const char*)h->gray + h->gray_fs * h->ref_idx + h->gray_pitch * ra), h->gray_pitch,
                                 (float*)((char*)h->kern + h->kern_pitch * ra), h->kern_pitch,
                                 w, rh, p.tensor_box_radius, p.Dth, p.Dtr, p.kDetail, p.kDenoise, p.kStretch, p.kShrink, st));
    // ---- G. robustness masks
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_ROBUST], st));
    // all frames in one launch; with a min filter the fused kernel (certainty + row / column minima in shared memory)
    RUN(launch_robustness((const float*)((const char*)h->rgb_half + h->rgbh_fs * h->ref_idx + h->rgbh_pitch * (ra / 2)),
                          (const float*)((const char*)h->rgb_half + h->rgbh_pitch * (ra / 2)), h->rgbh_pitch, h->rgbh_fs,
                          (const float*)((const char*)cur + h->flow_pitch * ra), h->flow_pitch, h->flow_fs,
                          (float*)((char*)h->mask + h->mask_pitch * (ra / 2)), h->mask_pitch, h->mask_fs,
                          nullptr, 0, n, hw2, rh / 2, p.alpha, p.beta, p.thresholdM, p.mask_erode_radius, st, h->rstats));
    h->launches += 1;
    // ---- fallback image (ApplyWeighting's inOutImg): demosaiced reference on the output grid
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_FALLBACK], st));
    // Row-band mode: the reference leaves the 1-pixel border of the merge WINDOW untouched (DeBayerKernels.cu:391).  At an
    // interior seam that border must not exist, so the window is grown by one output row there, merged into the handle's
    // own buffer, and the kept rows are copied out.
    mfsr_merge_geom mg = h->geom;
    int ext_top = 0, ext_bot = 0;
    if (p.band_global_h > 0) {
        const int keepn = h->geom.out_h / p.scale;
        ext_top = (p.band_row0 + p.band_keep_row0 > 0) ? 1 : 0;
        ext_bot = (p.band_row0 + p.band_keep_row0 + keepn < p.band_global_h) ? 1 : 0;
        mg.org_y -= ext_top; mg.out_h += ext_top + ext_bot;
    }
    const bool staged = out_on_host || ext_top || ext_bot || out_format != MFSR_OUT_F32;
    RUN(mfsr_stage_fallback_upsample(h->rgb_ref, h->rgb_pitch, w, hh, h->fallback, h->out_pitch_own, &mg, st));
    // ---- H+I. fused merge + normalise (+ gamma)
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_MERGE], st));
    float* dst = staged ? h->outbuf : out;
    const int64_t dst_pitch = staged ? h->out_pitch_own : out_pitch;
    RUN(mfsr_stage_merge(h->rawp, h->rawp_pitch, h->rawp_fs, (const float*)h->mask, h->mask_pitch, h->mask_fs,
                         (const float*)cur, h->flow_pitch, h->flow_fs, (const float*)h->kern, h->kern_pitch,
                         h->fallback, h->out_pitch_own, dst, dst_pitch, h->part_sum, h->part_weight, h->part_sum ? h->out_pitch_own : 0, n, &mg, cfa,
                         p.white_level, p.black_level, p.weight_threshold, p.merge_flags & ~MFSR_MERGE_NO_FALLBACK, st));
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_DOWNLOAD], st));
    if (staged) {
        const cudaMemcpyKind kind = out_on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        const char* src = (const char*)h->outbuf + h->out_pitch_own * ext_top;
        int64_t src_pitch = h->out_pitch_own, row_bytes = (int64_t)h->geom.out_w * 12;
        if (out_format != MFSR_OUT_F32) {
            // converted image goes into the (now dead) fallback buffer, dense rows
            const int64_t n2 = (int64_t)h->geom.out_w * h->geom.out_h * 3 / 2;
            const unsigned blocks = (unsigned)((n2 + 255) / 256);
            if (out_format == MFSR_OUT_F16) convert_out_kernel<MFSR_OUT_F16><<<blocks, 256, 0, st>>>((const float2*)src, h->fallback, n2);
            else convert_out_kernel<MFSR_OUT_U8><<<blocks, 256, 0, st>>>((const float2*)src, h->fallback, n2);
            MFSR_LAUNCH_CHECK();
            h->launches++;
            const int bpp = out_format == MFSR_OUT_F16 ? 6 : 3;
            src = (const char*)h->fallback; src_pitch = row_bytes = (int64_t)h->geom.out_w * bpp;
        }
        if (out_pitch < row_bytes) return MFSR_E_INVALID;
        if (!out_on_host) { RUN(copy_rows_d2d(src, src_pitch, out, out_pitch, row_bytes, h->geom.out_h, st)); }
        else if (out_pitch == src_pitch) MFSR_CUDA_TRY(cudaMemcpyAsync(out, src, (size_t)out_pitch * h->geom.out_h, kind, st));
        else MFSR_CUDA_TRY(cudaMemcpy2DAsync(out, out_pitch, src, src_pitch, (size_t)row_bytes, h->geom.out_h, kind, st));
    }
    MFSR_CUDA_TRY(cudaEventRecord(h->ev[ST_COUNT], st));
    h->ran = true;
    if (out_on_host && sync_host) MFSR_CUDA_TRY(cudaStreamSynchronize(st));
    return MFSR_OK;
}

extern "C" int mfsr_get_stage_ms(mfsr_handle h, float* ms, int capacity)
{
    if (!h || !ms) return MFSR_E_INVALID;
    if (!h->ran) return MFSR_E_STATE;
    MFSR_CUDA_TRY(cudaSetDevice(h->device));
    MFSR_CUDA_TRY(cudaStreamSynchronize(h->stream));
    int nwr = 0;
    for (int i = 0; i < ST_COUNT && i < capacity; i++) {
        float t = 0.f;
        MFSR_CUDA_TRY(cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]));
        h->stage_ms[i] = t; ms[i] = t; nwr++;
    }
    return nwr;
}

extern "C" const char* mfsr_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : nullptr; }

extern "C" int mfsr_get_tile_grid(mfsr_handle h, int* tilesX, int* tilesY, int* n_pairs)
{
    if (!h || !h->have_frames) return MFSR_E_STATE;
    if (tilesX) *tilesX = h->lv[0].tx;
    if (tilesY) *tilesY = h->lv[0].ty;
    if (n_pairs) *n_pairs = h->n > 1 ? h->m : 0;
    return MFSR_OK;
}

extern "C" int mfsr_get_tile_argmin(mfsr_handle h, int pair, int32_t* host_int2)
{
    if (!h || !host_int2) return MFSR_E_INVALID;
    if (!h->ran || h->n < 2) return MFSR_E_STATE;
    if (pair < 0 || pair >= h->m) return MFSR_E_INVALID;
    MFSR_CUDA_TRY(cudaSetDevice(h->device));
    const size_t nt = (size_t)h->lv[0].tx * h->lv[0].ty;
    MFSR_CUDA_TRY(cudaMemcpyAsync(host_int2, h->argmin + nt * pair, nt * 8, cudaMemcpyDeviceToHost, h->stream));
    MFSR_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MFSR_OK;
}

extern "C" int mfsr_get_tile_shifts(mfsr_handle h, int frame, float* host_float2)
{
    if (!h || !host_float2) return MFSR_E_INVALID;
    if (!h->ran) return MFSR_E_STATE;
    if (frame < 0 || frame >= h->n) return MFSR_E_INVALID;
    MFSR_CUDA_TRY(cudaSetDevice(h->device));
    const size_t nt = (size_t)h->lv[0].tx * h->lv[0].ty;
    MFSR_CUDA_TRY(cudaMemcpyAsync(host_float2, h->frame_shift + nt * frame, nt * 8, cudaMemcpyDeviceToHost, h->stream));
    MFSR_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MFSR_OK;
}

extern "C" int mfsr_get_buffer(mfsr_handle h, const char* name, void** dev_ptr, int64_t* pitch, int64_t* frame_stride)
{
    if (!h || !name || !dev_ptr) return MFSR_E_INVALID;
    if (!h->have_frames) return MFSR_E_STATE;
    int64_t pi = 0, fs = 0; void* p = nullptr;
    if (!strcmp(name, "raw")) { p = (void*)h->rawp; pi = h->rawp_pitch; fs = h->rawp_fs; }
    else if (!strcmp(name, "rgb_half")) { p = h->rgb_half; pi = h->rgbh_pitch; fs = h->rgbh_fs; }
    else if (!strcmp(name, "gray")) { p = h->gray; pi = h->gray_pitch; fs = h->gray_fs; }
    else if (!strcmp(name, "gray_q0")) { p = h->lv[0].img; pi = h->lv[0].pitch; fs = h->lv[0].frame_stride; }
    else if (!strcmp(name, "rgb_ref")) { p = h->rgb_ref; pi = h->rgb_pitch; fs = 0; }
    else if (!strcmp(name, "flow")) { if (!h->ran) return MFSR_E_STATE; p = h->flow_final; pi = h->flow_pitch; fs = h->flow_fs; }
    else if (!strcmp(name, "mask")) { p = h->mask; pi = h->mask_pitch; fs = h->mask_fs; }
    else if (!strcmp(name, "kernel")) { p = h->kern; pi = h->kern_pitch; fs = 0; }
    else if (!strcmp(name, "fallback")) { p = h->fallback; pi = h->out_pitch_own; fs = 0; }
    else if (!strcmp(name, "pose")) { if (!h->pose || !h->ran) return MFSR_E_STATE; p = h->pose; pi = 16; fs = 16; }
    else if (!strcmp(name, "prealign_result")) { if (!h->pa_result || !h->ran) return MFSR_E_STATE; p = h->pa_result; pi = 12; fs = 12; }
    else if (!strcmp(name, "tile_shift0")) { p = h->lv[0].shift; pi = (int64_t)h->lv[0].tx * 8; fs = (int64_t)h->lv[0].tx * h->lv[0].ty * 8; }
    else return MFSR_E_INVALID;
    *dev_ptr = p;
    if (pitch) *pitch = pi;
    if (frame_stride) *frame_stride = fs;
    return MFSR_OK;
}

extern "C" int mfsr_last_launch_count(mfsr_handle h) { return h ? h->launches : 0; }
