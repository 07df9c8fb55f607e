// common.cuh — shared device helpers for the sm_100a burst-SR kernels.
//
// Arithmetic contract: every place where a DISCRETE decision (round / floor /
// arg-min / quantise) hangs on an fp32 value uses the *_rn intrinsics so that
// nvcc cannot contract mul+add into FMA; those values are then bit-identical
// to a strict-IEEE CPU evaluation of the same formula.  Everything else is
// left to the compiler (FMA contraction on) and is checked by tolerance.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include "../../include/mfsr.h"

#define MFSR_CUDA_TRY(expr)                                   \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return (int)_e;                \
    } while (0)

#define MFSR_LAUNCH_CHECK()                                   \
    do {                                                      \
        cudaError_t _e = cudaGetLastError();                  \
        if (_e != cudaSuccess) return (int)_e;                \
    } while (0)

namespace mfsr {

struct Cfa { int c[4]; };   // c_cfaPattern[2][2] row-major (DeBayerKernels.cu:41), passed by value
struct F3 { float v[3]; };

__host__ __device__ inline int clampi(int v, int lo, int hi)
{
#ifdef __CUDA_ARCH__
    return max(lo, min(v, hi));          // two VIMNMX (or one 3-input form) instead of compare + select pairs; lo <= hi everywhere
#else
    return v < lo ? lo : (v > hi ? hi : v);
#endif
}
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

template <typename T>
__device__ __forceinline__ const T* row_ptr(const T* base, int64_t pitch_bytes, int y)
{
    return (const T*)((const char*)base + pitch_bytes * (int64_t)y);
}
template <typename T>
__device__ __forceinline__ T* row_ptr(T* base, int64_t pitch_bytes, int y)
{
    return (T*)((char*)base + pitch_bytes * (int64_t)y);
}

// Kernels that run once per frame take the frame from blockIdx.z: one launch per burst instead of one per frame.
template <typename T>
__device__ __forceinline__ T* frame_ptr(T* base, int64_t frame_stride_bytes, unsigned frame)
{
    return (T*)((char*)base + frame_stride_bytes * (int64_t)frame);
}
template <typename T>
__device__ __forceinline__ const T* frame_ptr(const T* base, int64_t frame_stride_bytes, unsigned frame)
{
    return (const T*)((const char*)base + frame_stride_bytes * (int64_t)frame);
}
struct FrameStrides { int64_t s[3]; };

// ---- texture model (linear filter, clamp address, 1.8 fixed-point fraction) ----
// Models the cudaTextureObject_t fetches of the reference (e.g.
// DeBayerKernels.cu:401-402, opticalFlow.cu:36-41,88, RobustnessModell.cu:58)
// with explicit, strictly rounded fp32 arithmetic.
__device__ __forceinline__ float q8(float a)
{
    return floorf(__fadd_rn(__fmul_rn(a, 256.0f), 0.5f)) * (1.0f / 256.0f);
}
struct TexAxis { int i0, i1; float a; };
__device__ __forceinline__ TexAxis tex_axis(float u, int n)
{
    TexAxis t;
    float xb = __fsub_rn(u, 0.5f);
    float f = floorf(xb);
    t.a = q8(__fsub_rn(xb, f));
    int i = (int)f;
    t.i0 = clampi(i, 0, n - 1);
    t.i1 = clampi(i + 1, 0, n - 1);
    return t;
}
__device__ __forceinline__ float tex_mix(float t00, float t10, float t01, float t11, float a, float b)
{
    float na = __fsub_rn(1.0f, a), nb = __fsub_rn(1.0f, b);
    float top = __fadd_rn(__fmul_rn(t00, na), __fmul_rn(t10, a));
    float bot = __fadd_rn(__fmul_rn(t01, na), __fmul_rn(t11, a));
    return __fadd_rn(__fmul_rn(top, nb), __fmul_rn(bot, b));
}
// normalised -> unnormalised as the texture unit does: (p + 0.5) / n_img * n_tex
__device__ __forceinline__ float tex_coord(float p_plus_half, int n_img, int n_tex)
{
    return __fmul_rn(__fdiv_rn(p_plus_half, (float)n_img), (float)n_tex);
}

__device__ __forceinline__ float srgb_gamma(float v)   // kernel.cu:380-390
{
    return v <= 0.0031308f ? 12.92f * v : (1.0f + 0.055f) * powf(v, 1.0f / 2.4f) - 0.055f;
}

}  // namespace mfsr
