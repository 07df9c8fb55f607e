// pipes.cu — instruction-issue microbenchmarks for the merge-kernel design (B200, sm_100a).
// Each kernel runs NW warps per SM on every SM and times an unrolled block of one instruction kind with clock64;
// prints warp-instructions per cycle per SM (4 schedulers: 4.0 = one per scheduler and cycle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2000
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

template <int KIND>
__global__ void __launch_bounds__(1024) bench(float* out, long long* cyc, float seed, int sel)
{
    extern __shared__ float sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = seed * i;
    __syncthreads();
    float a[16], b[4];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = seed * (i + 1 + tid);
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = seed + i;
    u64 p[8], q[4];
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = pack2(a[2 * i], a[2 * i + 1]);
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = pack2(b[i], b[(i + 1) & 3]);
    const int pr = (tid + sel) & 1, pr2 = sel & 1;
    unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + (tid & 31) * 4 + (tid >> 5) * 128 * 4 % 16384;
    unsigned saddr8 = (unsigned)__cvta_generic_to_shared(sm) + (tid & 31) * 8;
    unsigned saddr16 = (unsigned)__cvta_generic_to_shared(sm) + (tid & 31) * 16;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
        if (KIND == 0) {          // FFMA, three distinct register operands
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i & 3]), "f"(b[(i + 1) & 3]));
        } else if (KIND == 1) {   // FFMA2 (fma.rn.f32x2)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i & 7]) : "l"(q[i & 3]), "l"(q[(i + 1) & 3]));
        } else if (KIND == 2) {   // predicated FFMA, predicate true on half of the lanes
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("{.reg .pred t; setp.ne.s32 t, %3, 0; @t fma.rn.f32 %0, %1, %2, %0;}" : "+f"(a[i]) : "f"(b[i & 3]), "f"(b[(i + 1) & 3]), "r"(pr));
        } else if (KIND == 3) {   // FSEL
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("{.reg .pred t; setp.ne.s32 t, %3, 0; selp.f32 %0, %1, %2, t;}" : "=f"(a[i]) : "f"(a[(i + 1) & 15]), "f"(b[i & 3]), "r"(pr2));
        } else if (KIND == 4) {   // FADD2
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i & 7]) : "l"(q[i & 3]));
        } else if (KIND == 5) {   // FMUL2
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i & 7]) : "l"(q[i & 3]));
        } else if (KIND == 6) {   // LDS.32 conflict-free
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a[i]) : "r"(saddr + i * 128));
        } else if (KIND == 7) {   // LDS.64
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "r"(saddr8 + i * 256));
        } else if (KIND == 8) {   // LDS.128
#pragma unroll
            for (int i = 0; i < 4; i++) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[4 * i]), "=f"(a[4 * i + 1]), "=f"(a[4 * i + 2]), "=f"(a[4 * i + 3]) : "r"(saddr16 + i * 512));
        } else if (KIND == 9) {   // mix: 8 FFMA2 + 4 LDS.32 + 4 LOP3 per block of 16
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(q[i & 3]), "l"(q[(i + 1) & 3]));
#pragma unroll
            for (int i = 0; i < 4; i++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a[i]) : "r"(saddr + i * 128));
            int x = tid;
#pragma unroll
            for (int i = 0; i < 4; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(sel + i), "r"(it));
            a[8] += __int_as_float(x);
        } else if (KIND == 10) {  // FMUL (two register operands)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i & 3]));
        } else if (KIND == 11) {  // mix: 8 FFMA + 8 LOP3 (fma pipe + alu pipe dual use)
            int x = tid;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i & 3]), "f"(b[(i + 1) & 3]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(sel + i), "r"(it));
            }
            a[8] += __int_as_float(x);
        } else if (KIND == 12) {  // mix: 8 FFMA2 + 8 FSEL
#pragma unroll
            for (int i = 0; i < 8; i++) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(q[i & 3]), "l"(q[(i + 1) & 3]));
                asm volatile("{.reg .pred t; setp.ne.s32 t, %3, 0; selp.f32 %0, %1, %2, t;}" : "=f"(a[i]) : "f"(a[(i + 1) & 15]), "f"(b[i & 3]), "r"(pr2));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { float x, y; unpack2(p[i], x, y); s += x + y; }
    if (s == 1234.5f) out[0] = s;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int instr_per_iter, int threads)
{
    float* out; long long* cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(bench<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    bench<KIND><<<148, threads, 65536>>>(out, cyc, 1.0001f, 0);
    bench<KIND><<<148, threads, 65536>>>(out, cyc, 1.0001f, 0);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    const double wi = (double)(threads / 32) * ITER * instr_per_iter;
    printf("%-28s warps/SM %2d  warp-instr/clk/SM %.3f  (cycles %.0f)  err=%s\n", name, threads / 32, wi / avg, avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int th : {512, 1024}) {
        run<0>("FFMA 3-reg", 16, th);
        run<10>("FMUL 2-reg", 16, th);
        run<1>("FFMA2", 16, th);
        run<2>("FFMA predicated (half lanes)", 16, th);
        run<3>("FSEL", 16, th);
        run<4>("FADD2", 16, th);
        run<5>("FMUL2", 16, th);
        run<6>("LDS.32", 16, th);
        run<7>("LDS.64", 8, th);
        run<8>("LDS.128", 4, th);
        run<9>("mix 8 FFMA2+4 LDS+4 LOP3", 16, th);
        run<11>("mix 8 FFMA+8 LOP3", 16, th);
        run<12>("mix 8 FFMA2+8 FSEL", 16, th);
    }
    return 0;
}
