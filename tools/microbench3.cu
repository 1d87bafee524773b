// Does FFMA2 (packed fp32) leave issue slots free?  Mix N FFMA2 with M independent ALU/LDS ops per loop body.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int ITERS = 2048;
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// MODE 0: 16 FFMA2 ; 1: +8 LOP3 ; 2: +16 LOP3 ; 3: +16 MOV-like (prmt) ; 4: +8 LDS.128 uniform ; 5: +32 LOP3 ; 6: 16 FFMA2 + 16 FMUL (scalar, half rate) ; 7: 32 scalar FFMA + 16 LOP3
template <int MODE>
__global__ void k(float* out, float a, float b, unsigned m) {
    __shared__ float4 sm[256];
    sm[threadIdx.x & 255] = make_float4(a, b, a, b);
    __syncthreads();
    u64 acc[16];
    unsigned x[32];
    float f[32];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = pk(threadIdx.x * 0.001f + i, i);
#pragma unroll
    for (int i = 0; i < 32; i++) { x[i] = threadIdx.x + i; f[i] = threadIdx.x * 0.5f + i; }
    u64 A = pk(a, a), B = pk(b, b);
    for (int it = 0; it < ITERS; it++) {
        if (MODE != 7) {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = fma2(acc[i], A, B);
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
        }
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(m), "r"(x[i + 8]));
        }
        if (MODE == 2 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(m), "r"(x[i + 16]));
        }
        if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 32; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(m), "r"(x[(i + 7) & 31]));
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(x[i]) : "r"(x[i + 16]));
        }
        if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float4 v = sm[(it + i) & 255];
                f[i] += v.x;   // 1 scalar FADD per load keeps it alive
            }
        }
        if (MODE == 6) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(a));
        }
    }
    float s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
#pragma unroll
    for (int i = 0; i < 32; i++) { t ^= x[i]; s += f[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); for (int i = 0; i < 5; i++) f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / 5;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
    float* out; CK(cudaMalloc(&out, 4 * sms * 4 * 256));
    const char* names[] = {"16 FFMA2", "16 FFMA2 + 8 LOP3", "16 FFMA2 + 16 LOP3", "16 FFMA2 + 16 PRMT", "16 FFMA2 + 8 LDS.128 + 8 FADD", "16 FFMA2 + 32 LOP3", "16 FFMA2 + 16 FMUL", "32 FFMA + 16 LOP3"};
    for (int wps : {4, 2, 1}) {   // warps per SMSP
        int tpb = 128 * wps, blocks = sms;
        printf("--- %d warp(s) per SMSP\n", wps);
        float ms[8];
        ms[0] = timeit([&] { k<0><<<blocks, tpb>>>(out, 1.0001f, 0.5f, 3); });
        ms[1] = timeit([&] { k<1><<<blocks, tpb>>>(out, 1.0001f, 0.5f, 3); });
        ms[2] = timeit([&] { k<2><<<blocks, tpb>>>(out, 1.0001f, 0.5f, 3); });
        ms[3] = timeit([&] { k<3><<<blocks, tpb>>>(out, 1.0001f, 0.5f, 3); });
        ms[4] = timeit([&] { k<4><<<blocks, tpb>>>(out, 1.0001f, 0.5f, 3); });
        ms[5] = timeit([&] { k<5><<<blocks, tpb>>>(out, 1.0001f, 0.5f, 3); });
        ms[6] = timeit([&] { k<6><<<blocks, tpb>>>(out, 1.0001f, 0.5f, 3); });
        ms[7] = timeit([&] { k<7><<<blocks, tpb>>>(out, 1.0001f, 0.5f, 3); });
        for (int i = 0; i < 8; i++)
            printf("%-32s %.3f ms  -> %.2f cycles per loop body per SMSP-warp @1.965GHz (x%d warps = %.1f)\n", names[i], ms[i],
                   ms[i] * 1e-3 * 1.965e9 / ITERS / wps, wps, ms[i] * 1e-3 * 1.965e9 / ITERS);
    }
    return 0;
}
