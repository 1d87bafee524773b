// Which FP32 pipes do the packed (f32x2) and scalar forms use on sm_100a?  If packed ops run on one pipe only, scalar FFMA
// issued next to them is free.  Every op updates its own accumulator (32 independent chains per kind).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench4 microbench4.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int ITERS = 4096;
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

// NP2: packed fma per iter, NM2: packed mul, NA2: packed add, NS: scalar fma, NSM: scalar mul, NSA: scalar add
template <int NF2, int NM2, int NA2, int NS, int NSM, int NSA>
__global__ void k(float* out, float a, float b) {
    u64 f2[NF2 > 0 ? NF2 : 1], m2[NM2 > 0 ? NM2 : 1], a2[NA2 > 0 ? NA2 : 1];
    float s[NS > 0 ? NS : 1], sm[NSM > 0 ? NSM : 1], sa[NSA > 0 ? NSA : 1];
#pragma unroll
    for (int i = 0; i < NF2; i++) f2[i] = pk(threadIdx.x * 0.001f + i, i);
#pragma unroll
    for (int i = 0; i < NM2; i++) m2[i] = pk(1.0f + threadIdx.x * 1e-6f, 1.0f + i * 1e-6f);
#pragma unroll
    for (int i = 0; i < NA2; i++) a2[i] = pk(threadIdx.x * 0.01f, i);
#pragma unroll
    for (int i = 0; i < NS; i++) s[i] = threadIdx.x * 0.5f + i;
#pragma unroll
    for (int i = 0; i < NSM; i++) sm[i] = 1.0f + i * 1e-6f;
#pragma unroll
    for (int i = 0; i < NSA; i++) sa[i] = i;
    const u64 A = pk(a, a), B = pk(b, b);
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        constexpr int N = 32;
#pragma unroll
        for (int i = 0; i < N; i++) {          // interleave the kinds the way a compiler would schedule them
            if (i < NF2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(f2[i]) : "l"(A), "l"(B));
            if (i < NS) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a), "f"(b));
            if (i < NM2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(m2[i]) : "l"(A));
            if (i < NSM) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(sm[i]) : "f"(a));
            if (i < NA2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a2[i]) : "l"(B));
            if (i < NSA) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(sa[i]) : "f"(b));
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < NF2; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(f2[i])); r += lo + hi; }
#pragma unroll
    for (int i = 0; i < NM2; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(m2[i])); r += lo + hi; }
#pragma unroll
    for (int i = 0; i < NA2; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a2[i])); r += lo + hi; }
#pragma unroll
    for (int i = 0; i < NS; i++) r += s[i];
#pragma unroll
    for (int i = 0; i < NSM; i++) r += sm[i];
#pragma unroll
    for (int i = 0; i < NSA; i++) r += sa[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); for (int i = 0; i < 5; i++) f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / 5;
}
#define RUN(NAME, ...) { float ms = timeit([&] { k<__VA_ARGS__><<<blocks, tpb>>>(out, 1.0000001f, 1e-9f); }); CK(cudaGetLastError()); \
    printf("%-44s %7.3f ms -> %6.1f SMSP-cycles per iteration (all %d warps of the SMSP)\n", NAME, ms, ms * 1e-3 * 1.965e9 / ITERS, wps); }
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
    float* out; CK(cudaMalloc(&out, 4 * sms * 4 * 256 * 4));
    for (int wps : {3, 4}) {
        int tpb = 128 * wps, blocks = sms;
        printf("--- %d warps per SMSP; per warp and iteration:\n", wps);
        //                                         F2  M2  A2   S  SM  SA
        RUN("32 FFMA2",                            32,  0,  0,  0,  0,  0);
        RUN("32 FMUL2",                             0, 32,  0,  0,  0,  0);
        RUN("32 FADD2",                             0,  0, 32,  0,  0,  0);
        RUN("32 FFMA",                              0,  0,  0, 32,  0,  0);
        RUN("32 FMUL",                              0,  0,  0,  0, 32,  0);
        RUN("32 FADD",                              0,  0,  0,  0,  0, 32);
        RUN("32 FFMA2 + 32 FFMA",                  32,  0,  0, 32,  0,  0);
        RUN("32 FFMA2 + 16 FFMA",                  32,  0,  0, 16,  0,  0);
        RUN("32 FMUL2 + 32 FFMA",                   0, 32,  0, 32,  0,  0);
        RUN("32 FADD2 + 32 FFMA",                   0,  0, 32, 32,  0,  0);
        RUN("16 FMUL2 + 16 FFMA2 + 16 FADD2",      16, 16, 16,  0,  0,  0);
        RUN("16 FMUL2 + 16 FFMA2 + 16 FADD2 + 16 FFMA", 16, 16, 16, 16,  0,  0);
        RUN("16 FMUL2 + 16 FFMA2 + 16 FADD2 + 32 FFMA", 16, 16, 16, 32,  0,  0);
        RUN("16 FMUL2 + 32 FFMA2",                 32, 16,  0,  0,  0,  0);
        RUN("32 FFMA2 + 32 FMUL",                  32,  0,  0,  0, 32,  0);
        RUN("32 FFMA + 32 FMUL",                    0,  0,  0, 32, 32,  0);
        RUN("32 FFMA + 32 FADD",                    0,  0,  0, 32,  0, 32);
        RUN("32 FMUL + 32 FADD",                    0,  0,  0,  0, 32, 32);
    }
    return 0;
}
