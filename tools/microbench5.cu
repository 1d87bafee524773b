// Marginal issue cost of the non-FP32 instructions of the consumer loop next to packed FP32 (3 warps per scheduler, as in
// k_aggregate_ws): base = 16 FMUL2 + 16 FFMA2 + 16 FADD2 per iteration; add LDS.128 / LDS.32 / I2F.U8 / MOV / scalar FMUL.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench5 microbench5.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int ITERS = 4096;
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

template <int NLDS128, int NLDS32, int NI2F, int NMOV, int NFMUL, int PATTERN>
__global__ void k(float* out, float a, float b, int stride) {
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(a, b, a, b);
    __syncthreads();
    u64 f2[16], m2[16], a2[16];
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned iv[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    float fm[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { f2[i] = pk(threadIdx.x * 0.001f + i, i); m2[i] = pk(1.0f, 1.0f); a2[i] = pk(i, i); fm[i] = 1.0f + i; }
    const u64 A = pk(a, a), B = pk(b, b);
    const int lane = threadIdx.x & 31;
    // PATTERN 0: all lanes one address (broadcast); 1: our right-weight pattern (8 distinct quads, 4 lanes each);
    // 2: 14 distinct quads (unsheared mapping); 3: 32 distinct quads
    int base = PATTERN == 0 ? 0 : PATTERN == 1 ? (lane & 7) : PATTERN == 2 ? ((lane & 7) + 6 - 2 * (lane >> 3)) : lane;
    const float4* p = sm + base;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(m2[i]) : "l"(A));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(f2[i]) : "l"(A), "l"(B));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a2[i]) : "l"(B));
            if (i < NLDS128) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p + ((it * stride + i * 37) & 1023))));
                acc[i & 7] += v.x;            // one scalar add keeps the load alive (counted in the base of every variant? no: see NOTE)
            }
            if (i < NLDS32) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(reinterpret_cast<const float*>(sm) + ((lane + it * stride + i * 41) & 4095))));
                acc[i & 7] += v;
            }
            if (i < NI2F) { float v; asm volatile("cvt.rn.f32.u8 %0, %1;" : "=f"(v) : "r"(iv[i & 7] & 0xffu)); acc[i & 7] += v; }
            if (i < NMOV) asm volatile("mov.b32 %0, %1;" : "=r"(iv[i & 7]) : "r"(iv[(i + 1) & 7]));
            if (i < NFMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(fm[i]) : "f"(a));
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float lo, hi;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(f2[i])); r += lo + hi;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(m2[i])); r += lo + hi;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a2[i])); r += lo + hi;
        r += fm[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) r += acc[i] + iv[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); for (int i = 0; i < 5; i++) f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / 5;
}
static float g_base = 0;
#define RUN(NAME, N, ...) { float ms = timeit([&] { k<__VA_ARGS__><<<blocks, tpb>>>(out, 1.0000001f, 1e-9f, 1); }); CK(cudaGetLastError()); \
    float cyc = ms * 1e-3f * 1.965e9f / ITERS; if (N == 0) g_base = cyc; \
    printf("%-40s %7.1f SMSP-cycles per iteration", NAME, cyc); if (N > 0) printf("  -> %+.2f cycles per added instruction (x%d per warp, 3 warps)", (cyc - g_base) / (3 * N), N); printf("\n"); }
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
    float* out; CK(cudaMalloc(&out, 4 * sms * 4 * 256 * 4));
    int wps = 3, tpb = 128 * wps, blocks = sms;
    printf("3 warps per SMSP; base = 16 FMUL2 + 16 FFMA2 + 16 FADD2 per warp and iteration (NOTE: each load / convert also carries one scalar FADD)\n");
    //                                   L128 L32 I2F MOV FMUL PAT
    RUN("base (48 packed)",              0,  0,  0,  0,  0,  0, 0);
    RUN("+ 8 scalar FMUL",               8,  0,  0,  0,  0,  8, 0);
    RUN("+ 16 scalar FMUL",             16,  0,  0,  0,  0, 16, 0);
    RUN("+ 8 MOV",                       8,  0,  0,  0,  8,  0, 0);
    RUN("+ 16 MOV",                     16,  0,  0,  0, 16,  0, 0);
    RUN("+ 4 I2F.U8 (+4 FADD)",          4,  0,  0,  4,  0,  0, 0);
    RUN("+ 8 I2F.U8 (+8 FADD)",          8,  0,  0,  8,  0,  0, 0);
    RUN("+ 4 LDS.32 (+4 FADD)",          4,  0,  4,  0,  0,  0, 0);
    RUN("+ 5 LDS.128 broadcast (+5 FADD)", 5,  5,  0,  0,  0,  0, 0);
    RUN("+ 5 LDS.128 8 quads (+5 FADD)",   5,  5,  0,  0,  0,  0, 1);
    RUN("+ 5 LDS.128 14 quads (+5 FADD)",  5,  5,  0,  0,  0,  0, 2);
    RUN("+ 5 LDS.128 32 quads (+5 FADD)",  5,  5,  0,  0,  0,  0, 3);
    RUN("+ 8 LDS.128 8 quads (+8 FADD)",   8,  8,  0,  0,  0,  0, 1);
    return 0;
}
