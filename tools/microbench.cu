// Microbenchmarks that decide the shape of the ASW aggregation kernel on sm_100a.
//  (1) scalar FFMA / FMUL+FFMA+FADD issue rate  (2) packed fma.rn.f32x2 rate
//  (3) LDS.128 broadcast/wavefront behaviour    (4) MUFU ex2/sqrt rate
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0; 
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// our element pattern: ww = w1*w2 ; num += ww*e ; den += ww   (3 FP32-pipe instr / element)
__global__ void k_mix(float* out, float a, float b) {
    float num[16], den[16], w2[16], e[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { num[i] = 0; den[i] = 0; w2[i] = threadIdx.x * 0.001f + i; e[i] = i + b; }
    float w1 = a;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            float ww = w1 * w2[i];
            num[i] = fmaf(ww, e[i], num[i]);
            den[i] += ww;
        }
        w1 += b;   // keep the multiply from being hoisted
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += num[i] / den[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = pk(threadIdx.x * 0.001f + i, i);
    unsigned long long A = pk(a, a), B = pk(b, b);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma2(acc[i], A, B);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; upk(acc[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mix2(float* out, float a, float b) {
    unsigned long long num[8], den[8], w2[8], e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { num[i] = pk(0, 0); den[i] = pk(0, 0); w2[i] = pk(threadIdx.x * 0.001f + i, i * 0.5f); e[i] = pk(i + b, i); }
    float w1 = a;
    for (int it = 0; it < ITERS; it++) {
        unsigned long long W1 = pk(w1, w1);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned long long ww = mul2(W1, w2[i]);
            num[i] = fma2(ww, e[i], num[i]);
            den[i] = add2(den[i], ww);
        }
        w1 += b;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float a0, a1, b0, b1; upk(num[i], a0, a1); upk(den[i], b0, b1); s += a0 / b0 + a1 / b1; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MUFU: ex2 + sqrt per value
__global__ void k_mufu(float* out, float a) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 0.001f + i + a;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float s; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(acc[i]));
            float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-s));
            acc[i] = e + 1.0f;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS.128 patterns: mode 0 = all lanes same quad (broadcast), 1 = 32 distinct consecutive quads,
// 2 = 2 distinct quads, 3 = 17 distinct quads at stride 2 (overlapping-window pattern), 4 = stride-2 quads x 32 lanes
__global__ void k_lds(float* out, int mode) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    int lane = threadIdx.x & 31;
    int idx;
    switch (mode) {
        case 0: idx = 0; break;
        case 1: idx = lane; break;
        case 2: idx = lane >> 4; break;
        case 3: idx = 2 * ((lane & 1) + 16 - (lane >> 1)); break;
        case 4: idx = 2 * lane; break;
        case 5: idx = (lane & 1) + 16 - (lane >> 1); break;       // 17 distinct consecutive quads
        default: idx = lane; break;
    }
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float4 v = sm[(idx + k * 64 + it) & 1023];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; i++) f();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / 5;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024));
    for (int tpb : {256, 512, 1024}) {
        int blocks = sms * (2048 / tpb);
        double threads = (double)blocks * tpb;
        float ms;
        ms = timeit([&] { k_ffma<<<blocks, tpb>>>(out, 1.0001f, 0.5f); });
        printf("tpb %4d  FFMA   : %.3f ms  %.2f TFLOP/s (2 flop/FFMA)  %.1f lane-instr/clk/SM @1.965GHz\n", tpb, ms, threads * ITERS * 16 * 2 / ms / 1e9,
               threads * ITERS * 16 / (ms * 1e-3) / sms / 1.965e9);
        ms = timeit([&] { k_mix<<<blocks, tpb>>>(out, 1.0001f, 0.5f); });
        printf("tpb %4d  MIX    : %.3f ms  %.2f Gelem/s  %.2f TFLOP/s(4/elem)  %.1f lane-instr/clk/SM\n", tpb, ms, threads * ITERS * 16 / ms / 1e6, threads * ITERS * 16 * 4 / ms / 1e9,
               threads * ITERS * 16 * 3 / (ms * 1e-3) / sms / 1.965e9);
        ms = timeit([&] { k_ffma2<<<blocks, tpb>>>(out, 1.0001f, 0.5f); });
        printf("tpb %4d  FFMA2  : %.3f ms  %.2f TFLOP/s\n", tpb, ms, threads * ITERS * 16 * 2 / ms / 1e9);
        ms = timeit([&] { k_mix2<<<blocks, tpb>>>(out, 1.0001f, 0.5f); });
        printf("tpb %4d  MIX2   : %.3f ms  %.2f Gelem/s  %.2f TFLOP/s(4/elem)\n", tpb, ms, threads * ITERS * 16 / ms / 1e6, threads * ITERS * 16 * 4 / ms / 1e9);
        ms = timeit([&] { k_mufu<<<blocks, tpb>>>(out, 0.5f); });
        printf("tpb %4d  MUFU   : %.3f ms  %.2f G(sqrt+ex2)/s  %.2f mufu/clk/SM\n", tpb, ms, threads * ITERS * 8 / ms / 1e6, threads * ITERS * 8 * 2 / (ms * 1e-3) / sms / 1.965e9);
    }
    for (int mode = 0; mode <= 5; mode++) {
        int tpb = 256, blocks = sms * 8;
        float ms = timeit([&] { k_lds<<<blocks, tpb>>>(out, mode); });
        double warps = (double)blocks * tpb / 32;
        printf("LDS.128 mode %d: %.3f ms  %.2f cyc/warp-LDS/SM @1.965GHz\n", mode, ms, (ms * 1e-3) * 1.965e9 * sms / (warps * ITERS * 8));
    }
    return 0;
}
