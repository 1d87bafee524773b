// LDS.128 wavefront counter vs timed cost per address pattern (sm_100a): microbench2 reduced to the 128-bit loads.  One consumer IADD per load keeps the ALU floor low.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int ITERS = 2048;

template <int WIDTH>   // bytes per lane: 4, 8, 16
__global__ void k_lds(unsigned* out, int mode) {
    __shared__ __align__(16) unsigned sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    int lane = threadIdx.x & 31;
    int q;   // index in units of WIDTH bytes
    switch (mode) {
        case 0: q = 0; break;                       // all lanes same
        case 1: q = lane; break;                    // 32 distinct consecutive
        case 2: q = lane >> 4; break;               // 2 distinct
        case 3: q = lane >> 2; break;               // 8 distinct, each shared by 4 adjacent lanes
        case 4: q = lane & 7; break;                // 8 distinct, identical in every quarter-warp
        case 5: q = lane >> 3; break;               // 4 distinct, one per quarter-warp
        case 6: q = (lane & 1) + 16 - (lane >> 1); break;   // 17 distinct overlapping window pattern
        case 7: q = lane * 129; break;              // skewed rows (row stride 129 units)
        case 8: q = lane >> 1; break;               // 16 distinct, pairs of lanes share
        default: q = lane; break;
    }
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + q * WIDTH;
    unsigned acc = 0;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            unsigned a, b, c, d;
            unsigned addr = base + ((k * 512 + it * 16) & 8191);
            if (WIDTH == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
            else if (WIDTH == 8) { asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr)); }
            else { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(addr)); }
            acc += a;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); for (int i = 0; i < 5; i++) f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / 5;
}
int main() {
    // one launch per address pattern: run under `ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum`
    // to compare the hardware wavefront counter with the timed cycles per warp-load of microbench2
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
    unsigned* out; CK(cudaMalloc(&out, 4 * sms * 8 * 256));
    int blocks = sms * 8, tpb = 256; double warps = (double)blocks * tpb / 32;
    for (int mode = 0; mode <= 8; mode++) {
        float m16 = timeit([&] { k_lds<16><<<blocks, tpb>>>(out, mode); });
        printf("mode %d: LDS.128 %.2f cycles per warp-load per SM (%.0f warp-loads per launch)\n", mode,
               (m16 * 1e-3) * 1.965e9 * sms / (warps * ITERS * 16), warps * ITERS * 16);
    }
    return 0;
}
