// Shared-memory wavefronts of the EXACT load patterns of the k_aggregate_tc consumer loop (sm_100a), one launch per pattern:
//   run under ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum
// pattern 0: right weights  LDS.128 at  w2 + (88 - 32 wx + 32 wd + 4 dl) floats + s * 224 floats        (x-group rotation: same for the 4 quarter-warps)
// pattern 1: left weights   LDS.128 at  w1 + (4 wx + xl) * 144 bytes + a * 16                             (K-major operand, 144-byte group stride)
// pattern 2: raw costs      LDS.32  at  e + (32 wx + 8 xl) * 132 + 4 * ((8 wd + dl + 2 xl) % 32)           (byte tile, pitch 132)
// pattern 3: raw costs with pitch 140 instead of 132
// pattern 4: pattern 0 without the rotation (dg = 8 wd + dl): four overlapping spans
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int ITERS = 1024;
__global__ void k_pat(unsigned* out, int pat) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < 47000; i += blockDim.x) reinterpret_cast<unsigned*>(sm)[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) % 12;
    const int xl = lane >> 3, dl = lane & 7, wx = warp / 4, wd = warp % 4;
    const int dg = (8 * wd + dl + 2 * xl) % 32, xg = 4 * wx + xl;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (pat >= 5 ? 121504u : 0u);   // 5, 6: at the kernel's real offsets (W2 rows)
    if (pat == 6) base -= 121504u - 52384u;                                              // 6: W1 operand offset
    unsigned addr; int width = 16, step = 896;
    switch (pat) {
        case 0: addr = base + 4 * (96 - 8 - 8 * xg + 4 * dg); break;
        case 1: addr = base + xg * 144; step = 16; break;
        case 2: addr = base + 8 * xg * 132 + 4 * dg; width = 4; step = 132; break;
        case 5: addr = base + 4 * (96 - 8 - 8 * xg + 4 * dg); break;
        case 6: addr = base + xg * 144; step = 16; break;
        case 3: addr = base + 8 * xg * 140 + 4 * dg; width = 4; step = 140; break;
        default: addr = base + 4 * (96 - 8 - 8 * xg + 4 * (8 * wd + dl)); break;
    }
    unsigned acc = 0;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            unsigned a, b, c, d;
            const unsigned ad = addr + k * step + ((it & 1) << 15);   // two alternating regions: keeps the loads inside the loop
            if (width == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(ad));
            else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(ad));
            acc += a;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
    unsigned* out; CK(cudaMalloc(&out, 4 * sms * 384));
    CK(cudaFuncSetAttribute(k_pat, cudaFuncAttributeMaxDynamicSharedMemorySize, 190000));
    for (int pat = 0; pat <= 6; pat++) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        k_pat<<<sms, 384, 190000>>>(out, pat); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); k_pat<<<sms, 384, 190000>>>(out, pat); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("pattern %d: %.2f cycles per warp-load per SM\n", pat, ms * 1e-3 * 1.965e9 / (12.0 * ITERS * 16));
    }
    return 0;
}
