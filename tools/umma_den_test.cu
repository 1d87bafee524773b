// Standalone check of the tensor-core denominator: D[r][x] = sum_rows sum_j W2[row][j][r] * W1[row][j][x]
// with tcgen05.mma kind::tf32, 3xTF32 split (hi*hi + hi*lo + lo*hi), A = W2 in TMEM (written with tcgen05.st),
// B = W1 in shared memory (MN-major, no swizzle, N-quad stride SBO), accumulator in TMEM, read back with tcgen05.ld.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_den_test umma_den_test.cu
// Run:   umma_den_test [SBO bytes, default 144]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

constexpr int NX = 96;      // N: left-weight columns (x)
constexpr int NRr = 256;    // M total: right-weight columns (r), two halves of 128
constexpr int KP = 40;      // K padded (window columns j, 35 used)
constexpr int KG = KP / 8;  // K groups of 8 (one tcgen05.mma kind::tf32 each)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

// idesc: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), A K-major, B MN-major (1<<16), N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int b_mn_major = 1) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1
__device__ __forceinline__ uint64_t make_sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(160) k_test(const float *__restrict__ W1, const float *__restrict__ W2, float *__restrict__ Dout,
                                              int rows, int sbo, int dbgcol, int variant) {
    extern __shared__ __align__(128) unsigned char smem[];
    // layout: [tmem base 16 B][barrier 16 B] | B_hi | B_lo   (each KG * (NX/4) * sbo bytes)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem);
    const uint32_t bar = smem_u32(smem + 16);
    const int barr = (NX / 4) * sbo;                  // bytes per K group
    unsigned char *Bhi = smem + 128, *Blo = Bhi + KG * barr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero the B arrays once (K padding rows must be 0)
    for (int i = tid; i < 2 * KG * barr / 4; i += blockDim.x) reinterpret_cast<float *>(Bhi)[i] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = *tmem_slot;
    // TMEM columns: D half 0: [0,96)  D half 1: [96,192)  A: 192 + half*80 + (0:hi | 40:lo) + j
    const uint32_t colD[2] = {0u, 96u};
    auto colA = [&](int half, int lo) { return 192u + half * 80u + lo * 40u; };

    if (warp < 4) {
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const uint32_t sv = __float_as_uint(7.0f);
        for (int c = 0; c < 192; c += 4)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tbase + lane_base + c), "r"(sv), "r"(sv), "r"(sv), "r"(sv) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    uint32_t phase = 0;
    for (int row = 0; row < rows; ++row) {
        if (warp < 4) {
            // ---- "producers": warp q owns TMEM lanes [32q, 32q+32) = r rows 32q+lane of both halves ----
            for (int half = 0; half < 2; ++half) {
                const int r = half * 128 + warp * 32 + lane;
                for (int j0 = 0; j0 < KP; j0 += 4) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float w = W2[((size_t)row * KP + j0 + u) * NRr + r];
                        const float h = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
                        hi[u] = __float_as_uint(w);                // the MMA reads the top 19 bits
                        lo[u] = __float_as_uint(w - h);
                    }
                    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tbase + lane_base + colA(half, 0) + j0),
                                 "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tbase + lane_base + colA(half, 1) + j0),
                                 "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            // ---- B = W1 (x columns), MN-major no-swizzle: element (x, j) at (j/8)*barr + (x/4)*sbo + (j%8)*16 + (x%4)*4 ----
            for (int i = tid; i < KP * NX; i += 128) {
                const int j = i / NX, x = i % NX;
                const float w = W1[((size_t)row * KP + j) * NX + x];
                const float h = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
                // variant 2: K-major no-swizzle: (x%8)*16 + (x/8)*128 + (j%4)*4 + ((j%8)/4)*(NX/8)*128 within the K group
                const int off = variant == 2 ? (j / 8) * barr + ((j % 8) / 4) * (NX / 8) * sbo + (x / 8) * sbo + (x % 8) * 16 + (j % 4) * 4
                                             : (j / 8) * barr + (x / 4) * sbo + (j % 8) * 16 + (x % 4) * 4;
                *reinterpret_cast<float *>(Bhi + off) = w;
                *reinterpret_cast<float *>(Blo + off) = w - h;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async-proxy (UMMA) reads
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (warp == 4 && lane == 0) {
            const uint32_t idesc = make_idesc(128, NX, variant == 2 ? 0 : 1);
            for (int half = 0; half < 2; ++half) {
                for (int term = 0; term < 3; ++term) {          // hi*hi, hi*lo, lo*hi
                    const int alo = term == 2, blo = term == 1;
                    for (int kg = 0; kg < KG; ++kg) {
                        const uint32_t a = tbase + colA(half, alo) + kg * 8;
                        const uint32_t baddr = smem_u32((blo ? Blo : Bhi) + kg * barr);
                        const uint64_t b = variant == 2 ? make_sdesc(baddr, (NX / 8) * sbo, sbo)          // K-major: LBO between k chunks, SBO between 8-row groups
                                         : variant == 1 ? make_sdesc(baddr, (uint32_t)sbo, (uint32_t)barr)  // swapped roles
                                                        : make_sdesc(baddr, (uint32_t)barr, (uint32_t)sbo);
                        const uint32_t d = tbase + colD[half];
                        const uint32_t acc = !(row == 0 && term == 0 && kg == 0);
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc)
                                     : "memory");
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        mbar_wait(bar, phase);                                   // everyone: MMAs of this row are done (buffers reusable)
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    if (tid == 0) Dout[(size_t)NRr * NX] = __uint_as_float(tbase);
    // ---- read D back: warp q reads lanes [32q, 32q+32) of both halves (debug: dbgcol >= 0 reads that column range instead) ----
    if (warp < 4) {
        for (int half = 0; half < 2; ++half) {
            const int r = half * 128 + warp * 32 + lane;
            for (int c0 = 0; c0 < NX; c0 += 16) {
                if (dbgcol >= 0 && c0 >= 32) break;
                uint32_t v[16];
                const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                               "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(tbase + lane_base + (dbgcol >= 0 ? (uint32_t)dbgcol + half * 80u : colD[half]) + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int u = 0; u < 16; ++u) Dout[(size_t)r * NX + c0 + u] = __uint_as_float(v[u]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main(int argc, char **argv) {
    const int sbo = argc > 1 ? atoi(argv[1]) : 144;
    const int dbgcol = argc > 2 ? atoi(argv[2]) : -1;
    const int variant = argc > 3 ? atoi(argv[3]) : 0;
    const int rows = 5, J = 35;
    std::vector<float> W1((size_t)rows * KP * NX, 0.f), W2((size_t)rows * KP * NRr, 0.f);
    srand(1);
    for (int row = 0; row < rows; ++row)
        for (int j = 0; j < J; ++j) {
            for (int x = 0; x < NX; ++x) W1[((size_t)row * KP + j) * NX + x] = expf(-6.f * rand() / RAND_MAX);
            for (int r = 0; r < 224; ++r) W2[((size_t)row * KP + j) * NRr + r] = expf(-6.f * rand() / RAND_MAX);
        }
    float *d1, *d2, *dD;
    CK(cudaMalloc(&d1, W1.size() * 4)); CK(cudaMalloc(&d2, W2.size() * 4)); CK(cudaMalloc(&dD, (size_t)NRr * NX * 4 + 64));
    CK(cudaMemcpy(d1, W1.data(), W1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d2, W2.data(), W2.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, (size_t)NRr * NX * 4));
    const int smem = 128 + 2 * KG * (NX / 4) * sbo;
    CK(cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_test<<<1, 160, smem>>>(d1, d2, dD, rows, sbo, dbgcol, variant);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)NRr * NX + 16);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    { uint32_t tb; memcpy(&tb, &D[(size_t)NRr * NX], 4); printf("tmem base 0x%08x\n", tb); }
    if (dbgcol >= 0) {   // round trip: A hi of the LAST row, half 0/1, columns j = 0..31
        int badrt = 0;
        for (int r = 0; r < 224; ++r)
            for (int j = 0; j < 32; ++j) {
                const float want = W2[((size_t)(rows - 1) * KP + j) * NRr + r], got = D[(size_t)r * NX + j];
                if (want != got && badrt++ < 5) printf("  roundtrip r=%d j=%d got %g want %g\n", r, j, got, want);
            }
        printf("TMEM st->ld round trip of A_hi: %d mismatches\n", badrt);
        return badrt ? 1 : 0;
    }
    double maxrel = 0, maxrel_tf32 = 0;
    int bad = 0;
    for (int r = 0; r < 224; ++r)
        for (int x = 0; x < NX; ++x) {
            double ref = 0;
            for (int row = 0; row < rows; ++row)
                for (int j = 0; j < J; ++j) ref += (double)W2[((size_t)row * KP + j) * NRr + r] * (double)W1[((size_t)row * KP + j) * NX + x];
            const double rel = fabs(D[(size_t)r * NX + x] - ref) / ref;
            if (rel > maxrel) maxrel = rel;
            if (rel > 1e-5 && bad++ < 8) printf("  r=%d x=%d got %.8g want %.8g rel %.3g\n", r, x, D[(size_t)r * NX + x], ref, rel);
        }
    printf("variant %d ", variant);
    printf("sbo=%d: max relative error of the 3xTF32 tensor-core denominator vs double: %.3g (%d of %d above 1e-5)\n", sbo, maxrel, bad, 224 * NX);
    (void)maxrel_tf32;
    return bad ? 1 : 0;
}
