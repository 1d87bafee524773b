// ss_passive.cu -- libsspassive.so: ASW / GSW stereo matching for NVIDIA B200 (sm_100a).
//
// Replaces the CPU hot path of decadenza/SimpleStereo (simplestereo/_passive.cpp, citations are
// file:line in that checkout) behind the C ABI declared in include/ss_passive.h.
//
// Pipeline of one call (all on one stream, nothing synchronised inside):
//   k_prep_features   BGR u8 -> padded float4 "feature" rows (CIELab for ASW, BGR-as-float for GSW);
//                     out-of-image pixels carry a 1e18 sentinel so their support weight is exp(-huge)=0,
//                     which reproduces the reference's window clipping (_passive.cpp:39-45, :67-68)
//                     without a single bounds test in the hot loop.
//   k_cost_volume     raw per-pixel matching cost E[row][u][d] (truncated AD / capped colour distance,
//                     _passive.cpp:77-79, :528-531), chunk-major so an aggregation tile is one
//                     contiguous span that a single TMA bulk copy brings into shared memory.
//   k_aggregate_tc /  the hot kernels (98 % of the time): per (row, 96-column tile, 128-disparity chunk) producer warps
//   k_aggregate_ws    stream the window rows through shared memory (cp.async.bulk + mbarrier) and tabulate both
//                     support-weight rows once per (pixel, offset) -- the reference re-evaluates exp/sqrt/pow for every
//                     (x,d) pair, _passive.cpp:71-74 -- while consumer warps accumulate in registers with packed
//                     fma.rn.f32x2 (FFMA2/FMUL2/FADD2: two lanes' worth of work per issued instruction).  In
//                     k_aggregate_tc (ss_aggregate_tc.cuh; ASW, 128-disparity chunks) the denominators run on the
//                     tensor cores instead: tcgen05.mma kind::tf32, 3xTF32 split, right weights in tensor memory,
//                     accumulator in TMEM.  k_aggregate_ws is the all-CUDA-core form (GSW, short disparity ranges);
//                     k_aggregate the older single-role GSW fallback for windows whose float cost tiles do not fit.
//                     WTA over the disparity chunk is fused (warp shuffle + 64-bit atomicMin keys).
//   k_wta_right       right-reference WTA: minimum over diagonals of the SAME aggregated volume
//                     (C_R[xr,d] == C_L[xr+d,d]), so the "roughly doubled" second
//                     pass of the reference (passive.py:39) costs one read of the volume.
//   k_finalize        key decode, L-R invalidation (_passive.cpp:251-252), occlusion fill (:258-285).
//
// No CPU fallback exists: every entry point fails with SS_ERR_CUDA when no device is usable.

#include "../../include/ss_passive.h"
#include "../../include/ss_post.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------

typedef unsigned long long u64;

#ifndef SS_UNROLL
#define SS_UNROLL 2            // periods (8 window columns) per trip of the consumer loop: 1, 2 or 4
#endif
#ifndef SS_WAIT_HINT
#define SS_WAIT_HINT 1000000   // mbarrier.try_wait suspend-time hint (ns); 0 = plain polling
#endif
constexpr int TILE_X = 64;          // output columns per block of k_aggregate (GSW)
constexpr int TILE_WS = 96;         // output columns per block of k_aggregate_ws (ASW)
constexpr float SENTINEL = 1.0e18f; // feature value of out-of-image pixels
constexpr u64 KEY_NONE = ~0ull;

__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if SS_WAIT_HINT
    // the suspend-time hint keeps a waiting warp parked in hardware instead of re-issuing the poll: polling
    // loops were 18 % of all issued instructions (ncu) and compete with the warps that are being waited for
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"((uint32_t)SS_WAIT_HINT)
        : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
#endif
}
// waits that are expected to be long (producers waiting for a free buffer) -- same primitive: with the suspend-time hint a
// waiting warp is parked in hardware, a software back-off (nanosleep) measured no better
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ u64 make_key(float cost, int disp) {
    return ((u64)__float_as_uint(cost) << 32) | (uint32_t)disp;
}

// ------------------------------------------------------------------------------------------
// geometry shared by host and device
// ------------------------------------------------------------------------------------------

struct Geom {
    int W, H;          // image size
    int win, pad;      // window side, win/2
    int minD, maxD;    // inclusive disparity range of the call (_passive.cpp:56)
    int dLo, dHi;      // inclusive sub-range evaluated by this launch (disparity-range sharding)
    int DC, nch;       // disparity chunk, number of chunks covering [dLo, dHi]
    int row0, row1;    // output rows [row0,row1)
    int erow0, erow1;  // input rows needed [erow0,erow1) = output rows +- pad, clipped
    int T;             // output columns per block: 64 (k_aggregate, GSW) or 96 (k_aggregate_ws, ASW)
    int EP;            // ASW: bytes per cost-volume column (DC + 4: the +4 skews columns 8 apart onto different banks)
    int ntx;           // number of T-column tiles
    int UW;            // padded row pitch of the left feature image and of the cost volume (u' = u + pad)
    int VW, PL2;       // padded row pitch / left padding of the right feature image (xr' = xr + PL2)
    int NU;            // cost-volume columns per tile  = T + win - 1
    int NR, NRp;       // right centres per tile = T + DC - 1, padded pitch T + DC
    int NV;            // right feature columns per tile = NR + win - 1
};

// ------------------------------------------------------------------------------------------
// k_prep_features
// ------------------------------------------------------------------------------------------

__constant__ float c_lin100[256];   // sRGB byte -> linear*100 as float, built on the host with the same
                                    // powf as colorconversion.hpp:19-37 (exact)

// f(t) of colorconversion.hpp:55-65: powf(float(t), 0.33333334f) evaluated in double and rounded to float
__device__ __forceinline__ double lab_f(double t) {
    if (t > 0.008856) return (double)(float)pow((double)(float)t, (double)0.33333334f);
    return __dadd_rn(__dmul_rn(7.787, t), 16.0 / 116.0);
}

__device__ __forceinline__ float4 bgr_to_lab(uint8_t B, uint8_t G, uint8_t R) {
    const double r = c_lin100[R], g = c_lin100[G], b = c_lin100[B];
    // colorconversion.hpp:40-42, double arithmetic without contraction
    const double X = __dadd_rn(__dadd_rn(__dmul_rn(r, 0.4124), __dmul_rn(g, 0.3576)), __dmul_rn(b, 0.1805));
    const double Y = __dadd_rn(__dadd_rn(__dmul_rn(r, 0.2126), __dmul_rn(g, 0.7152)), __dmul_rn(b, 0.0722));
    const double Z = __dadd_rn(__dadd_rn(__dmul_rn(r, 0.0193), __dmul_rn(g, 0.1192)), __dmul_rn(b, 0.9505));
    const double fx = lab_f(__ddiv_rn(X, (double)95.047f));
    const double fy = lab_f(__ddiv_rn(Y, (double)100.0f));
    const double fz = lab_f(__ddiv_rn(Z, (double)108.883f));
    float4 o;
    o.x = (float)__dsub_rn(__dmul_rn(116.0, fy), 16.0);     // :67-69
    o.y = (float)__dmul_rn(500.0, __dsub_rn(fx, fy));
    o.z = (float)__dmul_rn(200.0, __dsub_rn(fy, fz));
    o.w = 0.f;
    return o;
}

// One thread per padded pixel.  pitch = padded width, lpad = columns of padding on the left.
template <bool GSW>
__global__ void k_prep_features(const uint8_t *__restrict__ img, float4 *__restrict__ feat, int W, int erow0,
                                int nrows, int pitch, int lpad) {
    const int xp = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (xp >= pitch || r >= nrows) return;
    const int x = xp - lpad, y = erow0 + r;
    float4 o = make_float4(SENTINEL, 0.f, 0.f, 0.f);
    if (x >= 0 && x < W) {
        const uint8_t *p = img + 3 * ((size_t)y * W + x);
        const uint8_t B = p[0], G = p[1], R = p[2];
        if (GSW) o = make_float4((float)B, (float)G, (float)R, 0.f);
        else o = bgr_to_lab(B, G, R);
    }
    feat[(size_t)r * pitch + xp] = o;
}

// ------------------------------------------------------------------------------------------
// k_cost_volume: E[ch][r][u'][k], k = 0..DC-1, disparity d = dLo + ch*DC + k
// ------------------------------------------------------------------------------------------

template <bool GSW>
__global__ void k_cost_volume(const uint8_t *__restrict__ img1, const uint8_t *__restrict__ img2,
                              void *__restrict__ Eout, Geom g, float f_max) {
    const int k4 = g.DC / 4;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;   // over UW * DC/4
    if (t >= g.UW * k4) return;
    const int up = t / k4, kq = (t % k4) * 4;
    const int r = blockIdx.y, ch = blockIdx.z;
    const int y = g.erow0 + r, u = up - g.pad;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    int vi[4] = {0, 0, 0, 0};
    if (u >= 0 && u < g.W) {
        const uint8_t *p = img1 + 3 * ((size_t)y * g.W + u);
        const int b1 = p[0], g1 = p[1], r1 = p[2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int d = g.dLo + ch * g.DC + kq + q;
            const int xr = u - d;
            if (d <= g.dHi && xr >= 0 && xr < g.W) {
                const uint8_t *s = img2 + 3 * ((size_t)y * g.W + xr);
                const int db = b1 - s[0], dg = g1 - s[1], dr = r1 - s[2];
                if (GSW) {
                    // min(fMax, (float)sqrt(int)) : _passive.cpp:528-531; IEEE sqrtf of an exact integer
                    const float e = __fsqrt_rn((float)(db * db + dg * dg + dr * dr));
                    v[q] = fminf(f_max, e);
                } else {
                    vi[q] = min(40, abs(db) + abs(dg) + abs(dr));            // _passive.cpp:77-79
                }
            }
        }
    }
    const size_t col = ((size_t)ch * (g.erow1 - g.erow0) + r) * g.UW + up;
    if (GSW) {
        *reinterpret_cast<float4 *>(static_cast<float *>(Eout) + col * g.DC + kq) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        // ASW: the truncated AD is an integer in [0,40] -> one byte; 4 disparities per 32-bit word
        *reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(Eout) + col * g.EP + kq) =
            (uint32_t)vi[0] | ((uint32_t)vi[1] << 8) | ((uint32_t)vi[2] << 16) | ((uint32_t)vi[3] << 24);
    }
}

// ------------------------------------------------------------------------------------------
// k_aggregate -- the hot kernel
// ------------------------------------------------------------------------------------------

struct AggParams {
    Geom g;
    const float4 *F1;     // left features  [erows][UW]
    const float4 *F2;     // right features [erows][VW]
    const void *E;        // raw cost volume: GSW float [nch][erows][UW][DC]; ASW uint8 [nch][erows][UW][EP]
    const float *proxarg; // ASW: -log2(e)*r/gammaP per window offset [win][(win+3)&~3]
    float kC;             // ASW: -log2(e)/gammaC ; GSW (k_aggregate): (float)gamma
    float kC2;            // GSW (k_aggregate_ws): -log2(e)/gamma
    int iterations;       // GSW only (<=0: centre weight only)
    u64 *bestL;           // [(row1-row0)*W] packed (cost,disp) keys, atomicMin
    float *vol0;          // optional: ASW cost / GSW right cost  [(rows)*W*Dp]
    float *vol1;          // optional: GSW left cost
    int Dp;               // pitch of vol0/vol1 (= nch*DC)
    int vol_export;       // the volumes are returned to the caller (debug export): unevaluated pairs must read +inf
    int freerun;          // timing experiment (SS_FREERUN=1): consumers ignore the barriers, producers idle; results are garbage
#ifdef SS_DEBUG_DUMP
    float *dbg;           // [0]=bx [1]=by [2]=step ; dump of W1s, W2s, Es of that block/step follows at dbg+16
#endif
};

template <bool GSW>
__device__ __forceinline__ float support_weight(const float4 c, const float4 n, float kC, float parg) {
    const float d0 = n.x - c.x, d1 = n.y - c.y, d2 = n.z - c.z;
    const float s = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
    if (GSW) {
        // exp(-(float)sqrt(n)/gamma), float throughout (_passive.cpp:494-496 after the closed form)
        return expf(-__fdiv_rn(__fsqrt_rn(s), kC));
    } else {
        // prox * exp(-dist/gammaC) == 2^(dist*kC + parg)   (_passive.cpp:47-50)
        return ex2_approx(fmaf(sqrt_approx(s), kC, parg));
    }
}

template <int V> struct IC { static constexpr int value = V; };

// A warp covers 4 x-groups (of 8 columns) x 8 disparity groups (of 4): shared-memory loads are deduplicated
// per warp instruction, so this shape minimises distinct bytes per step (W1 128 B, W2 3 x 224 B, E 512 B =
// 12 wavefronts, against 18 for a 1 x 32 arrangement; measured with tools/microbench2.cu).
template <int DC> struct AggCfg {
    static constexpr int ND = DC / 4;           // disparity groups (of 4) in the chunk
    static constexpr int NDB = ND / 8;          // blocks of 8 disparity groups
    static constexpr int NW = 2 * NDB;          // warps: 2 x-group blocks (of 4) x NDB
    static constexpr int NT = NW * 32;
    static constexpr int NRp = TILE_X + DC;
#ifdef SS_MINB1
    static constexpr int MINB = 1;
#else
    static constexpr int MINB = DC == 128 ? 2 : (DC == 64 ? 4 : 8);
#endif
};

// dynamic shared memory carve-up (bytes), mirrored on the host
struct SmemPlan {
    int es, f1, f2, pa, c1, c2, w1, w2, bars, total;
};
__host__ __device__ inline SmemPlan smem_plan(int win, int DC) {
    const int NU = TILE_X + win - 1, NR = TILE_X + DC - 1, NRp = TILE_X + DC, NV = NR + win - 1;
    SmemPlan p;
    int off = 0;
    p.es = off;  off += NU * DC * 4;            off = (off + 127) & ~127;
    p.f1 = off;  off += 2 * NU * 16;
    p.f2 = off;  off += 2 * NV * 16;
    p.pa = off;  off += 2 * ((win + 3) & ~3) * 4;
    p.c1 = off;  off += TILE_X * 16;
    p.c2 = off;  off += NRp * 16;
    p.w1 = off;  off += win * TILE_X * 4;       off = (off + 15) & ~15;
    p.w2 = off;  off += win * NRp * 4;          off = (off + 15) & ~15;
    p.bars = off; off += 4 * 8;
    p.total = off;
    return p;
}

// REM = win % 8: the window columns are walked in groups of 8 (one turn of the cost ring); the tail group
// is straight-line code so that no ring slot becomes a run-time phi.
template <bool GSW, int DC, int REM>
__global__ void __launch_bounds__(AggCfg<DC>::NT, AggCfg<DC>::MINB) k_aggregate(const AggParams P) {
    typedef AggCfg<DC> C;
    constexpr int T = TILE_X, NRp = C::NRp;
    extern __shared__ __align__(128) unsigned char smem[];

    const Geom &g = P.g;
    const int win = g.win, pad = g.pad;
    const int NU = g.NU, NR = g.NR, NV = g.NV;
    const SmemPlan sp = smem_plan(win, DC);
    float *Es = reinterpret_cast<float *>(smem + sp.es);
    float4 *F1s = reinterpret_cast<float4 *>(smem + sp.f1);
    float4 *F2s = reinterpret_cast<float4 *>(smem + sp.f2);
    float4 *C1s = reinterpret_cast<float4 *>(smem + sp.c1);
    float4 *C2s = reinterpret_cast<float4 *>(smem + sp.c2);
    float *W1s = reinterpret_cast<float *>(smem + sp.w1);
    float *W2s = reinterpret_cast<float *>(smem + sp.w2);
    // proximity-exponent rows are staged in 16-byte units like every other TMA destination
    // (ptxas 12.9 folded the stage offset of a 4-byte-unit destination into the 16-byte one: keep them uniform)
    float4 *PAs = reinterpret_cast<float4 *>(smem + sp.pa);
    const int winq = (win + 3) >> 2;                       // float4 per proximity row
    const int winp = winq * 4;                             // pitch of the proximity-exponent table in floats
    const uint32_t barC = smem_u32(smem + sp.bars), barF0 = barC + 8, barF1 = barC + 16, barE = barC + 24;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * T;
    const int y = g.row0 + blockIdx.y;
    const int ch = blockIdx.z;
    const int dlo = g.dLo + ch * DC;                       // first disparity of this chunk
    const int erows = g.erow1 - g.erow0;

    // window rows inside the image (_passive.cpp:59-62)
    const int i_lo = max(0, pad - y), i_hi = min(win - 1, g.H - 1 - y + pad);
    const int nsteps = i_hi - i_lo + 1;

    // per-tile source offsets
    const int f2_start = x0 - dlo - DC + 1 - pad + g.PL2;  // first right feature column (padded index)
    const int c2_start = x0 - dlo - DC + 1 + g.PL2;        // first right centre
    const size_t e_plane = (size_t)g.UW * DC;

    if (tid == 0) {
        mbar_init(barC, 1);
        mbar_init(barF0, 1);
        mbar_init(barF1, 1);
        mbar_init(barE, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_F = [&](int n) {       // feature rows (+ proximity exponents) of step n -> stage n&1
        const int i = i_lo + n, ii = y - pad + i;
        const int st = n & 1;
        const uint32_t bar = st ? barF1 : barF0;
        mbar_expect_tx(bar, (uint32_t)((NU + NV + (GSW ? 0 : winq)) * 16));
        tma_load_1d(smem_u32(F1s + st * NU), P.F1 + (size_t)(ii - g.erow0) * g.UW + x0, NU * 16, bar);
        tma_load_1d(smem_u32(F2s + st * NV), P.F2 + (size_t)(ii - g.erow0) * g.VW + f2_start, NV * 16, bar);
        if (!GSW) tma_load_1d(smem_u32(PAs + st * winq), P.proxarg + (size_t)i * winp, winq * 16, bar);
    };
    auto issue_E = [&](int n) {       // raw cost tile of step n
        const int ii = y - pad + i_lo + n;
        const uint32_t bytes = (uint32_t)(NU * DC * 4);
        mbar_expect_tx(barE, bytes);
        tma_load_1d(smem_u32(Es), static_cast<const float *>(P.E) + ((size_t)ch * erows + (ii - g.erow0)) * e_plane + (size_t)x0 * DC,
                    bytes, barE);
    };

    if (tid == 0) {
        mbar_expect_tx(barC, (uint32_t)((T + NR) * 16));
        tma_load_1d(smem_u32(C1s), P.F1 + (size_t)(y - g.erow0) * g.UW + x0 + pad, T * 16, barC);
        tma_load_1d(smem_u32(C2s), P.F2 + (size_t)(y - g.erow0) * g.VW + c2_start, NR * 16, barC);
        issue_F(0);
        issue_E(0);
    }

    // lane -> register tile: 8 consecutive x, 4 consecutive disparities
    const int dg = (warp % C::NDB) * 8 + (lane & 7);
    const int xb = 8 * ((warp / C::NDB) * 4 + (lane >> 3)); // tile-relative first column
    const int kb = 4 * dg;                                 // chunk-relative first disparity
    const int R0 = T - 8 - xb + kb;                        // first reversed right-centre index (multiple of 4)

    u64 acc0[8][2], acc1[8][2];                            // ASW: numerator, denominator; GSW: left, right cost
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) { acc0[a][b] = 0ull; acc1[a][b] = 0ull; }

    for (int n = 0; n < nsteps; ++n) {
        const int i = i_lo + n;
        if (tid == 0 && n + 1 < nsteps) issue_F(n + 1);
        if (n == 0) mbar_wait(barC, 0);
        mbar_wait((n & 1) ? barF1 : barF0, (n >> 1) & 1);

        // ---- phase A: tabulate the two support-weight rows of window row i -------------------
        // A warp owns a 32-column block (right-image blocks first, then left-image blocks) and walks all
        // window offsets j in batches of 4: loads first, stores last, so the four exp/sqrt chains overlap
        // (smem stores between them would otherwise serialise the loads of the next weight).
        {
            const float4 *f1 = F1s + (n & 1) * NU;
            const float4 *f2 = F2s + (n & 1) * NV;
            const float *parg = reinterpret_cast<const float *>(PAs + (n & 1) * winq);
            constexpr int NCBR = NRp / 32, NCB = NCBR + T / 32;
#pragma unroll 1
            for (int cb = warp; cb < NCB; cb += C::NW) {
                const bool right = cb < NCBR;                 // warp-uniform
                const int col = (right ? cb : cb - NCBR) * 32 + lane;
                // right: W2s[j][r], r reversed (xr = xr_max - r): centre NR-1-r, neighbour NR-1-r+j
                // left : W1s[j][x], centre (y, x0+x), neighbour (ii, x0+x-pad+j)
                const bool live = !right || col < NR;
                const int src = right ? (live ? NR - 1 - col : 0) : col;
                const float4 c = right ? C2s[src] : C1s[src];
                const float4 *nb = (right ? f2 : f1) + src;
                float *dst = (right ? W2s : W1s) + col;
                const int pitch = right ? NRp : T;
                // right-border abort of the LEFT pass of GSW only (_passive.cpp:445-446, :470-471)
                const bool quirk = GSW && !right && (x0 + col + pad >= g.W);
#pragma unroll 1
                for (int j0 = 0; j0 < win; j0 += 4) {
                    float pa[4] = {0.f, 0.f, 0.f, 0.f};
                    if (!GSW) {
                        const float4 t = *reinterpret_cast<const float4 *>(parg + j0);
                        pa[0] = t.x; pa[1] = t.y; pa[2] = t.z; pa[3] = t.w;
                    }
                    float w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = min(j0 + u, win - 1);
                        w[u] = support_weight<GSW>(c, nb[j], P.kC, pa[u]);
                        if (GSW) {
                            const bool centre = (i == pad) && (j == pad);
                            if (P.iterations <= 0) w[u] = centre ? 1.f : 0.f;
                            else if (quirk) {
                                const bool keep = (y == 0) ? (i == pad) : centre;
                                if (!keep) w[u] = 0.f;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (j0 + u < win) dst[(j0 + u) * pitch] = live ? w[u] : 0.f;
                }
            }
        }
        __syncthreads();
        mbar_wait(barE, n & 1);
#ifdef SS_DEBUG_DUMP
        if (P.dbg && blockIdx.x == (unsigned)P.dbg[0] && blockIdx.y == (unsigned)P.dbg[1] && blockIdx.z == 0 && n == (int)P.dbg[2]) {
            float *o = P.dbg + 16;
            for (int k = tid; k < win * T; k += C::NT) o[k] = W1s[k];
            o += win * T;
            for (int k = tid; k < win * NRp; k += C::NT) o[k] = W2s[k];
            o += win * NRp;
            for (int k = tid; k < NU * DC; k += C::NT) o[k] = Es[k];
        }
#endif

        // ---- phase B: accumulate window row i into the register tile ---------------------------
        {
            u64 ring[8][2];                                 // sliding window of 8 cost columns x 4 disparities
            const float *ep = Es + (size_t)xb * DC + kb;
#pragma unroll
            for (int a = 0; a < 7; ++a) {
                const float4 e = *reinterpret_cast<const float4 *>(ep + a * DC);
                ring[a][0] = pk(e.x, e.y);
                ring[a][1] = pk(e.z, e.w);
            }
            ep += 7 * DC;                                   // column consumed first by a = 7
            const float *w1p = W1s + xb;
            const float *w2p = W2s + R0;

            auto step = [&](auto sc) {
                constexpr int s = decltype(sc)::value;
                {
                    const float4 e = *reinterpret_cast<const float4 *>(ep + s * DC);
                    ring[(7 + s) & 7][0] = pk(e.x, e.y);
                    ring[(7 + s) & 7][1] = pk(e.z, e.w);
                }
                const float4 wa = *reinterpret_cast<const float4 *>(w1p + s * T);
                const float4 wb = *reinterpret_cast<const float4 *>(w1p + s * T + 4);
                const float4 v0 = *reinterpret_cast<const float4 *>(w2p + s * NRp);
                const float4 v1 = *reinterpret_cast<const float4 *>(w2p + s * NRp + 4);
                const float4 v2 = *reinterpret_cast<const float4 *>(w2p + s * NRp + 8);
                const float w1[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                const float v[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const u64 w1d = pk(w1[a], w1[a]);
#pragma unroll
                    for (int bp = 0; bp < 2; ++bp) {
                        // disparities kb+2bp, kb+2bp+1  <->  reversed right index 7-a+2bp, 8-a+2bp
                        const u64 w2d = pk(v[7 - a + 2 * bp], v[8 - a + 2 * bp]);
                        const u64 e2 = ring[(a + s) & 7][bp];
                        if (GSW) {
                            acc0[a][bp] = fma2(w1d, e2, acc0[a][bp]);       // _passive.cpp:528
                            acc1[a][bp] = fma2(w2d, e2, acc1[a][bp]);       // :644
                        } else {
                            const u64 ww = mul2(w1d, w2d);                  // w1*w2
                            acc0[a][bp] = fma2(ww, e2, acc0[a][bp]);        // cost += w1*w2*e  (:77)
                            acc1[a][bp] = add2(acc1[a][bp], ww);            // tot  += w1*w2    (:82)
                        }
                    }
                }
            };
            int j = 0;
#pragma unroll 1
            for (; j + 8 <= win; j += 8) {
                step(IC<0>{}); step(IC<1>{}); step(IC<2>{}); step(IC<3>{});
                step(IC<4>{}); step(IC<5>{}); step(IC<6>{}); step(IC<7>{});
                ep += 8 * DC;
                w1p += 8 * T;
                w2p += 8 * NRp;
            }
            if (REM > 0) step(IC<0>{});
            if (REM > 1) step(IC<1>{});
            if (REM > 2) step(IC<2>{});
            if (REM > 3) step(IC<3>{});
            if (REM > 4) step(IC<4>{});
            if (REM > 5) step(IC<5>{});
            if (REM > 6) step(IC<6>{});
        }
        __syncthreads();
        if (tid == 0 && n + 1 < nsteps) issue_E(n + 1);
    }

    // ---- epilogue: normalise, WTA over the chunk, optional volume store --------------------------
    const int rowo = y - g.row0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int x = x0 + xb + a;
        float c0[4], c1[4];
        upk(acc0[a][0], c0[0], c0[1]);
        upk(acc0[a][1], c0[2], c0[3]);
        upk(acc1[a][0], c1[0], c1[1]);
        upk(acc1[a][1], c1[2], c1[3]);
        u64 best = KEY_NONE;
        float out0[4], out1[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int d = dlo + kb + b;
            const bool valid = (x < g.W) && (d <= g.dHi) && (x - d >= 0);
            float costL, costR;
            if (GSW) { costL = c0[b]; costR = c1[b]; }
            else { costL = __fdiv_rn(c0[b], c1[b]); costR = costL; }   // cost / tot (:88)
            out0[b] = valid ? costR : INFINITY;
            out1[b] = valid ? costL : INFINITY;
            if (valid) {
                const u64 k = make_key(costL, d);
                best = k < best ? k : best;
            }
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {       // the 8 disparity groups of this warp share lane bits 0-2
            const u64 o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o < best ? o : best;
        }
        if ((lane & 7) == 0 && x < g.W && best != KEY_NONE) atomicMin(P.bestL + (size_t)rowo * g.W + x, best);
        if (x < g.W) {
            const size_t o = ((size_t)rowo * g.W + x) * P.Dp + (size_t)ch * DC + kb;
            if (P.vol0) *reinterpret_cast<float4 *>(P.vol0 + o) = make_float4(out0[0], out0[1], out0[2], out0[3]);
            if (P.vol1) *reinterpret_cast<float4 *>(P.vol1 + o) = make_float4(out1[0], out1[1], out1[2], out1[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_aggregate_ws -- warp-specialised ASW aggregation (the headline kernel)
//
// One block = (output row y, 96 columns, one chunk of DC disparities).  Consumer warps (3 column blocks x DC/32
// disparity blocks, each warp 4 x-groups x 8 disparity groups, lane tile 8 columns x 4 disparities) only run
// the packed-FP32 accumulation; producer warps tabulate the support-weight rows of the NEXT window row into the
// other half of a double buffer and drive the TMA copies.  mbarrier full/empty pairs order the three streams
// (features -> producers, weights -> consumers, raw costs -> consumers), so the FP32 pipe never waits for the
// exp/sqrt chains the way the single-role k_aggregate does.  Raw costs are bytes (0..40), 4 disparities per
// 32-bit shared load, converted with I2F.U8 on the otherwise idle XU pipe.
// At DC=128 the block is 16 warps, one block per SM: producers drop to 56 registers and consumers grow to 152
// (setmaxnreg), which removes the spills of the 128-register version.
// ------------------------------------------------------------------------------------------

template <bool GSW, int DC> struct WsCfg {
    static constexpr int T = GSW ? TILE_X : TILE_WS;   // GSW keeps float raw costs in shared memory: 64-column tiles
    static constexpr int XB = T / 32;            // blocks of 32 output columns
    static constexpr int NDB = DC / 32;          // blocks of 32 disparities (8 groups of 4)
    static constexpr int CW = XB * NDB;          // consumer warps
#ifndef SS_GSW_PW
#define SS_GSW_PW 2
#endif
    // producer warps: GSW has fewer consumer warps per block (2 per scheduler), so the weight rows are tabulated
    // by two producer warps per scheduler or the consumers wait for them (ncu: 45 % of consumer samples)
    static constexpr int PW = GSW ? SS_GSW_PW * NDB : NDB;
    static constexpr int NT = (CW + PW) * 32;    // ASW 512 / 256 / 128 threads, GSW 512 / 256 / 128
    static constexpr int MINB = GSW ? 1 : 4 / NDB;   // blocks per SM the register budget is sized for
    static constexpr int NRp = T + DC;
    static constexpr int EP = GSW ? DC * 4 : DC + 4;   // bytes per raw-cost column (ASW: bytes, skewed by 4)
    static constexpr bool SETREG = DC == 128 && NT == 512;   // 16 warps: producers give registers to the consumers
};

struct WsSmem {         // stage s of a double-buffered region lives at base + s * size
    int e, f1, f2, pa, c1, c2, w1, w2, bars, total;
    int ebytes, f1bytes, f2bytes, pabytes, w1bytes, w2bytes;
};
__host__ __device__ inline WsSmem ws_smem(int win, int DC, bool gsw) {
    const int T = gsw ? TILE_X : TILE_WS, NU = T + win - 1, NR = T + DC - 1, NRp = T + DC, NV = NR + win - 1;
    const int EP = gsw ? DC * 4 : DC + 4;
    const int winq = (win + 3) >> 2;
    WsSmem p;
    int off = 0;
    p.ebytes = (NU * EP + 15) & ~15;
    p.f1bytes = NU * 16;
    p.f2bytes = NV * 16;
    p.pabytes = winq * 16;
    // weight buffers hold (win + 3) & ~3 rows: the producers store batches of 4 window offsets without predicates
    // (rows past the window receive finite junk and are never read)
    const int winr = (win + 3) & ~3;
    p.w1bytes = (winr * T * 4 + 15) & ~15;
    p.w2bytes = (winr * NRp * 4 + 15) & ~15;
    p.e = off;  off += 2 * p.ebytes;
    p.f1 = off; off += 2 * p.f1bytes;
    p.f2 = off; off += 2 * p.f2bytes;
    p.pa = off; off += 2 * p.pabytes;
    p.c1 = off; off += T * 16;
    p.c2 = off; off += NRp * 16;
    p.w1 = off; off += 2 * p.w1bytes;
    p.w2 = off; off += 2 * p.w2bytes;
    p.bars = off; off += 16 * 8;
    p.total = off;
    return p;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float u8_to_f32(uint32_t w, int byte) {
    float f;
    // one I2F.U8 with a byte selector
    asm("cvt.rn.f32.u8 %0, %1;" : "=f"(f) : "r"((w >> (8 * byte)) & 0xffu));
    return f;
}

// Variants measured and dropped: a second, one-column-shifted copy of the right-weight rows (aligned
// odd pairs straight out of LDS.128, but 2x producer stores and 3 more loads per step), three weight stages, four
// periods per loop trip, PRMT+FADD byte conversion.
template <bool GSW, int DC, int REM>
__global__ void __launch_bounds__(WsCfg<GSW, DC>::NT, WsCfg<GSW, DC>::MINB) k_aggregate_ws(const AggParams P) {
    constexpr int NWS = 2;                                   // weight stages
    typedef WsCfg<GSW, DC> C;
    constexpr int T = C::T, NRp = C::NRp, EP = C::EP, CW = C::CW, PW = C::PW;
    extern __shared__ __align__(128) unsigned char smem[];

    const Geom &g = P.g;
    const int win = g.win, pad = g.pad;
    const int NU = g.NU, NR = g.NR, NV = g.NV;
    const WsSmem sp = ws_smem(win, DC, GSW);
    const int winq = (win + 3) >> 2, winp = winq * 4;
    const uint32_t bar0 = smem_u32(smem + sp.bars);
    // barrier slots: 0 centres | 1,2 fullF | 3,4 emptyF | 5,6 fullW | 8,9 emptyW | 11,12 fullE | 13,14 emptyE
    auto BAR = [&](int slot) { return bar0 + 8u * (uint32_t)slot; };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * T;
    const int y = g.row0 + blockIdx.y;
    const int ch = blockIdx.z;
    const int dlo = g.dLo + ch * DC;
    const int erows = g.erow1 - g.erow0;
    const int i_lo = max(0, pad - y), i_hi = min(win - 1, g.H - 1 - y + pad);
    const int nsteps = i_hi - i_lo + 1;

    // Tiles left of the chunk's first disparity hold no evaluated pair (x - d < 0 everywhere): at D = 512 that is 4 %
    // of the blocks.  Their winner keys stay KEY_NONE; the volumes are only touched when they are exported.
    if (x0 + T - 1 < dlo) {
        if (P.vol_export) {
            const int rowo = y - g.row0;
            for (int k = tid; k < T * (DC / 4); k += C::NT) {
                const int x = x0 + k / (DC / 4), kq = (k % (DC / 4)) * 4;
                if (x >= g.W) continue;
                const size_t o = ((size_t)rowo * g.W + x) * P.Dp + (size_t)ch * DC + kq;
                const float4 inf4 = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
                if (P.vol0) *reinterpret_cast<float4 *>(P.vol0 + o) = inf4;
                if (P.vol1) *reinterpret_cast<float4 *>(P.vol1 + o) = inf4;
            }
        }
        return;
    }

    if (tid == 0) {
        mbar_init(BAR(0), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(BAR(1 + s), 1);
            mbar_init(BAR(3 + s), PW);
            mbar_init(BAR(11 + s), 1);
            mbar_init(BAR(13 + s), CW);
        }
        for (int s = 0; s < NWS; ++s) {
            mbar_init(BAR(5 + s), PW);
            mbar_init(BAR(8 + s), CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= CW) {
        // =================================== producers ===================================
        if (C::SETREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");

        if (P.freerun) return;
        const int pw = warp - CW;
        const int f2_start = x0 - dlo - DC + 1 - pad + g.PL2;
        const int c2_start = x0 - dlo - DC + 1 + g.PL2;
        const size_t e_plane = (size_t)g.UW * EP;
        const float4 *C1s = reinterpret_cast<const float4 *>(smem + sp.c1);
        const float4 *C2s = reinterpret_cast<const float4 *>(smem + sp.c2);

        auto issue_F = [&](int n) {
            const int i = i_lo + n, ii = y - pad + i, st = n & 1;
            const uint32_t bar = BAR(1 + st);
            mbar_expect_tx(bar, (uint32_t)((NU + NV + (GSW ? 0 : winq)) * 16));
            tma_load_1d(smem_u32(smem + (sp.f1 + st * sp.f1bytes)), P.F1 + (size_t)(ii - g.erow0) * g.UW + x0, NU * 16, bar);
            tma_load_1d(smem_u32(smem + (sp.f2 + st * sp.f2bytes)), P.F2 + (size_t)(ii - g.erow0) * g.VW + f2_start, NV * 16, bar);
            if (!GSW) tma_load_1d(smem_u32(smem + (sp.pa + st * sp.pabytes)), P.proxarg + (size_t)i * winp, winq * 16, bar);
        };
        auto issue_E = [&](int n) {
            const int ii = y - pad + i_lo + n, st = n & 1;
            const uint32_t bar = BAR(11 + st);
            mbar_expect_tx(bar, (uint32_t)sp.ebytes);
            tma_load_1d(smem_u32(smem + (sp.e + st * sp.ebytes)),
                        static_cast<const uint8_t *>(P.E) + ((size_t)ch * erows + (ii - g.erow0)) * e_plane + (size_t)x0 * EP,
                        (uint32_t)sp.ebytes, bar);
        };

        if (pw == 0 && lane == 0) {
            mbar_expect_tx(BAR(0), (uint32_t)((T + NR) * 16));
            tma_load_1d(smem_u32(smem + sp.c1), P.F1 + (size_t)(y - g.erow0) * g.UW + x0 + pad, T * 16, BAR(0));
            tma_load_1d(smem_u32(smem + sp.c2), P.F2 + (size_t)(y - g.erow0) * g.VW + c2_start, NR * 16, BAR(0));
            issue_F(0);
        }
        constexpr int NCBR = NRp / 32, NCB = NCBR + T / 32;   // 32-column blocks: right image first, then left
        const int NB = winq;                                 // batches of 4 window offsets
        const int npairs = NCB * NB;                         // (column block, batch) pairs, split evenly over the producers
        const int p_begin = (pw * npairs) / PW, p_end = ((pw + 1) * npairs) / PW;
        const int cb_begin = p_begin / NB, jb_begin = p_begin - cb_begin * NB;
        // plain ints: keep the shared-memory plan out of local memory
        const int o_f1 = sp.f1, o_f2 = sp.f2, o_pa = sp.pa, o_w1 = sp.w1, o_w2 = sp.w2;
        const int b_f1 = sp.f1bytes, b_f2 = sp.f2bytes, b_pa = sp.pabytes, b_w1 = sp.w1bytes, b_w2 = sp.w2bytes;

        int sw = 0, phw = 0;                                // weight stage and its phase
        for (int n = 0; n < nsteps; ++n) {
            const int st = n & 1, ph = (n >> 1) & 1;
            if (pw == 0 && lane == 0) {
                if (n + 1 < nsteps) {
                    mbar_wait_long(BAR(3 + ((n + 1) & 1)), (((n + 1) >> 1) & 1) ^ 1);    // feature stage free
                    issue_F(n + 1);
                }
                mbar_wait_long(BAR(13 + st), ph ^ 1);                                // raw-cost stage free
                issue_E(n);
            }
            __syncwarp();
            if (n == 0) mbar_wait(BAR(0), 0);
            mbar_wait(BAR(1 + st), ph);          // features of this window row have landed
            mbar_wait_long(BAR(8 + sw), phw ^ 1); // consumers are done with this weight buffer

            const float4 *f1 = reinterpret_cast<const float4 *>(smem + o_f1 + st * b_f1);
            const float4 *f2 = reinterpret_cast<const float4 *>(smem + o_f2 + st * b_f2);
            const float *parg = reinterpret_cast<const float *>(smem + o_pa + st * b_pa);
            float *W1s = reinterpret_cast<float *>(smem + o_w1 + sw * b_w1);
            float *W2s = reinterpret_cast<float *>(smem + o_w2 + sw * b_w2);
            int cb = cb_begin, jb = jb_begin, left_pairs = p_end - p_begin;
#pragma unroll 1
            while (left_pairs > 0) {
                // ---- per column block: centre, neighbour row, destination column ----
                const bool right = cb < NCBR;                 // warp-uniform
                const int col = (right ? cb : cb - NCBR) * 32 + lane;
                // right: W2s[j][r], r reversed (xr = xr_max - r): centre NR-1-r, neighbour NR-1-r+j
                //        (column r = NR is padding: it reads the float4 in front of C2s and is never used)
                // left : W1s[j][x], centre (y, x0+x), neighbour (ii, x0+x-pad+j)
                const int src = right ? NR - 1 - col : col;
                const float4 c = right ? C2s[src] : C1s[src];
                const int pitch = right ? NRp : T;
                const float4 *nb = (right ? f2 : f1) + src + jb * 4;
                const float *pa = parg + jb * 4;
                float *dst = (right ? W2s : W1s) + col + jb * 4 * pitch;
                const int jend = min(NB, jb + left_pairs);
                left_pairs -= jend - jb;
                // batches of 4 window offsets, two batches in flight: all loads first, stores last, so eight
                // exp/sqrt chains overlap.  Offsets past the window (last batch) read finite padding of the staging
                // buffers and are not stored.
                auto store4 = [&](float *d, float w0, float w1, float w2, float w3) {
                    d[0] = w0;
                    d[pitch] = w1;
                    d[2 * pitch] = w2;
                    d[3 * pitch] = w3;
                };
#pragma unroll 1
                for (; jb < jend;) {
                    const float4 n0 = nb[0], n1 = nb[1], n2 = nb[2], n3 = nb[3];
                    float w0, w1, w2, w3;
                    if (GSW) {
                        // closed form of the relaxation: exp(-||I(q) - I(centre)|| / gamma)
                        w0 = support_weight<false>(c, n0, P.kC2, 0.f);
                        w1 = support_weight<false>(c, n1, P.kC2, 0.f);
                        w2 = support_weight<false>(c, n2, P.kC2, 0.f);
                        w3 = support_weight<false>(c, n3, P.kC2, 0.f);
                        // iterations <= 0: only the centre keeps weight 1.  Right-border abort of the LEFT pass
                        // (_passive.cpp:445-446, :470-471): centre only, or the centre row when y == 0.
                        const bool crow = (i_lo + n) == pad;
                        const bool qk = !right && (x0 + col + pad >= g.W);
                        const int j0 = jb * 4;
                        if (P.iterations <= 0 || (qk && !(y == 0 && crow))) {
                            w0 = (crow && j0 == pad) ? 1.f : 0.f;
                            w1 = (crow && j0 + 1 == pad) ? 1.f : 0.f;
                            w2 = (crow && j0 + 2 == pad) ? 1.f : 0.f;
                            w3 = (crow && j0 + 3 == pad) ? 1.f : 0.f;
                        }
                    } else {
                        const float4 t = *reinterpret_cast<const float4 *>(pa);
                        w0 = support_weight<false>(c, n0, P.kC, t.x);
                        w1 = support_weight<false>(c, n1, P.kC, t.y);
                        w2 = support_weight<false>(c, n2, P.kC, t.z);
                        w3 = support_weight<false>(c, n3, P.kC, t.w);
                    }
                    store4(dst, w0, w1, w2, w3);
                    nb += 4;
                    pa += 4;
                    dst += 4 * pitch;
                    ++jb;
                }
                jb = 0;
                ++cb;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(BAR(5 + sw));        // weights ready
                mbar_arrive(BAR(3 + st));        // feature stage may be refilled
            }
            if (++sw == NWS) { sw = 0; phw ^= 1; }
        }
        return;
    }

    // =================================== consumers ===================================
    if (C::SETREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");

    // Lane -> register tile: a warp is 4 x-groups (of 8 columns) x 8 disparity groups (of 4).
    // The d-groups of x-group xl are rotated by 2*xl (mod DC/4): the right-weight index R0 = T-8-xb+kb is then the
    // same for the four quarter-warps, so one LDS.128 of a right-weight row is ONE 128-byte wavefront for the warp
    // instead of four overlapping ones (shared-memory wavefronts, not issue slots, were the co-limiter: ncu 79 %).
    const int xl = lane >> 3, dl = lane & 7;
    const int xg = (warp / C::NDB) * 4 + xl;
    const int dg = ((warp % C::NDB) * 8 + dl + 2 * xl) % (DC / 4);
    const int xb = 8 * xg;                                   // tile-relative first column
    const int kb = 4 * dg;                                   // chunk-relative first disparity
    // A warp none of whose lane tiles holds an evaluated pair (x - d < 0 everywhere, at the left image border, or d
    // beyond the requested range) only keeps the barriers moving.
    const bool lane_live = (x0 + xb < g.W) && (dlo + kb <= g.dHi) && (x0 + xb + 7 >= dlo + kb);
    const bool warp_live = __any_sync(0xffffffffu, lane_live);
    const int R0 = T - 8 - xb + kb;                          // first reversed right-centre index (multiple of 4)

    u64 acc0[8][2], acc1[8][2];                              // numerator, denominator
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) { acc0[a][b] = 0ull; acc1[a][b] = 0ull; }

    int sw = 0, phw = 0;                                     // weight stage and its phase
    for (int n = 0; n < nsteps; ++n) {
        const int st = n & 1, ph = (n >> 1) & 1;
        if (!P.freerun) {
            mbar_wait(BAR(5 + sw), phw);         // weights of this window row
            mbar_wait(BAR(11 + st), ph);         // raw costs of this window row
        }
        if (warp_live) {
            u64 ring[8][2];
            // raw costs of column c, disparities kb..kb+3: ASW 4 bytes (I2F.U8 on the XU pipe), GSW one float4
            const uint8_t *ep = smem + (sp.e + st * sp.ebytes) + xb * EP + kb * (GSW ? 4 : 1);
            auto load_e = [&](const uint8_t *q, u64 &lo, u64 &hi) {
                if (GSW) {
                    const float4 e = *reinterpret_cast<const float4 *>(q);
                    lo = pk(e.x, e.y);
                    hi = pk(e.z, e.w);
                } else {
                    const uint32_t e = *reinterpret_cast<const uint32_t *>(q);
                    lo = pk(u8_to_f32(e, 0), u8_to_f32(e, 1));
                    hi = pk(u8_to_f32(e, 2), u8_to_f32(e, 3));
                }
            };
#pragma unroll
            for (int a = 0; a < 7; ++a) load_e(ep + a * EP, ring[a][0], ring[a][1]);
            ep += 7 * EP;
            const float *w1p = reinterpret_cast<const float *>(smem + (sp.w1 + sw * sp.w1bytes)) + xb;
            const float *w2p = reinterpret_cast<const float *>(smem + (sp.w2 + sw * sp.w2bytes)) + R0;

            auto step = [&](auto sc) {
                constexpr int s = decltype(sc)::value;
                load_e(ep + s * EP, ring[(7 + s) & 7][0], ring[(7 + s) & 7][1]);
                const float4 wa = *reinterpret_cast<const float4 *>(w1p + s * T);
                const float4 wb = *reinterpret_cast<const float4 *>(w1p + s * T + 4);
                const float4 v0 = *reinterpret_cast<const float4 *>(w2p + s * NRp);
                const float4 v1 = *reinterpret_cast<const float4 *>(w2p + s * NRp + 4);
                const float4 v2 = *reinterpret_cast<const float4 *>(w2p + s * NRp + 8);
                const float w1[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                // right-weight pairs (v[k], v[k+1]): even k are the register pairs the loads produced; odd k straddle two
                // loads (two register moves before a packed op could read them) and are multiplied with scalar ops instead
                const u64 VA[6] = {pk(v0.x, v0.y), pk(v0.z, v0.w), pk(v1.x, v1.y), pk(v1.z, v1.w), pk(v2.x, v2.y), pk(v2.z, v2.w)};
                const float vf[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const u64 w1d = pk(w1[a], w1[a]);
#pragma unroll
                    for (int bp = 0; bp < 2; ++bp) {
                        const int k = 7 - a + 2 * bp;                   // reversed right index of disparity kb+2bp
                        const u64 e2 = ring[(a + s) & 7][bp];
                        // Odd k: two scalar ops (1 issue cycle each, same rounding as the packed form) write the halves
                        // of an aligned pair directly: 2 cycles instead of 2 (packed) + 2 (pair-building moves).
                        if (GSW) {
                            acc0[a][bp] = fma2(w1d, e2, acc0[a][bp]);   // left-reference cost  (_passive.cpp:528)
                            if (k & 1) {                                // right-reference cost (:644)
                                float elo, ehi, alo, ahi;
                                upk(e2, elo, ehi);
                                upk(acc1[a][bp], alo, ahi);
                                acc1[a][bp] = pk(__fmaf_rn(vf[k], elo, alo), __fmaf_rn(vf[k + 1], ehi, ahi));
                            } else {
                                acc1[a][bp] = fma2(VA[k >> 1], e2, acc1[a][bp]);
                            }
                        } else {
                            const u64 ww = (k & 1) ? pk(__fmul_rn(w1[a], vf[k]), __fmul_rn(w1[a], vf[k + 1]))
                                                   : mul2(w1d, VA[k >> 1]);     // w1*w2
                            acc0[a][bp] = fma2(ww, e2, acc0[a][bp]);    // cost += w1*w2*e  (_passive.cpp:77)
                            acc1[a][bp] = add2(acc1[a][bp], ww);        // tot  += w1*w2    (:82)
                        }
                    }
                }
            };
            int j = 0;
            // SS_UNROLL periods of 8 window columns per loop trip: ptxas needs ~50 loop-carried register moves per
            // trip (ring + prefetched loads), so two periods per trip issue fewer instructions (-5 % time at C2)
            auto period = [&]() {
                step(IC<0>{}); step(IC<1>{}); step(IC<2>{}); step(IC<3>{});
                step(IC<4>{}); step(IC<5>{}); step(IC<6>{}); step(IC<7>{});
                ep += 8 * EP;
                w1p += 8 * T;
                w2p += 8 * NRp;
            };
#pragma unroll 1
            for (; j + 8 * SS_UNROLL <= win; j += 8 * SS_UNROLL) {
#pragma unroll
                for (int u = 0; u < SS_UNROLL; ++u) period();
            }
            if (SS_UNROLL > 2 && j + 16 <= win) { period(); period(); j += 16; }
            if (SS_UNROLL > 1 && j + 8 <= win) { period(); j += 8; }
            if (REM > 0) step(IC<0>{});
            if (REM > 1) step(IC<1>{});
            if (REM > 2) step(IC<2>{});
            if (REM > 3) step(IC<3>{});
            if (REM > 4) step(IC<4>{});
            if (REM > 5) step(IC<5>{});
            if (REM > 6) step(IC<6>{});
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(BAR(8 + sw));            // weight buffer free
            mbar_arrive(BAR(13 + st));           // raw-cost stage free
        }
        if (++sw == NWS) { sw = 0; phw ^= 1; }
    }

    // ---- epilogue: normalise, WTA over the chunk, optional volume store --------------------------
    const int rowo = y - g.row0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int x = x0 + xb + a;
        float c0[4], c1[4];
        upk(acc0[a][0], c0[0], c0[1]);
        upk(acc0[a][1], c0[2], c0[3]);
        upk(acc1[a][0], c1[0], c1[1]);
        upk(acc1[a][1], c1[2], c1[3]);
        u64 best = KEY_NONE;
        float out0[4], out1[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int d = dlo + kb + b;
            const bool valid = (x < g.W) && (d <= g.dHi) && (x - d >= 0);
            // ASW: cost / tot (:88), one volume serves both references; GSW: un-normalised left / right sums
            const float cost = GSW ? c0[b] : __fdiv_rn(c0[b], c1[b]);
            out0[b] = valid ? (GSW ? c1[b] : cost) : INFINITY;
            out1[b] = valid ? cost : INFINITY;
            if (valid) {
                const u64 k = make_key(cost, d);
                best = k < best ? k : best;
            }
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
            const u64 o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o < best ? o : best;
        }
        if ((lane & 7) == 0 && x < g.W && best != KEY_NONE) atomicMin(P.bestL + (size_t)rowo * g.W + x, best);
        if (x < g.W) {
            const size_t o = ((size_t)rowo * g.W + x) * P.Dp + (size_t)ch * DC + kb;
            if (P.vol0) *reinterpret_cast<float4 *>(P.vol0 + o) = make_float4(out0[0], out0[1], out0[2], out0[3]);
            if (GSW && P.vol1) *reinterpret_cast<float4 *>(P.vol1 + o) = make_float4(out1[0], out1[1], out1[2], out1[3]);
        }
    }
}

#include "ss_aggregate_tc.cuh"

// ------------------------------------------------------------------------------------------
// k_wta_right: right-reference winners = minimum over diagonals of the aggregated volume
// (_passive.cpp:209-248; d ascending, strict '<' => smallest disparity wins)
// ------------------------------------------------------------------------------------------

__global__ void k_wta_right(const float *__restrict__ vol, u64 *__restrict__ bestR, int W, int rows, int Dp,
                            int dLo, int dHi) {
    const int xr = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (xr >= W || r >= rows) return;
    u64 best = KEY_NONE;
    const int dmax = min(dHi, W - 1 - xr);
    for (int d = dLo; d <= dmax; ++d) {
        const float c = vol[((size_t)r * W + xr + d) * Dp + (d - dLo)];
        const u64 k = make_key(c, d);
        best = k < best ? k : best;
    }
    bestR[(size_t)r * W + xr] = best;
}

__global__ void k_merge_keys(u64 *__restrict__ keys, int nshards, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 b = keys[i];
    for (int s = 1; s < nshards; ++s) {
        const u64 o = keys[(long long)s * n + i];
        b = o < b ? o : b;
    }
    keys[i] = b;
}

// ------------------------------------------------------------------------------------------
// k_finalize: one block per row.  keys -> disparities, L-R invalidation, occlusion fill.
// ------------------------------------------------------------------------------------------

__global__ void k_finalize(const u64 *__restrict__ bestL, const u64 *__restrict__ bestR, int W,
                           int16_t *__restrict__ out, int16_t *__restrict__ out_left,
                           int16_t *__restrict__ out_right, uint8_t *__restrict__ out_invalid) {
    extern __shared__ int16_t sh[];
    int16_t *disp = sh;                                         // [W]
    uint8_t *inv = reinterpret_cast<uint8_t *>(sh + W);         // [W]
    const int r = blockIdx.x;
    const size_t base = (size_t)r * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const u64 k = bestL[base + x];
        // no candidate: dBest stays 0, output x - 0 (_passive.cpp:54, :98)
        const int d = (k == KEY_NONE) ? x : (int)(uint32_t)(k & 0xffffffffu);
        disp[x] = (int16_t)d;
        inv[x] = 0;
        if (out_left) out_left[base + x] = (int16_t)d;
    }
    __syncthreads();
    if (bestR) {
        for (int xr = threadIdx.x; xr < W; xr += blockDim.x) {
            const u64 k = bestR[base + xr];
            const int c = (k == KEY_NONE) ? 0 : xr + (int)(uint32_t)(k & 0xffffffffu);   // selected left column
            if (out_right) out_right[base + xr] = (int16_t)(c - xr);
            if ((int)disp[c] != c - xr) inv[c] = 1;              // _passive.cpp:251-252
        }
        __syncthreads();
        if (out_invalid)
            for (int x = threadIdx.x; x < W; x += blockDim.x) out_invalid[base + x] = inv[x];
        // occlusion fill (_passive.cpp:258-285): every run of invalid pixels takes min(left, right) valid
        // neighbour, or the only existing one at the borders.  Warp 0 scans the row in 32-wide chunks.
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            // pass 1: nearest valid index to the left (inclusive), stored temporarily in `out`
            int carry = -1;
            for (int x0 = 0; x0 < W; x0 += 32) {
                const int x = x0 + lane;
                int v = (x < W && !inv[x]) ? x : -1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v = max(v, t);
                }
                v = max(v, carry);
                carry = __shfl_sync(0xffffffffu, v, 31);
                if (x < W) out[base + x] = (int16_t)v;
            }
            __syncwarp();
            // pass 2 (right to left): nearest valid index to the right, then the fill value
            carry = W;
            for (int x0 = ((W - 1) / 32) * 32; x0 >= 0; x0 -= 32) {
                const int x = x0 + lane;
                int v = (x < W && !inv[x]) ? x : W;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_down_sync(0xffffffffu, v, o);
                    if (lane + o < 32) v = min(v, t);
                }
                v = min(v, carry);
                carry = __shfl_sync(0xffffffffu, v, 0);
                if (x < W) {
                    const int left = out[base + x], right = v;
                    int16_t val;
                    if (!inv[x]) val = disp[x];
                    else if (left < 0 && right >= W) val = 0;        // whole row invalid: reference reads past the
                                                                     // row (:272-275); clamped to 0
                    else if (left < 0) val = disp[right];
                    else if (right >= W) val = disp[left];
                    else val = min(disp[left], disp[right]);
                    out[base + x] = val;
                }
            }
        }
    } else {
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            out[base + x] = disp[x];
            if (out_invalid) out_invalid[base + x] = 0;
        }
    }
}

// register-only FFMA loop: measures the FP32 issue peak that bounds k_aggregate
__global__ void k_ffma_peak(float *out, float a, float b, int iters) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// vol (pitch Dp) -> dense [rows*W*D] host layout, D <= Dp
__global__ void k_compact_volume(const float *__restrict__ vol, float *__restrict__ dense, long long npx, int D, int Dp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx * D) return;
    const long long px = i / D;
    const int k = (int)(i % D);
    dense[i] = vol[px * Dp + k];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------

thread_local std::string t_err;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Ctx {
    std::mutex mu;
    bool ready = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    DevBuf img1, img2, out, f1, f2, evol, vol0, vol1, keysL, keysR, prox, stage_l, stage_r, stage_i, dense;
    DevBuf post_mm, post_pts, post_a;   // scratch of the pre/post steps (ss_post.cuh)
    // cached proximity table key
    int prox_win = -1;
    double prox_gp = -1;
    // instrumentation
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    double agg_ms_done = 0;
    long long agg_launches = 0, total_launches = 0;
    int smem_attr_val[2][12] = {};   // largest dynamic-smem opt-in set so far, per k_aggregate instantiation
    int smem_attr_ws[24] = {};       // same for k_aggregate_ws
    int smem_attr_tc[8] = {};        // same for k_aggregate_tc
};

Ctx g_ctx;
#ifdef SS_DEBUG_DUMP
float *g_dbg = nullptr;
#endif

int fail(int code, const std::string &msg) {
    t_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(SS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    } while (0)

int ensure(DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return SS_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SS_ERR_NOMEM, std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
    }
    b.cap = want;
    return SS_OK;
}

// sRGB byte -> linear*100 (float), exactly colorconversion.hpp:19-37
float srgb_linear100(int c) {
    float v = (float)(c / 255.0);
    if ((double)v > 0.04045) v = powf((float)(((double)v + 0.055) / 1.055), 2.4f);
    else v = (float)((double)v / 12.92);
    return v * 100.0f;
}

int ctx_init(int device) {
    Ctx &c = g_ctx;
    if (c.ready && (device < 0 || device == c.device)) return SS_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SS_ERR_CUDA, std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                     " (libsspassive has no CPU fallback)");
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= n) return fail(SS_ERR_CUDA, "device index out of range");
    CU_TRY(cudaSetDevice(device));
    if (c.ready && device != c.device) {
        // switching device: drop the cache
        DevBuf *bufs[] = {&c.img1, &c.img2, &c.out, &c.f1, &c.f2, &c.evol, &c.vol0, &c.vol1, &c.keysL, &c.keysR,
                          &c.prox, &c.stage_l, &c.stage_r, &c.stage_i, &c.dense, &c.post_mm, &c.post_pts, &c.post_a};
        for (DevBuf *b : bufs) { if (b->p) cudaFree(b->p); *b = DevBuf(); }
        if (c.stream) cudaStreamDestroy(c.stream);
        c.stream = nullptr;
        c.prox_win = -1;
        memset(c.smem_attr_val, 0, sizeof(c.smem_attr_val));
        memset(c.smem_attr_ws, 0, sizeof(c.smem_attr_ws));
        memset(c.smem_attr_tc, 0, sizeof(c.smem_attr_tc));
    }
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(SS_ERR_CUDA, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                     "; libsspassive is built for sm_100a only");
    float lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = srgb_linear100(i);
    CU_TRY(cudaMemcpyToSymbol(c_lin100, lut, sizeof(lut)));
    if (!c.stream) CU_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.device = device;
    c.ready = true;
    return SS_OK;
}

struct Call {
    bool gsw;
    int W, H, win, maxD, minD;
    double gammaC, gammaP;   // ASW
    int gamma, iterations;   // GSW
    float fMax;
    int consistent;          // GSW: always 1
    int row0, row1;
    int dBegin, dEnd;        // evaluated disparity sub-range (inclusive)
};

int validate(const Call &q) {
    if (q.W <= 0 || q.H <= 0) return fail(SS_ERR_DIMS, "Wrong image dimensions!");
    if (q.W > 32767) return fail(SS_ERR_DIMS, "image wider than 32767 columns does not fit int16 disparities");
    if (q.H > 65535) return fail(SS_ERR_DIMS, "image taller than 65535 rows is not supported (one grid row per image row)");
    if (q.maxD - q.minD > 65535) return fail(SS_ERR_PARAM, "more than 65536 disparity candidates are not supported");
    if (!(q.win > 0 && q.win % 2 == 1)) return fail(SS_ERR_WINSIZE, "winSize must be a positive odd number!");
    if (q.win > 255) return fail(SS_ERR_PARAM, "winSize > 255 is not supported");
    if (q.minD < 0) return fail(SS_ERR_PARAM, "minDisparity must be >= 0 (negative values read out of the row upstream)");
    if (q.gsw) {
        if (q.gamma <= 0) return fail(SS_ERR_PARAM, "gamma must be > 0");
    } else {
        if (!(q.gammaC > 0) || !(q.gammaP > 0)) return fail(SS_ERR_PARAM, "gammaC and gammaP must be > 0");
    }
    if (q.row0 < 0 || q.row1 > q.H || q.row0 > q.row1) return fail(SS_ERR_PARAM, "row range out of bounds");
    return SS_OK;
}

Geom make_geom(const Call &q) {
    Geom g;
    g.W = q.W; g.H = q.H; g.win = q.win; g.pad = q.win / 2;
    g.minD = q.minD; g.maxD = q.maxD;
    g.dLo = q.dBegin; g.dHi = q.dEnd;
    const int D = g.dHi - g.dLo + 1;
    g.DC = D <= 32 ? 32 : (D <= 64 ? 64 : 128);
    g.nch = (D + g.DC - 1) / g.DC;
    g.row0 = q.row0; g.row1 = q.row1;
    g.erow0 = q.row0 - g.pad < 0 ? 0 : q.row0 - g.pad;
    g.erow1 = q.row1 + g.pad > q.H ? q.H : q.row1 + g.pad;
    g.T = q.gsw ? TILE_X : TILE_WS;
    g.EP = g.DC + 4;
    g.ntx = (q.W + g.T - 1) / g.T;
    g.UW = (g.ntx * g.T + g.win - 1 + 3) & ~3;
    g.PL2 = g.dLo + g.nch * g.DC - 1 + g.pad;
    g.VW = g.ntx * g.T + g.pad + g.PL2;
    g.NU = g.T + g.win - 1;
    g.NR = g.T + g.DC - 1;
    g.NRp = g.T + g.DC;
    g.NV = g.NR + g.win - 1;
    return g;
}

template <bool GSW, int DC, int REM>
int launch_aggregate_rem(Ctx &c, const AggParams &P, cudaStream_t st) {
    typedef AggCfg<DC> C;
    const SmemPlan sp = smem_plan(P.g.win, DC);
    const int di = (DC == 128 ? 2 : (DC == 64 ? 1 : 0)) * 4 + REM / 2;
    if (c.smem_attr_val[GSW][di] < sp.total) {
        CU_TRY(cudaFuncSetAttribute(k_aggregate<GSW, DC, REM>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
        c.smem_attr_val[GSW][di] = sp.total;
    }
    dim3 grid(P.g.ntx, P.g.row1 - P.g.row0, P.g.nch);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c.profile) {
        CU_TRY(cudaEventCreate(&e0));
        CU_TRY(cudaEventCreate(&e1));
        CU_TRY(cudaEventRecord(e0, st));
    }
    k_aggregate<GSW, DC, REM><<<grid, C::NT, sp.total, st>>>(P);
    CU_TRY(cudaGetLastError());
    if (c.profile) {
        CU_TRY(cudaEventRecord(e1, st));
        c.events.emplace_back(e0, e1);
    }
    c.agg_launches++;
    c.total_launches++;
    return SS_OK;
}

template <bool GSW, int DC, int REM>
int launch_ws_rem(Ctx &c, const AggParams &P, cudaStream_t st) {
    typedef WsCfg<GSW, DC> C;
    const WsSmem sp = ws_smem(P.g.win, DC, GSW);
    if (sp.total > 227 * 1024) return fail(SS_ERR_PARAM, "winSize too large for the shared-memory tiling of k_aggregate_ws");
    const int di = ((DC == 128 ? 2 : (DC == 64 ? 1 : 0)) * 4 + REM / 2) * 2 + (GSW ? 1 : 0);
    if (c.smem_attr_ws[di] < sp.total) {
        CU_TRY(cudaFuncSetAttribute(k_aggregate_ws<GSW, DC, REM>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
        c.smem_attr_ws[di] = sp.total;
    }
    dim3 grid(P.g.ntx, P.g.row1 - P.g.row0, P.g.nch);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c.profile) {
        CU_TRY(cudaEventCreate(&e0));
        CU_TRY(cudaEventCreate(&e1));
        CU_TRY(cudaEventRecord(e0, st));
    }
    k_aggregate_ws<GSW, DC, REM><<<grid, C::NT, sp.total, st>>>(P);
    CU_TRY(cudaGetLastError());
    if (c.profile) {
        CU_TRY(cudaEventRecord(e1, st));
        c.events.emplace_back(e0, e1);
    }
    c.agg_launches++;
    c.total_launches++;
    return SS_OK;
}

// ASW, 128-disparity chunks: denominators on the tensor cores (ss_aggregate_tc.cuh).  win <= 39: two stages of operands in
// tensor memory; 39 < win <= 79: one stage (SINGLE).
template <int REM, bool SINGLE>
int launch_tc_rem(Ctx &c, const AggParams &P, cudaStream_t st) {
    const TcSmem sp = tc_smem(P.g.win, SINGLE);
    int &attr = c.smem_attr_tc[(REM / 2) * 2 + (SINGLE ? 1 : 0)];
    if (attr < sp.total) {
        CU_TRY(cudaFuncSetAttribute(k_aggregate_tc<REM, SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
        attr = sp.total;
    }
    dim3 grid(P.g.ntx, P.g.row1 - P.g.row0, P.g.nch);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c.profile) {
        CU_TRY(cudaEventCreate(&e0));
        CU_TRY(cudaEventCreate(&e1));
        CU_TRY(cudaEventRecord(e0, st));
    }
    k_aggregate_tc<REM, SINGLE><<<grid, 512, sp.total, st>>>(P);
    CU_TRY(cudaGetLastError());
    if (c.profile) {
        CU_TRY(cudaEventRecord(e1, st));
        c.events.emplace_back(e0, e1);
    }
    c.agg_launches++;
    c.total_launches++;
    return SS_OK;
}
template <bool SINGLE>
int launch_tc_s(Ctx &c, const AggParams &P, cudaStream_t st) {
    switch (P.g.win & 7) {
        case 1: return launch_tc_rem<1, SINGLE>(c, P, st);
        case 3: return launch_tc_rem<3, SINGLE>(c, P, st);
        case 5: return launch_tc_rem<5, SINGLE>(c, P, st);
        default: return launch_tc_rem<7, SINGLE>(c, P, st);
    }
}
int launch_tc(Ctx &c, const AggParams &P, cudaStream_t st) {
    return P.g.win <= 39 ? launch_tc_s<false>(c, P, st) : launch_tc_s<true>(c, P, st);
}
bool tc_usable(const Geom &g) {
    if (g.DC != 128 || g.win > 79 || tc_smem(g.win, g.win > 39).total > 227 * 1024) return false;
    const char *e = getenv("SS_TCDEN");
    return !(e && atoi(e) == 0);
}

// true when the warp-specialised kernel can stage this window in shared memory
bool ws_fits(int win, int DC, bool gsw) { return ws_smem(win, DC, gsw).total <= 227 * 1024; }

template <bool GSW, int DC>
int launch_ws(Ctx &c, const AggParams &P, cudaStream_t st) {
    switch (P.g.win & 7) {          // win is odd
        case 1: return launch_ws_rem<GSW, DC, 1>(c, P, st);
        case 3: return launch_ws_rem<GSW, DC, 3>(c, P, st);
        case 5: return launch_ws_rem<GSW, DC, 5>(c, P, st);
        default: return launch_ws_rem<GSW, DC, 7>(c, P, st);
    }
}

template <bool GSW, int DC>
int launch_aggregate(Ctx &c, const AggParams &P, cudaStream_t st) {
    switch (P.g.win & 7) {          // win is odd
        case 1: return launch_aggregate_rem<GSW, DC, 1>(c, P, st);
        case 3: return launch_aggregate_rem<GSW, DC, 3>(c, P, st);
        case 5: return launch_aggregate_rem<GSW, DC, 5>(c, P, st);
        default: return launch_aggregate_rem<GSW, DC, 7>(c, P, st);
    }
}

struct Outputs {
    int16_t *d_final = nullptr;     // [(rows)*W]
    int16_t *d_left = nullptr, *d_right = nullptr;
    uint8_t *d_invalid = nullptr;
    u64 *d_keysL = nullptr, *d_keysR = nullptr;   // when set, stop after WTA (partial / sharded call)
    bool want_vol0 = false, want_vol1 = false;    // keep aggregated volumes (debug export)
};

// Enqueue the whole pipeline for device-resident inputs.
int run_device(Ctx &c, const Call &q, const uint8_t *d_img1, const uint8_t *d_img2, const Outputs &o, cudaStream_t st) {
    const int rows = q.row1 - q.row0;
    if (rows == 0) return SS_OK;
    const bool partial = o.d_keysL != nullptr;
    const bool need_right = q.consistent != 0;
    u64 *keysL = o.d_keysL, *keysR = o.d_keysR;
    const size_t npx = (size_t)rows * q.W;
    if (!partial) {
        int rc = ensure(c.keysL, npx * 8);
        if (rc) return rc;
        keysL = (u64 *)c.keysL.p;
        if (need_right) {
            rc = ensure(c.keysR, npx * 8);
            if (rc) return rc;
            keysR = (u64 *)c.keysR.p;
        }
    }
    CU_TRY(cudaMemsetAsync(keysL, 0xff, npx * 8, st));
    if (need_right && keysR) CU_TRY(cudaMemsetAsync(keysR, 0xff, npx * 8, st));

    const int dB = q.dBegin < q.minD ? q.minD : q.dBegin;
    const int dE = q.dEnd > q.maxD ? q.maxD : q.dEnd;
    if (dE >= dB) {
        Call qq = q;
        qq.dBegin = dB;
        qq.dEnd = dE;
        const Geom g = make_geom(qq);
        const int erows = g.erow1 - g.erow0;
        int rc;
        if ((rc = ensure(c.f1, (size_t)erows * g.UW * 16))) return rc;
        if ((rc = ensure(c.f2, (size_t)erows * g.VW * 16))) return rc;
        if ((rc = ensure(c.evol, q.gsw ? (size_t)g.nch * erows * g.UW * g.DC * 4
                                       : (size_t)g.nch * erows * g.UW * g.EP + 64))) return rc;
        const int Dp = g.nch * g.DC;
        const bool vol0 = need_right || o.want_vol0;
        if (vol0 && (rc = ensure(c.vol0, npx * Dp * 4))) return rc;
        if (o.want_vol1 && (rc = ensure(c.vol1, npx * Dp * 4))) return rc;

        // proximity exponents: -log2(e) * sqrt(di^2+dj^2) / gammaP  (_passive.cpp:358-364)
        if (!q.gsw && (c.prox_win != q.win || c.prox_gp != q.gammaP)) {
            const int winp = (q.win + 3) & ~3;               // 16-byte rows: one TMA bulk copy per window row
            std::vector<float> h((size_t)q.win * winp, 0.f);
            const int p = q.win / 2;
            for (int i = 0; i < q.win; ++i)
                for (int j = 0; j < q.win; ++j) {
                    const double di = i - p, dj = j - p;
                    h[(size_t)i * winp + j] = (float)(-1.4426950408889634 * std::sqrt(di * di + dj * dj) / q.gammaP);
                }
            if ((rc = ensure(c.prox, h.size() * 4))) return rc;
            CU_TRY(cudaMemcpyAsync(c.prox.p, h.data(), h.size() * 4, cudaMemcpyHostToDevice, st));
            CU_TRY(cudaStreamSynchronize(st));   // h goes out of scope
            c.prox_win = q.win;
            c.prox_gp = q.gammaP;
        }

        {
            dim3 b(128), g1((g.UW + 127) / 128, erows), g2((g.VW + 127) / 128, erows);
            if (q.gsw) {
                k_prep_features<true><<<g1, b, 0, st>>>(d_img1, (float4 *)c.f1.p, g.W, g.erow0, erows, g.UW, g.pad);
                k_prep_features<true><<<g2, b, 0, st>>>(d_img2, (float4 *)c.f2.p, g.W, g.erow0, erows, g.VW, g.PL2);
            } else {
                k_prep_features<false><<<g1, b, 0, st>>>(d_img1, (float4 *)c.f1.p, g.W, g.erow0, erows, g.UW, g.pad);
                k_prep_features<false><<<g2, b, 0, st>>>(d_img2, (float4 *)c.f2.p, g.W, g.erow0, erows, g.VW, g.PL2);
            }
            CU_TRY(cudaGetLastError());
            c.total_launches += 2;
        }
        {
            const int n = g.UW * (g.DC / 4);
            dim3 b(256), gr((n + 255) / 256, erows, g.nch);
            if (q.gsw) k_cost_volume<true><<<gr, b, 0, st>>>(d_img1, d_img2, c.evol.p, g, q.fMax);
            else k_cost_volume<false><<<gr, b, 0, st>>>(d_img1, d_img2, c.evol.p, g, 0.f);
            CU_TRY(cudaGetLastError());
            c.total_launches += 1;
        }
        AggParams P;
        P.g = g;
        P.F1 = (const float4 *)c.f1.p;
        P.F2 = (const float4 *)c.f2.p;
        P.E = c.evol.p;
        P.proxarg = (const float *)c.prox.p;
        P.kC = q.gsw ? (float)q.gamma : (float)(-1.4426950408889634 / q.gammaC);
        P.kC2 = q.gsw ? (float)(-1.4426950408889634 / (double)q.gamma) : 0.f;
        P.iterations = q.iterations;
        P.bestL = keysL;
        P.vol0 = vol0 ? (float *)c.vol0.p : nullptr;
        P.vol1 = o.want_vol1 ? (float *)c.vol1.p : nullptr;
        P.Dp = Dp;
        P.vol_export = (o.want_vol0 || o.want_vol1) ? 1 : 0;
        P.freerun = getenv("SS_FREERUN") ? atoi(getenv("SS_FREERUN")) : 0;
#ifdef SS_DEBUG_DUMP
        P.dbg = g_dbg;
#endif
        const bool gsw_ws = q.gsw && ws_fits(q.win, g.DC, true) && !(getenv("SS_GSW_SINGLE") && atoi(getenv("SS_GSW_SINGLE")));
        if (q.gsw && !gsw_ws) {
            // single-role kernel: exact expf / IEEE sqrt weights; also the path for windows whose float raw-cost
            // tiles do not fit the double-buffered staging of the warp-specialised kernel
            if (g.DC == 128) rc = launch_aggregate<true, 128>(c, P, st);
            else if (g.DC == 64) rc = launch_aggregate<true, 64>(c, P, st);
            else rc = launch_aggregate<true, 32>(c, P, st);
        } else if (q.gsw) {
            if (g.DC == 128) rc = launch_ws<true, 128>(c, P, st);
            else if (g.DC == 64) rc = launch_ws<true, 64>(c, P, st);
            else rc = launch_ws<true, 32>(c, P, st);
        } else if (tc_usable(g)) {
            rc = launch_tc(c, P, st);
        } else {
            if (g.DC == 128) rc = launch_ws<false, 128>(c, P, st);
            else if (g.DC == 64) rc = launch_ws<false, 64>(c, P, st);
            else rc = launch_ws<false, 32>(c, P, st);
        }
        if (rc) return rc;
        if (need_right && keysR) {
            dim3 b(128), gr((q.W + 127) / 128, rows);
            k_wta_right<<<gr, b, 0, st>>>((const float *)c.vol0.p, keysR, q.W, rows, Dp, dB, dE);
            CU_TRY(cudaGetLastError());
            c.total_launches += 1;
        }
    }
    if (!partial) {
        const size_t sh = (size_t)q.W * 3 + 16;
        if (sh > 48 * 1024) CU_TRY(cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        k_finalize<<<rows, 128, sh, st>>>(keysL, need_right ? keysR : nullptr, q.W, o.d_final, o.d_left, o.d_right, o.d_invalid);
        CU_TRY(cudaGetLastError());
        c.total_launches += 1;
    }
    return SS_OK;
}

// host wrapper: H2D, run, D2H of the stripe
int run_host(const Call &q, const uint8_t *img1, const uint8_t *img2, int16_t *out, int16_t *out_left, int16_t *out_right,
             uint8_t *out_invalid, float *out_vol0, float *out_vol1) {
    if (!img1 || !img2) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = validate(q);
    if (rc) return rc;
    Ctx &c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    if ((rc = ctx_init(-1))) return rc;
    CU_TRY(cudaSetDevice(c.device));
    const int rows = q.row1 - q.row0;
    const size_t nimg = (size_t)q.W * q.H * 3, npx = (size_t)rows * q.W;
    if ((rc = ensure(c.img1, nimg))) return rc;
    if ((rc = ensure(c.img2, nimg))) return rc;
    if ((rc = ensure(c.out, npx * 2 + 2))) return rc;
    cudaStream_t st = c.stream;
    // only the rows the stripe needs travel
    const int pad = q.win / 2;
    const int er0 = q.row0 - pad < 0 ? 0 : q.row0 - pad, er1 = q.row1 + pad > q.H ? q.H : q.row1 + pad;
    const size_t off = (size_t)er0 * q.W * 3, len = (size_t)(er1 - er0) * q.W * 3;
    if (len) {
        CU_TRY(cudaMemcpyAsync((uint8_t *)c.img1.p + off, img1 + off, len, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync((uint8_t *)c.img2.p + off, img2 + off, len, cudaMemcpyHostToDevice, st));
    }
    Outputs o;
    o.d_final = (int16_t *)c.out.p;
    if (out_left) { if ((rc = ensure(c.stage_l, npx * 2 + 2))) return rc; o.d_left = (int16_t *)c.stage_l.p; }
    if (out_right) { if ((rc = ensure(c.stage_r, npx * 2 + 2))) return rc; o.d_right = (int16_t *)c.stage_r.p; }
    if (out_invalid) { if ((rc = ensure(c.stage_i, npx + 1))) return rc; o.d_invalid = (uint8_t *)c.stage_i.p; }
    o.want_vol0 = out_vol0 != nullptr;
    o.want_vol1 = out_vol1 != nullptr;
    if ((rc = run_device(c, q, (const uint8_t *)c.img1.p, (const uint8_t *)c.img2.p, o, st))) return rc;
    if (out && npx) CU_TRY(cudaMemcpyAsync(out, c.out.p, npx * 2, cudaMemcpyDeviceToHost, st));
    if (out_left && npx) CU_TRY(cudaMemcpyAsync(out_left, o.d_left, npx * 2, cudaMemcpyDeviceToHost, st));
    if (out_right && npx && q.consistent) CU_TRY(cudaMemcpyAsync(out_right, o.d_right, npx * 2, cudaMemcpyDeviceToHost, st));
    if (out_invalid && npx) CU_TRY(cudaMemcpyAsync(out_invalid, o.d_invalid, npx, cudaMemcpyDeviceToHost, st));
    const int D = q.maxD - q.minD + 1;
    if ((out_vol0 || out_vol1) && D > 0 && npx) {
        const int DC = D <= 32 ? 32 : (D <= 64 ? 64 : 128);
        const int Dp = ((D + DC - 1) / DC) * DC;
        if ((rc = ensure(c.dense, npx * D * 4))) return rc;
        const long long n = (long long)npx * D;
        for (int v = 0; v < 2; ++v) {
            float *dst = v == 0 ? out_vol0 : out_vol1;
            if (!dst) continue;
            k_compact_volume<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float *)(v == 0 ? c.vol0.p : c.vol1.p), (float *)c.dense.p,
                                                                             (long long)npx, D, Dp);
            CU_TRY(cudaGetLastError());
            c.total_launches += 1;
            CU_TRY(cudaMemcpyAsync(dst, c.dense.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
        }
    }
    CU_TRY(cudaStreamSynchronize(st));
    return SS_OK;
}

Call asw_call(int W, int H, int win, int maxD, int minD, double gc, double gp, int consistent, int r0, int r1) {
    Call q;
    q.gsw = false; q.W = W; q.H = H; q.win = win; q.maxD = maxD; q.minD = minD;
    q.gammaC = gc; q.gammaP = gp; q.gamma = 0; q.iterations = 0; q.fMax = 0.f;
    q.consistent = consistent ? 1 : 0;
    q.row0 = r0; q.row1 = r1; q.dBegin = minD; q.dEnd = maxD;
    return q;
}
Call gsw_call(int W, int H, int win, int maxD, int minD, int gamma, float fMax, int iterations, int r0, int r1) {
    Call q;
    q.gsw = true; q.W = W; q.H = H; q.win = win; q.maxD = maxD; q.minD = minD;
    q.gammaC = 0; q.gammaP = 0; q.gamma = gamma; q.iterations = iterations; q.fMax = fMax;
    q.consistent = 1;                       // workerGSW always runs both passes (_passive.cpp:428-665)
    q.row0 = r0; q.row1 = r1; q.dBegin = minD; q.dEnd = maxD;
    return q;
}

int device_entry(const Call &q, const uint8_t *d1, const uint8_t *d2, const Outputs &o, void *stream) {
    if (!d1 || !d2) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = validate(q);
    if (rc) return rc;
    Ctx &c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    if ((rc = ctx_init(-1))) return rc;
    CU_TRY(cudaSetDevice(c.device));
    return run_device(c, q, d1, d2, o, (cudaStream_t)stream);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------

extern "C" {

int ss_abi_version(void) { return 1; }

const char *ss_last_error(void) { return t_err.c_str(); }

int ss_init(int device) {
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    return ctx_init(device);
}

int ss_shutdown(void) {
    Ctx &c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.ready) return SS_OK;
    cudaSetDevice(c.device);
    DevBuf *bufs[] = {&c.img1, &c.img2, &c.out, &c.f1, &c.f2, &c.evol, &c.vol0, &c.vol1, &c.keysL, &c.keysR,
                      &c.prox, &c.stage_l, &c.stage_r, &c.stage_i, &c.dense, &c.post_mm, &c.post_pts, &c.post_a};
    for (DevBuf *b : bufs) { if (b->p) cudaFree(b->p); *b = DevBuf(); }
    for (auto &ev : c.events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    c.events.clear();
    if (c.stream) cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    c.prox_win = -1;
    c.ready = false;
    memset(c.smem_attr_val, 0, sizeof(c.smem_attr_val));
    memset(c.smem_attr_ws, 0, sizeof(c.smem_attr_ws));
    memset(c.smem_attr_tc, 0, sizeof(c.smem_attr_tc));
    return SS_OK;
}

int ss_asw_compute(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                   int min_disp, double gamma_c, double gamma_p, int consistent, int16_t *out_disp) {
    if (!out_disp) return fail(SS_ERR_FORMAT, "Invalid input format!");
    return run_host(asw_call(width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, 0, height), img1, img2,
                    out_disp, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int ss_gsw_compute(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                   int min_disp, int gamma, float f_max, int iterations, int bins, int16_t *out_disp) {
    (void)bins;
    if (!out_disp) return fail(SS_ERR_FORMAT, "Invalid input format!");
    return run_host(gsw_call(width, height, win_size, max_disp, min_disp, gamma, f_max, iterations, 0, height), img1, img2,
                    out_disp, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int ss_asw_compute_rows(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                        int min_disp, double gamma_c, double gamma_p, int consistent, int row_begin, int row_end,
                        int16_t *out_rows) {
    if (!out_rows) return fail(SS_ERR_FORMAT, "Invalid input format!");
    return run_host(asw_call(width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, row_begin, row_end), img1,
                    img2, out_rows, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int ss_gsw_compute_rows(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                        int min_disp, int gamma, float f_max, int iterations, int bins, int row_begin, int row_end,
                        int16_t *out_rows) {
    (void)bins;
    if (!out_rows) return fail(SS_ERR_FORMAT, "Invalid input format!");
    return run_host(gsw_call(width, height, win_size, max_disp, min_disp, gamma, f_max, iterations, row_begin, row_end), img1, img2,
                    out_rows, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int ss_asw_compute_device(const uint8_t *d_img1, const uint8_t *d_img2, int width, int height, int win_size,
                          int max_disp, int min_disp, double gamma_c, double gamma_p, int consistent, int row_begin,
                          int row_end, int16_t *d_out_rows, void *stream) {
    if (!d_out_rows) return fail(SS_ERR_FORMAT, "Invalid input format!");
    Outputs o;
    o.d_final = d_out_rows;
    return device_entry(asw_call(width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, row_begin, row_end),
                        d_img1, d_img2, o, stream);
}

int ss_gsw_compute_device(const uint8_t *d_img1, const uint8_t *d_img2, int width, int height, int win_size,
                          int max_disp, int min_disp, int gamma, float f_max, int iterations, int bins, int row_begin,
                          int row_end, int16_t *d_out_rows, void *stream) {
    (void)bins;
    if (!d_out_rows) return fail(SS_ERR_FORMAT, "Invalid input format!");
    Outputs o;
    o.d_final = d_out_rows;
    return device_entry(gsw_call(width, height, win_size, max_disp, min_disp, gamma, f_max, iterations, row_begin, row_end), d_img1,
                        d_img2, o, stream);
}

int ss_asw_partial_device(const uint8_t *d_img1, const uint8_t *d_img2, int width, int height, int win_size,
                          int max_disp, int min_disp, double gamma_c, double gamma_p, int consistent, int row_begin,
                          int row_end, int disp_begin, int disp_end, uint64_t *d_best_left, uint64_t *d_best_right,
                          void *stream) {
    if (!d_best_left || (consistent && !d_best_right)) return fail(SS_ERR_FORMAT, "Invalid input format!");
    Call q = asw_call(width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, row_begin, row_end);
    q.dBegin = disp_begin;
    q.dEnd = disp_end;
    Outputs o;
    o.d_keysL = (u64 *)d_best_left;
    o.d_keysR = (u64 *)d_best_right;
    return device_entry(q, d_img1, d_img2, o, stream);
}

int ss_merge_keys_device(uint64_t *d_keys, int n_shards, long long n, void *stream) {
    if (!d_keys || n_shards < 1 || n < 0) return fail(SS_ERR_FORMAT, "Invalid input format!");
    Ctx &c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    int rc = ctx_init(-1);
    if (rc) return rc;
    if (n == 0 || n_shards == 1) return SS_OK;
    k_merge_keys<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((u64 *)d_keys, n_shards, n);
    CU_TRY(cudaGetLastError());
    c.total_launches += 1;
    return SS_OK;
}

int ss_finalize_keys_device(const uint64_t *d_best_left, const uint64_t *d_best_right, int width, int rows,
                            int min_disp, int16_t *d_out_rows, void *stream) {
    (void)min_disp;
    if (!d_best_left || !d_out_rows || width <= 0 || rows < 0) return fail(SS_ERR_FORMAT, "Invalid input format!");
    Ctx &c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    int rc = ctx_init(-1);
    if (rc) return rc;
    if (rows == 0) return SS_OK;
    if ((size_t)width * 3 + 16 > 48 * 1024)
        CU_TRY(cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, width * 3 + 16));
    k_finalize<<<rows, 128, (size_t)width * 3 + 16, (cudaStream_t)stream>>>((const u64 *)d_best_left, (const u64 *)d_best_right, width,
                                                                           d_out_rows, nullptr, nullptr, nullptr);
    CU_TRY(cudaGetLastError());
    c.total_launches += 1;
    return SS_OK;
}

int ss_asw_stages(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                  int min_disp, double gamma_c, double gamma_p, int consistent, int16_t *out_left, int16_t *out_right,
                  uint8_t *out_invalid, int16_t *out_final, float *out_cost) {
    return run_host(asw_call(width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, 0, height), img1, img2,
                    out_final, out_left, out_right, out_invalid, out_cost, nullptr);
}

int ss_gsw_stages(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                  int min_disp, int gamma, float f_max, int iterations, int bins, int16_t *out_left, int16_t *out_right,
                  uint8_t *out_invalid, int16_t *out_final, float *out_cost_left, float *out_cost_right) {
    (void)bins;
    return run_host(gsw_call(width, height, win_size, max_disp, min_disp, gamma, f_max, iterations, 0, height), img1, img2, out_final,
                    out_left, out_right, out_invalid, out_cost_right, out_cost_left);
}

int ss_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    g_ctx.profile = on != 0;
    return SS_OK;
}

int ss_profile_reset(void) {
    Ctx &c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    for (auto &ev : c.events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    c.events.clear();
    c.agg_ms_done = 0;
    c.agg_launches = 0;
    c.total_launches = 0;
    return SS_OK;
}

int ss_profile_read(double *agg_ms, long long *agg_launches, long long *total_launches) {
    Ctx &c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    for (auto &ev : c.events) {
        CU_TRY(cudaEventSynchronize(ev.second));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, ev.first, ev.second));
        c.agg_ms_done += ms;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    c.events.clear();
    if (agg_ms) *agg_ms = c.agg_ms_done;
    if (agg_launches) *agg_launches = c.agg_launches;
    if (total_launches) *total_launches = c.total_launches;
    return SS_OK;
}

int ss_measure_fp32_peak(double *tflops, void *stream) {
    if (!tflops) return fail(SS_ERR_FORMAT, "Invalid input format!");
    Ctx &c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    int rc = ctx_init(-1);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, c.device));
    const int blocks = prop.multiProcessorCount * 8, tpb = 256, iters = 8192;
    if ((rc = ensure(c.dense, (size_t)blocks * tpb * 4))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CU_TRY(cudaEventRecord(e0, st));
        k_ffma_peak<<<blocks, tpb, 0, st>>>((float *)c.dense.p, 1.0001f, 0.5f, iters);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaEventRecord(e1, st));
        CU_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = (double)blocks * tpb * iters * 16 * 2 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return SS_OK;
}

#ifdef SS_DEBUG_DUMP
int ss_debug_set(int bx, int by, int step, int nfloats) {
    if (!g_dbg) cudaMalloc(&g_dbg, (size_t)(nfloats + 16) * 4);
    float h[3] = {(float)bx, (float)by, (float)step};
    cudaMemcpy(g_dbg, h, sizeof(h), cudaMemcpyHostToDevice);
    return 0;
}
int ss_debug_get(float *out, int nfloats) {
    cudaDeviceSynchronize();
    cudaMemcpy(out, g_dbg + 16, (size_t)nfloats * 4, cudaMemcpyDeviceToHost);
    return 0;
}
#endif

}  // extern "C"

#include "ss_post.cuh"
