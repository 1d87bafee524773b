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
//                     accumulator in TMEM.  k_aggregate_ws is the all-CUDA-core form (GSW, short disparity ranges, large
//                     windows).  Fused into the same launch: WTA over the disparity chunk for BOTH references -- the
//                     left one per pixel (warp shuffle), the right one per diagonal x - d of the block's cost tile
//                     (C_R[xr,d] == C_L[xr+d,d], SURVEY.md 3.3-5: the "roughly doubled" second pass of the reference,
//                     passive.py:39, costs one shared-memory pass) -- merged across blocks with 64-bit atomicMin keys.
//   k_finalize        key decode, L-R invalidation (_passive.cpp:251-252), occlusion fill (:258-285).
//
// No CPU fallback exists: every entry point fails with SS_ERR_CUDA when no device is usable.

#include "../../include/ss_passive.h"
#include "../../include/ss_post.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

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
    int dLo, dHi;      // first disparity of the first chunk this launch covers (minD + k * DC), last evaluated disparity
    int dVLo;          // first evaluated disparity (>= dLo; > dLo only for a disparity-range shard that starts inside a chunk)
    int DC, nch;       // disparity chunk, number of chunks covering [dLo, dHi]
    int row0, row1;    // output rows [row0,row1)
    int erow0, erow1;  // input rows needed [erow0,erow1) = output rows +- pad, clipped
    int tile0;         // first linear (chunk, row, x tile) index of this launch (blockIdx.x counts from here)
    int nsub;          // 1, or T/32 for the launch that covers the last partial wave: blockIdx.y then restricts a block
                       // to ONE 32-column block of its tile (same arithmetic per pair, a third of the work per SM)
    int T;             // output columns per block: 64 (GSW) or 96 (ASW)
    int EP, EG;        // ASW raw-cost plane: column c of a row starts at byte c * EP + (c >> 3) * EG.  k_aggregate_ws: EP = DC + 4,
                       // EG = 0; k_aggregate_tc: EP = 128, EG = 24 -- lane groups 8 columns apart then start 6 banks apart, which
                       // with the consumers' disparity-group rotation (2 per group) spreads a warp's 32 words over 32 banks
    int EPL;           // bytes per (chunk, row) plane of the ASW raw-cost volume (multiple of 16: TMA source alignment)
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
        // ASW: the truncated AD is an integer in [0,40] -> one byte; 4 disparities per 32-bit word.  (A bfloat16 tile -- exact
        // for these integers, byte permutes instead of I2F.U8 in the consumers -- measured 0.5 % slower at C2.)
        *reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(Eout) + ((size_t)ch * (g.erow1 - g.erow0) + r) * g.EPL + (size_t)up * g.EP +
                                      (size_t)(up >> 3) * g.EG + kq) =
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
    u64 *bestL;           // [(row1-row0)*W] packed (cost,disp) keys of the left-reference pass, atomicMin
    u64 *bestR;           // optional: the same for the right-reference pass, indexed by the RIGHT column xr = x - d
    float *vol0;          // optional (debug export only): ASW cost / GSW right cost  [(rows)*W*Dp]
    float *vol1;          // optional (debug export only): GSW left cost
    int Dp;               // pitch of vol0/vol1 (= nch*DC)
    int vol_export;       // the volumes are returned to the caller (debug export): unevaluated pairs must read +inf
    int freerun;          // timing experiments (SS_FREERUN bit mask, see run_device); results are garbage
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

// Right-reference WTA fused into the aggregation epilogue (_passive.cpp:191-248; north_star: "WTA argmin + left-right
// consistency fused into the same launch").  The right-pass cost of (xr, d) is the left-pass cost of (x = xr + d, d)
// (ASW: bit for bit; GSW: its own right-reference sum over the same pair), so a block that has staged its
// costs in shared memory as Cs[r][x], r = T-1-x+k (x tile-relative column, k chunk-relative disparity; invalid pairs hold
// +INF), owns in row r EVERY candidate this tile holds for right pixel xr = x0 - dlo + T-1-r.  One warp reduces one row with
// the packed (cost, disparity) keys -- unsigned min = ascending d with strict '<' (:209-248) -- and issues one atomicMin.
template <int T, int DC>
__device__ __forceinline__ void wta_right_rows(const float *Cs, int warp, int nwarps, int lane, int x0, int dlo, int W, u64 *bestR_row) {
    for (int r = warp; r < T + DC - 1; r += nwarps) {
        const int xr = x0 - dlo + T - 1 - r;
        if (xr < 0 || xr >= W) continue;                     // warp-uniform
        u64 best = KEY_NONE;
#pragma unroll
        for (int x = lane; x < T; x += 32) {
            const int k = r - (T - 1 - x);
            if (k >= 0 && k < DC) {
                const float c = Cs[r * T + x];
                const u64 key = make_key(c, dlo + k);
                if (c < INFINITY && key < best) best = key;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const u64 o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o < best ? o : best;
        }
        if (lane == 0 && best != KEY_NONE) atomicMin(bestR_row + xr, best);
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
    // blocks per SM the register budget is sized for.  DC = 32 (128 threads): 3 blocks -> 168 registers, no spills (4 blocks
    // = 128 registers spilled 72-120 bytes in the hot loop)
    static constexpr int MINB = GSW ? 1 : (NDB == 1 ? 3 : 4 / NDB);
    static constexpr int NRp = T + DC;
    static constexpr int EP = GSW ? DC * 4 : DC + 4;   // bytes per raw-cost column (ASW: bytes, skewed by 4)
    // 16 warps: the producers give registers to the consumers.  setmaxnreg must be executed by whole warpgroups (4 aligned
    // warps) with the same operand: only the DC = 128 block (12 consumer + 4 producer warps) is laid out that way -- a
    // 3 + 1 warp block (DC = 32) that tried it deadlocked on the B200.
    static constexpr bool SETREG = DC == 128 && NT == 512;   // ASW 12 + 4 warps, GSW 8 + 8 warps: whole warpgroups per role
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
    int off = 256;                                 // header: the 16 mbarriers (outside every region that is reused later)
    p.bars = 0;
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
    // the fused right-reference WTA stages the block's costs as float[T + DC - 1][T] over the (dead) streaming buffers
    const int stage = 256 + (T + DC - 1) * T * 4;
    p.total = off > stage ? off : stage;
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
    const int tile = g.tile0 + (int)blockIdx.x, per_ch = g.ntx * (g.row1 - g.row0);
    const int ch = tile / per_ch, rem = tile - ch * per_ch;
    const int x0 = (rem % g.ntx) * T;
    const int y = g.row0 + rem / g.ntx;
    const int xsub = g.nsub > 1 ? (int)blockIdx.y : -1;     // tail launch: only this 32-column block of the tile
    const int dlo = g.dLo + ch * DC;
    const int erows = g.erow1 - g.erow0;
    const int i_lo = max(0, pad - y), i_hi = min(win - 1, g.H - 1 - y + pad);
    const int nsteps = i_hi - i_lo + 1;

    // Tiles left of the chunk's first disparity hold no evaluated pair (x - d < 0 everywhere): at D = 512 that is 4 %
    // of the blocks.  Their winner keys stay KEY_NONE; the volumes are only touched when they are exported.
    if (x0 + T - 1 < dlo) {
        if (P.vol_export && xsub <= 0) {
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
        const size_t e_plane = GSW ? (size_t)g.UW * EP : (size_t)g.EPL;
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
        constexpr int NCBR = NRp / 32;                       // 32-column blocks: right image first, then left
        // a block restricted to x-block xsub needs right centres r = T-1-x+k, x in the block: blocks cbr0 .. cbr0+nbr-1, and
        // one left block
        const int cbr0 = xsub < 0 ? 0 : T / 32 - 1 - xsub;
        const int nbr = xsub < 0 ? NCBR : (T - 1 - 32 * xsub + DC - 1) / 32 - cbr0 + 1;
        const int NCB = nbr + (xsub < 0 ? T / 32 : 1);
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
                    mbar_wait(BAR(3 + ((n + 1) & 1)), (((n + 1) >> 1) & 1) ^ 1);    // feature stage free
                    issue_F(n + 1);
                }
                mbar_wait(BAR(13 + st), ph ^ 1);                                // raw-cost stage free
                issue_E(n);
            }
            __syncwarp();
            if (n == 0) mbar_wait(BAR(0), 0);
            mbar_wait(BAR(1 + st), ph);          // features of this window row have landed
            mbar_wait(BAR(8 + sw), phw ^ 1); // consumers are done with this weight buffer

            const float4 *f1 = reinterpret_cast<const float4 *>(smem + o_f1 + st * b_f1);
            const float4 *f2 = reinterpret_cast<const float4 *>(smem + o_f2 + st * b_f2);
            const float *parg = reinterpret_cast<const float *>(smem + o_pa + st * b_pa);
            float *W1s = reinterpret_cast<float *>(smem + o_w1 + sw * b_w1);
            float *W2s = reinterpret_cast<float *>(smem + o_w2 + sw * b_w2);
            int cb = cb_begin, jb = jb_begin, left_pairs = p_end - p_begin;
#pragma unroll 1
            while (left_pairs > 0) {
                // ---- per column block: centre, neighbour row, destination column ----
                const bool right = cb < nbr;                  // warp-uniform
                const int col = (right ? cbr0 + cb : (xsub < 0 ? cb - nbr : xsub)) * 32 + lane;
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
    const bool sub_ok = xsub < 0 || warp / C::NDB == xsub;   // warp-uniform
    const bool lane_live = sub_ok && (x0 + xb < g.W) && (dlo + kb <= g.dHi) && (dlo + kb + 3 >= g.dVLo) && (x0 + xb + 7 >= dlo + kb);
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

    // ---- epilogue: normalise, WTA over the chunk (both references), optional volume store ------------
    const int rowo = y - g.row0;
    float *Cs = reinterpret_cast<float *>(smem + 256);       // staged costs for the right-reference WTA
    if (P.bestR) asm volatile("bar.sync 2, %0;" ::"n"(CW * 32) : "memory");   // every consumer is done with the streaming buffers
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
            const bool valid = sub_ok && (x < g.W) && (d >= g.dVLo) && (d <= g.dHi) && (x - d >= 0);
            // ASW: cost / tot (:88), one volume serves both references; GSW: un-normalised left / right sums
            const float cost = GSW ? c0[b] : __fdiv_rn(c0[b], c1[b]);
            out0[b] = valid ? (GSW ? c1[b] : cost) : INFINITY;
            out1[b] = valid ? cost : INFINITY;
            if (valid) {
                const u64 k = make_key(cost, d);
                best = k < best ? k : best;
            }
            if (P.bestR) Cs[(T - 1 - (xb + a) + kb + b) * T + xb + a] = out0[b];
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
            const u64 o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o < best ? o : best;
        }
        if ((lane & 7) == 0 && x < g.W && best != KEY_NONE) atomicMin(P.bestL + (size_t)rowo * g.W + x, best);
        if (x < g.W && sub_ok) {
            const size_t o = ((size_t)rowo * g.W + x) * P.Dp + (size_t)ch * DC + kb;
            if (P.vol0) *reinterpret_cast<float4 *>(P.vol0 + o) = make_float4(out0[0], out0[1], out0[2], out0[3]);
            if (GSW && P.vol1) *reinterpret_cast<float4 *>(P.vol1 + o) = make_float4(out1[0], out1[1], out1[2], out1[3]);
        }
    }
    if (P.bestR) {
        asm volatile("bar.sync 2, %0;" ::"n"(CW * 32) : "memory");
        wta_right_rows<T, DC>(Cs, warp, CW, lane, x0, dlo, g.W, P.bestR + (size_t)rowo * g.W);
    }
}

#include "ss_aggregate_tc.cuh"

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

// reset != 0: every key that was read is set back to KEY_NONE, so the library's cached key planes need no memset before the
// next call.
__global__ void k_finalize(u64 *__restrict__ bestL, u64 *__restrict__ bestR, int W,
                           int16_t *__restrict__ out, int16_t *__restrict__ out_left,
                           int16_t *__restrict__ out_right, uint8_t *__restrict__ out_invalid, int reset) {
    extern __shared__ int16_t sh[];
    int16_t *disp = sh;                                         // [W]
    uint8_t *inv = reinterpret_cast<uint8_t *>(sh + W);         // [W]
    const int r = blockIdx.x;
    const size_t base = (size_t)r * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const u64 k = bestL[base + x];
        if (reset) bestL[base + x] = KEY_NONE;
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
            if (reset) bestR[base + xr] = KEY_NONE;
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

// One context per CUDA device: stream, cached scratch, instrumentation.  A context is only touched with its mutex held.
struct Ctx {
    std::mutex mu;
    bool ready = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    // The scratch below is shared by every call on this device, whatever stream the caller enqueues on: `done` is recorded
    // after the last enqueue of a call and the next call's stream waits for it, so two calls never overlap on the scratch.
    cudaEvent_t done = nullptr;
    bool done_valid = false;
    DevBuf img1, img2, out, f1, f2, evol, vol0, vol1, keys, prox, stage_l, stage_r, stage_i, dense;
    DevBuf post_mm, post_pts, post_a;   // scratch of the pre/post steps (ss_post.cuh)
    size_t keys_clean = 0;              // leading entries of `keys` known to hold KEY_NONE (k_finalize resets what it reads)
    // cached proximity table key
    int prox_win = -1;
    double prox_gp = -1;
    // instrumentation
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    double agg_ms_done = 0;
    long long agg_launches = 0, total_launches = 0;
    int sm_count = 148;
    int blocks_per_sm = 1;           // of the aggregation kernel about to be launched (launch_agg)
    int last_kernel = 0;             // aggregation kernel of the last call: 1 = k_aggregate_tc, 2 = k_aggregate_ws (0: none yet)
    int last_dc = 0;                 // and its disparity chunk
    int smem_attr_ws[24] = {};       // largest dynamic-smem opt-in set so far, per k_aggregate_ws instantiation
    int smem_attr_tc[8] = {};        // same for k_aggregate_tc
};

constexpr int SS_MAX_DEVICES = 64;
Ctx g_ctxs[SS_MAX_DEVICES];
std::mutex g_cfg_mu;
std::vector<int> g_devices;          // devices the host entry points run on (ss_init / ss_init_devices); empty: current device
bool g_profile = false;

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
    if (b.p) cudaFree(b.p);          // cudaFree synchronises the device: nothing in flight still reads the old buffer
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

void ctx_release(Ctx &c) {
    if (!c.ready) return;
    cudaSetDevice(c.device);
    DevBuf *bufs[] = {&c.img1, &c.img2, &c.out, &c.f1, &c.f2, &c.evol, &c.vol0, &c.vol1, &c.keys,
                      &c.prox, &c.stage_l, &c.stage_r, &c.stage_i, &c.dense, &c.post_mm, &c.post_pts, &c.post_a};
    for (DevBuf *b : bufs) { if (b->p) cudaFree(b->p); *b = DevBuf(); }
    for (auto &ev : c.events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    c.events.clear();
    if (c.done) cudaEventDestroy(c.done);
    c.done = nullptr;
    c.done_valid = false;
    if (c.stream) cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    c.prox_win = -1;
    c.keys_clean = 0;
    c.ready = false;
    memset(c.smem_attr_ws, 0, sizeof(c.smem_attr_ws));
    memset(c.smem_attr_tc, 0, sizeof(c.smem_attr_tc));
}

// c.mu held.  Leaves `device` current.
int ctx_init(Ctx &c, int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SS_ERR_CUDA, std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                     " (libsspassive has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(SS_ERR_CUDA, "device index out of range");
    CU_TRY(cudaSetDevice(device));
    if (c.ready) return SS_OK;
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(SS_ERR_CUDA, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                     "; libsspassive is built for sm_100a only");
    c.sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    float lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = srgb_linear100(i);
    CU_TRY(cudaMemcpyToSymbol(c_lin100, lut, sizeof(lut)));
    CU_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
    c.device = device;
    c.profile = g_profile;
    c.ready = true;
    return SS_OK;
}

// The device the host entry points use when no device list was configured: the caller's current device.
int default_device() {
    {
        std::lock_guard<std::mutex> lk(g_cfg_mu);
        if (!g_devices.empty()) return g_devices[0];
    }
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; }
    return d;
}
int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; }
    return d;
}

// RAII: the context of `device`, locked, initialised, its device current.
struct CtxLock {
    Ctx *c = nullptr;
    std::unique_lock<std::mutex> lk;
    int rc = SS_OK;
    explicit CtxLock(int device) {
        if (device < 0 || device >= SS_MAX_DEVICES) { rc = fail(SS_ERR_CUDA, "device index out of range"); return; }
        c = &g_ctxs[device];
        lk = std::unique_lock<std::mutex>(c->mu);
        rc = ctx_init(*c, device);
    }
};

// Cross-stream ordering on the cached scratch (see Ctx::done).
int scratch_begin(Ctx &c, cudaStream_t st) {
    if (c.done_valid) CU_TRY(cudaStreamWaitEvent(st, c.done, 0));
    return SS_OK;
}
int scratch_end(Ctx &c, cudaStream_t st) {
    CU_TRY(cudaEventRecord(c.done, st));
    c.done_valid = true;
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

// true when the warp-specialised kernel can stage this window in shared memory
bool ws_fits(int win, int DC, bool gsw) { return ws_smem(win, DC, gsw).total <= 227 * 1024; }

// ASW windows whose tensor-core denominators stay inside the 5e-5 cost tolerance BY CONSTRUCTION: the truncating TMEM
// accumulator loses at most n * 2^-23 of the sum over n MMAs, the epilogue centres that interval (ss_aggregate_tc.cuh), and
// n * 2^-24 = 3 * ceil(win/8) * win * 2^-24 <= 4.4e-5 up to win 41.  Larger windows run the all-CUDA-core kernel.
constexpr int TC_MAX_WIN = 41;
bool tc_enabled() {             // SS_TCDEN=0 forces the all-CUDA-core kernel (read per call: the tests flip it)
    const char *e = getenv("SS_TCDEN");
    return !(e && atoi(e) == 0);
}

// The kernel family and the disparity chunk DC are chosen from the CALL's range [minD, maxD], never from the evaluated
// sub-range: a disparity shard (ss_asw_partial_device) runs the same kernel on the same chunk grid -- chunks start at
// minD + k * DC -- as the unsharded call, so its costs are bit-identical and merged shards reproduce the unsharded map.
struct Plan {
    bool tc;     // k_aggregate_tc (ASW, 128-disparity chunks, win <= TC_MAX_WIN); else k_aggregate_ws
    int DC;      // 0: nothing fits
};
Plan make_plan(const Call &q) {
    const int D = q.maxD - q.minD + 1;
    Plan p;
    p.DC = D <= 32 ? 32 : (D <= 64 ? 64 : 128);
    p.tc = !q.gsw && p.DC == 128 && q.win <= TC_MAX_WIN && tc_enabled() && tc_smem(q.win, q.win > 39).total <= 227 * 1024;
    if (!p.tc) {
        // windows whose tiles do not fit at this chunk size run narrower chunks (several chunks merge through the atomicMin keys)
        while (p.DC > 32 && !ws_fits(q.win, p.DC, q.gsw)) p.DC /= 2;
        if (!ws_fits(q.win, p.DC, q.gsw)) p.DC = 0;
    }
    return p;
}

Geom make_geom(const Call &q, int DC, bool tc) {
    Geom g;
    g.W = q.W; g.H = q.H; g.win = q.win; g.pad = q.win / 2;
    g.minD = q.minD; g.maxD = q.maxD;
    g.DC = DC;
    const int ch0 = (q.dBegin - q.minD) / DC;               // chunk grid anchored at minD
    g.dLo = q.minD + ch0 * DC;
    g.dVLo = q.dBegin;
    g.dHi = q.dEnd;
    g.nch = (g.dHi - g.dLo) / DC + 1;
    g.row0 = q.row0; g.row1 = q.row1;
    g.erow0 = q.row0 - g.pad < 0 ? 0 : q.row0 - g.pad;
    g.erow1 = q.row1 + g.pad > q.H ? q.H : q.row1 + g.pad;
    g.T = q.gsw ? TILE_X : TILE_WS;
    g.EP = tc ? TC_EP : g.DC + 4;
    g.EG = tc ? TC_EG : 0;
    g.ntx = (q.W + g.T - 1) / g.T;
    g.UW = (g.ntx * g.T + g.win - 1 + 3) & ~3;
    g.EPL = (g.UW * g.EP + ((g.UW + 7) >> 3) * g.EG + 15) & ~15;
    g.PL2 = g.dLo + g.nch * g.DC - 1 + g.pad;
    g.VW = g.ntx * g.T + g.pad + g.PL2;
    g.NU = g.T + g.win - 1;
    g.NR = g.T + g.DC - 1;
    g.NRp = g.T + g.DC;
    g.NV = g.NR + g.win - 1;
    return g;
}

// One aggregation pass = one launch over every (chunk, row, x tile) block -- or two when the last wave would leave most SMs
// idle.  A block occupies a whole SM (shared memory, tensor memory), so `total` blocks take ceil(total / SMs) rounds; when
// the last round holds fewer than a third of the SMs' worth of blocks (8-way row stripes of a KITTI frame: 611 blocks =
// 4.13 rounds of 148) those tail tiles are launched separately with nsub = T/32: each block then handles ONE 32-column block
// of its tile (one consumer warp per scheduler instead of three, 6 of the 10 weight column blocks), so the tail costs about
// half a round.  Every (x, d) pair is still computed by the same instruction sequence: the maps do not change by a bit.
template <typename K>
int launch_agg(Ctx &c, K kernel, int &attr, int smem, AggParams P, int threads, cudaStream_t st) {
    if (smem > 227 * 1024) return fail(SS_ERR_PARAM, "winSize too large for the shared-memory tiling of the aggregation kernel");
    if (attr < smem) {
        CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = smem;
    }
    const int total = P.g.ntx * (P.g.row1 - P.g.row0) * P.g.nch;
    const int xb = P.g.T / 32;
    int tail = c.blocks_per_sm == 1 ? total % c.sm_count : 0;       // only the one-block-per-SM kernels are wave-quantised this way
    static const int split_off = [] { const char *e = getenv("SS_NO_TAIL_SPLIT"); return e ? atoi(e) : 0; }();
    if (split_off || tail * xb > c.sm_count || total < c.sm_count) tail = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c.profile) {
        CU_TRY(cudaEventCreate(&e0));
        CU_TRY(cudaEventCreate(&e1));
        CU_TRY(cudaEventRecord(e0, st));
    }
    P.g.tile0 = 0;
    P.g.nsub = 1;
    kernel<<<dim3(total - tail, 1, 1), threads, smem, st>>>(P);
    CU_TRY(cudaGetLastError());
    c.total_launches++;
    if (tail) {
        P.g.tile0 = total - tail;
        P.g.nsub = xb;
        kernel<<<dim3(tail, xb, 1), threads, smem, st>>>(P);
        CU_TRY(cudaGetLastError());
        c.total_launches++;
    }
    if (c.profile) {
        CU_TRY(cudaEventRecord(e1, st));
        c.events.emplace_back(e0, e1);
    }
    c.agg_launches++;
    return SS_OK;
}

template <bool GSW, int DC, int REM>
int launch_ws_rem(Ctx &c, const AggParams &P, cudaStream_t st) {
    const int di = ((DC == 128 ? 2 : (DC == 64 ? 1 : 0)) * 4 + REM / 2) * 2 + (GSW ? 1 : 0);
    const int smem = ws_smem(P.g.win, DC, GSW).total;
    // blocks per SM: register budget (WsCfg::MINB) and shared memory
    c.blocks_per_sm = std::max(1, std::min((int)WsCfg<GSW, DC>::MINB, (227 * 1024) / (smem + 1024)));
    return launch_agg(c, k_aggregate_ws<GSW, DC, REM>, c.smem_attr_ws[di], ws_smem(P.g.win, DC, GSW).total, P, WsCfg<GSW, DC>::NT, st);
}
template <bool GSW, int DC>
int launch_ws_dc(Ctx &c, const AggParams &P, cudaStream_t st) {
    switch (P.g.win & 7) {          // win is odd
        case 1: return launch_ws_rem<GSW, DC, 1>(c, P, st);
        case 3: return launch_ws_rem<GSW, DC, 3>(c, P, st);
        case 5: return launch_ws_rem<GSW, DC, 5>(c, P, st);
        default: return launch_ws_rem<GSW, DC, 7>(c, P, st);
    }
}
template <bool GSW>
int launch_ws(Ctx &c, const AggParams &P, cudaStream_t st) {
    if (P.g.DC == 128) return launch_ws_dc<GSW, 128>(c, P, st);
    if (P.g.DC == 64) return launch_ws_dc<GSW, 64>(c, P, st);
    return launch_ws_dc<GSW, 32>(c, P, st);
}

// ASW, 128-disparity chunks: denominators on the tensor cores (ss_aggregate_tc.cuh).  win <= 39: two stages of operands in
// tensor memory; 39 < win: one stage (SINGLE).
template <int REM, bool SINGLE>
int launch_tc_rem(Ctx &c, const AggParams &P, cudaStream_t st) {
    c.blocks_per_sm = 1;
    return launch_agg(c, k_aggregate_tc<REM, SINGLE>, c.smem_attr_tc[(REM / 2) * 2 + (SINGLE ? 1 : 0)], tc_smem(P.g.win, SINGLE).total, P,
                      TC_THREADS, st);
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
struct Outputs {
    int16_t *d_final = nullptr;     // [(rows)*W]
    int16_t *d_left = nullptr, *d_right = nullptr;
    uint8_t *d_invalid = nullptr;
    u64 *d_keysL = nullptr, *d_keysR = nullptr;   // when set, stop after WTA (partial / sharded call)
    bool want_vol0 = false, want_vol1 = false;    // keep aggregated volumes (debug export)
};

int finalize_launch(Ctx &c, u64 *keysL, u64 *keysR, int W, int rows, const Outputs &o, int reset, cudaStream_t st) {
    const size_t sh = (size_t)W * 3 + 16;
    if (sh > 48 * 1024) CU_TRY(cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    k_finalize<<<rows, 128, sh, st>>>(keysL, keysR, W, o.d_final, o.d_left, o.d_right, o.d_invalid, reset);
    CU_TRY(cudaGetLastError());
    c.total_launches += 1;
    return SS_OK;
}

// Enqueue the whole pipeline for device-resident inputs (c.mu held, c.device current).
int run_device(Ctx &c, const Call &q, const uint8_t *d_img1, const uint8_t *d_img2, const Outputs &o, cudaStream_t st) {
    const int rows = q.row1 - q.row0;
    if (rows == 0) return SS_OK;
    int rc;
    if ((rc = scratch_begin(c, st))) return rc;
    const bool partial = o.d_keysL != nullptr;
    const bool need_right = q.consistent != 0;
    u64 *keysL = o.d_keysL, *keysR = o.d_keysR;
    const size_t npx = (size_t)rows * q.W;
    if (!partial) {
        // one buffer [left | right]; k_finalize leaves every key it read at KEY_NONE, so a steady stream of calls needs no memset
        const size_t nkeys = need_right ? 2 * npx : npx;
        const size_t cap0 = c.keys.cap;
        if ((rc = ensure(c.keys, nkeys * 8))) return rc;
        if (c.keys.cap != cap0) c.keys_clean = 0;
        keysL = (u64 *)c.keys.p;
        keysR = need_right ? keysL + npx : nullptr;
        if (c.keys_clean < nkeys) CU_TRY(cudaMemsetAsync(keysL, 0xff, nkeys * 8, st));
        c.keys_clean = 0;
    } else {
        CU_TRY(cudaMemsetAsync(keysL, 0xff, npx * 8, st));
        if (need_right && keysR) CU_TRY(cudaMemsetAsync(keysR, 0xff, npx * 8, st));
    }

    const int dB = q.dBegin < q.minD ? q.minD : q.dBegin;
    const int dE = q.dEnd > q.maxD ? q.maxD : q.dEnd;
    if (dE >= dB) {
        const Plan plan = make_plan(q);
        if (plan.DC == 0) return fail(SS_ERR_PARAM, "winSize too large for the shared-memory tiling of the aggregation kernel");
        Call qq = q;
        qq.dBegin = dB;
        qq.dEnd = dE;
        const Geom g = make_geom(qq, plan.DC, plan.tc);
        const int erows = g.erow1 - g.erow0;
        if ((rc = ensure(c.f1, (size_t)erows * g.UW * 16))) return rc;
        if ((rc = ensure(c.f2, (size_t)erows * g.VW * 16))) return rc;
        if ((rc = ensure(c.evol, q.gsw ? (size_t)g.nch * erows * g.UW * g.DC * 4
                                       : (size_t)g.nch * erows * g.EPL + 256))) return rc;
        const int Dp = g.nch * g.DC;
        if (o.want_vol0 && (rc = ensure(c.vol0, npx * Dp * 4))) return rc;
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
        P.bestR = need_right ? keysR : nullptr;
        P.vol0 = o.want_vol0 ? (float *)c.vol0.p : nullptr;
        P.vol1 = o.want_vol1 ? (float *)c.vol1.p : nullptr;
        P.Dp = Dp;
        P.vol_export = (o.want_vol0 || o.want_vol1) ? 1 : 0;
        {
            // timing experiments (results are garbage): 1 k_aggregate_ws consumers free-running, producers idle;
            // k_aggregate_tc: 2 producers tabulate nothing, 4 consumers accumulate nothing, 8 no tcgen05.mma
            const char *e = getenv("SS_FREERUN");
            P.freerun = e ? atoi(e) : 0;
        }
        if (q.gsw) rc = launch_ws<true>(c, P, st);
        else if (plan.tc) rc = launch_tc(c, P, st);
        else rc = launch_ws<false>(c, P, st);
        if (rc) return rc;
        c.last_kernel = (!q.gsw && plan.tc) ? 1 : 2;
        c.last_dc = g.DC;
    }
    if (!partial) {
        if ((rc = finalize_launch(c, keysL, keysR, q.W, rows, o, 1, st))) return rc;
        c.keys_clean = need_right ? 2 * npx : npx;
    }
    return scratch_end(c, st);
}

// host wrapper on one device: H2D, run, D2H of the stripe (c.mu held, c.device current)
int run_host(Ctx &c, const Call &q, const uint8_t *img1, const uint8_t *img2, int16_t *out, int16_t *out_left, int16_t *out_right,
             uint8_t *out_invalid, float *out_vol0, float *out_vol1) {
    int rc;
    const int rows = q.row1 - q.row0;
    const size_t nimg = (size_t)q.W * q.H * 3, npx = (size_t)rows * q.W;
    cudaStream_t st = c.stream;
    if ((rc = scratch_begin(c, st))) return rc;
    if ((rc = ensure(c.img1, nimg))) return rc;
    if ((rc = ensure(c.img2, nimg))) return rc;
    if ((rc = ensure(c.out, npx * 2 + 2))) return rc;
    // only the rows the stripe needs travel
    const int pad = q.win / 2;
    const int er0 = q.row0 - pad < 0 ? 0 : q.row0 - pad, er1 = q.row1 + pad > q.H ? q.H : q.row1 + pad;
    const size_t off = (size_t)er0 * q.W * 3, len = (size_t)(er1 - er0) * q.W * 3;
    if (len && rows) {
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
        const int dc = make_plan(q).DC;
        const int Dp = ((D + dc - 1) / dc) * dc;
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
    if ((rc = scratch_end(c, st))) return rc;
    CU_TRY(cudaStreamSynchronize(st));
    return SS_OK;
}

// Host entry: one device, or -- when ss_init_devices configured several -- image-row stripes over all of them, one host
// thread per device (rows are independent jobs in the reference: a row index is what its thread pool pops,
// _passive.cpp:372-374, and the L-R check + fill are row-local, :251-285).  Every device reads the rows it needs (stripe +-
// win/2) straight from the caller's arrays and writes its stripe straight into the caller's output: no collective.
int host_entry(const Call &q, const uint8_t *img1, const uint8_t *img2, int16_t *out, int16_t *out_left = nullptr,
               int16_t *out_right = nullptr, uint8_t *out_invalid = nullptr, float *out_vol0 = nullptr, float *out_vol1 = nullptr) {
    if (!img1 || !img2) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = validate(q);
    if (rc) return rc;
    std::vector<int> devs;
    {
        std::lock_guard<std::mutex> lk(g_cfg_mu);
        devs = g_devices;
    }
    const int rows = q.row1 - q.row0;
    const bool staged = out_left || out_right || out_invalid || out_vol0 || out_vol1;
    const int n = (int)devs.size();
    if (n <= 1 || staged || rows < 2 * n) {
        CtxLock L(n ? devs[0] : default_device());
        if (L.rc) return L.rc;
        return run_host(*L.c, q, img1, img2, out, out_left, out_right, out_invalid, out_vol0, out_vol1);
    }
    const int S = (rows + n - 1) / n;
    std::vector<int> rcs(n, SS_OK);
    std::vector<std::string> errs(n);
    std::vector<std::thread> th;
    for (int k = 0; k < n; ++k) {
        th.emplace_back([&, k]() {
            Call qk = q;
            qk.row0 = std::min(q.row0 + k * S, q.row1);
            qk.row1 = std::min(q.row0 + (k + 1) * S, q.row1);
            if (qk.row0 >= qk.row1) return;
            CtxLock L(devs[k]);
            int r = L.rc;
            if (!r) r = run_host(*L.c, qk, img1, img2, out + (size_t)(qk.row0 - q.row0) * q.W, nullptr, nullptr, nullptr, nullptr, nullptr);
            rcs[k] = r;
            if (r) errs[k] = t_err;
        });
    }
    for (auto &t : th) t.join();
    for (int k = 0; k < n; ++k)
        if (rcs[k]) return fail(rcs[k], "device " + std::to_string(devs[k]) + ": " + errs[k]);
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

// device-resident entry: the pointers live on the caller's CURRENT device; work is enqueued on the caller's stream
int device_entry(const Call &q, const uint8_t *d1, const uint8_t *d2, const Outputs &o, void *stream) {
    if (!d1 || !d2) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = validate(q);
    if (rc) return rc;
    CtxLock L(current_device());
    if (L.rc) return L.rc;
    return run_device(*L.c, q, d1, d2, o, (cudaStream_t)stream);
}

// ---- NCCL, loaded at run time (the library must load on machines without it) ---------------------------------------------
// Single-process multi-GPU: ncclCommInitAll over the ss_init_devices list, one grouped in-place ncclAllGather per call.
struct Nccl {
    void *lib = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::vector<void *> comms;
    std::vector<int> devs;
    std::string err;
} g_nccl;

void nccl_teardown() {
    if (g_nccl.CommDestroy)
        for (void *cm : g_nccl.comms) g_nccl.CommDestroy(cm);
    g_nccl.comms.clear();
    g_nccl.devs.clear();
}

// g_cfg_mu held
int nccl_setup(const std::vector<int> &devs) {
    nccl_teardown();
    g_nccl.err.clear();
    if (devs.size() < 2) return SS_OK;
    if (!g_nccl.lib) {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            g_nccl.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.lib) break;
        }
        if (!g_nccl.lib) { g_nccl.err = "libnccl.so.2 not found"; return SS_ERR_CUDA; }
        g_nccl.CommInitAll = (int (*)(void **, int, const int *))dlsym(g_nccl.lib, "ncclCommInitAll");
        g_nccl.CommDestroy = (int (*)(void *))dlsym(g_nccl.lib, "ncclCommDestroy");
        g_nccl.GroupStart = (int (*)())dlsym(g_nccl.lib, "ncclGroupStart");
        g_nccl.GroupEnd = (int (*)())dlsym(g_nccl.lib, "ncclGroupEnd");
        g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
        g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
        if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.AllGather) {
            g_nccl.err = "libnccl.so.2 lacks the expected symbols";
            return SS_ERR_CUDA;
        }
    }
    g_nccl.comms.assign(devs.size(), nullptr);
    const int r = g_nccl.CommInitAll(g_nccl.comms.data(), (int)devs.size(), devs.data());
    if (r != 0) {
        g_nccl.err = std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "failed");
        g_nccl.comms.clear();
        return SS_ERR_CUDA;
    }
    g_nccl.devs = devs;
    return SS_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------

extern "C" {

int ss_abi_version(void) { return 2; }

const char *ss_last_error(void) { return t_err.c_str(); }

int ss_init_devices(const int *devices, int n) {
    std::vector<int> devs;
    if (!devices || n <= 0) {
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
            cudaGetLastError();
            return fail(SS_ERR_CUDA, "no usable CUDA device (libsspassive has no CPU fallback)");
        }
        for (int k = 0; k < cnt; ++k) devs.push_back(k);
    } else {
        devs.assign(devices, devices + n);
    }
    if (devs.size() > (size_t)SS_MAX_DEVICES) return fail(SS_ERR_PARAM, "too many devices");
    for (size_t a = 0; a < devs.size(); ++a)
        for (size_t b = a + 1; b < devs.size(); ++b)
            if (devs[a] == devs[b]) return fail(SS_ERR_PARAM, "duplicate device in the device list");
    for (int d : devs) {
        CtxLock L(d);
        if (L.rc) return L.rc;
    }
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    g_devices = devs;                               // the NCCL communicators are created by the first multi_device call
    cudaSetDevice(devs[0]);
    return SS_OK;
}

int ss_init(int device) {
    if (device < 0) device = default_device();
    return ss_init_devices(&device, 1);
}

int ss_device_count(void) {
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    return (int)g_devices.size();
}

int ss_shutdown(void) {
    {
        std::lock_guard<std::mutex> lk(g_cfg_mu);
        nccl_teardown();
        g_devices.clear();
    }
    for (Ctx &c : g_ctxs) {
        std::lock_guard<std::mutex> lk(c.mu);
        ctx_release(c);
    }
    return SS_OK;
}

int ss_asw_compute(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                   int min_disp, double gamma_c, double gamma_p, int consistent, int16_t *out_disp) {
    if (!out_disp) return fail(SS_ERR_FORMAT, "Invalid input format!");
    return host_entry(asw_call(width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, 0, height), img1, img2, out_disp);
}

int ss_gsw_compute(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                   int min_disp, int gamma, float f_max, int iterations, int bins, int16_t *out_disp) {
    (void)bins;
    if (!out_disp) return fail(SS_ERR_FORMAT, "Invalid input format!");
    return host_entry(gsw_call(width, height, win_size, max_disp, min_disp, gamma, f_max, iterations, 0, height), img1, img2, out_disp);
}

int ss_asw_compute_rows(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                        int min_disp, double gamma_c, double gamma_p, int consistent, int row_begin, int row_end,
                        int16_t *out_rows) {
    if (!out_rows) return fail(SS_ERR_FORMAT, "Invalid input format!");
    return host_entry(asw_call(width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, row_begin, row_end), img1,
                      img2, out_rows);
}

int ss_gsw_compute_rows(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                        int min_disp, int gamma, float f_max, int iterations, int bins, int row_begin, int row_end,
                        int16_t *out_rows) {
    (void)bins;
    if (!out_rows) return fail(SS_ERR_FORMAT, "Invalid input format!");
    return host_entry(gsw_call(width, height, win_size, max_disp, min_disp, gamma, f_max, iterations, row_begin, row_end), img1, img2,
                      out_rows);
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

// Single-process multi-GPU, device-resident: device k of the ss_init_devices list computes its row stripe into its slot of
// d_out[k] and ONE grouped, in-place ncclAllGather leaves the whole map on every device (north_star: "a single NCCL
// all-gather over NVLink to reassemble the final disparity map").
static int multi_device(const Call &q0, const uint8_t *const *d_img1, const uint8_t *const *d_img2, int16_t *const *d_out,
                        void *const *streams) {
    if (!d_img1 || !d_img2 || !d_out) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = validate(q0);
    if (rc) return rc;
    std::vector<int> devs;
    std::vector<void *> comms;
    {
        std::lock_guard<std::mutex> lk(g_cfg_mu);
        devs = g_devices;
        if (devs.size() > 1 && g_nccl.devs != devs && nccl_setup(devs) != SS_OK)
            return fail(SS_ERR_CUDA, "NCCL communicators are not available: " + g_nccl.err);
        comms = g_nccl.comms;
    }
    const int n = (int)devs.size();
    if (n < 1) return fail(SS_ERR_PARAM, "call ss_init_devices first");
    const int S = (q0.H + n - 1) / n;
    for (int k = 0; k < n; ++k) {
        if (!d_img1[k] || !d_img2[k] || !d_out[k]) return fail(SS_ERR_FORMAT, "Invalid input format!");
        Call q = q0;
        q.row0 = std::min(k * S, q0.H);
        q.row1 = std::min((k + 1) * S, q0.H);
        CtxLock L(devs[k]);
        if (L.rc) return L.rc;
        Outputs o;
        o.d_final = d_out[k] + (size_t)k * S * q0.W;
        if ((rc = run_device(*L.c, q, d_img1[k], d_img2[k], o, streams ? (cudaStream_t)streams[k] : L.c->stream))) return rc;
    }
    if (n > 1) {
        int r = g_nccl.GroupStart();
        for (int k = 0; k < n && r == 0; ++k) {
            cudaSetDevice(devs[k]);
            // int16 is not an NCCL type: gather the stripes as bytes (ncclUint8 = 1); in place: send = recv + rank * count
            r = g_nccl.AllGather(d_out[k] + (size_t)k * S * q0.W, d_out[k], (size_t)S * q0.W * 2, 1, comms[k],
                                 streams ? (cudaStream_t)streams[k] : g_ctxs[devs[k]].stream);
        }
        const int r2 = g_nccl.GroupEnd();
        if (r == 0) r = r2;
        if (r != 0) return fail(SS_ERR_CUDA, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "failed"));
    }
    if (!streams)
        for (int k = 0; k < n; ++k) {
            CU_TRY(cudaSetDevice(devs[k]));
            CU_TRY(cudaStreamSynchronize(g_ctxs[devs[k]].stream));
        }
    cudaSetDevice(devs[0]);
    return SS_OK;
}

int ss_asw_compute_multi_device(const uint8_t *const *d_img1, const uint8_t *const *d_img2, int width, int height, int win_size,
                                int max_disp, int min_disp, double gamma_c, double gamma_p, int consistent, int16_t *const *d_out,
                                void *const *streams) {
    return multi_device(asw_call(width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, 0, height), d_img1, d_img2,
                        d_out, streams);
}

int ss_gsw_compute_multi_device(const uint8_t *const *d_img1, const uint8_t *const *d_img2, int width, int height, int win_size,
                                int max_disp, int min_disp, int gamma, float f_max, int iterations, int bins, int16_t *const *d_out,
                                void *const *streams) {
    (void)bins;
    return multi_device(gsw_call(width, height, win_size, max_disp, min_disp, gamma, f_max, iterations, 0, height), d_img1, d_img2, d_out,
                        streams);
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
    CtxLock L(current_device());
    if (L.rc) return L.rc;
    if (n == 0 || n_shards == 1) return SS_OK;
    k_merge_keys<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((u64 *)d_keys, n_shards, n);
    CU_TRY(cudaGetLastError());
    L.c->total_launches += 1;
    return SS_OK;
}

int ss_finalize_keys_device(const uint64_t *d_best_left, const uint64_t *d_best_right, int width, int rows,
                            int min_disp, int16_t *d_out_rows, void *stream) {
    (void)min_disp;
    if (!d_best_left || !d_out_rows || width <= 0 || rows < 0) return fail(SS_ERR_FORMAT, "Invalid input format!");
    CtxLock L(current_device());
    if (L.rc) return L.rc;
    if (rows == 0) return SS_OK;
    Outputs o;
    o.d_final = d_out_rows;
    return finalize_launch(*L.c, (u64 *)d_best_left, (u64 *)d_best_right, width, rows, o, 0, (cudaStream_t)stream);
}

int ss_asw_stages_rows(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                       int min_disp, double gamma_c, double gamma_p, int consistent, int row_begin, int row_end,
                       int16_t *out_left, int16_t *out_right, uint8_t *out_invalid, int16_t *out_final, float *out_cost) {
    return host_entry(asw_call(width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, row_begin, row_end), img1,
                      img2, out_final, out_left, out_right, out_invalid, out_cost, nullptr);
}

int ss_asw_stages(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                  int min_disp, double gamma_c, double gamma_p, int consistent, int16_t *out_left, int16_t *out_right,
                  uint8_t *out_invalid, int16_t *out_final, float *out_cost) {
    return ss_asw_stages_rows(img1, img2, width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, 0, height,
                              out_left, out_right, out_invalid, out_final, out_cost);
}

int ss_gsw_stages_rows(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                       int min_disp, int gamma, float f_max, int iterations, int bins, int row_begin, int row_end,
                       int16_t *out_left, int16_t *out_right, uint8_t *out_invalid, int16_t *out_final, float *out_cost_left,
                       float *out_cost_right) {
    (void)bins;
    return host_entry(gsw_call(width, height, win_size, max_disp, min_disp, gamma, f_max, iterations, row_begin, row_end), img1, img2,
                      out_final, out_left, out_right, out_invalid, out_cost_right, out_cost_left);
}

int ss_gsw_stages(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                  int min_disp, int gamma, float f_max, int iterations, int bins, int16_t *out_left, int16_t *out_right,
                  uint8_t *out_invalid, int16_t *out_final, float *out_cost_left, float *out_cost_right) {
    return ss_gsw_stages_rows(img1, img2, width, height, win_size, max_disp, min_disp, gamma, f_max, iterations, bins, 0, height,
                              out_left, out_right, out_invalid, out_final, out_cost_left, out_cost_right);
}

// BGR -> CIELab of the reference (colorconversion.hpp:18-86) as the kernels see it: float32 L, a, b per pixel
int ss_debug_lab(const uint8_t *img, int width, int height, float *out_lab) {
    if (!img || !out_lab || width <= 0 || height <= 0) return fail(SS_ERR_FORMAT, "Invalid input format!");
    CtxLock L(default_device());
    if (L.rc) return L.rc;
    Ctx &c = *L.c;
    int rc;
    const size_t npx = (size_t)width * height;
    cudaStream_t st = c.stream;
    if ((rc = scratch_begin(c, st))) return rc;
    if ((rc = ensure(c.img1, npx * 3))) return rc;
    if ((rc = ensure(c.f1, npx * 16))) return rc;
    if ((rc = ensure(c.dense, npx * 12))) return rc;
    CU_TRY(cudaMemcpyAsync(c.img1.p, img, npx * 3, cudaMemcpyHostToDevice, st));
    dim3 b(128), g((width + 127) / 128, height);
    k_prep_features<false><<<g, b, 0, st>>>((const uint8_t *)c.img1.p, (float4 *)c.f1.p, width, 0, height, width, 0);
    k_compact_volume<<<(unsigned)((npx * 3 + 255) / 256), 256, 0, st>>>((const float *)c.f1.p, (float *)c.dense.p, (long long)npx, 3, 4);
    CU_TRY(cudaGetLastError());
    c.total_launches += 2;
    CU_TRY(cudaMemcpyAsync(out_lab, c.dense.p, npx * 12, cudaMemcpyDeviceToHost, st));
    if ((rc = scratch_end(c, st))) return rc;
    CU_TRY(cudaStreamSynchronize(st));
    return SS_OK;
}

// which aggregation kernel served the last call on `device` (< 0: the default device of the host entry points):
// 1 = k_aggregate_tc, 2 = k_aggregate_ws, 0 = none; *disp_chunk receives its disparity chunk
int ss_debug_last_kernel(int device, int *disp_chunk) {
    if (device < 0) device = default_device();
    if (device >= SS_MAX_DEVICES) return 0;
    Ctx &c = g_ctxs[device];
    std::lock_guard<std::mutex> lk(c.mu);
    if (disp_chunk) *disp_chunk = c.last_dc;
    return c.last_kernel;
}

int ss_profile_enable(int on) {
    g_profile = on != 0;
    for (Ctx &c : g_ctxs) {
        std::lock_guard<std::mutex> lk(c.mu);
        c.profile = g_profile;
    }
    return SS_OK;
}

int ss_profile_reset(void) {
    for (Ctx &c : g_ctxs) {
        std::lock_guard<std::mutex> lk(c.mu);
        for (auto &ev : c.events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
        c.events.clear();
        c.agg_ms_done = 0;
        c.agg_launches = 0;
        c.total_launches = 0;
    }
    return SS_OK;
}

// sums over every device this process has used
int ss_profile_read(double *agg_ms, long long *agg_launches, long long *total_launches) {
    double ms_sum = 0;
    long long nl = 0, nt = 0;
    for (Ctx &c : g_ctxs) {
        std::lock_guard<std::mutex> lk(c.mu);
        if (!c.ready) continue;
        for (auto &ev : c.events) {
            CU_TRY(cudaEventSynchronize(ev.second));
            float ms = 0.f;
            CU_TRY(cudaEventElapsedTime(&ms, ev.first, ev.second));
            c.agg_ms_done += ms;
            cudaEventDestroy(ev.first);
            cudaEventDestroy(ev.second);
        }
        c.events.clear();
        ms_sum += c.agg_ms_done;
        nl += c.agg_launches;
        nt += c.total_launches;
    }
    if (agg_ms) *agg_ms = ms_sum;
    if (agg_launches) *agg_launches = nl;
    if (total_launches) *total_launches = nt;
    return SS_OK;
}

int ss_measure_fp32_peak(double *tflops, void *stream) {
    if (!tflops) return fail(SS_ERR_FORMAT, "Invalid input format!");
    CtxLock L(current_device());
    if (L.rc) return L.rc;
    Ctx &c = *L.c;
    int rc;
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, c.device));
    const int blocks = prop.multiProcessorCount * 8, tpb = 256, iters = 8192;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = scratch_begin(c, st))) return rc;
    if ((rc = ensure(c.dense, (size_t)blocks * tpb * 4))) return rc;
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
    return scratch_end(c, st);
}


}  // extern "C"

#include "ss_post.cuh"
