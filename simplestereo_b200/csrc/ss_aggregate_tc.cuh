// ss_aggregate_tc.cuh -- k_aggregate_tc: the ASW aggregation kernel with the DENOMINATORS on the tensor cores.
// Included by ss_passive.cu after k_aggregate_ws (shares its helpers).  ASW only, DC = 128, T = 96, win <= 79.
//
// den[x][d] = sum_q w1[x][q] * w2[x-d][q] is a banded GEMM: per window row, D[r][x] += W2[r][j] * W1[x][j]
// with r the reversed right-centre index (den[x][d] = D[T-1-x+d][x]).  The numerator carries a per-(x,d,q) factor and
// stays on the CUDA cores, but its loop drops from 3 packed instructions per 2 elements (mul, fma, add) to 2.
//
//   * precision: 3xTF32.  Every weight w is split into hi = w (the MMA reads the top 19 bits) and lo = w - trunc19(w);
//     D += hi*hi + hi*lo + lo*hi.  Measured against float64 (tools/umma_den_test.cu): max relative error 7e-6, inside the
//     5e-5 cost tolerance the tests state (the float32 CUDA-core sum is ~1e-6).
//   * A = W2 (M = r: two halves of 128 rows) lives in TENSOR MEMORY: the producer lane that computed the weights of
//     right column r writes them with tcgen05.st (lane = r, columns = j).  Producer warp q owns TMEM lanes [32q, 32q+32),
//     so right-column block cb is tabulated by producer cb % 4.
//   * B = W1 (N = x = 96) lives in shared memory, K-major, no swizzle, 8-row groups at a padded stride of 144 bytes
//     (any multiple of 16 works; 144 keeps the consumers' reads of four row groups on distinct banks).  The SAME array
//     serves the consumers (one LDS.128 = one column's weights for 4 consecutive window offsets).
//   * the accumulator (2 x 128 lanes x 96 columns) stays in TMEM for the whole block; after the last window row the
//     producers read it back (tcgen05.ld) into shared memory in [r][x] order and the consumers pick their 32 values.
//   * the tensor core TRUNCATES on every accumulation: each tcgen05.mma loses at most one ulp of the running sum, so after
//     n accumulating MMAs the raw denominator is low by a relative 0 .. n * 2^-23 (measured against float64,
//     profiles/r01d_tc_denominator_error.txt: 0.9e-5 .. 4.4e-5 after the 525 MMAs of a 35x35 window, mean 0.83 * 2^-24 per
//     MMA).  The epilogue multiplies by 1 + n * 2^-24, the CENTRE of that interval, so |error| <= n * 2^-24 for ANY
//     input -- a bound, not a fit: 3.1e-5 at win 35 (n = 3 * 5 K-groups * 35 rows), 4.4e-5 at win 41 -- inside the 5e-5 cost
//     tolerance the tests state.  n counts only MMAs that can add something to this (x, d): window rows inside the image
//     and K-groups that intersect the pair's valid window columns (adding exact zeros truncates nothing), so border pixels
//     are not over-compensated.  Windows above TC_MAX_WIN = 41 (ss_passive.cu) would exceed the bound and run
//     k_aggregate_ws.  (Banking partial sums every 8 rows in float32 would make the bound independent of the window, but
//     there is no room for a second accumulator in TMEM -- D 192 + operands 160..224 columns of 512 -- and red.global
//     runs at 1.3 cycles per lane: measured +24 %.)
//   * TMEM map (512 columns): D half h at 96 h; A at 192 + 160 stage + 80 half + 40 (hi|lo) + j  (win <= 39).
//     SINGLE (39 < win <= 79): one stage of A, 192 + 2 KC half + KC (hi|lo) + j with KC = 8 ceil(win / 8), and one stage of
//     the left-weight residual in shared memory; the producers then wait for the previous row's MMAs (a second commit onto
//     barrier 7) before they overwrite them -- a stall of one MMA batch per window row instead of a second buffer.
//   * one thread (producer 3, lane 0) issues the 30 tcgen05.mma per window row and commits them onto the "weight stage
//     free" mbarrier, whose count is consumers + 1.

#ifndef SS_TC_CREG
#define SS_TC_CREG 136         // registers per consumer / producer thread after setmaxnreg: 12 * CREG + 4 * PREG = 16 * 128
#define SS_TC_PREG 96
#endif
#ifndef SS_TC_PUNROLL
#define SS_TC_PUNROLL 2        // weight batches (of 4) in flight per producer lane in the right-column loop
#endif

constexpr int TC_SBO = 144;                       // bytes between 8-column groups of the left-weight operand
constexpr int TC_LBO = (TILE_WS / 8) * TC_SBO;    // bytes between the two 4-offset chunks of a K group
constexpr int TC_KGB = 2 * TC_LBO;                // bytes per K group (8 window offsets)
constexpr int TC_DS_BYTES = 224 * TILE_WS * 4;    // denominators read back: [r][x]
// Raw-cost tile: 128 bytes per column (4 disparities per word, 32 words = one pass over the banks) plus 24 bytes after every
// 8 columns.  A consumer lane reads the word of (column 8 xg + a, disparity group dg), dg = dl + 2 xl rotated: its bank is
// 6 xg + dg = 8 xl + dl (+ a warp constant), 32 distinct banks -- one wavefront per load.  (Round 2 measured two with the
// uniform 132-byte pitch: 10 xl + dl collides for xl = 0 / 3.)
constexpr int TC_EP = 128, TC_EG = 24;
constexpr float TC_TRUNC_PER_MMA = 5.9604645e-8f;  // 2^-24: half the worst-case relative truncation loss of one tcgen05.mma

struct TcSmem {
    int e, f1, f2, pa, c1, c2, w1, w2, ds, total;
    int ebytes, f1bytes, f2bytes, pabytes, w1arr, w2bytes;
};
// left-weight operand region: [hi stage 0 | hi stage 1 | lo stage 0 | lo stage 1 (absent when single)]
__host__ __device__ inline TcSmem tc_smem(int win, bool single) {
    const int T = TILE_WS, DC = 128, NU = T + win - 1, NR = T + DC - 1, NRp = T + DC, NV = NR + win - 1;
    const int winq = (win + 3) >> 2, winr = winq * 4, KG = (win + 7) >> 3;
    TcSmem p;
    p.ebytes = (NU * TC_EP + ((NU + 7) >> 3) * TC_EG + 15) & ~15;   // tiles start at a multiple of 8 columns
    p.f1bytes = NU * 16;
    p.f2bytes = NV * 16;
    p.pabytes = winq * 16;
    p.w1arr = KG * TC_KGB;
    p.w2bytes = (winr * NRp * 4 + 15) & ~15;
    int off = 256;                                 // header: 16 mbarriers + the TMEM base address
    p.ds = off;
    p.e = off;  off += 2 * p.ebytes;
    p.f1 = off; off += 2 * p.f1bytes;
    p.f2 = off; off += 2 * p.f2bytes;
    p.pa = off; off += 2 * p.pabytes;
    p.c1 = off; off += T * 16;
    p.c2 = off; off += NRp * 16;
    p.w1 = off; off += (single ? 3 : 4) * p.w1arr;
    p.w2 = off; off += 2 * p.w2bytes;
    p.total = off > p.ds + TC_DS_BYTES ? off : p.ds + TC_DS_BYTES;
    return p;
}

// idesc of tcgen05.mma kind::tf32: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor (SWIZZLE_NONE, version 1): start address, leading / stride byte offsets in 16-byte units
__device__ __forceinline__ u64 tc_sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (u64)((addr & 0x3FFFF) >> 4) | ((u64)(lbo >> 4) << 16) | ((u64)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float tc_lo(float w) { return __fsub_rn(w, __uint_as_float(__float_as_uint(w) & 0xffffe000u)); }

// 12 consumer + 4 producer warps = the whole register file at 128 per thread, re-split 136 / 96 by setmaxnreg.  (A 17th warp
// for the TMA / MMA issue was tried: ptxas sizes a 544-thread launch at 96 registers per thread, the pool then holds
// 17 * 32 * 96 and the consumers' setmaxnreg.inc to 136 never completes -- the kernel hangs.)
constexpr int TC_THREADS = 512;

template <int REM, bool SINGLE>
__global__ void __launch_bounds__(TC_THREADS, 1) k_aggregate_tc(const AggParams P) {
    constexpr int DC = 128, T = TILE_WS, NRp = T + DC, EP = TC_EP, EG = TC_EG, CW = 12, PW = 4, NDB = 4, NT = TC_THREADS;
    extern __shared__ __align__(128) unsigned char smem[];

    const Geom &g = P.g;
    const int win = g.win, pad = g.pad;
    const int NU = g.NU, NR = g.NR, NV = g.NV;
    const TcSmem sp = tc_smem(win, SINGLE);
    const int winq = (win + 3) >> 2, winp = winq * 4, KG = (win + 7) >> 3;
    const uint32_t KC = SINGLE ? 8u * (uint32_t)KG : 40u;     // TMEM columns per (half, hi|lo) block of A
    const uint32_t bar0 = smem_u32(smem);
    // barrier slots: 0 centres | 1,2 fullF | 3,4 emptyF | 5,6 fullW | 8,9 emptyW (consumers + MMA commit) | 11,12 fullE |
    //                13,14 emptyE | 15 denominators read back | 7 (SINGLE) MMAs of the previous window row done
    auto BAR = [&](int slot) { return bar0 + 8u * (uint32_t)slot; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 128);

    const int tid = threadIdx.x, lane = tid & 31;
    // The warp index comes out of a shuffle so that ptxas treats it -- and every role branch on it -- as warp-uniform: the
    // producers' loop control and the whole tcgen05.mma issue loop then run in the uniform datapath (11 instructions per MMA
    // instead of 19 with an ELECT / 4 x R2UR.BROADCAST / VOTEU group around each; C2 7.87 -> 7.76 ms).  The same change made
    // the GSW instantiation of k_aggregate_ws 10 % slower and left the ASW one unchanged, so it is applied here only.
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int tile = g.tile0 + (int)blockIdx.x, per_ch = g.ntx * (g.row1 - g.row0);
    const int ch = tile / per_ch, rem = tile - ch * per_ch;
    const int x0 = (rem % g.ntx) * T;
    const int y = g.row0 + rem / g.ntx;
    const int xsub = g.nsub > 1 ? (int)blockIdx.y : -1;     // tail launch: only this 32-column block of the tile (ss_passive.cu)
    const int dlo = g.dLo + ch * DC;
    const int erows = g.erow1 - g.erow0;
    const int i_lo = max(0, pad - y), i_hi = min(win - 1, g.H - 1 - y + pad);
    const int nsteps = i_hi - i_lo + 1;

    if (x0 + T - 1 < dlo) {                        // no evaluated pair in this tile (see k_aggregate_ws)
        if (P.vol_export && xsub <= 0) {
            const int rowo = y - g.row0;
            for (int k = tid; k < T * (DC / 4); k += NT) {
                const int x = x0 + k / (DC / 4), kq = (k % (DC / 4)) * 4;
                if (x >= g.W) continue;
                const size_t o = ((size_t)rowo * g.W + x) * P.Dp + (size_t)ch * DC + kq;
                if (P.vol0) *reinterpret_cast<float4 *>(P.vol0 + o) = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
            }
        }
        return;
    }

    if (warp == 0) {                               // the whole 512-column tensor memory of this SM (one block per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(BAR(0), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(BAR(1 + s), 1);
            mbar_init(BAR(3 + s), PW);
            mbar_init(BAR(5 + s), PW);
            mbar_init(BAR(8 + s), CW + 1);
            mbar_init(BAR(11 + s), 1);
            mbar_init(BAR(13 + s), CW);
        }
        mbar_init(BAR(15), PW);
        mbar_init(BAR(7), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tmem_slot;
    if (tbase != 0u) __trap();   // one block per SM allocates all 512 columns: the only placement is 0 (the issue loop relies on it)
    auto colA = [&](int stage, int half, int lo) {
        return SINGLE ? 192u + (uint32_t)half * 2u * KC + (uint32_t)lo * KC
                      : 192u + (uint32_t)stage * 160u + (uint32_t)half * 80u + (uint32_t)lo * 40u;
    };

    if (warp >= CW) {
        // =================================== producers ===================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SS_TC_PREG));
        const int pw = warp - CW;
        const uint32_t lane_base = (uint32_t)(pw * 32) << 16;       // this warp's quarter of the TMEM lanes
        const int f2_start = x0 - dlo - DC + 1 - pad + g.PL2;
        const int c2_start = x0 - dlo - DC + 1 + g.PL2;
        const size_t e_plane = (size_t)g.EPL;
        const float4 *C1s = reinterpret_cast<const float4 *>(smem + sp.c1);
        const float4 *C2s = reinterpret_cast<const float4 *>(smem + sp.c2);

        auto issue_F = [&](int n) {
            const int i = i_lo + n, ii = y - pad + i, st = n & 1;
            const uint32_t bar = BAR(1 + st);
            mbar_expect_tx(bar, (uint32_t)((NU + NV + winq) * 16));
            tma_load_1d(smem_u32(smem + (sp.f1 + st * sp.f1bytes)), P.F1 + (size_t)(ii - g.erow0) * g.UW + x0, NU * 16, bar);
            tma_load_1d(smem_u32(smem + (sp.f2 + st * sp.f2bytes)), P.F2 + (size_t)(ii - g.erow0) * g.VW + f2_start, NV * 16, bar);
            tma_load_1d(smem_u32(smem + (sp.pa + st * sp.pabytes)), P.proxarg + (size_t)i * winp, winq * 16, bar);
        };
        auto issue_E = [&](int n) {
            const int ii = y - pad + i_lo + n, st = n & 1;
            const uint32_t bar = BAR(11 + st);
            mbar_expect_tx(bar, (uint32_t)sp.ebytes);
            tma_load_1d(smem_u32(smem + (sp.e + st * sp.ebytes)),
                        static_cast<const uint8_t *>(P.E) + ((size_t)ch * erows + (ii - g.erow0)) * e_plane + (size_t)x0 * EP + (size_t)(x0 >> 3) * EG,
                        (uint32_t)sp.ebytes, bar);
        };
        if (pw == 0 && lane == 0) {
            mbar_expect_tx(BAR(0), (uint32_t)((T + NR) * 16));
            tma_load_1d(smem_u32(smem + sp.c1), P.F1 + (size_t)(y - g.erow0) * g.UW + x0 + pad, T * 16, BAR(0));
            tma_load_1d(smem_u32(smem + sp.c2), P.F2 + (size_t)(y - g.erow0) * g.VW + c2_start, NR * 16, BAR(0));
            issue_F(0);
        }
        // K padding: window offsets in [win, 8 KG) must contribute 0.  Zero every A column of this warp's TMEM lanes and
        // the whole left-weight operand once; offsets inside the last written batch are zeroed when they are written.
        for (int c = 0; c < (SINGLE ? (int)(4u * KC) : 320); c += 4) tc_st4(tbase + lane_base + 192u + c, 0u, 0u, 0u, 0u);
        for (int k = (pw * 32 + lane) * 16; k < (SINGLE ? 3 : 4) * sp.w1arr; k += PW * 32 * 16)
            *reinterpret_cast<float4 *>(smem + sp.w1 + k) = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");              // the producers' own barrier: zeroing done everywhere

        const int NB = winq;                                        // batches of 4 window offsets per column block
        // A block restricted to x-block xsub needs right centres r = T-1-x+k with x in the block (column blocks cbr0..cbr1)
        // and one left block; everything else stays at the zeros written below.
        const int cbr0 = xsub < 0 ? 0 : T / 32 - 1 - xsub, cbr1 = xsub < 0 ? NRp / 32 - 1 : (T - 1 - 32 * xsub + DC - 1) / 32;
        // full tile: 7 right blocks (producers 0-2 two each, producer 3 one) + 3 left blocks; producer 3 also issues the MMAs and
        // takes 20 of the 90 batches of a 35-wide window (the others 23-24): left shares 5/9, 6/9, 5/9, 11/9 NB
        const int l0 = xsub >= 0 ? xsub * NB + (pw * NB) / 4
                                 : pw == 0 ? 0 : pw == 1 ? (NB * 5) / 9 : pw == 2 ? (NB * 11) / 9 : (NB * 16) / 9;
        const int l1 = xsub >= 0 ? xsub * NB + ((pw + 1) * NB) / 4
                                 : pw == 0 ? (NB * 5) / 9 : pw == 1 ? (NB * 11) / 9 : pw == 2 ? (NB * 16) / 9 : 3 * NB;
        const int l0_blk = l0 / NB, l0_jb = l0 - l0_blk * NB;
        const int o_f1 = sp.f1, o_f2 = sp.f2, o_pa = sp.pa, o_w1 = sp.w1, o_w2 = sp.w2;
        const int b_f1 = sp.f1bytes, b_f2 = sp.f2bytes, b_pa = sp.pabytes, b_w2 = sp.w2bytes, w1arr = sp.w1arr;

        constexpr int PUNROLL = SS_TC_PUNROLL;
        int sw = 0, phw = 0;
        for (int n = 0; n < nsteps; ++n) {
            const int st = n & 1, ph = (n >> 1) & 1;
            if (pw == 0 && lane == 0) {
                if (n + 1 < nsteps) {
                    mbar_wait(BAR(3 + ((n + 1) & 1)), (((n + 1) >> 1) & 1) ^ 1);
                    issue_F(n + 1);
                }
                mbar_wait(BAR(13 + st), ph ^ 1);
                issue_E(n);
            }
            __syncwarp();
            if (n == 0) mbar_wait(BAR(0), 0);
            mbar_wait(BAR(1 + st), ph);            // features of this window row have landed
            mbar_wait(BAR(8 + sw), phw ^ 1);  // consumers AND the tensor core are done with this weight stage
            if (SINGLE && n > 0) mbar_wait(BAR(7), (n - 1) & 1);   // single-stage operands: the previous row's MMAs have read them
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

            const float4 *f1 = reinterpret_cast<const float4 *>(smem + o_f1 + st * b_f1);
            const float4 *f2 = reinterpret_cast<const float4 *>(smem + o_f2 + st * b_f2);
            const float *parg = reinterpret_cast<const float *>(smem + o_pa + st * b_pa);
            unsigned char *W1hi = smem + o_w1 + sw * w1arr;
            unsigned char *W1lo = smem + o_w1 + (2 + (SINGLE ? 0 : sw)) * w1arr;
            float *W2s = reinterpret_cast<float *>(smem + o_w2 + sw * b_w2);

            // ---- right columns: consumer copy (row-major, reversed index) + tensor-memory copy (hi, lo) ----
#pragma unroll 1
            for (int cb = pw; cb < ((P.freerun & 2) ? 0 : NRp / 32); cb += 4) {
                if (cb < cbr0 || cb > cbr1) continue;
                const int col = cb * 32 + lane;                      // r; column NR is padding (never read)
                const int src = NR - 1 - col;
                const float4 c = C2s[src];
                const float4 *nb = f2 + src;
                float *dst = W2s + col;
                const uint32_t ta = tbase + lane_base + colA(sw, cb >> 2, 0);
#pragma unroll PUNROLL
                for (int jb = 0; jb < NB; ++jb) {
                    const float4 t = *reinterpret_cast<const float4 *>(parg + 4 * jb);
                    const float4 n0 = nb[0], n1 = nb[1], n2 = nb[2], n3 = nb[3];
                    float w0 = support_weight<false>(c, n0, P.kC, t.x);
                    float w1 = support_weight<false>(c, n1, P.kC, t.y);
                    float w2 = support_weight<false>(c, n2, P.kC, t.z);
                    float w3 = support_weight<false>(c, n3, P.kC, t.w);
                    if (4 * jb + 3 >= win) {                          // offsets past the window (last batch only)
                        if (4 * jb + 1 >= win) w1 = 0.f;
                        if (4 * jb + 2 >= win) w2 = 0.f;
                        w3 = 0.f;
                    }
                    dst[0] = w0; dst[NRp] = w1; dst[2 * NRp] = w2; dst[3 * NRp] = w3;
                    tc_st4(ta + 4 * jb, __float_as_uint(w0), __float_as_uint(w1), __float_as_uint(w2), __float_as_uint(w3));
                    tc_st4(ta + KC + 4 * jb, __float_as_uint(tc_lo(w0)), __float_as_uint(tc_lo(w1)), __float_as_uint(tc_lo(w2)),
                           __float_as_uint(tc_lo(w3)));
                    nb += 4;
                    dst += 4 * NRp;
                }
            }
            // ---- left columns: K-major operand shared by the consumers and the tensor core (hi) + its residual (lo) ----
            int blk = l0_blk, jb = l0_jb;                            // (column block, batch) of lb, kept without a division
#pragma unroll 2
            for (int lb = l0; lb < ((P.freerun & 2) ? l0 : l1); ++lb) {
                const int col = blk * 32 + lane;                     // x
                const float4 c = C1s[col];
                const float4 *nb = f1 + col + 4 * jb;
                const float4 t = *reinterpret_cast<const float4 *>(parg + 4 * jb);
                const float4 n0 = nb[0], n1 = nb[1], n2 = nb[2], n3 = nb[3];
                float w0 = support_weight<false>(c, n0, P.kC, t.x);
                float w1 = support_weight<false>(c, n1, P.kC, t.y);
                float w2 = support_weight<false>(c, n2, P.kC, t.z);
                float w3 = support_weight<false>(c, n3, P.kC, t.w);
                if (4 * jb + 3 >= win) {
                    if (4 * jb + 1 >= win) w1 = 0.f;
                    if (4 * jb + 2 >= win) w2 = 0.f;
                    w3 = 0.f;
                }
                const int qo = (jb >> 1) * TC_KGB + (jb & 1) * TC_LBO + (col >> 3) * TC_SBO + (col & 7) * 16;
                *reinterpret_cast<float4 *>(W1hi + qo) = make_float4(w0, w1, w2, w3);
                *reinterpret_cast<float4 *>(W1lo + qo) = make_float4(tc_lo(w0), tc_lo(w1), tc_lo(w2), tc_lo(w3));
                if (++jb == NB) { jb = 0; ++blk; }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> tensor-core (async proxy) reads
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(BAR(5 + sw));        // weights ready
                mbar_arrive(BAR(3 + st));        // feature stage may be refilled
            }
            // ---- the tensor core: D[r][x] += W2hi*W1hi + W2hi*W1lo + W2lo*W1hi over this window row ----
            // Producer 3 issues the row's 6 KG MMAs (30 for a 35-wide window).  The whole warp runs the loop (the instruction is
            // predicated on elect.sync) and, because the role branch is warp-uniform (see `warp`), entirely in the uniform
            // datapath: 11 instructions per MMA.  The issue is still not free: without the MMAs the kernel is faster
            // (SS_FREERUN=8) -- issue slots of a warp whose weights the whole block waits for.  Tried and rejected:
            // the issue split over two producers (+2 %), issued two at a time between the next row's weight batches (the
            // bookkeeping costs more than it hides), a 17th warp (see TC_THREADS).
            if (pw == 3) {
                mbar_wait(BAR(5 + sw), phw);     // every producer has arrived: the operands of this row are in place
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int KGx = (P.freerun & 8) ? 0 : KG;         // timing experiment: no MMAs (the commits still arrive)
                // Every operand of the issue loop is derived from launch parameters, loop counters and constants only (the TMEM base
                // is 0: the block owns all 512 columns, checked after the allocation), so ptxas can keep them in uniform registers
                const uint32_t bhi = smem_u32(smem) + (uint32_t)(o_w1 + sw * w1arr);
                const uint32_t blo = smem_u32(smem) + (uint32_t)(o_w1 + (2 + (SINGLE ? 0 : sw)) * w1arr);
                const uint32_t tb = 0u, ca = colA(sw, 0, 0);
                const int KGu = KGx, first = n == 0 ? 1 : 0;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        uint32_t a = tb + ca + (uint32_t)half * (SINGLE ? 2u * KC : 80u) + (term == 2 ? (SINGLE ? KC : 40u) : 0u);
                        u64 bd = tc_sdesc(term == 1 ? blo : bhi, TC_LBO, TC_SBO);
#pragma unroll 1
                        for (int kg = 0; kg < KGu; ++kg) {
                            const uint32_t acc = !(first && term == 0 && kg == 0);
                            asm volatile("{\n.reg .pred p, pe;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|pe, 0xffffffff;\n"
                                         "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tb + 96u * half),
                                         "r"(a), "l"(bd), "r"(tc_idesc(128, T)), "r"(acc)
                                         : "memory");
                            a += 8u;
                            bd += (u64)(TC_KGB >> 4);
                        }
                    }
                }
                // completion of everything issued so far arrives on "weight stage free" (and on barrier 7 when the operands are
                // single-staged)
                asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
                             "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(BAR(8 + sw)) : "memory");
                if (SINGLE)
                    asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
                                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(BAR(7)) : "memory");
            }
            if (++sw == 2) { sw = 0; phw ^= 1; }
        }
        // ---- denominators: TMEM -> shared memory [r][x] once the last window row is fully consumed and accumulated ----
        {
            const int swl = sw ^ 1, phl = sw == 0 ? phw ^ 1 : phw;   // stage / phase of the last window row

            mbar_wait(BAR(8 + swl), phl);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float *Ds = reinterpret_cast<float *>(smem + sp.ds);
            for (int half = 0; half < 2; ++half) {
                const int r = half * 128 + pw * 32 + lane;
                for (int c0 = 0; c0 < T; c0 += 16) {
                    uint32_t v[16];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(tbase + lane_base + 96u * half + c0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (r < 224) {
                        uint4 *o = reinterpret_cast<uint4 *>(Ds + (size_t)r * T + c0);
                        o[0] = make_uint4(v[0], v[1], v[2], v[3]);
                        o[1] = make_uint4(v[4], v[5], v[6], v[7]);
                        o[2] = make_uint4(v[8], v[9], v[10], v[11]);
                        o[3] = make_uint4(v[12], v[13], v[14], v[15]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(15));
        }
        return;
    }

    // =================================== consumers ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SS_TC_CREG));
    const int xl = lane >> 3, dl = lane & 7;
    const int xg = (warp / NDB) * 4 + xl;
    const int dg = ((warp % NDB) * 8 + dl + 2 * xl) % (DC / 4);
    const int xb = 8 * xg, kb = 4 * dg;
    const bool sub_ok = xsub < 0 || warp / NDB == xsub;     // warp-uniform
    const bool lane_live = sub_ok && (x0 + xb < g.W) && (dlo + kb <= g.dHi) && (dlo + kb + 3 >= g.dVLo) && (x0 + xb + 7 >= dlo + kb);
    const bool warp_live = __any_sync(0xffffffffu, lane_live);
    const int R0 = T - 8 - xb + kb;

    u64 acc0[8][2];                                            // numerators
#pragma unroll
    for (int a = 0; a < 8; ++a) { acc0[a][0] = 0ull; acc0[a][1] = 0ull; }

    int sw = 0, phw = 0;
    for (int n = 0; n < nsteps; ++n) {
        const int st = n & 1, ph = (n >> 1) & 1;
        mbar_wait(BAR(5 + sw), phw);             // weights of this window row
        mbar_wait(BAR(11 + st), ph);             // raw costs of this window row
        if (warp_live && !(P.freerun & 4)) {
            u64 ring[8][2];
            const uint8_t *ep = smem + (sp.e + st * sp.ebytes) + xb * EP + xg * EG + kb;
            auto load_e = [&](const uint8_t *q, u64 &lo, u64 &hi) {
                const uint32_t e = *reinterpret_cast<const uint32_t *>(q);
                lo = pk(u8_to_f32(e, 0), u8_to_f32(e, 1));
                hi = pk(u8_to_f32(e, 2), u8_to_f32(e, 3));
            };
#pragma unroll
            for (int a = 0; a < 7; ++a) load_e(ep + a * EP, ring[a][0], ring[a][1]);
            ep += 7 * EP;
            // left weights: K-major operand, one 16-byte chunk = one column x, 4 consecutive window offsets
            const unsigned char *w1q = smem + (sp.w1 + sw * sp.w1arr) + xg * TC_SBO;
            const float *w2p = reinterpret_cast<const float *>(smem + (sp.w2 + sw * sp.w2bytes)) + R0;
            float4 wq[8];
#ifdef SS_DEBUG_ADDR
            // shared-memory byte addresses of this lane's three load streams (first window row of one interior block)
            if (n == 0 && blockIdx.x == 5)
                printf("ADDR %d %d %u %u %u\n", warp, lane, smem_u32(w2p), smem_u32(w1q), smem_u32(ep));
#endif

            auto step = [&](auto sc) {
                constexpr int s = decltype(sc)::value;
                load_e(ep + s * EP + (s > 0 ? EG : 0), ring[(7 + s) & 7][0], ring[(7 + s) & 7][1]);   // column xb + 7 + s: next group
                if (s % 4 == 0) {
#pragma unroll
                    for (int a = 0; a < 8; ++a) wq[a] = *reinterpret_cast<const float4 *>(w1q + (s / 4) * TC_LBO + a * 16);
                }
                const float4 v0 = *reinterpret_cast<const float4 *>(w2p + s * NRp);
                const float4 v1 = *reinterpret_cast<const float4 *>(w2p + s * NRp + 4);
                const float4 v2 = *reinterpret_cast<const float4 *>(w2p + s * NRp + 8);
                const u64 VA[6] = {pk(v0.x, v0.y), pk(v0.z, v0.w), pk(v1.x, v1.y), pk(v1.z, v1.w), pk(v2.x, v2.y), pk(v2.z, v2.w)};
                const float vf[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const float w1 = (s % 4 == 0) ? wq[a].x : (s % 4 == 1) ? wq[a].y : (s % 4 == 2) ? wq[a].z : wq[a].w;
                    const u64 w1d = pk(w1, w1);
#pragma unroll
                    for (int bp = 0; bp < 2; ++bp) {
                        const int k = 7 - a + 2 * bp;
                        const u64 e2 = ring[(a + s) & 7][bp];
                        const u64 ww = (k & 1) ? pk(__fmul_rn(w1, vf[k]), __fmul_rn(w1, vf[k + 1])) : mul2(w1d, VA[k >> 1]);
                        acc0[a][bp] = fma2(ww, e2, acc0[a][bp]);        // cost += w1*w2*e  (_passive.cpp:77)
                    }
                }
            };
            int j = 0;
            auto period = [&]() {
                step(IC<0>{}); step(IC<1>{}); step(IC<2>{}); step(IC<3>{});
                step(IC<4>{}); step(IC<5>{}); step(IC<6>{}); step(IC<7>{});
                ep += 8 * EP + EG;
                w1q += TC_KGB;
                w2p += 8 * NRp;
            };
#pragma unroll 1
            for (; j + 16 <= win; j += 16) { period(); period(); }
            if (j + 8 <= win) { period(); j += 8; }
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
            mbar_arrive(BAR(8 + sw));            // weight stage free (the MMA commit is the other arrival)
            mbar_arrive(BAR(13 + st));           // raw-cost stage free
        }
        if (++sw == 2) { sw = 0; phw ^= 1; }
    }

    // ---- epilogue: denominators from the tensor core, normalise, WTA over the chunk, optional volume store ----
    mbar_wait(BAR(15), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
    float *Ds = reinterpret_cast<float *>(smem + sp.ds);
    const float fix_kg = TC_TRUNC_PER_MMA * (float)(3 * nsteps);   // per K-group that contributes; see the header
    const int rowo = y - g.row0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int x = x0 + xb + a;
        float c0[4];
        upk(acc0[a][0], c0[0], c0[1]);
        upk(acc0[a][1], c0[2], c0[3]);
        u64 best = KEY_NONE;
        float out0[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int d = dlo + kb + b;
            const bool valid = sub_ok && (x < g.W) && (d >= g.dVLo) && (d <= g.dHi) && (x - d >= 0);
            float *dp = Ds + (T - 1 - (xb + a) + kb + b) * T + xb + a;
            // valid window columns of this pair: right column x-d-pad+j >= 0 and left column x-pad+j < W (_passive.cpp:67-68)
            const int jlo = max(0, pad - (x - d)), jhi = min(win - 1, g.W - 1 - x + pad);
            const float den = __fmul_rn(*dp, fmaf(fix_kg, (float)((jhi >> 3) - (jlo >> 3) + 1), 1.0f));
            const float cost = __fdiv_rn(c0[b], den);                   // cost / tot (:88)
            out0[b] = valid ? cost : INFINITY;
            if (valid) {
                const u64 k = make_key(cost, d);
                best = k < best ? k : best;
            }
            if (P.bestR) *dp = out0[b];          // in place: the slot of den[x][d] now holds cost[x][d] (right-reference WTA below)
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
            const u64 o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o < best ? o : best;
        }
        if ((lane & 7) == 0 && x < g.W && best != KEY_NONE) atomicMin(P.bestL + (size_t)rowo * g.W + x, best);
        if (x < g.W && P.vol0 && sub_ok) {
            const size_t o = ((size_t)rowo * g.W + x) * P.Dp + (size_t)ch * DC + kb;
            *reinterpret_cast<float4 *>(P.vol0 + o) = make_float4(out0[0], out0[1], out0[2], out0[3]);
        }
    }
    if (P.bestR) {
        asm volatile("bar.sync 2, %0;" ::"n"(CW * 32) : "memory");
        wta_right_rows<T, DC>(Ds, warp, CW, lane, x0, dlo, g.W, P.bestR + (size_t)rowo * g.W);
    }
}
