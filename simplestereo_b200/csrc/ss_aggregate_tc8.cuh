// ss_aggregate_tc8.cuh -- k_aggregate_tc8: k_aggregate_tc with 128-column tiles and 8 x 8 lane tiles.
// Included by ss_passive.cu after ss_aggregate_tc.cuh (shares its helpers).  ASW only, DC = 128, T = 128, win <= 35.
//
// Why a second tile shape.  ncu on k_aggregate_tc (profiles/r02_*): the shared-memory data pipe is 80 % busy and the FP32
// pipe 56 % -- the kernel is bound by shared-memory DELIVERY, not by issue slots.  A 128-bit shared load hands 512 bytes to
// the register file and costs 4 data-pipe wavefronts however few distinct bytes the warp touches (measured on the B200: the
// right-weight loads, 8 distinct 16-byte chunks per warp, identical in every quarter-warp, take exactly 4.0; the "one
// wavefront" microbenchmark of round 1 had its v4 loads narrowed to 32 bits by ptxas).  An 8-column x 4-disparity lane tile
// needs 8 + 11 weights and 4 raw costs per 32 multiply-adds: 18 wavefronts per warp and window offset, 216 per block = 60 % of
// the pipe for the consumers alone.  An 8 x 8 lane tile needs 8 + 15 + 8 for 64: 22 wavefronts per 64 multiply-adds, 39 % less,
// which puts the kernel back under the FP32 pipe.
//
// What changes against k_aggregate_tc:
//   * block = (row, 128 columns, 128 disparities); 8 consumer warps (4 x-blocks of 32 columns x 2 disparity blocks of 64),
//     lane tile 8 x 8: 64 numerator accumulators + a 64-register ring of raw costs per lane; registers 208 / 88 per consumer /
//     producer thread (setmaxnreg; 384 threads are launched at 168).
//   * tensor memory (all 512 columns): accumulator halves at 0 and 128 (r = T-1-x+k spans 255 rows: two M = 128 halves, N = 128
//     columns), right weights hi at 256 + 80 stage + 40 half (two stages), right-weight residuals lo at 416 + 40 half (ONE
//     stage: there is no room for a second).  The row's MMAs are therefore ordered lo x hi FIRST and committed onto barrier 7
//     on their own; the producers tabulate the LEFT weights of the next row (shared memory, double-staged) before they wait
//     for that barrier and start on the right columns, so the single stage costs no stall.
//   * the right-weight rows in shared memory are swizzled at 16-byte granularity (chunk q -> q ^ ((q >> 3) & 1)): lanes of a
//     quarter-warp read chunks 8 floats apart (R0 = ... + 8 dl), which would put lanes dl and dl + 4 on the same banks.
//   * raw-cost column pitch 136 bytes (8-byte aligned: a lane reads its 8 disparities with one 64-bit load).
//   * shared memory: 231 KB at win 35 (both left-weight residual stages included) -- larger windows run k_aggregate_tc.

constexpr int TC8_T = 128;
constexpr int TC8_EP = 128 + 8;
constexpr int TC8_LBO = (TC8_T / 8) * TC_SBO;     // bytes between the two 4-offset chunks of a K group
constexpr int TC8_KGB = 2 * TC8_LBO;              // bytes per K group (8 window offsets)
constexpr int TC8_THREADS = 384;                  // 8 consumer + 4 producer warps
constexpr int TC8_CREG = 208, TC8_PREG = 88;      // 8 * 32 * 208 + 4 * 32 * 88 = 384 * 168

__host__ __device__ inline TcSmem tc8_smem(int win) {
    const int T = TC8_T, DC = 128, NU = T + win - 1, NR = T + DC - 1, NRp = T + DC, NV = NR + win - 1;
    const int winq = (win + 3) >> 2, winr = winq * 4, KG = (win + 7) >> 3;
    TcSmem p;
    p.ebytes = (NU * TC8_EP + 15) & ~15;
    p.f1bytes = NU * 16;
    p.f2bytes = NV * 16;
    p.pabytes = winq * 16;
    p.w1arr = KG * TC8_KGB;
    p.w2bytes = (winr * NRp * 4 + 15) & ~15;
    int off = 256;                                 // header: 16 mbarriers + the TMEM base address
    p.ds = off;
    p.e = off;  off += 2 * p.ebytes;
    p.f1 = off; off += 2 * p.f1bytes;
    p.f2 = off; off += 2 * p.f2bytes;
    p.pa = off; off += 2 * p.pabytes;
    p.c1 = off; off += T * 16;
    p.c2 = off; off += NRp * 16;
    p.w1 = off; off += 4 * p.w1arr;                // [hi stage 0 | hi stage 1 | lo stage 0 | lo stage 1]
    p.w2 = off; off += 2 * p.w2bytes;
    const int ds_end = p.ds + NRp * T * 4;         // denominators / staged costs [r][x]
    p.total = off > ds_end ? off : ds_end;
    return p;
}

// float index inside a right-weight row of column r (16-byte chunk swizzle, see the header)
__host__ __device__ __forceinline__ int tc8_w2idx(int r) {
    const int q = r >> 2;
    return ((q ^ ((q >> 3) & 1)) << 2) | (r & 3);
}

template <int REM>
__global__ void __launch_bounds__(TC8_THREADS, 1) k_aggregate_tc8(const AggParams P) {
    constexpr int DC = 128, T = TC8_T, NRp = T + DC, EP = TC8_EP, CW = 8, PW = 4, NT = TC8_THREADS;
    constexpr uint32_t KC = 40u;                   // TMEM columns per (half, hi|lo) block of the right-weight operand
    extern __shared__ __align__(128) unsigned char smem[];

    const Geom &g = P.g;
    const int win = g.win, pad = g.pad;
    const int NU = g.NU, NR = g.NR, NV = g.NV;
    const TcSmem sp = tc8_smem(win);
    const int winq = (win + 3) >> 2, winp = winq * 4, KG = (win + 7) >> 3;
    const uint32_t bar0 = smem_u32(smem);
    // barrier slots: 0 centres | 1,2 fullF | 3,4 emptyF | 5,6 fullW | 8,9 emptyW (consumers + MMA commit) | 11,12 fullE |
    //                13,14 emptyE | 15 denominators read back | 7 the row's lo x hi MMAs have read the residual stage
    auto BAR = [&](int slot) { return bar0 + 8u * (uint32_t)slot; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 128);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    auto colHi = [&](int stage, int half) { return 256u + (uint32_t)stage * 80u + (uint32_t)half * KC; };
    auto colLo = [&](int half) { return 416u + (uint32_t)half * KC; };

    if (warp >= CW) {
        // =================================== producers ===================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC8_PREG));
        const int pw = warp - CW;
        const uint32_t lane_base = (uint32_t)(pw * 32) << 16;       // this warp's quarter of the TMEM lanes
        const int f2_start = x0 - dlo - DC + 1 - pad + g.PL2;
        const int c2_start = x0 - dlo - DC + 1 + g.PL2;
        const size_t e_plane = (size_t)g.UW * EP;
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
                        static_cast<const uint8_t *>(P.E) + ((size_t)ch * erows + (ii - g.erow0)) * e_plane + (size_t)x0 * EP,
                        (uint32_t)sp.ebytes, bar);
        };
        if (pw == 0 && lane == 0) {
            mbar_expect_tx(BAR(0), (uint32_t)((T + NR) * 16));
            tma_load_1d(smem_u32(smem + sp.c1), P.F1 + (size_t)(y - g.erow0) * g.UW + x0 + pad, T * 16, BAR(0));
            tma_load_1d(smem_u32(smem + sp.c2), P.F2 + (size_t)(y - g.erow0) * g.VW + c2_start, NR * 16, BAR(0));
            issue_F(0);
        }
        // K padding: window offsets in [win, 8 KG) must contribute 0.  Zero every operand column of this warp's TMEM lanes and
        // the whole left-weight operand once; offsets inside the last written batch are zeroed when they are written.
        for (int c = 0; c < 240; c += 4) tc_st4(tbase + lane_base + 256u + c, 0u, 0u, 0u, 0u);
        for (int k = (pw * 32 + lane) * 16; k < 4 * sp.w1arr; k += PW * 32 * 16)
            *reinterpret_cast<float4 *>(smem + sp.w1 + k) = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");              // the producers' own barrier: zeroing done everywhere

        const int NB = winq;                                        // batches of 4 window offsets per column block
        // A block restricted to x-block xsub needs right centres r = T-1-x+k with x in the block (column blocks cbr0..cbr1)
        // and one left block; everything else stays at the zeros written above.
        const int cbr0 = xsub < 0 ? 0 : T / 32 - 1 - xsub, cbr1 = xsub < 0 ? NRp / 32 - 1 : (T - 1 - 32 * xsub + DC - 1) / 32;
        // full tile: 8 right blocks (two per producer: cb = pw, pw + 4) + 4 left blocks of NB batches; producer 3 also issues the
        // MMAs and takes about 15 % fewer batches than the others: left shares 31 : 31 : 31 : 15 of 4 NB
        const int l0 = xsub >= 0 ? xsub * NB + (pw * NB) / 4 : (4 * NB * 31 * pw) / 108;
        const int l1 = xsub >= 0 ? xsub * NB + ((pw + 1) * NB) / 4 : (pw == 3 ? 4 * NB : (4 * NB * 31 * (pw + 1)) / 108);
        const int l0_blk = l0 / NB, l0_jb = l0 - l0_blk * NB;
        const int o_f1 = sp.f1, o_f2 = sp.f2, o_pa = sp.pa, o_w1 = sp.w1, o_w2 = sp.w2;
        const int b_f1 = sp.f1bytes, b_f2 = sp.f2bytes, b_pa = sp.pabytes, b_w2 = sp.w2bytes, w1arr = sp.w1arr;

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
            mbar_wait(BAR(8 + sw), phw ^ 1);       // consumers AND the tensor core are done with this weight stage

            const float4 *f1 = reinterpret_cast<const float4 *>(smem + o_f1 + st * b_f1);
            const float4 *f2 = reinterpret_cast<const float4 *>(smem + o_f2 + st * b_f2);
            const float *parg = reinterpret_cast<const float *>(smem + o_pa + st * b_pa);
            unsigned char *W1hi = smem + o_w1 + sw * w1arr;
            unsigned char *W1lo = smem + o_w1 + (2 + sw) * w1arr;
            float *W2s = reinterpret_cast<float *>(smem + o_w2 + sw * b_w2);

            // ---- left columns first: K-major operand shared by the consumers and the tensor core (hi) + its residual (lo),
            //      both double-staged in shared memory ----
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
                if (4 * jb + 3 >= win) {                              // offsets past the window (last batch only)
                    if (4 * jb + 1 >= win) w1 = 0.f;
                    if (4 * jb + 2 >= win) w2 = 0.f;
                    w3 = 0.f;
                }
                const int qo = (jb >> 1) * TC8_KGB + (jb & 1) * TC8_LBO + (col >> 3) * TC_SBO + (col & 7) * 16;
                *reinterpret_cast<float4 *>(W1hi + qo) = make_float4(w0, w1, w2, w3);
                *reinterpret_cast<float4 *>(W1lo + qo) = make_float4(tc_lo(w0), tc_lo(w1), tc_lo(w2), tc_lo(w3));
                if (++jb == NB) { jb = 0; ++blk; }
            }
            // the residual stage in tensor memory is single: the previous row's lo x hi MMAs must have read it
            if (n > 0) mbar_wait(BAR(7), (n - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

            // ---- right columns: consumer copy (row-major, reversed index, swizzled chunks) + tensor-memory copy (hi, lo) ----
#pragma unroll 1
            for (int cb = pw; cb < ((P.freerun & 2) ? 0 : NRp / 32); cb += 4) {
                if (cb < cbr0 || cb > cbr1) continue;
                const int col = cb * 32 + lane;                      // r; column NR is padding (never read)
                const int src = NR - 1 - col;
                const float4 c = C2s[src];
                const float4 *nb = f2 + src;
                float *dst = W2s + tc8_w2idx(col);
                const uint32_t ta = tbase + lane_base + colHi(sw, cb >> 2);
                const uint32_t tl = tbase + lane_base + colLo(cb >> 2);
#pragma unroll 2
                for (int jb2 = 0; jb2 < NB; ++jb2) {
                    const float4 t = *reinterpret_cast<const float4 *>(parg + 4 * jb2);
                    const float4 n0 = nb[0], n1 = nb[1], n2 = nb[2], n3 = nb[3];
                    float w0 = support_weight<false>(c, n0, P.kC, t.x);
                    float w1 = support_weight<false>(c, n1, P.kC, t.y);
                    float w2 = support_weight<false>(c, n2, P.kC, t.z);
                    float w3 = support_weight<false>(c, n3, P.kC, t.w);
                    if (4 * jb2 + 3 >= win) {
                        if (4 * jb2 + 1 >= win) w1 = 0.f;
                        if (4 * jb2 + 2 >= win) w2 = 0.f;
                        w3 = 0.f;
                    }
                    dst[0] = w0; dst[NRp] = w1; dst[2 * NRp] = w2; dst[3 * NRp] = w3;
                    tc_st4(ta + 4 * jb2, __float_as_uint(w0), __float_as_uint(w1), __float_as_uint(w2), __float_as_uint(w3));
                    tc_st4(tl + 4 * jb2, __float_as_uint(tc_lo(w0)), __float_as_uint(tc_lo(w1)), __float_as_uint(tc_lo(w2)),
                           __float_as_uint(tc_lo(w3)));
                    nb += 4;
                    dst += 4 * NRp;
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> tensor-core (async proxy) reads
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(BAR(5 + sw));        // weights ready
                mbar_arrive(BAR(3 + st));        // feature stage may be refilled
            }
            // ---- the tensor core: D[r][x] += W2lo*W1hi (first: it frees the single residual stage) + W2hi*W1hi + W2hi*W1lo ----
            // Producer 3 issues the row's 6 KG MMAs; the whole warp runs the loop in convergent flow and the instruction itself is
            // predicated on elect.sync (see k_aggregate_tc).
            if (pw == 3) {
                mbar_wait(BAR(5 + sw), phw);     // every producer has arrived: the operands of this row are in place
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int KGx = (P.freerun & 8) ? 0 : KG;         // timing experiment: no MMAs (the commits still arrive)
                const uint32_t bhi = __shfl_sync(0xffffffffu, smem_u32(smem) + (uint32_t)(o_w1 + sw * w1arr), 0);
                const uint32_t blo = __shfl_sync(0xffffffffu, smem_u32(smem) + (uint32_t)(o_w1 + (2 + sw) * w1arr), 0);
                const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
                const uint32_t ahi = __shfl_sync(0xffffffffu, colHi(sw, 0), 0);
                const int KGu = __shfl_sync(0xffffffffu, KGx, 0), first = __shfl_sync(0xffffffffu, n == 0 ? 1 : 0, 0);
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {            // pass 0: lo x hi; pass 1: hi x hi, hi x lo
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
#pragma unroll
                        for (int term = 0; term < (pass == 0 ? 1 : 2); ++term) {
                            uint32_t a = tb + (pass == 0 ? 416u : ahi) + (uint32_t)half * KC;
                            u64 bd = tc_sdesc((pass == 1 && term == 1) ? blo : bhi, TC8_LBO, TC_SBO);
#pragma unroll 1
                            for (int kg = 0; kg < KGu; ++kg) {
                                const uint32_t acc = !(first && pass == 0 && kg == 0);
                                asm volatile("{\n.reg .pred p, pe;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|pe, 0xffffffff;\n"
                                             "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tb + 128u * half),
                                             "r"(a), "l"(bd), "r"(tc_idesc(128, T)), "r"(acc)
                                             : "memory");
                                a += 8u;
                                bd += (u64)(TC8_KGB >> 4);
                            }
                        }
                    }
                    // completion of everything issued so far: pass 0 -> "residual stage read" (barrier 7), pass 1 -> "weight
                    // stage free"
                    asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
                                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(pass == 0 ? BAR(7) : BAR(8 + sw))
                                 : "memory");
                }
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
                                 : "r"(tbase + lane_base + 128u * half + c0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    uint4 *o = reinterpret_cast<uint4 *>(Ds + (size_t)r * T + c0);
                    o[0] = make_uint4(v[0], v[1], v[2], v[3]);
                    o[1] = make_uint4(v[4], v[5], v[6], v[7]);
                    o[2] = make_uint4(v[8], v[9], v[10], v[11]);
                    o[3] = make_uint4(v[12], v[13], v[14], v[15]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(15));
        }
        return;
    }

    // =================================== consumers ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC8_CREG));
    // warp = 2 * (x-block of 32 columns) + (disparity block of 64); lane = 8 * xl + dl: x-group (of 8 columns) 4 wx + xl, disparity
    // group (of 8) 8 wd + dl, rotated by xl so that the right-weight index R0 is the same in the four quarter-warps
    const int xl = lane >> 3, dl = lane & 7;
    const int wx = warp >> 1, wd = warp & 1;
    const int xg = wx * 4 + xl;
    const int dg = (wd * 8 + dl + xl) % (DC / 8);
    const int xb = 8 * xg, kb = 8 * dg;
    const bool sub_ok = xsub < 0 || wx == xsub;                // warp-uniform
    const bool lane_live = sub_ok && (x0 + xb < g.W) && (dlo + kb <= g.dHi) && (dlo + kb + 7 >= g.dVLo) && (x0 + xb + 7 >= dlo + kb);
    const bool warp_live = __any_sync(0xffffffffu, lane_live);
    const int R0 = T - 8 - xb + kb;                            // first reversed right-centre index (multiple of 8)
    int w2o[4];                                                // float offsets of the lane's four right-weight chunks
#pragma unroll
    for (int c = 0; c < 4; ++c) w2o[c] = tc8_w2idx(R0 + 4 * c);

    u64 acc0[8][4];                                            // numerators: 8 columns x 4 disparity pairs
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc0[a][b] = 0ull;

    int sw = 0, phw = 0;
    for (int n = 0; n < nsteps; ++n) {
        const int st = n & 1, ph = (n >> 1) & 1;
        mbar_wait(BAR(5 + sw), phw);             // weights of this window row
        mbar_wait(BAR(11 + st), ph);             // raw costs of this window row
        if (warp_live && !(P.freerun & 4)) {
            u64 ring[8][4];                      // sliding window of 8 cost columns x 8 disparities
            const uint8_t *ep = smem + (sp.e + st * sp.ebytes) + xb * EP + kb;
            auto load_e = [&](const uint8_t *q, u64 (&r)[4]) {
                const uint2 e = *reinterpret_cast<const uint2 *>(q);
                r[0] = pk(u8_to_f32(e.x, 0), u8_to_f32(e.x, 1));
                r[1] = pk(u8_to_f32(e.x, 2), u8_to_f32(e.x, 3));
                r[2] = pk(u8_to_f32(e.y, 0), u8_to_f32(e.y, 1));
                r[3] = pk(u8_to_f32(e.y, 2), u8_to_f32(e.y, 3));
            };
#pragma unroll
            for (int a = 0; a < 7; ++a) load_e(ep + a * EP, ring[a]);
            ep += 7 * EP;
            // left weights: K-major operand, one 16-byte chunk = one column x, 4 consecutive window offsets
            const unsigned char *w1q = smem + (sp.w1 + sw * sp.w1arr) + xg * TC_SBO;
            const float *w2p = reinterpret_cast<const float *>(smem + (sp.w2 + sw * sp.w2bytes));
            float4 wq[8];

            auto step = [&](auto sc) {
                constexpr int s = decltype(sc)::value;
                load_e(ep + s * EP, ring[(7 + s) & 7]);
                if (s % 4 == 0) {
#pragma unroll
                    for (int a = 0; a < 8; ++a) wq[a] = *reinterpret_cast<const float4 *>(w1q + (s / 4) * TC8_LBO + a * 16);
                }
                const float4 v0 = *reinterpret_cast<const float4 *>(w2p + s * NRp + w2o[0]);
                const float4 v1 = *reinterpret_cast<const float4 *>(w2p + s * NRp + w2o[1]);
                const float4 v2 = *reinterpret_cast<const float4 *>(w2p + s * NRp + w2o[2]);
                const float4 v3 = *reinterpret_cast<const float4 *>(w2p + s * NRp + w2o[3]);
                const u64 VA[8] = {pk(v0.x, v0.y), pk(v0.z, v0.w), pk(v1.x, v1.y), pk(v1.z, v1.w),
                                   pk(v2.x, v2.y), pk(v2.z, v2.w), pk(v3.x, v3.y), pk(v3.z, v3.w)};
                const float vf[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const float w1 = (s % 4 == 0) ? wq[a].x : (s % 4 == 1) ? wq[a].y : (s % 4 == 2) ? wq[a].z : wq[a].w;
                    const u64 w1d = pk(w1, w1);
#pragma unroll
                    for (int bp = 0; bp < 4; ++bp) {
                        const int k = 7 - a + 2 * bp;                   // reversed right index of disparity kb + 2 bp
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
                ep += 8 * EP;
                w1q += TC8_KGB;
                w2p += 8 * NRp;
            };
            // two periods per loop trip: the loop-carried register moves (ring + prefetched loads) are paid once per trip
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

    // ---- epilogue: denominators from the tensor core, normalise, WTA over the chunk (both references), optional volume ----
    mbar_wait(BAR(15), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
    float *Ds = reinterpret_cast<float *>(smem + sp.ds);
    const float fix_kg = TC_TRUNC_PER_MMA * (float)(3 * nsteps);   // per K-group that contributes; see ss_aggregate_tc.cuh
    const int rowo = y - g.row0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int x = x0 + xb + a;
        float c0[8];
#pragma unroll
        for (int b = 0; b < 4; ++b) upk(acc0[a][b], c0[2 * b], c0[2 * b + 1]);
        u64 best = KEY_NONE;
        float out0[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
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
            *reinterpret_cast<float4 *>(P.vol0 + o + 4) = make_float4(out0[4], out0[5], out0[6], out0[7]);
        }
    }
    if (P.bestR) {
        asm volatile("bar.sync 2, %0;" ::"n"(CW * 32) : "memory");
        wta_right_rows<T, DC>(Ds, warp, CW, lane, x0, dlo, g.W, P.bestR + (size_t)rowo * g.W);
    }
}
