// Tensor-core path (sm_100a): one persistent kernel per block of MH updates.
//
//   rows      one surrogate row = (chain, time point); a tile is 128 rows = the 128 TMEM lanes
//   layers    all six tanh layers run as tcgen05.mma kind::f16 GEMMs (M=128, N=64) with FP32
//             accumulators in TMEM.  FP32 accuracy comes from operand splitting
//             (SURVEY.md App. C):  a = a_hi + a_lo,  W = W_hi + W_lo  (fp16 each)
//                 D = a_hi*W_hi + a_lo*W_hi + a_hi*W_lo            (3 MMAs per logical MMA)
//             Activations never leave the SM: the epilogue reads D with tcgen05.ld, applies tanh,
//             splits, and writes the next A operand straight back to TMEM (tcgen05.st); A is consumed
//             from TMEM (".ts" form), weights from shared memory (128B-swizzled, K-major, resident
//             for the whole run, staged once with a bulk-TMA copy).
//   bias      added by the tensor core: one extra K=16 MMA of a constant ones-block (shared memory)
//             against [b_hi; b_mid; b_lo]; 2*log2(e) is folded into weights and biases on the host so the
//             accumulator is directly the exp2 argument of tanh(z) = 1 - 2/(2^(2z*log2 e) + 1).
//   layer 1   the (P+1)->64 layer is a K=32 MMA over three-way-split scaled inputs.
//   pipeline  4 tiles in flight per SM (4 x 128 TMEM columns), one per team of 8 warps (mems_tc_device.cuh: two groups
//             of two teams); the warp that completes a team's A operand issues the team's MMAs, so every warp of
//             the CTA is an epilogue worker and the teams cover each other's accumulator waits.
//   tail      64->1 output layer, residual against y_obs, squared-sum reduction in FP64, accept/reject
//             and chain-state update (mems_common.cuh) all happen in the same kernel.
//
// Reference semantics: see mems_common.cuh header (src/InverseProblems/utils.py:30-40 etc.).
#include "mems_tc_device.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// MMA issue (one thread)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void issue_layer(uint32_t wimg_saddr, uint32_t d, int layer) {
    if (layer == 0) {
        const uint64_t b = make_desc(wimg_saddr + BLK_B1 * BLK);
        umma_ts(d, d + 64, b, 0u);
        umma_ts(d, d + 64 + 8, b + 2, 1u);
    } else {
        const int h = layer - 1;
        // Accumulation order matters: the tensor core aligns every product to the largest exponent in
        // flight (accumulator included) and truncates ~25 bits below it (measured: tools/probe_numerics.py).
        // Small correction terms therefore go first, among themselves; the large hi*hi products and the
        // bias come last, so the corrections are truncated once per instruction instead of once per product.
        const uint64_t whi = make_desc(wimg_saddr + (BLK_WHI + h) * BLK);
        const uint64_t wlo = make_desc(wimg_saddr + (BLK_WLO + h) * BLK);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 96 + 8 * ks, whi + 2 * ks, ks > 0 ? 1u : 0u);   // a_lo * W_hi
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 64 + 8 * ks, wlo + 2 * ks, 1u);                 // a_hi * W_lo
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 64 + 8 * ks, whi + 2 * ks, 1u);                 // a_hi * W_hi
        const uint64_t ones = make_desc(wimg_saddr + BLK_ONES * BLK);
        const uint64_t bias = make_desc(wimg_saddr + (BLK_BIAS + (h >> 2)) * BLK) + 2 * (h & 3);
        umma_ss(d, ones, bias, 1u);                                                                    // + bias
    }
}

// Who am I: see the CTA organisation in mems_tc_device.cuh.
struct Role {
    int g;      // group
    int s;      // team of the group = TMEM tile slot
    int h;      // column half
    int q;      // TMEM lane quarter
    int row;    // TMEM lane = row of the team's tile
    int gt;     // thread index within the group, 0..GTHREADS-1 (gt < 128: team 0, column half 0)
};
__device__ __forceinline__ Role my_role() {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Role r;
    r.q = warp & 3;
    const int j = warp >> 2;
    r.g = j & 1;
    const int gi = j >> 1;               // 0..3 within the group
    r.s = gi & 1;
    r.h = gi >> 1;
    r.row = r.q * 32 + lane;
    r.gt = gi * 128 + r.row;
    return r;
}

__device__ __forceinline__ void group_bar_sync(int g) { named_bar_sync(1 + g, GTHREADS); }

// Every thread of a team calls this when its part of the team's next A operand is in TMEM (tcgen05.st complete).
// The warp whose arrival completes the operand (the eighth) issues the MMAs of `layer` on the team's slot and commits
// them to the team's accumulator barrier.  The arrival counter only ever grows (no reset races): "eighth" = count % 8.
__device__ __forceinline__ void team_arrive_issue(TcSmem& sm, GroupSmem& gs, const Role& me, int layer) {
    tc_fence_before();
    __syncwarp();
    uint32_t last = 0;
    if ((threadIdx.x & 31) == 0) {
        uint32_t old;
        asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;"
                     : "=r"(old) : "r"(smem_u32(&gs.arrive_cnt[me.s])) : "memory");
        last = ((old & (uint32_t)(TEAM_WARPS - 1)) == (uint32_t)(TEAM_WARPS - 1)) ? 1u : 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {                                               // warp-uniform
        tc_fence_after();
        const uint32_t slot = __shfl_sync(0xffffffffu, sm.tmem_base, 0) + (uint32_t)((me.g * GTEAM + me.s) * 128);
        if (elect_one()) {
            issue_layer(smem_u32(sm.wimg), slot, layer);
            umma_commit(&gs.bar_D[me.s]);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Worker side: evaluate S[j] = sum_t (y_obs[t] - f(x_j)[t])^2 for the nc proposals in cb.xs
// ------------------------------------------------------------------------------------------
// K-slot of (term, input) in the layer-1 operand: terms 0..2 multiply W_hi, 3..4 multiply W_lo
template <int P>
__device__ __forceinline__ constexpr int k1_slot(int term, int i) { return term * (P + 1) + i; }

template <int P>
__device__ __forceinline__ void set_half(uint32_t (&a1)[16], int slot, float v) {
    const unsigned short hb = __half_as_ushort(__float2half_rn(v));
    const int w = slot >> 1;
    if (slot & 1) a1[w] = (a1[w] & 0x0000FFFFu) | ((uint32_t)hb << 16);
    else a1[w] = (a1[w] & 0xFFFF0000u) | (uint32_t)hb;
}

template <int P>
__device__ __forceinline__ void put_input(uint32_t (&a1)[16], int i, float v) {
    float h, m, l;
    split3(v, h, m, l);
    set_half<P>(a1, k1_slot<P>(0, i), h);
    set_half<P>(a1, k1_slot<P>(1, i), m);
    set_half<P>(a1, k1_slot<P>(2, i), l);
    set_half<P>(a1, k1_slot<P>(3, i), h);
    set_half<P>(a1, k1_slot<P>(4, i), m);
}

// The threads of one group evaluate one batch; team s takes tiles s, s + GTEAM, ...  Thread (row, h) of a team owns TMEM
// lane `row` of the team's slot and columns [32 h, 32 h + 32) of it; row `row` of every tile is (chain slot row mod nce,
// time sub-row row / nce).
// compact: only the gs.nlive chains listed in gs.map take part (nce = gs.nlive): proposals outside the prior box and,
// in the delayed-acceptance second stage, proposals the first stage did not promote cost nothing -- the tiles are
// rebuilt from the live chains, so the pass shrinks with the promotion rate.  Result: gs.Sfin[chain].
// DENSE (compact passes only): (time, chain) pairs are laid out back to back, row i of tile k is pair f = 128k + i
// = (time f / n_live, chain slot f mod n_live), so every tile is full even when n_live does not divide 128.  A thread
// then meets a different chain in every tile: the squared residuals of a tile go through shared memory (gs.R2) and
// thread c < n_live of the team adds the entries of chain slot c in time order (deterministic).
template <int P, bool DENSE>
__device__ __forceinline__ void tc_eval_group(TcSmem& sm, GroupSmem& gs, const Role& me, const double* s_yobs,
                                              const float* s_ts, int T, int stride, int nc, int ncv,
                                              bool compact, float* y_out, int c0, uint32_t& ph) {
    constexpr int NCOL = 64 / NH;                     // accumulator columns per thread
    const int row = me.row, h = me.h, q = me.q, s = me.s, gt = me.gt;
    const int nce = compact ? gs.nlive : nc;          // chains laid out in the tiles (any value in 1..128)
    const int jc = row % nce;                         // rows >= nce*tpt of a tile idle
    const int sub = row / nce;
    const int tpt = 128 / nce;                        // time points per tile
    const int j = compact ? (int)gs.map[jc] : jc;     // batch-local chain index
    const int Tn = (T + stride - 1) / stride;         // time points of this pass: t = 0, stride, 2*stride, ...
    const int n_tiles = DENSE ? (nce * Tn + 127) / 128 : (Tn + tpt - 1) / tpt;
    const uint32_t tD = sm.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((me.g * GTEAM + s) * 128);
    const bool chain_live = (sub < tpt) && (compact || jc < ncv);

    uint32_t a1[16];
    if (h == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a1[i] = 0u;
        if (!DENSE) {
#pragma unroll
            for (int p = 0; p < P; ++p) put_input<P>(a1, p, (compact || jc < ncv) ? gs.cb.xs[p][j] : 0.0f);
        }
        set_half<P>(a1, 5 * (P + 1) + 0, 1.0f);       // ones that pick up b_hi, b_mid, b_lo
        set_half<P>(a1, 5 * (P + 1) + 1, 1.0f);
        set_half<P>(a1, 5 * (P + 1) + 2, 1.0f);
    }
    // layer-1 operand of `tile` into the team's slot (column-half 0 threads)
    auto stage_tile = [&](int tile) {
        if (DENSE) {
            const int f = tile * 128 + row;
            const int ti = f / nce;
            const bool ok = ti < Tn;
            const int jj = ok ? (int)gs.map[f - ti * nce] : 0;
#pragma unroll
            for (int p = 0; p < P; ++p) put_input<P>(a1, p, ok ? gs.cb.xs[p][jj] : 0.0f);
            put_input<P>(a1, P, s_ts[ok ? ti * stride : 0]);
        } else {
            const int ti = tile * tpt + sub;
            put_input<P>(a1, P, s_ts[(ti < Tn && sub < tpt) ? ti * stride : 0]);
        }
        tmem_st16(tD + 64, a1);
        tc_wait_st();
    };

    double Sp = 0.0;
    for (int tile = s; tile < n_tiles; tile += GTEAM) {
        if (h == 0) stage_tile(tile);
        team_arrive_issue(sm, gs, me, 0);
#pragma unroll 1
        for (int layer = 0; layer < MEMS_HIDDEN; ++layer) {
            mbar_wait(&gs.bar_D[s], ph & 1u);
            ph ^= 1u;
            tc_fence_after();
            // the thread's NCOL accumulator columns in pieces of EC: load, tanh, hi/lo split, store (a finer
            // grain than the whole row keeps XU, FMA and ALU work of one warp interleaved: profiles/r02_microbench_*)
            constexpr int EC = (MEMS_EPI_COLS < NCOL) ? MEMS_EPI_COLS : NCOL;
            if (layer < MEMS_HIDDEN - 1) {
#pragma unroll
                for (int pc = 0; pc < NCOL / EC; ++pc) {
                    uint32_t w[EC], hi2[EC / 2], lo2[EC / 2];
                    tmem_ld_n<EC>(tD + NCOL * h + EC * pc, w);
                    tc_wait_ld();
                    epilogue_split_v2<EC / 8, (MEMS_EPI_R16 != 0) && (EC >= 16), MEMS_EPI_POLY8>(w, hi2, lo2);
                    tmem_st_n<EC / 2>(tD + 64 + (NCOL / 2) * h + (EC / 2) * pc, hi2);
                    tmem_st_n<EC / 2>(tD + 96 + (NCOL / 2) * h + (EC / 2) * pc, lo2);
                }
                tc_wait_st();
                team_arrive_issue(sm, gs, me, layer + 1);
            } else {
                uint64_t acc = pk(0.0f, 0.0f);
#pragma unroll
                for (int pc = 0; pc < NCOL / EC; ++pc) {
                    uint32_t w[EC];
                    tmem_ld_n<EC>(tD + NCOL * h + EC * pc, w);
                    tc_wait_ld();
#pragma unroll
                    for (int g8 = 0; g8 < EC / 8; ++g8)
                        acc = epilogue_dot8(w + 8 * g8, sm.w7 + NCOL * h + EC * pc + 8 * g8, acc);
                }
                float oa, ob;
                upk(acc, oa, ob);
                float o = oa + ob;
                if (h == 1) {                              // the upper column half hands its partial dot to the lower one
                    gs.xchg[s][row] = o;
                    mbar_arrive(&gs.bar_X[s][q]);          // release: the store above is visible to the waiter
                } else {
                    mbar_wait(&gs.bar_X[s][q], (ph >> 1) & 1u);   // bit 1 of ph: bar_X parity
                    ph ^= 2u;
                    o += gs.xchg[s][row];
                    o += sm.b7;
                    if (DENSE) {
                        const int f = tile * 128 + row;
                        const int ti = f / nce;
                        double r2 = 0.0;
                        if (ti < Tn) {
                            const double r = s_yobs[ti * stride] - (double)o;
                            r2 = r * r;
                        }
                        gs.R2[s][row] = r2;
                        named_bar_sync(3 + me.g * GTEAM + s, 128);     // the tile's 128 residuals are in shared memory
                        if (row < nce) {                               // chain slot `row`: its rows of this tile, in time order
                            int r0 = row - (tile * 128) % nce;
                            if (r0 < 0) r0 += nce;
                            for (int rr = r0; rr < 128; rr += nce) Sp += gs.R2[s][rr];
                        }
                        // R2[s] is rewritten only after the team's next tile has gone through six accumulator barriers,
                        // which every reader above has to pass as well
                    } else {
                        const int ti = tile * tpt + sub;
                        const int t = ti * stride;
                        if (chain_live && ti < Tn) {
                            if (y_out) y_out[(size_t)(c0 + j) * T + t] = o;
                            const double r = s_yobs[t] - (double)o;
                            Sp = fma(r, r, Sp);
                        }
                    }
                }
                tc_fence_before();                         // the slot's D and A are free again (next tile: stage + arrive)
            }
        }
    }
    if (DENSE) {
        if (h == 0 && row < nce) gs.Sdense[s][row] = Sp;   // thread `row` of each team summed chain slot `row` over the team's tiles
        group_bar_sync(me.g);
        if (gt < nce) gs.Sfin[gs.map[gt]] = gs.Sdense[0][gt] + gs.Sdense[1][gt];
    } else {
        if (h == 0) gs.Spart[s][row] = Sp;
        group_bar_sync(me.g);
        if (gt < nce) {
            double S = 0.0;                                // fixed order: time sub-rows ascending, teams inside
            for (int u = 0; u < tpt; ++u) S += gs.Spart[0][u * nce + gt] + gs.Spart[1][u * nce + gt];
            gs.Sfin[compact ? (int)gs.map[gt] : gt] = S;
        }
    }
    group_bar_sync(me.g);                                  // Sfin[chain] may have been written by another thread
}

// Compaction of the chains that need the coming pass: gs.map / gs.nlive (first warp of the group; ballot prefix sums).
__device__ __forceinline__ void group_compact_live(GroupSmem& gs, int g, int gt, int ncv) {
    group_bar_sync(g);                                 // live[] of every chain is written
    if (gt < 32) {
        int base = 0;
#pragma unroll
        for (int c = 0; c < MEMS_NCMAX / 32; ++c) {
            const int jj = 32 * c + gt;
            const bool lv = jj < ncv && gs.cb.live[jj] == 1;
            const unsigned m = __ballot_sync(0xffffffffu, lv);
            if (lv) gs.map[base + __popc(m & ((1u << gt) - 1u))] = (unsigned char)jj;
            base += __popc(m);
        }
        if (gt == 0) gs.nlive = base;
    }
    group_bar_sync(g);
}

// Common prologue: barriers, TMEM, weight image (bulk TMA), y_obs / time columns.
__device__ __forceinline__ void tc_prologue(TcSmem& sm, const MemsProblem& pb, double* s_yobs, float* s_ts) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    if (tid == 0) {
        for (int g = 0; g < NGROUP; ++g)
            for (int s = 0; s < GTEAM; ++s) {
                mbar_init(&sm.grp[g].bar_D[s], 1);
                for (int q = 0; q < 4; ++q) mbar_init(&sm.grp[g].bar_X[s][q], 32);
                sm.grp[g].arrive_cnt[s] = 0u;
            }
        mbar_init(&sm.bar_w, 1);
        sm.c = pb.c;
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&sm.tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        if ((tid & 31) == 0) {
            mbar_arrive_expect_tx(&sm.bar_w, (uint32_t)WIMG_BYTES);
            const unsigned char* src = reinterpret_cast<const unsigned char*>(pb.wimg);
            for (int b = 0; b < NBLK; ++b) bulk_g2s(sm.wimg + b * BLK, src + (size_t)b * BLK, BLK, &sm.bar_w);
            bulk_g2s(sm.w7, src + (size_t)NBLK * BLK, WIMG_TAIL, &sm.bar_w);
        }
    }
    for (int i = tid; i < pb.c.T; i += NTHREADS) { s_yobs[i] = pb.yobs[i]; s_ts[i] = pb.ts[i]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(&sm.bar_w, 0);
}

__device__ __forceinline__ void tc_epilogue_free(TcSmem& sm) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem_base), "r"(512) : "memory");
    }
}

template <int P>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_run_kernel(const MemsProblem* __restrict__ pbp, MemsChains ch, MemsRunArgs ra) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
    const MemsProblem& pb = *pbp;
    double* s_yobs = reinterpret_cast<double*>(smem_raw + sizeof(TcSmem));
    float* s_ts = reinterpret_cast<float*>(s_yobs + pb.c.T);
    tc_prologue(sm, pb, s_yobs, s_ts);

    const Role me = my_role();
    const int T = sm.c.T, nc = ra.nc;
    const int nbatch = (ch.C + nc - 1) / nc;
    const int g = me.g, gt = me.gt;
    GroupSmem& gs = sm.grp[g];
    const int gid = blockIdx.x * NGROUP + g, gstride = gridDim.x * NGROUP;
    uint32_t ph = 0;
    for (int batch = gid; batch < nbatch; batch += gstride) {
        const int c0 = batch * nc;
        const int ncv = min(nc, ch.C - c0);
        group_bar_sync(g);
        batch_load(gs.cb, ch, P, ra.mode, c0, gt, ncv);
        for (int s = 0; s < ra.n_steps; ++s) {
            for (int sub = 0; sub < ra.n_sub; ++sub) {
                if (gt < ncv) batch_pre(gs.cb, sm.c, ch, ra, c0, gt, s, sub);
                // only the chains that need this pass are laid out in tiles (proposals outside the prior box and
                // proposals the DA first stage did not promote drop out); no live chain -> the pass is skipped
                group_compact_live(gs, g, gt, ncv);
                const bool any = gs.nlive > 0;
                if (any && batch_pass_is_coarse_model(ra, sub))   // DA first stage on the FFT surrogate (FP32, CUDA cores)
                    coarse_fft_pass(ra.coarse, P, T, &gs.cb.xp[0][0], MEMS_NCMAX, gs.map, gs.nlive, gs.Sfin, coarse_scratch(gs.cb), gt, GTHREADS,
                                    [g] { group_bar_sync(g); });
                else if (any) {
                    // dense packing when it saves tiles (n_live does not divide 128 well)
                    const int stride = batch_pass_stride(ra, sub);
                    const int Tn = (T + stride - 1) / stride, tpt = 128 / gs.nlive;
                    if ((gs.nlive * Tn + 127) / 128 < (Tn + tpt - 1) / tpt)
                        tc_eval_group<P, true>(sm, gs, me, s_yobs, s_ts, T, stride, nc, ncv, true, nullptr, c0, ph);
                    else
                        tc_eval_group<P, false>(sm, gs, me, s_yobs, s_ts, T, stride, nc, ncv, true, nullptr, c0, ph);
                }
                if (gt < ncv) batch_post(gs.cb, sm.c, ch, ra, c0, gt, s, sub, (any && gs.cb.live[gt] == 1) ? gs.Sfin[gt] : 0.0);
            }
        }
        batch_store(gs.cb, ch, P, ra.mode, c0, gt, ncv);
    }
    tc_epilogue_free(sm);
}

template <int P>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_forward_kernel(const MemsProblem* __restrict__ pbp, const double* __restrict__ x, int n, int nc, float* y_out,
                  double* ll_out, int apply_prior, int stride) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
    const MemsProblem& pb = *pbp;
    double* s_yobs = reinterpret_cast<double*>(smem_raw + sizeof(TcSmem));
    float* s_ts = reinterpret_cast<float*>(s_yobs + pb.c.T);
    tc_prologue(sm, pb, s_yobs, s_ts);

    const Role me = my_role();
    const int T = sm.c.T;
    const int nbatch = (n + nc - 1) / nc;
    const int g = me.g, gt = me.gt;
    GroupSmem& gs = sm.grp[g];
    const int gid = blockIdx.x * NGROUP + g, gstride = gridDim.x * NGROUP;
    uint32_t ph = 0;
    for (int batch = gid; batch < nbatch; batch += gstride) {
        const int c0 = batch * nc;
        const int ncv = min(nc, n - c0);
        group_bar_sync(g);
        if (gt < ncv) {
            bool inb = true;
            for (int p = 0; p < P; ++p) {
                const double v = x[(size_t)p * n + c0 + gt];
                gs.cb.xs[p][gt] = scale_param(sm.c, p, v);
                inb = inb && !(v < sm.c.low[p]) && !(v > sm.c.high[p]);
            }
            gs.cb.inb[gt] = inb ? 1 : 0;
            gs.cb.live[gt] = 1;
        }
        group_bar_sync(g);
        tc_eval_group<P, false>(sm, gs, me, s_yobs, s_ts, T, stride, nc, ncv, false, y_out, c0, ph);
        if (gt < ncv && ll_out) {
            // a strided (coarse) pass is rescaled to the full trace length, as in the DA first stage
            const int Tn = (T + stride - 1) / stride;
            const double ll = sm.c.neg_half_inv_sigma2 * gs.Sfin[gt] * ((double)T / (double)Tn);
            ll_out[c0 + gt] = (apply_prior && !gs.cb.inb[gt]) ? -CUDART_INF : ll;
        }
    }
    tc_epilogue_free(sm);
}

size_t tc_smem_bytes(int T) { return sizeof(TcSmem) + (size_t)T * (sizeof(double) + sizeof(float)) + 16; }

// host-side fp16 helpers (round to nearest even through the CUDA host implementation)
inline unsigned short h_bits(double v) {
    const __half h = __float2half_rn((float)v);
    unsigned short b;
    memcpy(&b, &h, 2);
    return b;
}
inline double h_val(double v) { return (double)__half2float(__float2half_rn((float)v)); }

inline void put(unsigned char* img, int blk, int row, int k, double v) {
    const size_t off = (size_t)blk * BLK + (size_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
    const unsigned short b = h_bits(v);
    memcpy(img + off, &b, 2);
}

template <typename K1, typename K2, typename K3, typename K4, typename... Args>
int launch_by_p(int Pv, K1 k1, K2 k2, K3 k3, K4 k4, int grid, size_t smem, cudaStream_t st, Args... args) {
    auto go = [&](auto k) -> int {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<grid, NTHREADS, smem, st>>>(args...);
        return (int)cudaGetLastError();
    };
    switch (Pv) {
        case 1: return go(k1);
        case 2: return go(k2);
        case 3: return go(k3);
        default: return go(k4);
    }
}

}  // namespace

size_t mems_tc_wimg_bytes() { return WIMG_BYTES; }

// largest trace length whose observation / time columns fit next to the resident weights in `smem_optin` bytes
int mems_tc_max_time_points(size_t smem_optin) {
    const size_t fixed = tc_smem_bytes(0);
    if (smem_optin <= fixed) return 0;
    return (int)((smem_optin - fixed) / (sizeof(double) + sizeof(float)));
}

// Chains per group batch (any value 1..128): minimise (waves of batches over the 2*SMs groups) x (tile pairs per step).
int mems_tc_choose_nc(int C, int T, int sm_count) {
    int best = 1;
    long long best_cost = -1;
    for (int nc = 1; nc <= MEMS_NCMAX; ++nc) {
        const long long nbatch = (C + nc - 1) / nc;
        const long long waves = (nbatch + (long long)NGROUP * sm_count - 1) / ((long long)NGROUP * sm_count);
        const int tpt = 128 / nc;
        const long long tiles = (T + tpt - 1) / tpt;
        const long long cost = waves * ((tiles + 1) / 2) * 2;
        if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = nc; }
    }
    return best;
}

// Build the shared-memory image: 128B-swizzled fp16 operand blocks + the FP32 output layer.
void mems_tc_build_wimg(const MemsProblem& pb, void* host_img) {
    unsigned char* img = static_cast<unsigned char*>(host_img);
    memset(img, 0, WIMG_BYTES);
    const int P = pb.c.P, nin = P + 1;
    // layer 1: rows n (output neuron), K slots as in k1_slot(): terms 0..2 carry W_hi, 3..4 W_lo
    for (int n = 0; n < MEMS_WIDTH; ++n) {
        for (int i = 0; i < nin; ++i) {
            const double w = C2 * (double)pb.W1[i * MEMS_WIDTH + n];
            const double whi = h_val(w), wlo = h_val(w - whi);
            put(img, BLK_B1, n, 0 * nin + i, whi);
            put(img, BLK_B1, n, 1 * nin + i, whi);
            put(img, BLK_B1, n, 2 * nin + i, whi);
            put(img, BLK_B1, n, 3 * nin + i, wlo);
            put(img, BLK_B1, n, 4 * nin + i, wlo);
        }
        const double b = C2 * (double)pb.b1[n];
        const double bh = h_val(b), bm = h_val(b - bh), bl = h_val(b - bh - bm);
        put(img, BLK_B1, n, 5 * nin + 0, bh);
        put(img, BLK_B1, n, 5 * nin + 1, bm);
        put(img, BLK_B1, n, 5 * nin + 2, bl);
    }
    // hidden layers: B[n][k] = 2log2e * W[k][n]
    for (int h = 0; h < MEMS_NLAYER_H; ++h) {
        for (int n = 0; n < MEMS_WIDTH; ++n) {
            for (int k = 0; k < MEMS_WIDTH; ++k) {
                const double w = C2 * (double)pb.Wh[(h * MEMS_WIDTH + k) * MEMS_WIDTH + n];
                const double whi = h_val(w);
                put(img, BLK_WHI + h, n, k, whi);
                put(img, BLK_WLO + h, n, k, w - whi);
            }
            const double b = C2 * (double)pb.bh[h * MEMS_WIDTH + n];
            const double bh = h_val(b), bm = h_val(b - bh), bl = h_val(b - bh - bm);
            const int blk = BLK_BIAS + (h >> 2), k0 = 16 * (h & 3);
            put(img, blk, n, k0 + 0, bh);
            put(img, blk, n, k0 + 1, bm);
            put(img, blk, n, k0 + 2, bl);
        }
    }
    for (int m = 0; m < 128; ++m)
        for (int k = 0; k < 3; ++k) put(img, BLK_ONES + (m >> 6), m & 63, k, 1.0);
    float* tail = reinterpret_cast<float*>(img + (size_t)NBLK * BLK);
    for (int k = 0; k < MEMS_WIDTH; ++k) {          // per group of eight: (w0,w4, w1,w5, w2,w6, w3,w7), see epilogue_dot
        const int g8 = k & ~7, i = k & 7;
        const int perm = (i < 4) ? 2 * i : 2 * (i - 4) + 1;
        tail[g8 + perm] = pb.w7[k];
    }
    tail[MEMS_WIDTH] = pb.b7;
}

int mems_tc_run(const MemsLaunchCtx& ctx, const MemsRunArgs& ra) {
    const size_t smem = tc_smem_bytes(ctx.pb_host->c.T);
    const int nbatch = (ctx.chains.C + ra.nc - 1) / ra.nc;
    const int want = (nbatch + NGROUP - 1) / NGROUP;
    const int grid = want < ctx.sm_count ? want : ctx.sm_count;
    return launch_by_p(ctx.pb_host->c.P, tc_run_kernel<1>, tc_run_kernel<2>, tc_run_kernel<3>, tc_run_kernel<4>,
                       grid, smem, ctx.stream, ctx.pb_dev, ctx.chains, ra);
}

int mems_tc_forward(const MemsLaunchCtx& ctx, const double* x_dev, int n, float* y_out, double* ll_out,
                    int apply_prior, int t_stride) {
    if (t_stride < 1) t_stride = 1;
    const size_t smem = tc_smem_bytes(ctx.pb_host->c.T);
    const int nc = mems_tc_choose_nc(n, ctx.pb_host->c.T, ctx.sm_count);
    const int nbatch = (n + nc - 1) / nc;
    const int want = (nbatch + NGROUP - 1) / NGROUP;
    const int grid = want < ctx.sm_count ? want : ctx.sm_count;
    const int gr = grid > 0 ? grid : 1;
    return launch_by_p(ctx.pb_host->c.P, tc_forward_kernel<1>, tc_forward_kernel<2>, tc_forward_kernel<3>,
                       tc_forward_kernel<4>, gr, smem, ctx.stream, ctx.pb_dev, x_dev, n, nc, y_out, ll_out, apply_prior, t_stride);
}
