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
//   pipeline  4 tile slots in flight per SM (4 x 128 TMEM columns) in two independent groups; per group eight
//             hidden-layer warps (all six tanh epilogues; the last one fused with the 64->1 output layer), four tail
//             warps (residual + next tile's layer-1 operand + per-chain bookkeeping) and one MMA-issuer warp per slot
//             (mems_tc_device.cuh), handing a slot round by mbarriers.
//   tail      residual against y_obs, squared-sum reduction in FP64, accept/reject and chain-state update
//             (mems_common.cuh) all happen in the same kernel.
//   coarse    delayed acceptance with the FFT surrogate: its one-row-per-proposal network runs through the same
//             pipeline, operand blocks streamed from L2 layer by layer (issue_coarse_pass / tc_coarse_group).
//   switches  MEMS_NGROUP, MEMS_TAIL_SHARE, MEMS_PRESTAGE, MEMS_BIAS_PRELOAD, MEMS_ISSUER_PER_SLOT, MEMS_DENSE_MIN_SAVING:
//             organisations that were measured against the default (profiles/r02_ab_variants.log);
//             MEMS_TRACE: development builds that record a hand-off timeline (tools/trace_timeline.py).
//
// Reference semantics: see mems_common.cuh header (src/InverseProblems/utils.py:30-40 etc.).
#include "mems_tc_device.cuh"

#ifndef MEMS_TRACE
#define MEMS_TRACE 0           // 1 (development builds only): block 0 records a timeline of its hand-offs, read with mems_trace_read
#endif
#if MEMS_TRACE
__device__ unsigned long long g_trace[32][8192];
__device__ unsigned int g_trace_n[32];
__device__ __forceinline__ void trace_ev(int tag) {         // ~40 cycles: clock read, shared-memory counter, fire-and-forget store
    __shared__ unsigned int s_n[32];
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        unsigned int n = s_n[w];
        if (n >= 0x80000000u || n == 0xFFFFFFFFu) n = 0;      // (uninitialised shared memory is handled by the marker below)
        if ((tag & 15) == 15) { s_n[w] = 0; return; }         // kind 15: reset at kernel start
        if (n < 8192u) { g_trace[w][n] = ((unsigned long long)clock64() << 12) | (unsigned long long)(tag & 0xFFF); s_n[w] = n + 1u; g_trace_n[w] = n + 1u; }
    }
}
extern "C" int mems_trace_read(unsigned long long* host_buf, unsigned int* host_n) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_buf, g_trace, sizeof(g_trace));
    cudaMemcpyFromSymbol(host_n, g_trace_n, sizeof(g_trace_n));
    unsigned int zero[32] = {0};
    cudaMemcpyToSymbol(g_trace_n, zero, sizeof(zero));
    return 0;
}
#define TRACE(kind, s, layer) trace_ev((kind) | ((s) << 4) | ((layer) << 6))
#else
#define TRACE(kind, s, layer)
#endif

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
        for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 96 + 8 * ks, whi + 2 * ks, (ks > 0 || MEMS_BIAS_PRELOAD) ? 1u : 0u);   // a_lo * W_hi
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 64 + 8 * ks, wlo + 2 * ks, 1u);                 // a_hi * W_lo
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 64 + 8 * ks, whi + 2 * ks, 1u);                 // a_hi * W_hi
        if (!MEMS_BIAS_PRELOAD) {
            const uint64_t ones = make_desc(wimg_saddr + BLK_ONES * BLK);
            const uint64_t bias = make_desc(wimg_saddr + (BLK_BIAS + (h >> 2)) * BLK) + 2 * (h & 3);
            umma_ss(d, ones, bias, 1u);                                                                // + bias
        }
    }
}

// MMA issuer warp (all 32 lanes run the loop convergently; one elected lane issues).  It serves `nslot` slots of group g
// starting at slot0 (one slot with MEMS_ISSUER_PER_SLOT, else both) in the fixed order the epilogue warps produce them: for every
// surrogate pass the group announces (ctrl_tiles via bar_S) tiles slot, slot + GSLOT, ...; ctrl_tiles < 0 ends the loop.
// Per tile: layer 0 waits for the tail warps' layer-1 operand, layers 1..5 for the hidden warps' A operands; every
// accumulator goes to the hidden warps (bar_Dh).  g, slot0 must be warp-uniform.
// Coarse pass (FFT surrogate of delayed acceptance), MMA side: one tile in slot 0, seven GEMM layers whose B operands are
// streamed from the global image `cimg` into the staging blocks just before they are needed:
//   layer 0          K = 32 over the split inputs (bias rides on the ones columns)            stage A <- layer-1 block
//   layers 1..5      a_lo W_hi + a_hi W_lo + a_hi W_hi ON TOP of the bias the hidden warps     stage A <- W_hi, stage B <- W_lo
//                    wrote into the accumulator (tcgen05.st) -- no bias MMA
//   output layer     N = 16 (15 features), no bias (added in FP32 by the tail warps)           stage A <- hi | lo blocks
// A staging block is overwritten only after the MMAs that read it have completed (the issuer waits for its own previous
// commit on bar_Dh); the copy overlaps the hidden warps' epilogue.  ph bit 5: bar_W parity.
__device__ __forceinline__ void issue_coarse_pass(GroupSmem& gs, uint32_t d, const unsigned char* cimg, uint32_t stageA,
                                                  uint32_t& ph, uint32_t& nD, uint32_t aAt, uint32_t aAh, uint32_t aDh) {
    const uint32_t stageB = smem_u32(&gs.cb.scale_cw[0][0]);
    const uint32_t aW = smem_u32(&gs.bar_W);
#pragma unroll 1
    for (int layer = 0; layer <= MEMS_HIDDEN; ++layer) {
        if (layer > 0) mbar_wait_a(aDh, (nD - 1u) & 1u);                 // the previous layer's MMAs are done with the staging blocks
        if (elect_one()) {
            if (layer == 0) {
                mbar_arrive_expect_tx(&gs.bar_W, (uint32_t)BLK);
                bulk_g2s_a(stageA, cimg + (size_t)CBLK_B1 * BLK, BLK, aW);
            } else if (layer < MEMS_HIDDEN) {
                mbar_arrive_expect_tx(&gs.bar_W, (uint32_t)(2 * BLK));
                bulk_g2s_a(stageA, cimg + (size_t)(CBLK_WHI + layer - 1) * BLK, BLK, aW);
                bulk_g2s_a(stageB, cimg + (size_t)(CBLK_WLO + layer - 1) * BLK, BLK, aW);
            } else {
                mbar_arrive_expect_tx(&gs.bar_W, 4096u);
                bulk_g2s_a(stageA, cimg + (size_t)CBLK_WO * BLK, 4096u, aW);
            }
        }
        __syncwarp();
        if (layer == 0) { mbar_wait_a(aAt, (ph >> GSLOT) & 1u); ph ^= (1u << GSLOT); }
        else { mbar_wait_a(aAh, ph & 1u); ph ^= 1u; }
        mbar_wait_a(aW, (ph >> (2 * GSLOT + 1)) & 1u);
        ph ^= (1u << (2 * GSLOT + 1));
        tc_fence_after();
        if (elect_one()) {
            const uint64_t bA = make_desc(stageA), bB = make_desc(stageB);
            if (layer == 0) {
                umma_ts(d, d + 64, bA, 0u);
                umma_ts(d, d + 64 + 8, bA + 2, 1u);
            } else if (layer < MEMS_HIDDEN) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 96 + 8 * ks, bA + 2 * ks, 1u);     // a_lo * W_hi (+ preloaded bias)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 64 + 8 * ks, bB + 2 * ks, 1u);     // a_hi * W_lo
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 64 + 8 * ks, bA + 2 * ks, 1u);     // a_hi * W_hi
            } else {
                const uint64_t bL = make_desc(stageA + 2048u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts_n16(d, d + 96 + 8 * ks, bA + 2 * ks, ks > 0 ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts_n16(d, d + 64 + 8 * ks, bL + 2 * ks, 1u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts_n16(d, d + 64 + 8 * ks, bA + 2 * ks, 1u);
            }
            umma_commit_a(aDh);
        }
        ++nD;
        __syncwarp();
    }
}

template <int NSLOT>
__device__ __forceinline__ void mma_issuer_loop(TcSmem& sm, GroupSmem& gs, int g, int slot0, const unsigned char* cimg = nullptr,
                                                uint32_t stageA = 0u) {
    const uint32_t wimg = smem_u32(sm.wimg);
    const uint32_t base = __shfl_sync(0xffffffffu, sm.tmem_base, 0) + (uint32_t)(g * GSLOT * 128);
    uint32_t ph = 0;                  // bits [0, GSLOT): bar_Ah parity per slot; [GSLOT, 2 GSLOT): bar_At; then bar_S, bar_W
    uint32_t nD = 0;                  // commits issued on slot 0 so far (phase of bar_Dh[0] the next commit completes)
    const uint32_t aAt = keep_in_register(smem_u32(&gs.bar_At[0]));
    const uint32_t aAh = aAt + (uint32_t)(offsetof(GroupSmem, bar_Ah) - offsetof(GroupSmem, bar_At));
    const uint32_t aDh = aAt + (uint32_t)(offsetof(GroupSmem, bar_Dh) - offsetof(GroupSmem, bar_At));
    const uint32_t aX = aAt + (uint32_t)(offsetof(GroupSmem, bar_X) - offsetof(GroupSmem, bar_At));
    const uint32_t aF = aAt + (uint32_t)(offsetof(GroupSmem, bar_F) - offsetof(GroupSmem, bar_At));
    uint32_t xph = 0;                 // bit s: parity of the number of tiles issued on slot s = phases of bar_X[s] started so far
    for (;;) {
        mbar_wait(&gs.bar_S[slot0], (ph >> (2 * GSLOT)) & 1u);
        ph ^= (1u << (2 * GSLOT));
        const int n_tiles = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile int*>(&gs.ctrl_tiles), 0);
        if (n_tiles < 0) break;
        if (n_tiles == COARSE_TILES) {                                   // only the issuer of slot 0 is woken for a coarse pass
            issue_coarse_pass(gs, base, cimg, stageA, ph, nD, aAt, aAh, aDh);
            xph ^= 1u;                                                   // the coarse pass completes one phase of bar_X[0]
            continue;
        }
        for (int t0 = 0; t0 < n_tiles; t0 += GSLOT) {
#pragma unroll 1
            for (int layer = 0; layer < MEMS_HIDDEN; ++layer) {
#pragma unroll
                for (int k = 0; k < NSLOT; ++k) {
                    const int s = slot0 + k;
                    if (t0 + s < n_tiles) {
                        TRACE(4, s, layer);
                        if (layer == 0) {
                            mbar_wait_a(aAt + 8 * s, (ph >> (GSLOT + s)) & 1u);
                            ph ^= (1u << (GSLOT + s));
                            // prestaged operands arrive before the previous tile of the slot has been read out of the accumulator
                            if (MEMS_PRESTAGE && t0 > 0) mbar_wait_a(aX + 8 * s, ((xph >> s) & 1u) ^ 1u);
                        } else {
                            mbar_wait_a(aAh + 8 * s, (ph >> s) & 1u);
                            ph ^= (1u << s);
                        }
                        TRACE(5, s, layer);
                        tc_fence_after();
                        if (elect_one()) {
                            issue_layer(wimg, base + s * 128, layer);
                            umma_commit_a(aDh + 8 * s);
                            if (MEMS_PRESTAGE && layer == MEMS_HIDDEN - 1) umma_commit_a(aF + 8 * s);
                        }
                        if (layer == MEMS_HIDDEN - 1) xph ^= (1u << s);   // one more tile of this slot will complete a phase of bar_X
                        TRACE(6, s, layer);
                        if (s == 0) ++nD;
                        __syncwarp();
                    }
                }
            }
        }
    }
}

// Tell the group's MMA issuer(s) how many tiles follow (n_tiles < 0: no more work).  One thread of the group.
__device__ __forceinline__ void announce_tiles(GroupSmem& gs, int n_tiles) {
    *reinterpret_cast<volatile int*>(&gs.ctrl_tiles) = n_tiles;
    if (MEMS_ISSUER_PER_SLOT) {
        const int n_slots = (n_tiles == COARSE_TILES) ? 1 : n_tiles;       // a coarse pass is one tile in slot 0
#pragma unroll
        for (int s = 0; s < GSLOT; ++s)
            if (n_tiles < 0 || s < n_slots) mbar_arrive(&gs.bar_S[s]);     // an issuer without a tile in this pass is not woken
    } else {
        mbar_arrive(&gs.bar_S[0]);
    }
}

// ------------------------------------------------------------------------------------------
// Epilogue side: evaluate S[j] = sum_t (y_obs[t] - f(x_j)[t])^2 for the nc proposals in cb.xs
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

__device__ __forceinline__ void group_bar_sync(int g) { named_bar_sync(1 + g, GTHREADS); }

// The epilogue threads of group g evaluate one batch: gt < HTHREADS are the hidden-layer threads (row = gt % 128, column
// half gt / 128), the rest the tail threads (row = gt - HTHREADS).  Row `row` of every tile is (chain slot row mod nce,
// time sub-row row / nce).
// compact: only the gs.nlive chains listed in gs.map take part (nce = gs.nlive): proposals outside the prior box and,
// in the delayed-acceptance second stage, proposals the first stage did not promote cost nothing -- the tiles are
// rebuilt from the live chains, so the pass shrinks with the promotion rate.  Result: gs.Sfin[chain].
// DENSE (compact passes only): (time, chain) pairs are laid out back to back, row i of tile k is pair f = 128k + i
// = (time f / n_live, chain slot f mod n_live), so every tile is full even when n_live does not divide 128.  A tail
// thread then meets a different chain in every tile: the squared residuals of a tile go through shared memory (gs.R2)
// and tail thread c < n_live adds the entries of chain slot c in time order (deterministic).
// One thread's share of a tile-layer epilogue: NCH pieces of eight accumulator columns from column c0 of the slot at tD.
// Piece by piece (a finer grain than the whole row keeps XU, FMA and ALU work of one warp interleaved,
// profiles/r02_microbench_*): tcgen05.ld -> tanh -> fp16 hi/lo split -> tcgen05.st of the next layer's A operand words.
template <int NCH>
__device__ __forceinline__ void epi_split_chunks(uint32_t tD, int c0) {
#pragma unroll
    for (int pc = 0; pc < NCH; ++pc) {
        uint32_t w[8], hi2[4], lo2[4];
        tmem_ld8(tD + c0 + 8 * pc, w);
        tc_wait_ld();
        epilogue_split_v2<1, false, 0>(w, hi2, lo2);
        tmem_st4(tD + 64 + (c0 >> 1) + 4 * pc, hi2);
        tmem_st4(tD + 96 + (c0 >> 1) + 4 * pc, lo2);
    }
}

// The same for the last tanh layer: its activations go straight into the 64 -> 1 output layer (partial dot over the
// thread's columns; w7p: output weights permuted per group of eight, see epilogue_dot8).
template <int NCH>
__device__ __forceinline__ float epi_dot_chunks(uint32_t tD, int c0, const float* w7p) {
    uint64_t acc = pk(0.0f, 0.0f);
#pragma unroll
    for (int pc = 0; pc < NCH; ++pc) {
        uint32_t w[8];
        tmem_ld8(tD + c0 + 8 * pc, w);
        tc_wait_ld();
        acc = epilogue_dot8(w, w7p + c0 + 8 * pc, acc);
    }
    float oa, ob;
    upk(acc, oa, ob);
    return oa + ob;
}

template <int P, bool DENSE>
__device__ __forceinline__ void tc_eval_group(TcSmem& sm, GroupSmem& gs, int g, int gt, const double* s_yobs,
                                              const float* s_ts, int T, int stride, int nc, int ncv,
                                              bool compact, float* y_out, int c0, uint32_t& ph) {
    const bool is_tail = gt >= HTHREADS;
    const int row = is_tail ? gt - HTHREADS : (gt & 127);
    const int q = row >> 5;
    const int nce = compact ? gs.nlive : nc;          // chains laid out in the tiles (any value in 1..128)
    const int tpt = 128 / nce;                        // time points per tile
    const int Tn = (T + stride - 1) / stride;         // time points of this pass: t = 0, stride, 2*stride, ...
    const int n_tiles = DENSE ? (nce * Tn + 127) / 128 : (Tn + tpt - 1) / tpt;
    const int n_rounds = (n_tiles + GSLOT - 1) / GSLOT;      // every slot takes one tile per round
    const uint32_t tD0 = sm.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * GSLOT * 128);
    if (gt == 0) announce_tiles(gs, n_tiles);

    if (!is_tail) {
        // ---- hidden-layer warps: accumulator -> tanh -> fp16 hi/lo split -> next A operand; last layer -> partial output dot
        const int cb0 = HCOLS * (gt >> 7);            // first accumulator column of this thread
        const uint32_t aDh = keep_in_register(smem_u32(&gs.bar_Dh[0]));
        const uint32_t aAh = aDh - (uint32_t)(offsetof(GroupSmem, bar_Dh) - offsetof(GroupSmem, bar_Ah));
        const uint32_t aX = aDh + (uint32_t)(offsetof(GroupSmem, bar_X) - offsetof(GroupSmem, bar_Dh));
        for (int round = 0; round < n_rounds; ++round) {
#pragma unroll 1
            for (int layer = 0; layer < MEMS_HIDDEN; ++layer) {
#pragma unroll
                for (int s = 0; s < GSLOT; ++s) {
                    if (round * GSLOT + s < n_tiles) {
                        const uint32_t tD = tD0 + s * 128;
                        TRACE(1, s, layer);
                        mbar_wait_a(aDh + 8 * s, (ph >> s) & 1u);
                        ph ^= (1u << s);
                        TRACE(2, s, layer);
                        tc_fence_after();
                        if (layer < MEMS_HIDDEN - 1) {
                            epi_split_chunks<HCOLS / 8>(tD, cb0);
                            if (MEMS_BIAS_PRELOAD) {           // bias of the next GEMM into the accumulator columns just read
#pragma unroll
                                for (int c8 = 0; c8 < HCOLS / 8; ++c8) {
                                    const float4* b4 = reinterpret_cast<const float4*>(sm.tcbias + layer * MEMS_WIDTH + cb0 + 8 * c8);
                                    const float4 u = b4[0], v = b4[1];
                                    const uint32_t bw[8] = {__float_as_uint(u.x), __float_as_uint(u.y), __float_as_uint(u.z), __float_as_uint(u.w),
                                                            __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
                                    tmem_st8(tD + cb0 + 8 * c8, bw);
                                }
                            }
                            tc_wait_st();
                            tc_fence_before();
                            mbar_arrive_a(aAh + 8 * s);
                            TRACE(3, s, layer);
                        } else {                               // last tanh layer . output weights: partial dot to the tail warps
                            gs.xchg[s][gt >> 7][row] = epi_dot_chunks<HCOLS / 8>(tD, cb0, sm.w7);
                            tc_fence_before();                 // the accumulator has been read: the slot may be restaged
                            mbar_arrive_a(aX + 8 * s);         // release: the store above is visible to the tail warps
                            TRACE(3, s, layer);
                        }
                    }
                }
            }
        }
    } else {
        // ---- tail warps: layer-1 operands, residuals -- and (MEMS_TAIL_SHARE) the last TCOLS columns of every tile-layer
        const int jc = row % nce;                     // rows >= nce*tpt of a tile idle
        const int sub = row / nce;
        const int j = compact ? (int)gs.map[jc] : jc; // batch-local chain index
        const bool chain_live = (sub < tpt) && (compact || jc < ncv);
        const uint32_t aAt = keep_in_register(smem_u32(&gs.bar_At[0]));
        const uint32_t aX = aAt + (uint32_t)(offsetof(GroupSmem, bar_X) - offsetof(GroupSmem, bar_At));
        const uint32_t aDh = aAt + (uint32_t)(offsetof(GroupSmem, bar_Dh) - offsetof(GroupSmem, bar_At));
        const uint32_t aAh = aAt + (uint32_t)(offsetof(GroupSmem, bar_Ah) - offsetof(GroupSmem, bar_At));
        const uint32_t aF = aAt + (uint32_t)(offsetof(GroupSmem, bar_F) - offsetof(GroupSmem, bar_At));
        uint32_t a1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a1[i] = 0u;
        if (!DENSE) {
#pragma unroll
            for (int p = 0; p < P; ++p) put_input<P>(a1, p, (compact || jc < ncv) ? gs.cb.xs[p][j] : 0.0f);
        }
        set_half<P>(a1, 5 * (P + 1) + 0, 1.0f);       // ones that pick up b_hi, b_mid, b_lo
        set_half<P>(a1, 5 * (P + 1) + 1, 1.0f);
        set_half<P>(a1, 5 * (P + 1) + 2, 1.0f);
        // layer-1 operand of `tile` into slot s
        auto stage_tile = [&](int s, int tile) {
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
            tmem_st16(tD0 + s * 128 + 64, a1);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive_a(aAt + 8 * s);
        };
        double Sp = 0.0;
        // a tile's output is complete (own columns in `own`, the hidden warps' partial dots in xchg): residual, then the
        // layer-1 operand of the slot's next tile
        auto finish_tile = [&](int s, int tile, float own) {
            TRACE(7, s, 0);
            mbar_wait_a(aX + 8 * s, (ph >> s) & 1u);
            ph ^= (1u << s);
            TRACE(8, s, 0);
            tc_fence_after();                              // the hidden warps' accumulator reads precede the restaging below
            float o = own + sm.b7;
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) o += gs.xchg[s][hh][row];
            if (DENSE) {
                const int f = tile * 128 + row;
                const int ti = f / nce;
                double r2 = 0.0;
                if (ti < Tn) {
                    const double r = s_yobs[ti * stride] - (double)o;
                    r2 = r * r;
                }
                gs.R2[s][row] = r2;
                named_bar_sync(3 + g, TTHREADS);               // the tile's 128 residuals are in shared memory
                if (row < nce) {                               // chain slot `row`: its rows of this tile, in time order
                    int r0 = row - (tile * 128) % nce;
                    if (r0 < 0) r0 += nce;
                    for (int rr = r0; rr < 128; rr += nce) Sp += gs.R2[s][rr];
                }
                // R2[s] is written again GSLOT tiles later, behind the barriers of all tail threads on the other slots' R2
            } else {
                const int ti = tile * tpt + sub;
                const int t = ti * stride;
                if (chain_live && ti < Tn) {
                    if (y_out) y_out[(size_t)(c0 + j) * T + t] = o;
                    const double r = s_yobs[t] - (double)o;
                    Sp = fma(r, r, Sp);
                }
            }
            // this slot's D and A are free again: stage the layer-1 operand of its next tile (unless it was staged early)
            const int nt = tile + GSLOT;
            if (!MEMS_PRESTAGE && nt < n_tiles) stage_tile(s, nt);
            TRACE(9, s, 0);
        };
#pragma unroll
        for (int s = 0; s < GSLOT; ++s)
            if (s < n_tiles) stage_tile(s, s);

        for (int round = 0; round < n_rounds; ++round) {
            if (MEMS_TAIL_SHARE) {
#pragma unroll 1
                for (int layer = 0; layer < MEMS_HIDDEN; ++layer) {
#pragma unroll
                    for (int s = 0; s < GSLOT; ++s) {
                        const int tile = round * GSLOT + s;
                        if (tile < n_tiles) {
                            const uint32_t tD = tD0 + s * 128;
                            mbar_wait_a(aDh + 8 * s, (ph >> (GSLOT + s)) & 1u);
                            ph ^= (1u << (GSLOT + s));
                            tc_fence_after();
                            if (layer < MEMS_HIDDEN - 1) {
                                epi_split_chunks<TCOLS / 8>(tD, 2 * HCOLS);
                                tc_wait_st();
                                tc_fence_before();
                                mbar_arrive_a(aAh + 8 * s);
                            } else {
                                const float own = epi_dot_chunks<TCOLS / 8>(tD, 2 * HCOLS, sm.w7);
                                tc_fence_before();
                                finish_tile(s, tile, own);
                            }
                        }
                    }
                }
            } else {
                if (MEMS_PRESTAGE) {
                    // the operand columns of a slot are free once its last layer's MMAs have completed (bar_F) -- long before the
                    // hidden warps have turned that accumulator into the tile's output: the next tile's layer-1 operand goes in now
#pragma unroll 1
                    for (int s = 0; s < GSLOT; ++s) {
                        const int tile = round * GSLOT + s;
                        if (tile < n_tiles) {
                            mbar_wait_a(aF + 8 * s, (ph >> (GSLOT + s)) & 1u);
                            ph ^= (1u << (GSLOT + s));
                            tc_fence_after();
                            if (tile + GSLOT < n_tiles) stage_tile(s, tile + GSLOT);
                        }
                    }
                }
#pragma unroll 1
                for (int s = 0; s < GSLOT; ++s) {
                    const int tile = round * GSLOT + s;
                    if (tile < n_tiles) finish_tile(s, tile, 0.0f);
                }
            }
        }
        if (DENSE) {
            if (row < nce) gs.Sfin[gs.map[row]] = Sp; // tail thread `row` summed chain slot `row` over all tiles
        } else {
            gs.Spart[row] = Sp;
        }
    }
    if (!DENSE) {
        group_bar_sync(g);
        if (gt < nce) {
            double S = 0.0;                            // fixed order: time sub-rows ascending
            for (int u = 0; u < tpt; ++u) S += gs.Spart[u * nce + gt];
            gs.Sfin[compact ? (int)gs.map[gt] : gt] = S;
        }
    }
    group_bar_sync(g);                                 // Sfin[chain] may have been written by another thread
}

// K-slot layout of a layer-1 operand with NIN inputs (coarse model: the P parameters, no time column)
template <int NIN>
__device__ __forceinline__ void put_input_n(uint32_t (&a1)[16], int i, float v) {
    float h, m, l;
    split3(v, h, m, l);
    set_half<0>(a1, 0 * NIN + i, h);
    set_half<0>(a1, 1 * NIN + i, m);
    set_half<0>(a1, 2 * NIN + i, l);
    set_half<0>(a1, 3 * NIN + i, h);
    set_half<0>(a1, 4 * NIN + i, m);
}

// Coarse pass (FFT surrogate, delayed-acceptance first stage) on the tensor cores, epilogue side: the gs.nlive proposals
// listed in gs.map are the rows of ONE tile in slot 0 (row i = live chain i); result gs.Sfin[chain] = ||y_obs - y_c||^2.
//   tail warps     scaled inputs (f64 scale, f32 cast: utils.py:38-39 convention) -> three-way split layer-1 operand;
//                  after the output layer: 15 features + bias -> band-limited misfit (same arithmetic as coarse_fft_pass)
//   hidden warps   six tanh epilogues (identical to the fine model's) and, for layers 1..5, the NEXT layer's bias
//                  written into the accumulator columns they have just read (the MMAs of that layer accumulate onto it)
template <int P>
__device__ __forceinline__ void tc_coarse_group(TcSmem& sm, GroupSmem& gs, int g, int gt, const MemsCoarse* __restrict__ cm,
                                                int T, uint32_t& ph) {
    const bool is_tail = gt >= HTHREADS;
    const int row = is_tail ? gt - HTHREADS : (gt & 127);
    const int q = row >> 5;
    const uint32_t tD = sm.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * GSLOT * 128);     // slot 0 of the group
    if (gt == 0) announce_tiles(gs, COARSE_TILES);
    // this thread's accumulator columns: hidden threads HCOLS from HCOLS * half, tail threads (MEMS_TAIL_SHARE) the last TCOLS
    const int cb0 = is_tail ? 2 * HCOLS : HCOLS * (gt >> 7);
    const int nch = is_tail ? TCOLS / 8 : HCOLS / 8;
    const uint32_t aDh = smem_u32(&gs.bar_Dh[0]), aAh = smem_u32(&gs.bar_Ah[0]), aX = smem_u32(&gs.bar_X[0]);
    // six tanh epilogues; hidden threads wait on ph bit 0, tail threads (when they share the columns) on bit 2
    auto tanh_layers = [&](uint32_t bit) {
#pragma unroll 1
        for (int layer = 0; layer < MEMS_HIDDEN; ++layer) {
            mbar_wait_a(aDh, (ph & bit) ? 1u : 0u);
            ph ^= bit;
            tc_fence_after();
            if (nch == 3) epi_split_chunks<3>(tD, cb0);
            else if (nch == 4) epi_split_chunks<4>(tD, cb0);
            else epi_split_chunks<2>(tD, cb0);
            if (layer < MEMS_NLAYER_H) {                   // bias of hidden layer `layer` (the next GEMM) into D
                for (int c8 = 0; c8 < nch; ++c8) {
                    const float4* b4 = reinterpret_cast<const float4*>(cm->tcbias + layer * MEMS_WIDTH + cb0 + 8 * c8);
                    const float4 u = __ldg(b4), v = __ldg(b4 + 1);
                    const uint32_t bw[8] = {__float_as_uint(u.x), __float_as_uint(u.y), __float_as_uint(u.z), __float_as_uint(u.w),
                                            __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
                    tmem_st8(tD + cb0 + 8 * c8, bw);
                }
            }
            tc_wait_st();
            tc_fence_before();
            mbar_arrive_a(aAh);
        }
    };
    if (!is_tail) {
        tanh_layers(1u);
        mbar_wait_a(aDh, ph & 1u);                         // output layer complete: relay to the tail warps
        ph ^= 1u;
        tc_fence_after();
        tc_fence_before();
        mbar_arrive_a(aX);
    } else {
        const bool live = row < gs.nlive;
        const int j = live ? (int)gs.map[row] : 0;
        uint32_t a1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a1[i] = 0u;
        if (live) {
#pragma unroll
            for (int p = 0; p < P; ++p)
                put_input_n<P>(a1, p, (float)((gs.cb.xp[p][j] - cm->mean[p]) / cm->sscale[p]));
        }
        set_half<0>(a1, 5 * P + 0, 1.0f);                   // ones that pick up b_hi, b_mid, b_lo
        set_half<0>(a1, 5 * P + 1, 1.0f);
        set_half<0>(a1, 5 * P + 2, 1.0f);
        tmem_st16(tD + 64, a1);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&gs.bar_At[0]);
        if (MEMS_TAIL_SHARE) {
            tanh_layers(1u << GSLOT);
            ph ^= (1u << GSLOT);                            // the output layer's phase of bar_Dh[0], observed through bar_X below
        }
        mbar_wait(&gs.bar_X[0], ph & 1u);
        ph ^= 1u;
        tc_fence_after();
        uint32_t fw[16];
        tmem_ld16(tD, fw);
        tc_wait_ld();
        tc_fence_before();                                  // the accumulator has been read: slot 0 may be restaged
        if (live) {
            float f[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) f[o] = __uint_as_float(fw[o]) + __ldg(&cm->bo[o]);
            const double m = 0.1 * (double)f[0];
            double dot = m * cm->sum_y, nrm = (double)T * m * m;
            const double w2 = 2.0 / (double)T;
#pragma unroll 1
            for (int k = 0; k < MEMS_COARSE_HARM; ++k) {
                const double A = (double)f[1 + k], phi = 0.01 * (double)f[1 + MEMS_COARSE_HARM + k];
                double sn, cs;
                sincos(phi, &sn, &cs);
                dot += w2 * A * (cs * cm->cy[k] - sn * cm->sy[k]);
                nrm += w2 * A * A;
            }
            gs.Sfin[j] = cm->sum_y2 - 2.0 * dot + nrm;
        }
    }
    group_bar_sync(g);                                      // Sfin[] is complete
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

// thread -> (group, index within the group's epilogue threads); issuer warps get gt = -1
//   tid <  NGROUP*HTHREADS                : hidden threads, group = tid / HTHREADS, gt = tid % HTHREADS
//   tid <  NWORKER                        : tail threads,   group = (tid - 512) / TTHREADS, gt = HTHREADS + row
//   else                                  : MMA issuer warps; gt = -1 - (first slot served); one warp per (group, slot) or per group
// (warp % 4 is the TMEM lane quarter of row = 32 * (warp % 4) + lane in both worker kinds)
__device__ __forceinline__ void my_role(int& g, int& gt) {
    const int tid = threadIdx.x;
    if (tid < NGROUP * HTHREADS) { g = tid / HTHREADS; gt = tid % HTHREADS; }
    else if (tid < NWORKER) { g = (tid - NGROUP * HTHREADS) / TTHREADS; gt = HTHREADS + (tid - NGROUP * HTHREADS) % TTHREADS; }
    else {
        const int iw = (tid - NWORKER) >> 5;
        if (MEMS_ISSUER_PER_SLOT) { g = iw / GSLOT; gt = -1 - iw % GSLOT; }
        else { g = iw; gt = -1; }
    }
}

// Common prologue: barriers, TMEM, weight image (bulk TMA), y_obs / time columns.
__device__ __forceinline__ void tc_prologue(TcSmem& sm, const MemsProblem& pb, double* s_yobs, float* s_ts) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    if (tid == 0) {
        for (int g = 0; g < NGROUP; ++g) {
            for (int s = 0; s < GSLOT; ++s) {
                mbar_init(&sm.grp[g].bar_At[s], TTHREADS);
                mbar_init(&sm.grp[g].bar_Ah[s], HTHREADS + (MEMS_TAIL_SHARE ? TTHREADS : 0));
                mbar_init(&sm.grp[g].bar_Dh[s], 1);
                mbar_init(&sm.grp[g].bar_X[s], HTHREADS);
                mbar_init(&sm.grp[g].bar_F[s], 1);
            }
            for (int s = 0; s < GSLOT; ++s) mbar_init(&sm.grp[g].bar_S[s], 1);
            mbar_init(&sm.grp[g].bar_W, 1);
            sm.grp[g].ctrl_tiles = 0;
        }
        mbar_init(&sm.bar_w, 1);
        sm.c = pb.c;
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == NWORKER / 32) {
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
    if (MEMS_BIAS_PRELOAD)
        for (int i = tid; i < MEMS_NLAYER_H * MEMS_WIDTH; i += NTHREADS) sm.tcbias[i] = (float)(C2 * (double)pb.bh[i]);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(&sm.bar_w, 0);
}

__device__ __forceinline__ void tc_epilogue_free(TcSmem& sm) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == NWORKER / 32) {
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

    TRACE(15, 0, 0);
    const int T = sm.c.T, nc = ra.nc;
    const int nbatch = (ch.C + nc - 1) / nc;
    int g, gt;
    my_role(g, gt);
    GroupSmem& gs = sm.grp[g];
    const int gid = blockIdx.x * NGROUP + g, gstride = gridDim.x * NGROUP;

    if (gt < 0) {
        const int gw = __shfl_sync(0xffffffffu, g, 0);                   // provably warp-uniform group index
        const int sw = __shfl_sync(0xffffffffu, -1 - gt, 0);             // and first slot
        // coarse pass on the tensor cores: stage A of the group = extra dynamic shared memory behind the time column
        const uint32_t stage0 = (smem_u32(s_ts + T) + 1023u) & ~1023u;
        const unsigned char* cimg = ra.coarse_tc ? ra.coarse->wimg : nullptr;
        mma_issuer_loop<MEMS_ISSUER_PER_SLOT ? 1 : GSLOT>(sm, sm.grp[gw], gw, sw, cimg, stage0 + (uint32_t)(gw * BLK));
    } else {
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
                    if (any && batch_pass_is_coarse_model(ra, sub) && ra.coarse_tc)   // DA first stage on the FFT surrogate
                        tc_coarse_group<P>(sm, gs, g, gt, ra.coarse, T, ph);
                    else if (any && batch_pass_is_coarse_model(ra, sub))   // ... on CUDA cores when the staging blocks do not fit
                        coarse_fft_pass(ra.coarse, P, T, &gs.cb.xp[0][0], MEMS_NCMAX, gs.map, gs.nlive, gs.Sfin, coarse_scratch(gs), gt,
                                        GTHREADS, [g] { group_bar_sync(g); });
                    else if (any) {
                        // dense packing when it saves tiles (n_live does not divide 128 well)
                        const int stride = batch_pass_stride(ra, sub);
                        const int Tn = (T + stride - 1) / stride, tpt = 128 / gs.nlive;
                        if (gt == 0) atomicAdd(ch.rows_eval, (unsigned long long)gs.nlive * (unsigned long long)Tn);
                        // (a dense tile costs the tail warps more -- every row changes its chain from tile to tile -- so the packing
                        //  must save more than MEMS_DENSE_MIN_SAVING percent of the tiles to be chosen)
                        const int t_dense = (gs.nlive * Tn + 127) / 128, t_plain = (Tn + tpt - 1) / tpt;
                        if (100 * t_dense < (100 - MEMS_DENSE_MIN_SAVING) * t_plain)
                            tc_eval_group<P, true>(sm, gs, g, gt, s_yobs, s_ts, T, stride, nc, ncv, true, nullptr, c0, ph);
                        else
                            tc_eval_group<P, false>(sm, gs, g, gt, s_yobs, s_ts, T, stride, nc, ncv, true, nullptr, c0, ph);
                    }
                    if (gt < ncv) batch_post(gs.cb, sm.c, ch, ra, c0, gt, s, sub, (any && gs.cb.live[gt] == 1) ? gs.Sfin[gt] : 0.0);
                }
            }
            batch_store(gs.cb, ch, P, ra.mode, c0, gt, ncv);
        }
        group_bar_sync(g);
        if (gt == 0) announce_tiles(gs, -1);
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

    const int T = sm.c.T;
    const int nbatch = (n + nc - 1) / nc;
    int g, gt;
    my_role(g, gt);
    GroupSmem& gs = sm.grp[g];
    const int gid = blockIdx.x * NGROUP + g, gstride = gridDim.x * NGROUP;

    if (gt < 0) {
        const int gw = __shfl_sync(0xffffffffu, g, 0);                   // provably warp-uniform group index
        const int sw = __shfl_sync(0xffffffffu, -1 - gt, 0);             // and first slot
        mma_issuer_loop<MEMS_ISSUER_PER_SLOT ? 1 : GSLOT>(sm, sm.grp[gw], gw, sw);
    } else {
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
            tc_eval_group<P, false>(sm, gs, g, gt, s_yobs, s_ts, T, stride, nc, ncv, false, y_out, c0, ph);
            if (gt < ncv && ll_out) {
                // a strided (coarse) pass is rescaled to the full trace length, as in the DA first stage
                const int Tn = (T + stride - 1) / stride;
                const double ll = sm.c.neg_half_inv_sigma2 * gs.Sfin[gt] * ((double)T / (double)Tn);
                ll_out[c0 + gt] = (apply_prior && !gs.cb.inb[gt]) ? -CUDART_INF : ll;
            }
        }
        group_bar_sync(g);
        if (gt == 0) announce_tiles(gs, -1);
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

// ---- coarse (FFT) surrogate on the tensor cores: global-memory operand image, streamed by the MMA issuer -------------
size_t mems_tc_coarse_wimg_bytes() { return COARSE_WIMG_BYTES; }

// Does the extra staging block of every group fit behind the observation / time columns?
bool mems_tc_coarse_fits(int T, size_t smem_optin) { return tc_smem_bytes(T) + COARSE_STAGE_BYTES <= smem_optin; }

// cm: the FP32 coarse model (Keras layout).  host_img: COARSE_WIMG_BYTES.  Also fills cm.tcbias (2 log2(e) * bh).
void mems_tc_build_coarse_wimg(MemsCoarse& cm, int P, void* host_img) {
    unsigned char* img = static_cast<unsigned char*>(host_img);
    memset(img, 0, COARSE_WIMG_BYTES);
    for (int n = 0; n < MEMS_WIDTH; ++n) {              // layer 1: K slots term * P + i, terms 0..2 carry W_hi, 3..4 W_lo
        for (int i = 0; i < P; ++i) {
            const double w = C2 * (double)cm.W1[i * MEMS_WIDTH + n];
            const double whi = h_val(w), wlo = h_val(w - whi);
            put(img, CBLK_B1, n, 0 * P + i, whi);
            put(img, CBLK_B1, n, 1 * P + i, whi);
            put(img, CBLK_B1, n, 2 * P + i, whi);
            put(img, CBLK_B1, n, 3 * P + i, wlo);
            put(img, CBLK_B1, n, 4 * P + i, wlo);
        }
        const double b = C2 * (double)cm.b1[n];
        const double bh = h_val(b), bm = h_val(b - bh), bl = h_val(b - bh - bm);
        put(img, CBLK_B1, n, 5 * P + 0, bh);
        put(img, CBLK_B1, n, 5 * P + 1, bm);
        put(img, CBLK_B1, n, 5 * P + 2, bl);
    }
    for (int h = 0; h < MEMS_NLAYER_H; ++h)
        for (int n = 0; n < MEMS_WIDTH; ++n) {
            for (int k = 0; k < MEMS_WIDTH; ++k) {
                const double w = C2 * (double)cm.Wh[(h * MEMS_WIDTH + k) * MEMS_WIDTH + n];
                const double whi = h_val(w);
                put(img, CBLK_WHI + h, n, k, whi);
                put(img, CBLK_WLO + h, n, k, w - whi);
            }
            cm.tcbias[h * MEMS_WIDTH + n] = (float)(C2 * (double)cm.bh[h * MEMS_WIDTH + n]);
        }
    for (int o = 0; o < 16; ++o)                          // output layer, linear: B[o][k] = Wo[k][o]; lo block 2 KB further
        for (int k = 0; k < MEMS_WIDTH; ++k) {
            const double w = (o < MEMS_COARSE_OUT) ? (double)cm.Wo[k * 16 + o] : 0.0;
            const double whi = h_val(w);
            put(img, CBLK_WO, o, k, whi);
            put(img + 2048, CBLK_WO, o, k, w - whi);
        }
}

// largest trace length whose observation / time columns fit next to the resident weights in `smem_optin` bytes
int mems_tc_max_time_points(size_t smem_optin) {
    const size_t fixed = tc_smem_bytes(0);
    if (smem_optin <= fixed) return 0;
    return (int)((smem_optin - fixed) / (sizeof(double) + sizeof(float)));
}

// Chains per group batch (any value 1..128): minimise (waves of batches over the NGROUP*SMs groups) x (tile rounds per step).
int mems_tc_choose_nc(int C, int T, int sm_count) {
    int best = 1;
    long long best_cost = -1;
    for (int nc = 1; nc <= MEMS_NCMAX; ++nc) {
        const long long nbatch = (C + nc - 1) / nc;
        const long long waves = (nbatch + (long long)NGROUP * sm_count - 1) / ((long long)NGROUP * sm_count);
        const int tpt = 128 / nc;
        const long long tiles = (T + tpt - 1) / tpt;
        const long long cost = waves * ((tiles + GSLOT - 1) / GSLOT) * GSLOT;
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
    const size_t smem = tc_smem_bytes(ctx.pb_host->c.T) + (ra.coarse_tc ? COARSE_STAGE_BYTES : 0);
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
