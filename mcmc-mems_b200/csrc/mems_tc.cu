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
//   pipeline  4 tile slots in flight per SM (4 x 128 TMEM columns); warps 0..15 are epilogue workers
//             (slot = warp/4, TMEM lane quarter = warp%4), warp 16 issues every MMA.
//   tail      64->1 output layer, residual against y_obs, squared-sum reduction in FP64, accept/reject
//             and chain-state update (mems_common.cuh) all happen in the same kernel.
//
// Reference semantics: see mems_common.cuh header (src/InverseProblems/utils.py:30-40 etc.).
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "mems_common.cuh"

#ifndef MEMS_TRYWAIT_HINT
#define MEMS_TRYWAIT_HINT 0
#endif
#ifndef MEMS_EXP_EPI
#define MEMS_EXP_EPI 0         // != 0: timing experiments with deliberately wrong numerics (never shipped)
#endif
#ifndef MEMS_EX2_MIX
#define MEMS_EX2_MIX 0         // share of exponentials evaluated on the FMA pipe: 0 none, 1 -> 1/4, 2 -> 3/8, 3 -> 1/2
#endif
#ifndef MEMS_TANH_GROUP
#define MEMS_TANH_GROUP 8      // activations sharing one MUFU.RCP: 8 (1.125 MUFU/activation) or 4 (1.25, shorter chain)
#endif

namespace {

// CTA organisation: two independent worker groups; each owns one batch of chains at a time, two TMEM
// tile slots (ping-pong: the epilogue of one slot overlaps the MMAs of the other) and its own MMA
// issuer warp.  Groups never synchronise with each other, so one group's accept/reject section
// overlaps the other group's tensor work.
constexpr int NGROUP = 2;
constexpr int GSLOT = 2;                       // tile slots per group
// NH = worker warps per TMEM lane quarter: 1 -> a thread owns a whole 64-column row of the tile,
// 2 -> two threads split the row (32 columns each), doubling the warps the scheduler can interleave.
__host__ __device__ constexpr int gthreads(int NH) { return 128 * NH; }                 // worker threads per group
__host__ __device__ constexpr int nworker(int NH) { return NGROUP * gthreads(NH); }
__host__ __device__ constexpr int nthreads(int NH) { return nworker(NH) + 32 * NGROUP; } // + one MMA issuer warp per group
constexpr int BLK = 8192;             // one [64 rows][64 halves] 128B-swizzled operand block
constexpr int BLK_B1 = 0;             // layer-1 B operand
constexpr int BLK_WHI = 1;            // 5 blocks
constexpr int BLK_WLO = 6;            // 5 blocks
constexpr int BLK_BIAS = 11;          // 2 blocks: hidden layer h<4 at k-step h of block 0, h=4 at k-step 0 of block 1
constexpr int BLK_ONES = 13;          // 2 blocks = 128 rows of [1,1,1,0,...]
constexpr int NBLK = 15;
constexpr int WIMG_TAIL = 64 * 4 + 16;                 // w7[64] f32, b7 f32, pad
constexpr int WIMG_BYTES = NBLK * BLK + WIMG_TAIL;
constexpr double C2 = 2.0 * 1.4426950408889634074;     // 2*log2(e)

// instruction descriptor, kind::f16: D=f32, A=B=f16, K-major both, N=64, M=128
constexpr uint32_t IDESC = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

struct GroupSmem {
    ChainBatch cb;
    double Spart[128];
    float xchg[GSLOT][128];            // NH=2: partial output-layer dot of the upper column half
    unsigned long long bar_A[GSLOT];   // workers -> MMA issuer: A operand of the slot is in TMEM
    unsigned long long bar_D[GSLOT];   // tensor core -> workers: accumulator of the slot is complete
    unsigned long long bar_X[GSLOT][4]; // NH=2: upper-half warp of a lane quarter -> lower-half warp (partial dot ready)
    unsigned long long bar_S;          // workers -> MMA issuer: ctrl_tiles of the next pass is valid
    int ctrl_tiles;                    // tiles of the coming pass, -1 = no more work
    int pad_;
};

struct TcSmem {
    __align__(1024) unsigned char wimg[NBLK * BLK];
    float w7[64];
    float b7;
    float pad_[3];
    MemsConsts c;
    GroupSmem grp[NGROUP];
    unsigned long long bar_w;
    uint32_t tmem_base;
    uint32_t pad2_;
    // followed by: double yobs[T]; float ts[T];
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(void* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
#if MEMS_TRYWAIT_HINT
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");   // hardware-suspended wait, wakes on completion
#endif
    return ok != 0;
}
// Blocking wait with a watchdog: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// one lane of a converged warp (CUTLASS idiom): the warp stays convergent around the elected instruction, so
// warp-uniform operands remain in uniform registers (no per-lane "waterfall" loop around UTCHMMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(IDESC), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(void* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);   // .x (low 16 bits) = lo
    return *reinterpret_cast<const uint32_t*>(&h);
}

// packed FP32x2 arithmetic (FADD2 / FMUL2 / FFMA2): two lanes of work per issue slot
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// tanh of eight accumulator values that already hold y = 2*log2(e)*z:
//     q = 2^-|y| in (0, 1],   |tanh z| = (1 - q)/(1 + q) = 2/(1 + q) - 1,   sign from y.
// Because 1 + q lies in (1, 2] the eight denominators can be inverted with ONE MUFU.RCP (product tree,
// no overflow, no clamp): 1.125 MUFU ops per activation.  Pairs are (element i, element i+4) so that every
// level of the tree is a packed FMUL2 with no register shuffles.  Results: T[i] = (t_i, t_{i+4}).
__device__ __forceinline__ void tanh8(const uint32_t* v, uint64_t (&T)[4]) {
    float q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = ex2_approx(-fabsf(__uint_as_float(v[i])));
    const uint64_t one2 = pk(1.0f, 1.0f);
    const uint64_t A0 = add2(pk(q[0], q[4]), one2), A1 = add2(pk(q[1], q[5]), one2);
    const uint64_t A2 = add2(pk(q[2], q[6]), one2), A3 = add2(pk(q[3], q[7]), one2);
    const uint64_t B0 = mul2(A0, A1), B1 = mul2(A2, A3);       // (d0 d1, d4 d5), (d2 d3, d6 d7)
    float Q0, Q1;
    upk(mul2(B0, B1), Q0, Q1);                                  // products of each quad
    const float r = rcp_approx(Q0 * Q1);
    const uint64_t Rq = pk(r * Q1, r * Q0);                     // (1/Q0, 1/Q1)
    const uint64_t Rb0 = mul2(Rq, B1), Rb1 = mul2(Rq, B0);      // (1/(d0 d1), 1/(d4 d5)), (1/(d2 d3), 1/(d6 d7))
    const uint64_t two2 = pk(2.0f, 2.0f), m1 = pk(-1.0f, -1.0f);
    const uint64_t U0 = fma2(mul2(Rb0, A1), two2, m1);          // (|t0|, |t4|)
    const uint64_t U1 = fma2(mul2(Rb0, A0), two2, m1);          // (|t1|, |t5|)
    const uint64_t U2 = fma2(mul2(Rb1, A3), two2, m1);          // (|t2|, |t6|)
    const uint64_t U3 = fma2(mul2(Rb1, A2), two2, m1);          // (|t3|, |t7|)
    float a, b;
    upk(U0, a, b); T[0] = pk(copysignf(a, __uint_as_float(v[0])), copysignf(b, __uint_as_float(v[4])));
    upk(U1, a, b); T[1] = pk(copysignf(a, __uint_as_float(v[1])), copysignf(b, __uint_as_float(v[5])));
    upk(U2, a, b); T[2] = pk(copysignf(a, __uint_as_float(v[2])), copysignf(b, __uint_as_float(v[6])));
    upk(U3, a, b); T[3] = pk(copysignf(a, __uint_as_float(v[3])), copysignf(b, __uint_as_float(v[7])));
}

// 2^-|y| for a pair on the FMA pipe (no MUFU): round-to-nearest range reduction with the 1.5*2^23 magic constant,
// degree-5 minimax polynomial of 2^f on [-0.5, 0.5] (max rel. error 7.5e-8, the same grade as ex2.approx), exponent
// inserted with one integer shift-add per element.  Offloading part of the exponentials relieves the XU pipe, which
// also executes the F2FP packs (FlashAttention-4 uses the same trick for softmax).
__device__ __forceinline__ uint64_t ex2_neg_abs_poly2(uint32_t ya, uint32_t yb) {
    const float wa = fmaxf(-fabsf(__uint_as_float(ya)), -125.0f), wb = fmaxf(-fabsf(__uint_as_float(yb)), -125.0f);
    const uint64_t w2 = pk(wa, wb);
    const uint64_t magic2 = pk(12582912.0f, 12582912.0f);
    const uint64_t r2 = add2(w2, magic2);
    const uint64_t f2 = sub2(w2, sub2(r2, magic2));
    uint64_t p2 = pk(0.001327647129073739f, 0.001327647129073739f);
    p2 = fma2(p2, f2, pk(0.009675540961325169f, 0.009675540961325169f));
    p2 = fma2(p2, f2, pk(0.05550713092088699f, 0.05550713092088699f));
    p2 = fma2(p2, f2, pk(0.24022120237350464f, 0.24022120237350464f));
    p2 = fma2(p2, f2, pk(0.6931469440460205f, 0.6931469440460205f));
    p2 = fma2(p2, f2, pk(1.0000001192092896f, 1.0000001192092896f));
    float pa, pb, ra, rb;
    upk(p2, pa, pb);
    upk(r2, ra, rb);
    const float qa = __uint_as_float(__float_as_uint(pa) + (__float_as_uint(ra) << 23));
    const float qb = __uint_as_float(__float_as_uint(pb) + (__float_as_uint(rb) << 23));
    return pk(qa, qb);
}

// tanh8 with NPOLY (0..2) of the four element pairs taking their exponential from the FMA pipe.
template <int NPOLY>
__device__ __forceinline__ void tanh8_mix(const uint32_t* v, uint64_t (&T)[4]) {
    uint64_t Q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i < NPOLY) Q[i] = ex2_neg_abs_poly2(v[i], v[i + 4]);
        else Q[i] = pk(ex2_approx(-fabsf(__uint_as_float(v[i]))), ex2_approx(-fabsf(__uint_as_float(v[i + 4]))));
    }
    const uint64_t one2 = pk(1.0f, 1.0f);
    const uint64_t A0 = add2(Q[0], one2), A1 = add2(Q[1], one2), A2 = add2(Q[2], one2), A3 = add2(Q[3], one2);
    const uint64_t B0 = mul2(A0, A1), B1 = mul2(A2, A3);
    float Q0, Q1;
    upk(mul2(B0, B1), Q0, Q1);
    const float r = rcp_approx(Q0 * Q1);
    const uint64_t Rq = pk(r * Q1, r * Q0);
    const uint64_t Rb0 = mul2(Rq, B1), Rb1 = mul2(Rq, B0);
    const uint64_t two2 = pk(2.0f, 2.0f), m1 = pk(-1.0f, -1.0f);
    const uint64_t U0 = fma2(mul2(Rb0, A1), two2, m1), U1 = fma2(mul2(Rb0, A0), two2, m1);
    const uint64_t U2 = fma2(mul2(Rb1, A3), two2, m1), U3 = fma2(mul2(Rb1, A2), two2, m1);
    float a, b;
    upk(U0, a, b); T[0] = pk(copysignf(a, __uint_as_float(v[0])), copysignf(b, __uint_as_float(v[4])));
    upk(U1, a, b); T[1] = pk(copysignf(a, __uint_as_float(v[1])), copysignf(b, __uint_as_float(v[5])));
    upk(U2, a, b); T[2] = pk(copysignf(a, __uint_as_float(v[2])), copysignf(b, __uint_as_float(v[6])));
    upk(U3, a, b); T[3] = pk(copysignf(a, __uint_as_float(v[3])), copysignf(b, __uint_as_float(v[7])));
}

// hi/lo split + pack of eight tanh values (T[i] = (t_i, t_{i+4})) into operand words 4g..4g+3
__device__ __forceinline__ void split_pack8(const uint64_t (&T)[4], uint32_t* hi2, uint32_t* lo2) {
    float t[8], h[8], l[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) upk(T[i], t[i], t[i + 4]);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = __uint_as_float(__float_as_uint(t[i]) & 0xFFFFE000u);
#pragma unroll
    for (int i = 0; i < 4; ++i) upk(sub2(T[i], pk(h[i], h[i + 4])), l[i], l[i + 4]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hi2[i] = pack_half2(h[2 * i], h[2 * i + 1]);
        lo2[i] = pack_half2(l[2 * i], l[2 * i + 1]);
    }
}

// MIX: share of exponentials on the FMA pipe: 0 -> none, 1 -> 1/4, 2 -> 3/8 (groups alternate 2,1 pairs), 3 -> 1/2
template <int MIX>
__device__ __forceinline__ void epilogue_split_mix(const uint32_t* v, uint32_t (&hi2)[16], uint32_t (&lo2)[16]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint64_t T[4];
        if (MIX == 0) tanh8_mix<0>(&v[8 * g], T);
        else if (MIX == 1) tanh8_mix<1>(&v[8 * g], T);
        else if (MIX == 2) { if (g & 1) tanh8_mix<1>(&v[8 * g], T); else tanh8_mix<2>(&v[8 * g], T); }
        else tanh8_mix<2>(&v[8 * g], T);
        split_pack8(T, &hi2[4 * g], &lo2[4 * g]);
    }
}

#if MEMS_TANH_GROUP == 4
// Same contract as tanh8 with two independent 4-way reciprocal trees (shorter dependency chain, 1.25 MUFU/activation):
// lane 0 of every pair belongs to quad {0,1,2,3}, lane 1 to quad {4,5,6,7}; each lane is inverted separately.
__device__ __forceinline__ void tanh8_g4(const uint32_t* v, uint64_t (&T)[4]) {
    float q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = ex2_approx(-fabsf(__uint_as_float(v[i])));
    const uint64_t one2 = pk(1.0f, 1.0f);
    const uint64_t A0 = add2(pk(q[0], q[4]), one2), A1 = add2(pk(q[1], q[5]), one2);
    const uint64_t A2 = add2(pk(q[2], q[6]), one2), A3 = add2(pk(q[3], q[7]), one2);
    const uint64_t B0 = mul2(A0, A1), B1 = mul2(A2, A3);
    float Q0, Q1;
    upk(mul2(B0, B1), Q0, Q1);
    const uint64_t Rq = pk(rcp_approx(Q0), rcp_approx(Q1));    // one reciprocal per quad
    const uint64_t Rb0 = mul2(Rq, B1), Rb1 = mul2(Rq, B0);
    const uint64_t two2 = pk(2.0f, 2.0f), m1 = pk(-1.0f, -1.0f);
    const uint64_t U0 = fma2(mul2(Rb0, A1), two2, m1), U1 = fma2(mul2(Rb0, A0), two2, m1);
    const uint64_t U2 = fma2(mul2(Rb1, A3), two2, m1), U3 = fma2(mul2(Rb1, A2), two2, m1);
    float a, b;
    upk(U0, a, b); T[0] = pk(copysignf(a, __uint_as_float(v[0])), copysignf(b, __uint_as_float(v[4])));
    upk(U1, a, b); T[1] = pk(copysignf(a, __uint_as_float(v[1])), copysignf(b, __uint_as_float(v[5])));
    upk(U2, a, b); T[2] = pk(copysignf(a, __uint_as_float(v[2])), copysignf(b, __uint_as_float(v[6])));
    upk(U3, a, b); T[3] = pk(copysignf(a, __uint_as_float(v[3])), copysignf(b, __uint_as_float(v[7])));
}
#define TANH8 tanh8_g4
#else
#define TANH8 tanh8
#endif

// 32 accumulator columns -> tanh -> fp16 hi/lo operand words (16 + 16 packed pairs)
__device__ __forceinline__ void epilogue_split(const uint32_t* v, uint32_t (&hi2)[16], uint32_t (&lo2)[16]) {
#if MEMS_EXP_EPI == 1      // timing experiment only (wrong numerics): MUFU.TANH + one pack, no lo operand
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float a, b;
        asm("tanh.approx.f32 %0, %1;" : "=f"(a) : "f"(__uint_as_float(v[2 * i])));
        asm("tanh.approx.f32 %0, %1;" : "=f"(b) : "f"(__uint_as_float(v[2 * i + 1])));
        hi2[i] = pack_half2(a, b);
        lo2[i] = 0u;
    }
    return;
#elif MEMS_EXP_EPI == 2    // timing experiment only: split + pack without tanh
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float t0_ = __uint_as_float(v[2 * i]), t1_ = __uint_as_float(v[2 * i + 1]);
        const float h0 = __uint_as_float(v[2 * i] & 0xFFFFE000u), h1 = __uint_as_float(v[2 * i + 1] & 0xFFFFE000u);
        float l0, l1;
        upk(sub2(pk(t0_, t1_), pk(h0, h1)), l0, l1);
        hi2[i] = pack_half2(h0, h1);
        lo2[i] = pack_half2(l0, l1);
    }
    return;
#endif
    epilogue_split_mix<MEMS_EX2_MIX>(v, hi2, lo2);
}

// 32 accumulator columns of the last tanh layer -> dot with the 64->1 output weights.
// w7p holds the weights permuted per group of eight as (w0,w4, w1,w5, w2,w6, w3,w7) to match tanh8's pairing.
__device__ __forceinline__ uint64_t epilogue_dot(const uint32_t* v, const float* w7p, uint64_t acc) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint64_t T[4];
        tanh8_mix<(MEMS_EX2_MIX > 0) ? 1 : 0>(&v[8 * g], T);
        const float4 wa = *reinterpret_cast<const float4*>(w7p + 8 * g);
        const float4 wb = *reinterpret_cast<const float4*>(w7p + 8 * g + 4);
        acc = fma2(T[0], pk(wa.x, wa.y), acc);
        acc = fma2(T[1], pk(wa.z, wa.w), acc);
        acc = fma2(T[2], pk(wb.x, wb.y), acc);
        acc = fma2(T[3], pk(wb.z, wb.w), acc);
    }
    return acc;
}

__device__ __forceinline__ void split3(float v, float& h, float& m, float& l) {
    h = __half2float(__float2half_rn(v));
    const float r1 = v - h;
    m = __half2float(__float2half_rn(r1));
    l = r1 - m;
}

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

// MMA issuer warp of one group (all 32 lanes run the loop convergently; one elected lane issues): for every
// surrogate pass the workers announce (ctrl_tiles via bar_S) it serves the group's two slots in the fixed order
// the workers produce them; ctrl_tiles < 0 ends the loop.  `g` must be warp-uniform.
__device__ __forceinline__ void mma_issuer_loop(TcSmem& sm, GroupSmem& gs, int g) {
    const uint32_t wimg = smem_u32(sm.wimg);
    const uint32_t base = __shfl_sync(0xffffffffu, sm.tmem_base, 0) + (uint32_t)(g * GSLOT * 128);
    uint32_t phA = 0, phS = 0;
    for (;;) {
        mbar_wait(&gs.bar_S, phS);
        phS ^= 1u;
        const int n_tiles = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile int*>(&gs.ctrl_tiles), 0);
        if (n_tiles < 0) break;
        const int n_pairs = (n_tiles + 1) >> 1;
        for (int pair = 0; pair < n_pairs; ++pair) {
#pragma unroll 1
            for (int layer = 0; layer < MEMS_HIDDEN; ++layer) {
#pragma unroll
                for (int s = 0; s < GSLOT; ++s) {
                    if (pair * 2 + s < n_tiles) {
                        mbar_wait(&gs.bar_A[s], (phA >> s) & 1u);
                        phA ^= (1u << s);
                        tc_fence_after();
                        if (elect_one()) {
                            issue_layer(wimg, base + s * 128, layer);
                            umma_commit(&gs.bar_D[s]);
                        }
                        __syncwarp();
                    }
                }
            }
        }
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

template <int NH>
__device__ __forceinline__ void group_bar_sync(int g) { named_bar_sync(1 + g, gthreads(NH)); }

// group barrier + OR-reduction of a predicate over the group's worker threads
template <int NH>
__device__ __forceinline__ bool group_any(int g, bool pred) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\t"
        "barrier.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(out) : "r"((uint32_t)pred), "r"(1 + g), "r"(gthreads(NH)) : "memory");
    return out != 0;
}

// The worker threads of group g evaluate one batch.  Thread (row, h) owns TMEM lane `row` of both slots and
// columns [64/NH * h, 64/NH * (h+1)) of it; row `row` of every tile is (chain j = row mod nc, time sub-row row / nc).
template <int P, int NH>
__device__ __forceinline__ void tc_eval_group(TcSmem& sm, GroupSmem& gs, int g, const double* s_yobs,
                                              const float* s_ts, int T, int stride, int nc, int ncv,
                                              bool skip_dead, float* y_out, int c0, uint32_t& phD) {
    constexpr int NCOL = 64 / NH;                     // accumulator columns per thread
    const int gt = threadIdx.x & (gthreads(NH) - 1);
    const int row = gt & 127;
    const int h = gt >> 7;                            // column half (always 0 when NH == 1)
    const int q = row >> 5;
    const int j = row % nc;                           // nc is any value in 1..128: rows >= nc*tpt of a tile idle
    const int sub = row / nc;
    const int tpt = 128 / nc;                         // time points per tile
    const int Tn = (T + stride - 1) / stride;         // time points of this pass: t = 0, stride, 2*stride, ...
    const int n_tiles = (Tn + tpt - 1) / tpt;
    const int n_pairs = (n_tiles + 1) >> 1;
    const uint32_t tD0 = sm.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * GSLOT * 128);
    const bool chain_live = (j < ncv) && (sub < tpt) && (!skip_dead || gs.cb.live[j]);
    if (gt == 0) {                                    // tell the group's MMA issuer how many tiles follow
        *reinterpret_cast<volatile int*>(&gs.ctrl_tiles) = n_tiles;
        mbar_arrive(&gs.bar_S);
    }

    uint32_t a1[16];
    if (h == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a1[i] = 0u;
#pragma unroll
        for (int p = 0; p < P; ++p) put_input<P>(a1, p, (j < ncv) ? gs.cb.xs[p][j] : 0.0f);
        set_half<P>(a1, 5 * (P + 1) + 0, 1.0f);       // ones that pick up b_hi, b_mid, b_lo
        set_half<P>(a1, 5 * (P + 1) + 1, 1.0f);
        set_half<P>(a1, 5 * (P + 1) + 2, 1.0f);
    }

    // prologue: layer-1 operands of the first tile of each slot
#pragma unroll
    for (int s = 0; s < GSLOT; ++s) {
        if (s < n_tiles) {
            if (h == 0) {
                const int ti = s * tpt + sub;
                put_input<P>(a1, P, s_ts[(ti < Tn && sub < tpt) ? ti * stride : 0]);
                tmem_st16(tD0 + s * 128 + 64, a1);
                tc_wait_st();
            }
            tc_fence_before();
            mbar_arrive(&gs.bar_A[s]);
        }
    }

    double Sp = 0.0;
    for (int pair = 0; pair < n_pairs; ++pair) {
#pragma unroll 1
        for (int layer = 0; layer < MEMS_HIDDEN; ++layer) {
#pragma unroll
            for (int s = 0; s < GSLOT; ++s) {
                const int tile = pair * 2 + s;
                if (tile < n_tiles) {
                    const uint32_t tD = tD0 + s * 128;
                    mbar_wait(&gs.bar_D[s], (phD >> s) & 1u);
                    phD ^= (1u << s);
                    tc_fence_after();
                    uint32_t v[NCOL];
#pragma unroll
                    for (int c = 0; c < NCOL / 32; ++c) tmem_ld32(tD + NCOL * h + 32 * c, v + 32 * c);
                    tc_wait_ld();
                    if (layer < MEMS_HIDDEN - 1) {
#pragma unroll
                        for (int c = 0; c < NCOL / 32; ++c) {
                            uint32_t hi2[16], lo2[16];
                            epilogue_split(v + 32 * c, hi2, lo2);
                            tmem_st16(tD + 64 + (NCOL / 2) * h + 16 * c, hi2);
                            tmem_st16(tD + 96 + (NCOL / 2) * h + 16 * c, lo2);
                        }
                        tc_wait_st();
                        tc_fence_before();
                        mbar_arrive(&gs.bar_A[s]);
                    } else {
                        uint64_t acc = pk(0.0f, 0.0f);
#pragma unroll
                        for (int c = 0; c < NCOL / 32; ++c) acc = epilogue_dot(v + 32 * c, sm.w7 + NCOL * h + 32 * c, acc);
                        float oa, ob;
                        upk(acc, oa, ob);
                        float o = oa + ob;
                        if (NH == 2) {                 // the upper column half hands its partial dot to the lower one
                            if (h == 1) {
                                gs.xchg[s][row] = o;
                                mbar_arrive(&gs.bar_X[s][q]);          // release: the store above is visible to the waiter
                            } else {
                                mbar_wait(&gs.bar_X[s][q], (phD >> (2 + s)) & 1u);   // bits 2,3 of phD: bar_X parity
                                phD ^= (4u << s);
                                o += gs.xchg[s][row];
                            }
                        }
                        if (h == 0) {
                            o += sm.b7;
                            const int ti = tile * tpt + sub;
                            const int t = ti * stride;
                            if (chain_live && ti < Tn) {
                                if (y_out) y_out[(size_t)(c0 + j) * T + t] = o;
                                const double r = s_yobs[t] - (double)o;
                                Sp = fma(r, r, Sp);
                            }
                        }
                        // this slot's D and A are free again: stage the layer-1 operand of its next tile
                        const int nt = tile + GSLOT;
                        if (nt < n_tiles) {
                            if (h == 0) {
                                const int ti2 = nt * tpt + sub;
                                put_input<P>(a1, P, s_ts[(ti2 < Tn && sub < tpt) ? ti2 * stride : 0]);
                                tmem_st16(tD + 64, a1);
                                tc_wait_st();
                            }
                            tc_fence_before();
                            mbar_arrive(&gs.bar_A[s]);
                        } else {
                            tc_fence_before();
                        }
                    }
                }
            }
        }
    }
    if (h == 0) gs.Spart[row] = Sp;
    group_bar_sync<NH>(g);
    if (gt < nc) {
        double S = 0.0;                                // fixed order: time sub-rows ascending
        for (int u = 0; u < tpt; ++u) S += gs.Spart[u * nc + gt];
        gs.Spart[gt] = S;   // thread gt only ever reads slots {u*nc + gt}: no cross-thread hazard
    }
}

// Common prologue: barriers, TMEM, weight image (bulk TMA), y_obs / time columns.
template <int NH>
__device__ __forceinline__ void tc_prologue(TcSmem& sm, const MemsProblem& pb, double* s_yobs, float* s_ts) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    if (tid == 0) {
        for (int g = 0; g < NGROUP; ++g)
            for (int s = 0; s < GSLOT; ++s) {
                mbar_init(&sm.grp[g].bar_A[s], gthreads(NH));
                mbar_init(&sm.grp[g].bar_D[s], 1);
                for (int q = 0; q < 4; ++q) mbar_init(&sm.grp[g].bar_X[s][q], 32);
            }
        for (int g = 0; g < NGROUP; ++g) { mbar_init(&sm.grp[g].bar_S, 1); sm.grp[g].ctrl_tiles = 0; }
        mbar_init(&sm.bar_w, 1);
        sm.c = pb.c;
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == nworker(NH) / 32) {
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
    for (int i = tid; i < pb.c.T; i += nthreads(NH)) { s_yobs[i] = pb.yobs[i]; s_ts[i] = pb.ts[i]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(&sm.bar_w, 0);
}

template <int NH>
__device__ __forceinline__ void tc_epilogue_free(TcSmem& sm) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == nworker(NH) / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem_base), "r"(512) : "memory");
    }
}

template <int P, int NH>
__global__ void __launch_bounds__(nthreads(NH), 1)
tc_run_kernel(const MemsProblem* __restrict__ pbp, MemsChains ch, MemsRunArgs ra) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
    const MemsProblem& pb = *pbp;
    double* s_yobs = reinterpret_cast<double*>(smem_raw + sizeof(TcSmem));
    float* s_ts = reinterpret_cast<float*>(s_yobs + pb.c.T);
    tc_prologue<NH>(sm, pb, s_yobs, s_ts);

    constexpr int NW = nworker(NH), GT = gthreads(NH);
    const int T = sm.c.T, nc = ra.nc;
    const int tid = threadIdx.x;
    const int nbatch = (ch.C + nc - 1) / nc;
    const int g = (tid < NW) ? (tid / GT) : ((tid - NW) >> 5);
    GroupSmem& gs = sm.grp[g];
    const int gid = blockIdx.x * NGROUP + g, gstride = gridDim.x * NGROUP;

    if (tid >= NW) {
        const int gw = __shfl_sync(0xffffffffu, (tid - NW) >> 5, 0);     // provably warp-uniform group index
        mma_issuer_loop(sm, sm.grp[gw], gw);
    } else {
        const int gt = tid & (GT - 1);
        uint32_t phD = 0;
        for (int batch = gid; batch < nbatch; batch += gstride) {
            const int c0 = batch * nc;
            const int ncv = min(nc, ch.C - c0);
            group_bar_sync<NH>(g);
            batch_load(gs.cb, ch, P, ra.mode, c0, gt, ncv);
            for (int s = 0; s < ra.n_steps; ++s) {
                for (int sub = 0; sub < ra.n_sub; ++sub) {
                    if (gt < ncv) batch_pre(gs.cb, sm.c, ch, ra, c0, gt, s, sub);
                    // barrier + "does any chain of the batch need this pass?" (all proposals outside the prior box,
                    // or no proposal promoted by the DA first stage -> the pass is skipped entirely)
                    const bool any = group_any<NH>(g, gt < ncv && gs.cb.live[gt]);
                    if (any)
                        tc_eval_group<P, NH>(sm, gs, g, s_yobs, s_ts, T, batch_pass_stride(ra, sub), nc, ncv, true,
                                             nullptr, c0, phD);
                    if (gt < ncv) batch_post(gs.cb, sm.c, ch, ra, c0, gt, s, sub, any ? gs.Spart[gt] : 0.0);
                }
            }
            batch_store(gs.cb, ch, P, ra.mode, c0, gt, ncv);
        }
        group_bar_sync<NH>(g);
        if (gt == 0) {
            *reinterpret_cast<volatile int*>(&gs.ctrl_tiles) = -1;
            mbar_arrive(&gs.bar_S);
        }
    }
    tc_epilogue_free<NH>(sm);
}

template <int P, int NH>
__global__ void __launch_bounds__(nthreads(NH), 1)
tc_forward_kernel(const MemsProblem* __restrict__ pbp, const double* __restrict__ x, int n, int nc, float* y_out,
                  double* ll_out, int apply_prior, int stride) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
    const MemsProblem& pb = *pbp;
    double* s_yobs = reinterpret_cast<double*>(smem_raw + sizeof(TcSmem));
    float* s_ts = reinterpret_cast<float*>(s_yobs + pb.c.T);
    tc_prologue<NH>(sm, pb, s_yobs, s_ts);

    constexpr int NW = nworker(NH), GT = gthreads(NH);
    const int T = sm.c.T;
    const int tid = threadIdx.x;
    const int nbatch = (n + nc - 1) / nc;
    const int g = (tid < NW) ? (tid / GT) : ((tid - NW) >> 5);
    GroupSmem& gs = sm.grp[g];
    const int gid = blockIdx.x * NGROUP + g, gstride = gridDim.x * NGROUP;

    if (tid >= NW) {
        const int gw = __shfl_sync(0xffffffffu, (tid - NW) >> 5, 0);     // provably warp-uniform group index
        mma_issuer_loop(sm, sm.grp[gw], gw);
    } else {
        const int gt = tid & (GT - 1);
        uint32_t phD = 0;
        for (int batch = gid; batch < nbatch; batch += gstride) {
            const int c0 = batch * nc;
            const int ncv = min(nc, n - c0);
            group_bar_sync<NH>(g);
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
            group_bar_sync<NH>(g);
            tc_eval_group<P, NH>(sm, gs, g, s_yobs, s_ts, T, stride, nc, ncv, false, y_out, c0, phD);
            if (gt < ncv && ll_out) {
                // a strided (coarse) pass is rescaled to the full trace length, as in the DA first stage
                const int Tn = (T + stride - 1) / stride;
                const double ll = sm.c.neg_half_inv_sigma2 * gs.Spart[gt] * ((double)T / (double)Tn);
                ll_out[c0 + gt] = (apply_prior && !gs.cb.inb[gt]) ? -CUDART_INF : ll;
            }
        }
        group_bar_sync<NH>(g);
        if (gt == 0) {
            *reinterpret_cast<volatile int*>(&gs.ctrl_tiles) = -1;
            mbar_arrive(&gs.bar_S);
        }
    }
    tc_epilogue_free<NH>(sm);
}

// ------------------------------------------------------------------------------------------
// Self-test: D[128][64] = A[128][64] * B[64][64]^T through the same descriptors / TMEM paths.
//   variant 0: A from TMEM (tcgen05.st packed fp16), variant 1: A from shared memory (SS form).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sw128_offset(int row, int k) {   // byte offset of element (row, k) in a block
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

__global__ void __launch_bounds__(160, 1)
umma_selftest_kernel(int variant, const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sB = smem_raw;                    // 8 KB
    unsigned char* sA = smem_raw + BLK;              // 16 KB
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + 3 * BLK);
    uint32_t* tbase = reinterpret_cast<uint32_t*>(smem_raw + 3 * BLK + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    for (int i = tid; i < 64 * 64; i += blockDim.x) {
        const int n = i >> 6, k = i & 63;
        *reinterpret_cast<__half*>(sB + sw128_offset(n, k)) = B[i];
    }
    if (variant == 2) {          // split chain: A = [A_hi; A_lo] (256 x 64), B = [W_hi; W_lo] (128 x 64)
        for (int i = tid; i < 64 * 64; i += blockDim.x) {
            const int n = i >> 6, k = i & 63;
            *reinterpret_cast<__half*>(sA + sw128_offset(n, k)) = B[64 * 64 + i];
        }
    } else {
        for (int i = tid; i < 128 * 64; i += blockDim.x) {
            const int m = i >> 6, k = i & 63;
            *reinterpret_cast<__half*>(sA + sw128_offset(m, k)) = A[i];
        }
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tbase)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = *tbase;
    if (warp < 4) {
        const uint32_t tl = base + ((uint32_t)(warp * 32) << 16);
        if (variant == 0 || variant == 2) {
            const int m = warp * 32 + lane;
            const int nparts = (variant == 2) ? 2 : 1;
            for (int part = 0; part < nparts; ++part) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t w[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int k = 32 * c + 2 * i;
                        const __half* src = A + (size_t)(part * 128 + m) * 64 + k;
                        w[i] = (uint32_t)__half_as_ushort(src[0]) | ((uint32_t)__half_as_ushort(src[1]) << 16);
                    }
                    tmem_st16(tl + 64 + 32 * part + 16 * c, w);
                }
            }
            tc_wait_st();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4 && lane == 0) {
        tc_fence_after();
        const uint64_t b = make_desc(smem_u32(sB));
        const uint64_t a = make_desc(smem_u32(sA));
        if (variant == 2) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ts(base, base + 64 + 8 * ks, b + 2 * ks, ks > 0 ? 1u : 0u);   // hi*hi
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ts(base, base + 96 + 8 * ks, b + 2 * ks, 1u);                 // lo*hi
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ts(base, base + 64 + 8 * ks, a + 2 * ks, 1u);                 // hi*lo
        } else {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (variant == 0) umma_ts(base, base + 64 + 8 * ks, b + 2 * ks, ks > 0 ? 1u : 0u);
                else umma_ss(base, a + 2 * ks, b + 2 * ks, ks > 0 ? 1u : 0u);
            }
        }
        umma_commit(bar);
    }
    if (warp < 4) {
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t tl = base + ((uint32_t)(warp * 32) << 16);
        const int m = warp * 32 + lane;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(tl + 32 * c, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) D[m * 64 + 32 * c + i] = __uint_as_float(v[i]);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(128) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// Micro-benchmarks of the epilogue's building blocks (tools/microbench.py): cycles for `iters` repetitions of
//   0: tcgen05.ld 32x32b.x32 + wait          1: 2 x tcgen05.st x16 + wait       2: tanh + split arithmetic (registers)
//   3: ld + arithmetic + st                  4: 2 x ld x32, one wait            5: ld 32x32b.x64 + wait
//   6: 2 x ld 16x256b.x8 (both lane halves)  7: tanh arithmetic only            8: split + pack only
//   9: tanh.approx + single pack             10: ld x32 without consuming wait per iteration (4 in flight)
// One CTA per SM, blockDim = 32*warps.  All register arrays are indexed statically (no local memory).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#ifndef MEMS_MB_THREADS
#define MEMS_MB_THREADS 512    // 576 -> the register budget of the product kernel (112 per thread)
#endif
template <int MODE>
__global__ void __launch_bounds__(MEMS_MB_THREADS, 1)
microbench_kernel(int iters, long long* cycles_out, unsigned int* sink) {
    __shared__ uint32_t tbase_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tbase_s;
    const uint32_t tl = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) * 128) & 511);
    uint32_t v[64], hi2[16], lo2[16];
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __float_as_uint(0.01f * (float)(i + tid % 7) - 0.1f);
#pragma unroll
    for (int i = 0; i < 16; ++i) { hi2[i] = i; lo2[i] = 2 * i; }
    tmem_st16(tl, hi2);
    tmem_st16(tl + 16, lo2);
    tmem_st16(tl + 32, hi2);
    tmem_st16(tl + 48, lo2);
    tc_wait_st();
    __syncthreads();
    const long long t0 = clock64();
    if (iters & 1) {                       // odd iteration counts: stagger the warps of an SMSP by ~a quarter iteration each
        const long long until = t0 + (long long)(warp >> 2) * ((MODE == 17) ? 900 : 450);
        while (clock64() < until) {}
    }
    if (MODE == 10) {
        for (int it = 0; it < iters; it += 4) {
            tmem_ld32(tl, v);
            tmem_ld32(tl + 32, v + 32);
            tmem_ld32(tl, v);
            tmem_ld32(tl + 32, v + 32);
            tc_wait_ld();
            acc ^= v[0] ^ v[31] ^ v[32] ^ v[63];
        }
    } else {
        for (int it = 0; it < iters; ++it) {
            if (MODE == 0) {
                tmem_ld32(tl, v);
                tc_wait_ld();
                acc ^= v[0] ^ v[31];
            } else if (MODE == 1) {
                hi2[0] ^= acc;
                tmem_st16(tl + 32, hi2);
                tmem_st16(tl + 48, lo2);
                tc_wait_st();
                acc += 1;
            } else if (MODE == 2) {
                v[0] ^= (acc & 0x3FF);
                v[9] ^= (acc & 0x1FF);
                v[18] ^= (acc & 0x3FF);
                v[27] ^= (acc & 0x1FF);
                epilogue_split(v, hi2, lo2);
                acc ^= hi2[0] + lo2[3] + hi2[5] + lo2[6] + hi2[9] + lo2[10] + hi2[12] + lo2[15];
            } else if (MODE == 3) {
                tmem_ld32(tl, v);
                tc_wait_ld();
                epilogue_split(v, hi2, lo2);
                tmem_st16(tl + 32, hi2);
                tmem_st16(tl + 48, lo2);
                tc_wait_st();
                acc ^= hi2[0];
            } else if (MODE >= 11 && MODE <= 14) {
                tmem_ld32(tl, v);
                tc_wait_ld();
                epilogue_split_mix<MODE - 11>(v, hi2, lo2);
                tmem_st16(tl + 32, hi2);
                tmem_st16(tl + 48, lo2);
                tc_wait_st();
                acc ^= hi2[0];
            } else if (MODE == 15 || MODE == 16) {
                // software pipeline across chunks: exponential/tanh stage of chunk it+1 in the same basic block as the
                // split/pack/store stage of chunk it (Tc = carried tanh values)
                uint64_t* Tc = reinterpret_cast<uint64_t*>(v + 32);       // 16 packed pairs = v[32..63]
                if (it == 0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) tanh8_mix<MODE - 15>(&v[8 * g], *reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]));
                }
                tmem_ld32(tl, v);
                tc_wait_ld();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    split_pack8(*reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]), &hi2[4 * g], &lo2[4 * g]);
                    tanh8_mix<MODE - 15>(&v[8 * g], *reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]));
                }
                tmem_st16(tl + 32, hi2);
                tmem_st16(tl + 48, lo2);
                tc_wait_st();
                acc ^= hi2[0];
            } else if (MODE == 17) {
                // one thread owns a full 64-column row: intra-thread software pipeline over the 8 groups of eight
                tmem_ld32(tl, v);
                tmem_ld32(tl + 32, v + 32);
                tc_wait_ld();
                uint64_t Ta[4], Tb[4];
                tanh8_mix<0>(&v[0], Ta);
#pragma unroll
                for (int g = 0; g < 8; g += 2) {
                    tanh8_mix<0>(&v[8 * (g + 1)], Tb);
                    split_pack8(Ta, &hi2[(4 * g) & 15], &lo2[(4 * g) & 15]);
                    if (g + 2 < 8) tanh8_mix<0>(&v[8 * (g + 2)], Ta);
                    split_pack8(Tb, &hi2[(4 * g + 4) & 15], &lo2[(4 * g + 4) & 15]);
                    if (g == 2) { tmem_st16(tl, hi2); tmem_st16(tl + 32, lo2); }
                }
                tmem_st16(tl + 16, hi2);
                tmem_st16(tl + 48, lo2);
                tc_wait_st();
                acc ^= hi2[0];
            } else if (MODE == 18 || MODE == 19) {
                // chunk of 32 (mode 18) / 64 (mode 19) columns as a ROLLED loop over groups of eight with the tanh values
                // carried: body = { ld(g+1); split/pack/st(g); wait::ld; tanh(g+1) }
                constexpr int NG = (MODE == 18) ? 4 : 8;
                uint64_t T[4];
                uint32_t w8[8], h4[4], l4[4];
                tmem_ld8(tl, w8);
                tc_wait_ld();
                tanh8_mix<0>(w8, T);
#pragma unroll 1
                for (int g = 1; g < NG; ++g) {
                    tmem_ld8(tl + 8 * g, w8);
                    split_pack8(T, h4, l4);
                    tmem_st4(tl + 64 + 4 * (g - 1), h4);
                    tmem_st4(tl + 96 + 4 * (g - 1), l4);
                    tc_wait_ld();
                    tanh8_mix<0>(w8, T);
                }
                split_pack8(T, h4, l4);
                tmem_st4(tl + 64 + 4 * (NG - 1), h4);
                tmem_st4(tl + 96 + 4 * (NG - 1), l4);
                tc_wait_st();
                acc ^= h4[0];
            } else if (MODE == 29) {
                // mode 15 with the operand words stored per group of eight (st.x4): shorter live ranges
                uint64_t* Tc = reinterpret_cast<uint64_t*>(v + 32);
                if (it == 0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) tanh8_mix<0>(&v[8 * g], *reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]));
                }
                tmem_ld32(tl, v);
                tc_wait_ld();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t h4[4], l4[4];
                    split_pack8(*reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]), h4, l4);
                    tmem_st4(tl + 64 + 4 * g, h4);
                    tmem_st4(tl + 96 + 4 * g, l4);
                    tanh8_mix<0>(&v[8 * g], *reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]));
                    acc ^= h4[0];
                }
                tc_wait_st();
            } else if (MODE == 28) {
                // mode 15 at 16-column granularity (2 + 2 groups of eight in flight per thread)
                uint64_t* Tc = reinterpret_cast<uint64_t*>(v + 32);       // 8 packed pairs
                if (it == 0) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) tanh8_mix<0>(&v[8 * g], *reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]));
                }
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(tl) : "memory");
                tc_wait_ld();
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    split_pack8(*reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]), &hi2[4 * g], &lo2[4 * g]);
                    tanh8_mix<0>(&v[8 * g], *reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]));
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             ::"r"(tl + 64), "r"(hi2[0]), "r"(hi2[1]), "r"(hi2[2]), "r"(hi2[3]), "r"(hi2[4]), "r"(hi2[5]), "r"(hi2[6]), "r"(hi2[7]) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             ::"r"(tl + 96), "r"(lo2[0]), "r"(lo2[1]), "r"(lo2[2]), "r"(lo2[3]), "r"(lo2[4]), "r"(lo2[5]), "r"(lo2[6]), "r"(lo2[7]) : "memory");
                tc_wait_st();
                acc ^= hi2[0];
            } else if (MODE == 4) {
                tmem_ld32(tl, v);
                tmem_ld32(tl + 32, v + 32);
                tc_wait_ld();
                acc ^= v[0] ^ v[31] ^ v[32] ^ v[63];
            } else if (MODE == 5) {
                tmem_ld64(tl, v);
                tc_wait_ld();
                acc ^= v[0] ^ v[31] ^ v[32] ^ v[63];
            } else if (MODE == 6) {
                tmem_ld_16x256b_x8(tl, v);
                tmem_ld_16x256b_x8(tl + (16u << 16), v + 32);
                tc_wait_ld();
                acc ^= v[0] ^ v[31] ^ v[32] ^ v[63];
            } else if (MODE == 7) {
                v[0] ^= (acc & 0x3FF);
                v[9] ^= (acc & 0x1FF);
                v[18] ^= (acc & 0x3FF);
                v[27] ^= (acc & 0x1FF);
                uint32_t x = 0;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint64_t T[4];
                    TANH8(&v[8 * g], T);
                    const uint64_t s01 = add2(T[0], T[1]), s23 = add2(T[2], T[3]);
                    float a, b;
                    upk(add2(s01, s23), a, b);
                    x ^= __float_as_uint(a + b);
                }
                acc ^= x;
            } else if (MODE == 8) {
                v[0] ^= (acc & 0x3FF);
                v[9] ^= (acc & 0x1FF);
                v[18] ^= (acc & 0x3FF);
                v[27] ^= (acc & 0x1FF);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float t0_ = __uint_as_float(v[2 * i]), t1_ = __uint_as_float(v[2 * i + 1]);
                    const float h0 = __uint_as_float(v[2 * i] & 0xFFFFE000u), h1 = __uint_as_float(v[2 * i + 1] & 0xFFFFE000u);
                    float l0, l1;
                    upk(sub2(pk(t0_, t1_), pk(h0, h1)), l0, l1);
                    hi2[i] = pack_half2(h0, h1);
                    lo2[i] = pack_half2(l0, l1);
                }
                acc ^= hi2[0] + lo2[3] + hi2[5] + lo2[6] + hi2[9] + lo2[10] + hi2[12] + lo2[15];
            } else if (MODE == 9) {
                v[0] ^= (acc & 0x3FF);
                v[9] ^= (acc & 0x1FF);
                v[18] ^= (acc & 0x3FF);
                v[27] ^= (acc & 0x1FF);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    hi2[i] = pack_half2(tanh_approx(__uint_as_float(v[2 * i])), tanh_approx(__uint_as_float(v[2 * i + 1])));
                acc ^= hi2[0] + hi2[3] + hi2[5] + hi2[6] + hi2[9] + hi2[10] + hi2[12] + hi2[15];
            }
        }
    }
    const long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + tid] = acc;
    __shared__ long long tmax;
    if (tid == 0) tmax = 0;
    __syncthreads();
    atomicMax(reinterpret_cast<unsigned long long*>(&tmax), (unsigned long long)(t1 - t0));
    tc_fence_before();
    __syncthreads();
    if (tid == 0) cycles_out[blockIdx.x] = tmax;
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
    }
}

// Pipe-throughput probes (16 independent dependency chains per thread, `iters` rounds):
//   20: MUFU.EX2   21: F2FP (cvt.rn.f16x2.f32)   22: MUFU.EX2 + F2FP interleaved   23: FFMA2   24: LOP3
//   25: MUFU.EX2 + 4 FFMA2 each   26: MUFU.RCP     27: FADD2
template <int MODE>
__global__ void __launch_bounds__(512, 1)
pipe_probe_kernel(int iters, long long* cycles_out, unsigned int* sink) {
    const int tid = threadIdx.x;
    float x[16];
    uint64_t p[16];
    uint32_t u[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        x[i] = -0.01f * (float)(i + (tid & 7)) - 0.1f;
        p[i] = pk(x[i], 0.5f * x[i]);
        u[i] = 0x3c003c00u + i + tid;
    }
    const uint64_t c2 = pk(0.999f, 1.001f), d2 = pk(1e-3f, -1e-3f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 20 || MODE == 22 || MODE == 25) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(-fabsf(x[i])));
            if (MODE == 26) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(x[i]));
            if (MODE == 21 || MODE == 22) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])), "f"(__uint_as_float(u[i] ^ 0x1000u)));
            if (MODE == 23) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(c2), "l"(d2));
            if (MODE == 25) {
#pragma unroll
                for (int k = 0; k < 4; ++k) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[(i + 4 * k) & 15]) : "l"(p[(i + 4 * k) & 15]), "l"(c2), "l"(d2));
            }
            if (MODE == 27) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(p[i]) : "l"(p[i]), "l"(d2));
            if (MODE == 24) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(u[i]) : "r"(u[i]), "r"(u[(i + 1) & 15]), "r"(0x5555u));
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float a, b; upk(p[i], a, b); acc ^= __float_as_uint(x[i]) ^ u[i] ^ __float_as_uint(a) ^ __float_as_uint(b); }
    sink[blockIdx.x * blockDim.x + tid] = acc;
    __shared__ long long tmax;
    if (tid == 0) tmax = 0;
    __syncthreads();
    atomicMax(reinterpret_cast<unsigned long long*>(&tmax), (unsigned long long)(t1 - t0));
    __syncthreads();
    if (tid == 0) cycles_out[blockIdx.x] = tmax;
}

// MMA issue-rate probe: `iters` x 4 back-to-back tcgen05.mma (M=128, N=n, K=16, kind::f16) with the A operand in
// TMEM (ts != 0) or shared memory; reports SM cycles from first issue to commit completion.
__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int n, int ts, int iters, long long* cycles_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sB = smem_raw;                    // 32 KB: up to 256 rows x 128 B
    unsigned char* sA = smem_raw + 4 * BLK;          // 16 KB
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + 6 * BLK);
    uint32_t* tbase = reinterpret_cast<uint32_t*>(smem_raw + 6 * BLK + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 6 * BLK / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tbase)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        const uint32_t base = __shfl_sync(0xffffffffu, *tbase, 0);
        const uint32_t idesc = (1u << 4) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t b = make_desc(smem_u32(sB));
        const uint64_t a = make_desc(smem_u32(sA));
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (ts) {
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(base), "r"(base + 256 + 8 * ks), "l"(b + 2 * ks), "r"(idesc), "r"(1u) : "memory");
                    } else {
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(base), "l"(a + 2 * ks), "l"(b + 2 * ks), "r"(idesc), "r"(1u) : "memory");
                    }
                }
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(bar);
        __syncwarp();
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        if (tid == 0) cycles_out[blockIdx.x] = t1 - t0;
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tbase), "r"(512) : "memory");
    }
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

template <int NH, typename K1, typename K2, typename K3, typename K4, typename... Args>
int launch_by_p(int Pv, K1 k1, K2 k2, K3 k3, K4 k4, int grid, size_t smem, cudaStream_t st, Args... args) {
    auto go = [&](auto k) -> int {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<grid, nthreads(NH), smem, st>>>(args...);
        return (int)cudaGetLastError();
    };
    switch (Pv) {
        case 1: return go(k1);
        case 2: return go(k2);
        case 3: return go(k3);
        default: return go(k4);
    }
}

int g_tc_variant = -1;   // -1: read MEMS_TC_NH from the environment once
int tc_nh() {
    if (g_tc_variant < 0) {
        const char* e = getenv("MEMS_TC_NH");
        g_tc_variant = (e && e[0] == '1') ? 1 : 2;
    }
    return g_tc_variant;
}

}  // namespace

size_t mems_tc_wimg_bytes() { return WIMG_BYTES; }

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
    if (tc_nh() == 1)
        return launch_by_p<1>(ctx.pb_host->c.P, tc_run_kernel<1, 1>, tc_run_kernel<2, 1>, tc_run_kernel<3, 1>, tc_run_kernel<4, 1>,
                              grid, smem, ctx.stream, ctx.pb_dev, ctx.chains, ra);
    return launch_by_p<2>(ctx.pb_host->c.P, tc_run_kernel<1, 2>, tc_run_kernel<2, 2>, tc_run_kernel<3, 2>, tc_run_kernel<4, 2>,
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
    if (tc_nh() == 1)
        return launch_by_p<1>(ctx.pb_host->c.P, tc_forward_kernel<1, 1>, tc_forward_kernel<2, 1>, tc_forward_kernel<3, 1>,
                              tc_forward_kernel<4, 1>, gr, smem, ctx.stream, ctx.pb_dev, x_dev, n, nc, y_out, ll_out, apply_prior, t_stride);
    return launch_by_p<2>(ctx.pb_host->c.P, tc_forward_kernel<1, 2>, tc_forward_kernel<2, 2>, tc_forward_kernel<3, 2>,
                          tc_forward_kernel<4, 2>, gr, smem, ctx.stream, ctx.pb_dev, x_dev, n, nc, y_out, ll_out, apply_prior, t_stride);
}

int mems_tc_microbench(int mode, int warps, int iters, int blocks, long long* cycles_out, unsigned int* sink, cudaStream_t st) {
    switch (mode) {
#define MB_CASE(M) case M: microbench_kernel<M><<<blocks, 32 * warps, 0, st>>>(iters, cycles_out, sink); break;
        MB_CASE(0) MB_CASE(1) MB_CASE(2) MB_CASE(3) MB_CASE(4) MB_CASE(5) MB_CASE(6) MB_CASE(7) MB_CASE(8) MB_CASE(9) MB_CASE(10) MB_CASE(11) MB_CASE(12) MB_CASE(13) MB_CASE(14) MB_CASE(15) MB_CASE(16) MB_CASE(17) MB_CASE(18) MB_CASE(19)
#undef MB_CASE
#define MB_CASE(M) case M: microbench_kernel<M><<<blocks, 32 * warps, 0, st>>>(iters, cycles_out, sink); break;
#define PP_CASE(M) case M: pipe_probe_kernel<M><<<blocks, 32 * warps, 0, st>>>(iters, cycles_out, sink); break;
        PP_CASE(20) PP_CASE(21) PP_CASE(22) PP_CASE(23) PP_CASE(24) PP_CASE(25) PP_CASE(26) PP_CASE(27) MB_CASE(28) MB_CASE(29)
#undef PP_CASE
        case 100: case 101: {          // MMA rate probe: `warps` carries N/8 (8, 16 or 32), mode 100 = A in TMEM, 101 = A in smem
            const size_t smem = 6 * BLK + 64;
            cudaError_t e = cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            mma_rate_kernel<<<blocks, 128, smem, st>>>(warps * 8, mode == 100 ? 1 : 0, iters, cycles_out);
            break;
        }
        default: return (int)cudaErrorInvalidValue;
    }
    return (int)cudaGetLastError();
}

int mems_tc_selftest(int variant, const void* a_f16, const void* b_f16, float* d_out, cudaStream_t st) {
    const size_t smem = 3 * BLK + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    umma_selftest_kernel<<<1, 160, smem, st>>>(variant, static_cast<const __half*>(a_f16),
                                               static_cast<const __half*>(b_f16), d_out);
    return (int)cudaGetLastError();
}
