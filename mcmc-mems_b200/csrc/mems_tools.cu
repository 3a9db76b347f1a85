// Development tools of the tensor-core path, built into their OWN library (tools/libmems_b200_tools.so, header
// tools/mems_b200_tools.h) -- nothing here is on the product path or in the public C ABI:
//   * measurement kernels (tools/microbench.py): epilogue building blocks, pipe throughput probes, tcgen05.mma issue
//     rate; the numbers they produce are quoted in DESIGN.md and stored under profiles/;
//   * the UMMA plumbing self-test (tests/test_gpu.py::test_umma_selftest, tools/probe_numerics.py).
#include <cstdarg>
#include <cstdio>

#include "mems_tc_device.cuh"

namespace {

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
// MAXT: launch bound = register budget (512 threads: 128 registers, 768: 80, 1024: 64)
template <int MODE, int MAXT = MEMS_MB_THREADS>
__global__ void __launch_bounds__(MAXT, 1)
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
            } else if (MODE == 30) {
                // 32 columns as two sequential 16-column halves (ld.x16 -> tanh -> split -> 2 x st.x8): the body that fits
                // a 64-register budget (32 warps per SM)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t w[16], h8[8], l8[8];
                    tmem_ld16(tl + 16 * hf, w);
                    tc_wait_ld();
                    epilogue_split16(w, h8, l8);
                    tmem_st8(tl + 64 + 8 * hf, h8);
                    tmem_st8(tl + 96 + 8 * hf, l8);
                    tc_wait_st();
                    acc ^= h8[0];
                }
            } else if (MODE == 31) {
                // mode 30 with ONE wait::st per 32 columns and the second half's load issued before the first half's math
                uint32_t w0[16], w1[16], h8[8], l8[8];
                tmem_ld16(tl, w0);
                tmem_ld16(tl + 16, w1);
                tc_wait_ld();
                epilogue_split16(w0, h8, l8);
                tmem_st8(tl + 64, h8);
                tmem_st8(tl + 96, l8);
                acc ^= h8[0];
                epilogue_split16(w1, h8, l8);
                tmem_st8(tl + 72, h8);
                tmem_st8(tl + 104, l8);
                tc_wait_st();
                acc ^= h8[0];
            } else if (MODE == 50) {
                // mode 11 (32 columns at once) with the round-2 arithmetic (folded factor 2, FHADD split)
                tmem_ld32(tl, v);
                tc_wait_ld();
                epilogue_split_v2<4>(v, hi2, lo2);
                tmem_st16(tl + 64, hi2);
                tmem_st16(tl + 96, lo2);
                tc_wait_st();
                acc ^= hi2[0];
            } else if (MODE == 51) {
                // mode 30 (two sequential 16-column halves) with the round-2 arithmetic
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t w[16], h8[8], l8[8];
                    tmem_ld16(tl + 16 * hf, w);
                    tc_wait_ld();
                    epilogue_split_v2<2>(w, h8, l8);
                    tmem_st8(tl + 64 + 8 * hf, h8);
                    tmem_st8(tl + 96 + 8 * hf, l8);
                    tc_wait_st();
                    acc ^= h8[0];
                }
            } else if (MODE == 52) {
                // mode 15 (tanh of the next chunk in one block with split/pack/store of the current) with the round-2 arithmetic
                uint64_t* Tc = reinterpret_cast<uint64_t*>(v + 32);
                if (it == 0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) tanh8_fold(&v[8 * g], *reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]));
                }
                tmem_ld32(tl, v);
                tc_wait_ld();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    split_pack8_rn(*reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]), &hi2[4 * g], &lo2[4 * g]);
                    tanh8_fold(&v[8 * g], *reinterpret_cast<uint64_t(*)[4]>(&Tc[4 * g]));
                }
                tmem_st16(tl + 64, hi2);
                tmem_st16(tl + 96, lo2);
                tc_wait_st();
                acc ^= hi2[0];
            } else if (MODE == 53) {
                // four sequential 8-column pieces (ld.x8 -> tanh -> split -> 2 x st.x4), one wait::st per 32 columns
                uint32_t w[8], h4[4], l4[4];
#pragma unroll
                for (int q8 = 0; q8 < 4; ++q8) {
                    tmem_ld8(tl + 8 * q8, w);
                    tc_wait_ld();
                    epilogue_split_v2<1>(w, h4, l4);
                    tmem_st4(tl + 64 + 4 * q8, h4);
                    tmem_st4(tl + 96 + 4 * q8, l4);
                    acc ^= h4[0];
                }
                tc_wait_st();
            } else if (MODE == 60 || MODE == 61) {
                // mode 53 (four 8-column pieces) with 1/8 (60) or 1/4 (61) of the exponentials on the FMA pipe
                uint32_t w[8], h4[4], l4[4];
#pragma unroll
                for (int q8 = 0; q8 < 4; ++q8) {
                    tmem_ld8(tl + 8 * q8, w);
                    tc_wait_ld();
                    uint64_t T[4];
                    if (MODE == 61 || (q8 & 1)) tanh8_fold<1>(w, T); else tanh8_fold<0>(w, T);
                    split_pack8_rn(T, h4, l4);
                    tmem_st4(tl + 64 + 4 * q8, h4);
                    tmem_st4(tl + 96 + 4 * q8, l4);
                    acc ^= h4[0];
                }
                tc_wait_st();
            } else if (MODE >= 62 && MODE <= 66) {
                // mode 51 (two 16-column halves): 62..64, 66 one reciprocal per sixteen activations with 0, 1/8, 1/4, 1/2 of the
                // exponentials on the FMA pipe; 65 one reciprocal per eight with 1/4
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t w[16], h8[8], l8[8];
                    tmem_ld16(tl + 16 * hf, w);
                    tc_wait_ld();
                    if (MODE == 62) epilogue_split_v2<2, true, 0>(w, h8, l8);
                    if (MODE == 63) epilogue_split_v2<2, true, 1>(w, h8, l8);
                    if (MODE == 64) epilogue_split_v2<2, true, 2>(w, h8, l8);
                    if (MODE == 66) epilogue_split_v2<2, true, 4>(w, h8, l8);
                    if (MODE == 65) epilogue_split_v2<2, false, 2>(w, h8, l8);
                    tmem_st8(tl + 64 + 8 * hf, h8);
                    tmem_st8(tl + 96 + 8 * hf, l8);
                    tc_wait_st();
                    acc ^= h8[0];
                }
            } else if (MODE == 67) {
                // mode 53 with the next piece's load in flight during the current piece's arithmetic
                uint32_t wa[8], wb[8], h4[4], l4[4];
                tmem_ld8(tl, wa);
                tc_wait_ld();
#pragma unroll
                for (int q8 = 0; q8 < 4; ++q8) {
                    uint32_t* cur = (q8 & 1) ? wb : wa;
                    uint32_t* nxt = (q8 & 1) ? wa : wb;
                    if (q8 < 3) tmem_ld8(tl + 8 * (q8 + 1), nxt);
                    epilogue_split_v2<1>(cur, h4, l4);
                    tmem_st4(tl + 64 + 4 * q8, h4);
                    tmem_st4(tl + 96 + 4 * q8, l4);
                    if (q8 < 3) tc_wait_ld();
                    acc ^= h4[0];
                }
                tc_wait_st();
            } else if (MODE == 68) {
                // mode 62 (16-column halves, one reciprocal per sixteen) with one wait::st per 32 columns and the second half's
                // load in flight during the first half's arithmetic
                uint32_t wa[16], wb[16], h8[8], l8[8];
                tmem_ld16(tl, wa);
                tc_wait_ld();
                tmem_ld16(tl + 16, wb);
                epilogue_split_v2<2, true, 0>(wa, h8, l8);
                tmem_st8(tl + 64, h8);
                tmem_st8(tl + 96, l8);
                tc_wait_ld();
                acc ^= h8[0];
                epilogue_split_v2<2, true, 0>(wb, h8, l8);
                tmem_st8(tl + 72, h8);
                tmem_st8(tl + 104, l8);
                tc_wait_st();
                acc ^= h8[0];
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
                    tanh8_mix<0>(&v[8 * g], T);
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
    float x[16], y[16], z[16];
    uint64_t p[16];
    uint32_t u[16], w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        x[i] = -0.01f * (float)(i + (tid & 7)) - 0.1f;
        y[i] = 1.0f + 0.01f * (float)(i + (tid & 3));
        z[i] = 1.0f + 0.02f * (float)(i + (tid & 3));
        p[i] = pk(x[i], 0.5f * x[i]);
        u[i] = 0x3c003c00u + i + tid;
        w[i] = 0x3c003c00u + 3 * i + tid;
    }
    const uint64_t c2 = pk(0.999f, 1.001f), d2 = pk(1e-3f, -1e-3f);
    const float cs = 0.999f + 1e-6f * (float)(tid & 1), ds = 1e-3f;
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
            // pipe-sharing probes (round 2): pairs of independent instructions of two classes, 16 pairs per round
            if (MODE == 32 || MODE == 34 || MODE == 36) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(u[i]) : "r"(u[i]), "r"(u[(i + 1) & 15]), "r"(0x5555u));
            if (MODE == 32 || MODE == 33 || MODE == 45) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(c2), "l"(d2));
            if (MODE == 33 || MODE == 34) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(__uint_as_float(w[i])), "f"(__uint_as_float(w[(i + 1) & 15])));
            if (MODE == 35 || MODE == 36 || MODE == 45) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i]) : "f"(x[i]), "f"(cs), "f"(ds));
            if (MODE == 48 || MODE == 49 || MODE == 54)
                asm volatile("{\n\t.reg .b16 h0, h1;\n\tmov.b32 {h0, h1}, %1;\n\tadd.rn.f32.f16 %0, h0, %0;\n\t}" : "+f"(x[i]) : "r"(w[i]));
            if (MODE == 49) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(c2), "l"(d2));
            if (MODE == 54) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(u[i]) : "r"(u[i]), "r"(u[(i + 1) & 15]), "r"(0x5555u));
            if (MODE == 46) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(u[i]), "r"(u[(i + 1) & 15]));
            if (MODE == 47) asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(u[i]) : "r"(u[i]), "r"(0x3c003c00u));
        }
        if (MODE == 55 || MODE == 56) {
            // the round-2 epilogue's instruction mix per 16 activations without its dependencies: 16 ex2 + 2 rcp, 26 packed
            // FP32x2, 8 scalar FP32, 16 LOP3, 16 F2FP, 16 FHADD.   56: the same without the LOP3 (sign copies)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(-fabsf(x[i])));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(c2), "l"(d2));
                if (MODE == 55) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(u[i]) : "r"(u[i]), "r"(u[(i + 1) & 15]), "r"(0x5555u));
                if ((i % 8) < 5) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p[(i + 5) & 15]) : "l"(p[(i + 5) & 15]), "l"(c2));
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(__uint_as_float(w[i])), "f"(__uint_as_float(w[(i + 1) & 15])));
                asm volatile("{\n\t.reg .b16 h0, h1;\n\tmov.b32 {h0, h1}, %1;\n\tadd.rn.f32.f16 %0, h0, %0;\n\t}" : "+f"(y[i]) : "r"(w[(i + 3) & 15]));
                if ((i % 8) < 4) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(z[i]) : "f"(z[i]), "f"(cs));
                if ((i % 8) == 7) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(z[i]) : "f"(z[i]));
            }
        }
        if (MODE >= 37 && MODE <= 44) {
            // the epilogue's instruction mix per 16 activations without its dependencies: 16 ex2 + 2 rcp, 42 packed FP32x2,
            // 6 scalar FMUL, 32 LOP3, 16 F2FP.   37: all, 38: no LOP3, 39: no F2FP, 40: no MUFU, 41: no packed FP,
            // 42: half the LOP3, 43: LOP3 + F2FP only, 44: MUFU + packed FP only
            constexpr bool MU = MODE != 40 && MODE != 43, PK = MODE != 41 && MODE != 43, LO = MODE != 38 && MODE != 44, CV = MODE != 39 && MODE != 44;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MU) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(-fabsf(x[i])));
                if (PK) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(c2), "l"(d2));
                if (LO) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(u[i]) : "r"(u[i]), "r"(u[(i + 1) & 15]), "r"(0x5555u));
                if (PK) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p[(i + 5) & 15]) : "l"(p[(i + 5) & 15]), "l"(c2));
                if (CV) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(__uint_as_float(w[i])), "f"(__uint_as_float(w[(i + 1) & 15])));
                if (LO && (MODE != 42)) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(u[(i + 7) & 15]) : "r"(u[(i + 7) & 15]), "r"(u[(i + 8) & 15]), "r"(0x3333u));
                if (PK && (i % 8) < 5) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(p[(i + 11) & 15]) : "l"(p[(i + 11) & 15]), "l"(d2));
                if (PK && (i % 8) < 3) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(y[i]) : "f"(y[i]), "f"(cs));
                if (MU && (i % 8) == 7) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y[i]) : "f"(y[i]));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float a, b; upk(p[i], a, b); acc ^= __float_as_uint(x[i]) ^ u[i] ^ w[i] ^ __float_as_uint(y[i]) ^ __float_as_uint(z[i]) ^ __float_as_uint(a) ^ __float_as_uint(b); }
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

// Epilogue + tensor core together (round 2): 16 epilogue warps run the mode-53 body (four 8-column pieces per 32 columns,
// round-2 arithmetic) on the four 128-column slots while warp 16 issues the product kernel's MMA batches (12 TS-form
// + 1 SS-form tcgen05.mma, N = 64) on the same slots: `duty` = 0 no MMAs, 1 one batch then a pause of ~2 batch times
// (the product kernel's tensor duty), 2 back to back.  Tells whether tensor-core TMEM / shared-memory traffic slows the
// epilogue down.  cycles_out[blk] = slowest epilogue thread.
__global__ void __launch_bounds__(544, 1)
epi_mma_kernel(int duty, int iters, long long* cycles_out, unsigned int* sink, int random_data) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sB = smem_raw;                    // B (8 KB) + A of the SS-form MMA (16 KB), zero
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + 3 * BLK);
    uint32_t* tbase = reinterpret_cast<uint32_t*>(smem_raw + 3 * BLK + 16);
    volatile int* stop = reinterpret_cast<volatile int*>(smem_raw + 3 * BLK + 32);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 3 * BLK / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
    if (random_data) {
        // weights in (-0.5, 0.5) and activations in (-1, 1) from a hash: constant (zero) operands hardly toggle the
        // datapaths, which makes the power drawn by modes 200..202 meaningless (tools/power_probe.py uses 210..212)
        for (int i = tid; i < BLK / 2; i += blockDim.x) {
            const unsigned int hsh = (unsigned int)(i + 1) * 2654435761u;
            reinterpret_cast<__half*>(sB)[i] = __float2half_rn((float)(hsh >> 8) * (1.0f / 16777216.0f) - 0.5f);   // gain > 1: the activations neither die out nor all saturate
        }
    }
    if (tid == 0) { mbar_init(bar, 1); *stop = 0; fence_barrier_init(); }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tbase)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = *tbase;
    __shared__ long long tmax;
    if (tid == 0) tmax = 0;
    uint32_t acc = 0;
    if (warp < 16) {
        const uint32_t tl = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            z[i] = 0x3c003c00u;
            if (random_data) {
                const unsigned int hsh = (unsigned int)(tid * 16 + i + 7) * 2246822519u;
                z[i] = pack_half2(((float)(hsh >> 16) * (1.0f / 65536.0f)) * 2.0f - 1.0f, ((float)(hsh & 0xFFFFu) * (1.0f / 65536.0f)) * 2.0f - 1.0f);
            }
        }
        tmem_st16(tl, z); tmem_st16(tl + 16, z); tmem_st16(tl + 32, z); tmem_st16(tl + 48, z);
        tmem_st16(tl + 64, z); tmem_st16(tl + 80, z); tmem_st16(tl + 96, z); tmem_st16(tl + 112, z);
        tc_wait_st();
        tc_fence_before();
        named_bar_sync(1, 544);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t w[8], h4[4], l4[4];
#pragma unroll
            for (int q8 = 0; q8 < 4; ++q8) {
                tmem_ld8(tl + 8 * q8 + 32 * (it & 1), w);
                tc_wait_ld();
                epilogue_split_v2<1>(w, h4, l4);
                tmem_st4(tl + 64 + 4 * q8 + 16 * (it & 1), h4);
                tmem_st4(tl + 96 + 4 * q8 + 16 * (it & 1), l4);
                acc ^= h4[0];
            }
            tc_wait_st();
        }
        const long long t1 = clock64();
        atomicMax(reinterpret_cast<unsigned long long*>(&tmax), (unsigned long long)(t1 - t0));
        tc_fence_before();
        named_bar_sync(2, 512);
        if (tid == 0) *stop = 1;
    } else {
        named_bar_sync(1, 544);
        tc_fence_after();
        const uint64_t b = make_desc(smem_u32(sB));
        const uint64_t a = make_desc(smem_u32(sB + BLK));
        uint32_t par = 0;
        int slot = 0;
        while (duty > 0 && !*stop) {
            const uint32_t d = base + (uint32_t)(slot * 128);
            if (elect_one()) {
#pragma unroll
                // (random data: the first MMA overwrites the accumulator, as in the product -- accumulating for ever saturates
                //  every tanh and the operands degenerate to +-1 / 0 constants, which understates the power)
                for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 96 + 8 * ks, b + 2 * ks, (ks > 0 || !random_data) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 64 + 8 * ks, b + 2 * ks, 1u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts(d, d + 64 + 8 * ks, b + 2 * ks, 1u);
                umma_ss(d, a, b, 1u);
                umma_commit(bar);
            }
            __syncwarp();
            mbar_wait(bar, par);
            par ^= 1u;
            slot = (slot + 1) & 3;
            if (duty == 1) { const long long until = clock64() + 1100; while (clock64() < until) {} }
        }
        tc_fence_before();
    }
    __syncthreads();
    sink[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 0) cycles_out[blockIdx.x] = tmax;
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
    }
}

}  // namespace

static int microbench_launch(int mode, int warps, int iters, int blocks, long long* cycles_out, unsigned int* sink, cudaStream_t st) {
    switch (mode) {
#define MB_CASE(M) case M: microbench_kernel<M><<<blocks, 32 * warps, 0, st>>>(iters, cycles_out, sink); break;
        MB_CASE(0) MB_CASE(1) MB_CASE(2) MB_CASE(3) MB_CASE(4) MB_CASE(5) MB_CASE(6) MB_CASE(7) MB_CASE(8) MB_CASE(9) MB_CASE(10) MB_CASE(11) MB_CASE(12) MB_CASE(13) MB_CASE(14) MB_CASE(15) MB_CASE(16) MB_CASE(17) MB_CASE(18) MB_CASE(19)
#undef MB_CASE
#define MB_CASE(M) case M: microbench_kernel<M><<<blocks, 32 * warps, 0, st>>>(iters, cycles_out, sink); break;
#define PP_CASE(M) case M: pipe_probe_kernel<M><<<blocks, 32 * warps, 0, st>>>(iters, cycles_out, sink); break;
        PP_CASE(20) PP_CASE(21) PP_CASE(22) PP_CASE(23) PP_CASE(24) PP_CASE(25) PP_CASE(26) PP_CASE(27)
        PP_CASE(32) PP_CASE(33) PP_CASE(34) PP_CASE(35) PP_CASE(36) PP_CASE(37) PP_CASE(38) PP_CASE(39) PP_CASE(40) PP_CASE(41) PP_CASE(42) PP_CASE(43) PP_CASE(44) PP_CASE(45) PP_CASE(46) PP_CASE(47) PP_CASE(48) PP_CASE(49) PP_CASE(54) PP_CASE(55) PP_CASE(56)
        MB_CASE(28) MB_CASE(29) MB_CASE(30) MB_CASE(31)
#undef PP_CASE
#undef MB_CASE
        // the same bodies under the register budget of 24 warps (80 registers) and 32 warps (64 registers) per SM
#define MB_CASE(M) case 1000 + M: microbench_kernel<M, 768><<<blocks, 32 * warps, 0, st>>>(iters, cycles_out, sink); break; \
                   case 2000 + M: microbench_kernel<M, 1024><<<blocks, 32 * warps, 0, st>>>(iters, cycles_out, sink); break;
        MB_CASE(11) MB_CASE(15) MB_CASE(28) MB_CASE(30) MB_CASE(31) MB_CASE(50) MB_CASE(51) MB_CASE(52) MB_CASE(53)
        MB_CASE(60) MB_CASE(61) MB_CASE(62) MB_CASE(63) MB_CASE(64) MB_CASE(65) MB_CASE(66) MB_CASE(67) MB_CASE(68)
#undef MB_CASE
#define MB_CASE(M) case M: microbench_kernel<M><<<blocks, 32 * warps, 0, st>>>(iters, cycles_out, sink); break;
        MB_CASE(50) MB_CASE(51) MB_CASE(52) MB_CASE(53) MB_CASE(60) MB_CASE(61) MB_CASE(62) MB_CASE(63) MB_CASE(64) MB_CASE(65) MB_CASE(66) MB_CASE(67) MB_CASE(68)
#undef MB_CASE
        case 200: case 201: case 202: {          // epilogue (16 warps) + MMA stream: duty = mode - 200
            const size_t smem = 3 * BLK + 64;
            epi_mma_kernel<<<blocks, 544, smem, st>>>(mode - 200, iters, cycles_out, sink, 0);
            break;
        }
        case 210: case 211: case 212: {          // the same on pseudo-random weights / activations (power measurements)
            const size_t smem = 3 * BLK + 64;
            epi_mma_kernel<<<blocks, 544, smem, st>>>(mode - 210, iters, cycles_out, sink, 1);
            break;
        }
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


static int selftest_launch(int variant, const void* a_f16, const void* b_f16, float* d_out, cudaStream_t st) {
    const size_t smem = 3 * BLK + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    umma_selftest_kernel<<<1, 160, smem, st>>>(variant, static_cast<const __half*>(a_f16),
                                               static_cast<const __half*>(b_f16), d_out);
    return (int)cudaGetLastError();
}

namespace {
thread_local char g_tools_err[512] = "";
int tools_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_tools_err, sizeof(g_tools_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace

extern "C" {

const char* mems_tools_last_error(void) { return g_tools_err; }

int mems_tools_microbench(int32_t device, int32_t mode, int32_t warps, int32_t iters, int32_t blocks, int64_t* cycles_out_dev,
                          uint32_t* sink_dev, void* stream) {
    if (!cycles_out_dev || !sink_dev) return tools_fail(MEMS_E_INVALID, "null argument");
    if (mode < 0 || warps < 1 || warps > 32 || iters < 1 || blocks < 1) return tools_fail(MEMS_E_INVALID, "bad microbench arguments");
    if (cudaSetDevice(device) != cudaSuccess) return tools_fail(MEMS_E_CUDA, "cudaSetDevice(%d) failed", device);
    const int e = microbench_launch(mode, warps, iters, blocks, reinterpret_cast<long long*>(cycles_out_dev), sink_dev,
                                    reinterpret_cast<cudaStream_t>(stream));
    if (e != 0) return tools_fail(MEMS_E_CUDA, "microbench launch failed (mode %d, %d warps): %s", mode, warps, cudaGetErrorString((cudaError_t)e));
    return MEMS_OK;
}

int mems_tools_selftest_umma(int32_t device, int32_t variant, const void* a_f16_dev, const void* b_f16_dev, float* d_out_dev,
                             void* stream) {
    if (!a_f16_dev || !b_f16_dev || !d_out_dev) return tools_fail(MEMS_E_INVALID, "null argument");
    if (variant < 0 || variant > 2) return tools_fail(MEMS_E_INVALID, "variant must be 0 (A in TMEM), 1 (A in shared memory) or 2 (split chain)");
    if (cudaSetDevice(device) != cudaSuccess) return tools_fail(MEMS_E_CUDA, "cudaSetDevice(%d) failed", device);
    const int e = selftest_launch(variant, a_f16_dev, b_f16_dev, d_out_dev, reinterpret_cast<cudaStream_t>(stream));
    if (e != 0) return tools_fail(MEMS_E_CUDA, "selftest launch failed: %s", cudaGetErrorString((cudaError_t)e));
    return MEMS_OK;
}

}  // extern "C"
