// Measurement kernels for the tensor-core path (tools/microbench.py): epilogue building blocks, pipe throughput
// probes and the tcgen05.mma issue rate.  Nothing here is on the product path; the numbers they produce are quoted in
// DESIGN.md and stored under profiles/.
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

}  // namespace

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

