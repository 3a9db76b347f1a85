/*
 * mems_b200_tools.h -- development tools of the tensor-core path (tools/libmems_b200_tools.so, built from
 * mcmc-mems_b200/csrc/mems_tools.cu).  NOT part of the product's C ABI (include/mems_b200.h): nothing on the
 * hot path links or loads this library.
 */
#ifndef MEMS_B200_TOOLS_H
#define MEMS_B200_TOOLS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* mems_tools_last_error(void);

/* Hardware self-test of the tensor-core plumbing the MH kernels rely on (UMMA descriptors, 128B swizzle, TMEM
 * operand packing, tcgen05.ld/st, commit->mbarrier):
 * D[128][64] (f32) = A[128][64] (f16) * B[64][64]^T (f16), all device pointers, row-major.
 * variant 0: A operand written to TMEM by the threads (the product path), 1: A from shared memory,
 * 2: the split chain A_hi*W_hi + A_lo*W_hi + A_hi*W_lo with A = [A_hi; A_lo] (256x64), B = [W_hi; W_lo] (128x64). */
int mems_tools_selftest_umma(int32_t device, int32_t variant, const void* a_f16_dev, const void* b_f16_dev,
                             float* d_out_dev, void* stream);

/* Micro-benchmarks of the epilogue building blocks (tools/microbench.py; modes are listed above the kernels in
 * mems_tools.cu).  `blocks` CTAs of `warps` warps run `iters` repetitions; cycles_out_dev [blocks] i64 = slowest thread
 * of each CTA (SM clocks), sink_dev [blocks*32*warps] u32 = scratch.  mode + 1000 / + 2000: the same body compiled
 * under the register budget of 24 / 32 warps per SM. */
int mems_tools_microbench(int32_t device, int32_t mode, int32_t warps, int32_t iters, int32_t blocks,
                          int64_t* cycles_out_dev, uint32_t* sink_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
