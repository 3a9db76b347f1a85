// Shared device-side definitions for the batched surrogate-MH kernels (sm_100a).
//
// What one Metropolis-Hastings update computes, and where the reference does it:
//   propose  x* = x + scale*xi                       cuqi MH.single_update  (identification.ipynb:302)
//   prior    -inf if any(x*<low) or any(x*>high)     cuqi Uniform.logpdf    (identification.ipynb:165)
//   scale    (x*-mean_)/scale_ in f64, cast to f32   src/InverseProblems/utils.py:38-39
//   forward  [P+1]->64x6 tanh->1 per time point      src/SurrogateModeling/model.py:56-59
//   loglik   -0.5*sum((y_obs-f)^2)/sigma2            cuqi Gaussian.logd     (identification.ipynb:169)
//   accept   log u <= min(0, l*-l)                   cuqi MH.single_update
//   adapt    every Na updates, target 0.234          cuqi MH._sample_adapt
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "mems_b200.h"

#define MEMS_NCMAX 128            // chains per block batch (upper bound)
#define MEMS_NLAYER_H (MEMS_HIDDEN - 1)  // 64x64 hidden layers (5)

// ---------------------------------------------------------------------------------------
// Read-only problem description, one copy per handle in global memory.
// ---------------------------------------------------------------------------------------
struct MemsConsts {        // small scalar part: copied to shared memory by the kernels
    int P, T;
    double mean[MEMS_PMAX + 1], sscale[MEMS_PMAX + 1]; // StandardScaler mean_/scale_
    double low[MEMS_PMAX], high[MEMS_PMAX];
    double chol[MEMS_PMAX * MEMS_PMAX];                // lower Cholesky factor of the proposal cov
    double sigma2, neg_half_inv_sigma2;
};

struct MemsProblem {
    MemsConsts c;
    float W1[(MEMS_PMAX + 1) * MEMS_WIDTH];            // [P+1][64], Keras layout [in][out]
    float b1[MEMS_WIDTH];
    __align__(16) float Wh[MEMS_NLAYER_H * MEMS_WIDTH * MEMS_WIDTH]; // [l][k][n]; float4 loads
    float bh[MEMS_NLAYER_H * MEMS_WIDTH];
    float w7[MEMS_WIDTH];
    float b7;
    const float* ts;                                   // [T] scaled time column, f32
    const double* yobs;                                // [T]
    const uint4* wimg;                                 // tensor-core path: pre-swizzled fp16 weight image
};

// Mutable chain state, SoA over chains.
struct MemsChains {
    int C;
    double* x;            // [P][C]
    double* ll;           // [C] log-likelihood core (-0.5*S/sigma2), -inf outside the prior box
    double* scale;        // [C] cuqi self.scale
    double* lambda;       // [C] cuqi lambd
    int* win_acc;         // [C] accepted updates in the open adaptation window
    int* adapt_i;         // [C] adaptation counter i
    unsigned int* acc_count;  // [C]
    unsigned long long seed;
    unsigned long long chain_id0;
};

struct MemsRunArgs {
    int n_steps, adapt_window, thin;
    unsigned long long step0;          // updates already done (global sample index of the state)
    const double* xi;                  // explicit mode: [n_steps][P][C], else nullptr
    const double* log_u;               // explicit mode: [n_steps][C]
    double* samples_out;               // [n_keep][P][C] or nullptr
    double* loglik_out;                // [n_keep][C] or nullptr
    double* llprop_out;                // explicit mode: [n_steps][C] or nullptr
    unsigned char* decisions_out;      // explicit mode: [n_steps][C] or nullptr
    int nc;                            // chains per block batch (power of two <= MEMS_NCMAX)
};

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  counter = (chain_lo, chain_hi, step_lo, step_hi | purpose<<24)
// ---------------------------------------------------------------------------------------
struct Philox4 { unsigned int x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2,
                                                 unsigned int c3, unsigned int k0, unsigned int k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Two standard normals from two 32-bit words (Box-Muller, f32 transcendental, f64 result).
__device__ __forceinline__ void box_muller(unsigned int a, unsigned int b, double& z0, double& z1) {
    const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;   // (0, 1]
    const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    z0 = (double)(r * c);
    z1 = (double)(r * s);
}

// The randomness of update `step` of global chain `gid`: xi = chol*z (P values) and log u.
// purpose 0: normals 0..3; purpose 1: word x = uniform of the accept test (words y,z,w reserved
// for second-stage / per-component uniforms).
__device__ __forceinline__ void draw_update(const MemsConsts& pb, unsigned long long seed,
                                            unsigned long long gid, unsigned long long step,
                                            double* xi, double& log_u) {
    const unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
    const unsigned int c0 = (unsigned int)gid, c1 = (unsigned int)(gid >> 32);
    const unsigned int c2 = (unsigned int)step, c3 = (unsigned int)(step >> 32) & 0x00FFFFFFu;
    const Philox4 n = philox4x32_10(c0, c1, c2, c3, k0, k1);
    const Philox4 u = philox4x32_10(c0, c1, c2, c3 | (1u << 24), k0, k1);
    double z[4];
    box_muller(n.x, n.y, z[0], z[1]);
    box_muller(n.z, n.w, z[2], z[3]);
    const int P = pb.P;
    for (int p = 0; p < P; ++p) {
        double acc = 0.0;
        for (int q = 0; q <= p; ++q) acc = fma(pb.chol[p * P + q], z[q], acc);
        xi[p] = acc;
    }
    log_u = log(((double)u.x + 1.0) * 2.3283064365386963e-10);      // u in (0, 1]
}

// cuqi MH.single_update acceptance rule; min(0, NaN) == 0 in Python is reproduced.
__device__ __forceinline__ bool mh_accept(double log_u, double ll_star, double ll_cur) {
    const double ratio = ll_star - ll_cur;
    const double alpha = (ratio < 0.0) ? ratio : 0.0;
    return log_u <= alpha;
}

// ---------------------------------------------------------------------------------------
// Per-block chain batch kept in shared memory while the block iterates over MH updates.
// ---------------------------------------------------------------------------------------
struct ChainBatch {
    double x[MEMS_PMAX][MEMS_NCMAX];      // current state
    double xp[MEMS_PMAX][MEMS_NCMAX];     // proposal
    double ll[MEMS_NCMAX];
    double scale[MEMS_NCMAX];
    double lambda[MEMS_NCMAX];
    double logu[MEMS_NCMAX];
    float xs[MEMS_PMAX][MEMS_NCMAX];      // scaled proposal, f32 network input
    int win_acc[MEMS_NCMAX];
    int adapt_i[MEMS_NCMAX];
    unsigned int acc_count[MEMS_NCMAX];
    unsigned char inb[MEMS_NCMAX];        // proposal inside the prior box
};

__device__ __forceinline__ void batch_load(ChainBatch& cb, const MemsChains& ch, int P, int c0, int j, int ncv) {
    if (j < ncv) {
        const int c = c0 + j;
        for (int p = 0; p < P; ++p) cb.x[p][j] = ch.x[(size_t)p * ch.C + c];
        cb.ll[j] = ch.ll[c];
        cb.scale[j] = ch.scale[c];
        cb.lambda[j] = ch.lambda[c];
        cb.win_acc[j] = ch.win_acc[c];
        cb.adapt_i[j] = ch.adapt_i[c];
        cb.acc_count[j] = ch.acc_count[c];
    }
}

__device__ __forceinline__ void batch_store(const ChainBatch& cb, const MemsChains& ch, int P, int c0, int j, int ncv) {
    if (j < ncv) {
        const int c = c0 + j;
        for (int p = 0; p < P; ++p) ch.x[(size_t)p * ch.C + c] = cb.x[p][j];
        ch.ll[c] = cb.ll[j];
        ch.scale[c] = cb.scale[j];
        ch.lambda[c] = cb.lambda[j];
        ch.win_acc[c] = cb.win_acc[j];
        ch.adapt_i[c] = cb.adapt_i[j];
        ch.acc_count[c] = cb.acc_count[j];
    }
}

// Scale a parameter vector the way the reference does before the network call:
// StandardScaler.transform in f64 (utils.py:38), then the Keras cast to f32 (utils.py:39).
__device__ __forceinline__ float scale_param(const MemsConsts& pb, int p, double v) {
    return (float)((v - pb.mean[p]) / pb.sscale[p]);
}

// Thread j proposes for chain c0+j at local update index s (global index step0+s).
__device__ __forceinline__ void batch_propose(ChainBatch& cb, const MemsConsts& pb, const MemsChains& ch,
                                              const MemsRunArgs& ra, int c0, int j, int s) {
    const int P = pb.P;
    const int c = c0 + j;
    double xi[MEMS_PMAX];
    double lu;
    if (ra.xi != nullptr) {
        for (int p = 0; p < P; ++p) xi[p] = ra.xi[((size_t)s * P + p) * ch.C + c];
        lu = ra.log_u[(size_t)s * ch.C + c];
    } else {
        draw_update(pb, ch.seed, ch.chain_id0 + (unsigned long long)c, ra.step0 + (unsigned long long)s, xi, lu);
    }
    const double sc = cb.scale[j];
    bool inb = true;
    for (int p = 0; p < P; ++p) {
        // two roundings like numpy's `x_t + self.scale*xi` (no FMA contraction): keeps states bit-exact
        const double v = __dadd_rn(cb.x[p][j], __dmul_rn(sc, xi[p]));
        cb.xp[p][j] = v;
        cb.xs[p][j] = scale_param(pb, p, v);
        inb = inb && !(v < pb.low[p]) && !(v > pb.high[p]);
    }
    cb.inb[j] = inb ? 1 : 0;
    cb.logu[j] = lu;
}

// Thread j finishes update s of chain c0+j given S = sum_t (y_obs - f(x*))^2.
__device__ __forceinline__ void batch_finish(ChainBatch& cb, const MemsConsts& pb, const MemsChains& ch,
                                             const MemsRunArgs& ra, int c0, int j, int s, double S) {
    const int P = pb.P;
    const int c = c0 + j;
    const double ll_star = cb.inb[j] ? pb.neg_half_inv_sigma2 * S : -CUDART_INF;
    const bool acc = mh_accept(cb.logu[j], ll_star, cb.ll[j]);
    if (acc) {
        for (int p = 0; p < P; ++p) cb.x[p][j] = cb.xp[p][j];
        cb.ll[j] = ll_star;
        cb.acc_count[j] += 1u;
    }
    // cuqi MH._sample_adapt: the window closed at (s+1) % Na == 0 holds acc[idx:idx+Na], i.e. it
    // excludes the decision just taken, which opens the next window.
    const unsigned long long k = ra.step0 + (unsigned long long)s + 1ull;   // global sample index
    if (ra.adapt_window > 0 && (k % (unsigned long long)ra.adapt_window) == 0ull) {
        const double hat = (double)cb.win_acc[j] / (double)ra.adapt_window;
        const double zeta = 1.0 / sqrt((double)(cb.adapt_i[j] + 1));
        const double lam = exp(log(cb.lambda[j]) + zeta * (hat - 0.234));
        cb.lambda[j] = lam;
        cb.scale[j] = fmin(lam, 1.0);
        cb.adapt_i[j] += 1;
        cb.win_acc[j] = 0;
    }
    cb.win_acc[j] += acc ? 1 : 0;
    if (ra.xi != nullptr) {                         // parity mode: everything, every update
        if (ra.llprop_out) ra.llprop_out[(size_t)s * ch.C + c] = ll_star;
        if (ra.decisions_out) ra.decisions_out[(size_t)s * ch.C + c] = acc ? 1 : 0;
    }
    if (((s + 1) % ra.thin) == 0) {
        const size_t row = (size_t)((s + 1) / ra.thin - 1);
        if (ra.samples_out)
            for (int p = 0; p < P; ++p) ra.samples_out[(row * P + p) * ch.C + c] = cb.x[p][j];
        if (ra.loglik_out) ra.loglik_out[row * ch.C + c] = cb.ll[j];
    }
}

// Host-side launch helpers implemented per path.
struct MemsLaunchCtx {
    const MemsProblem* pb_dev;   // device copy
    MemsProblem* pb_host;        // host mirror (P, T, ...)
    MemsChains chains;
    int sm_count;
    cudaStream_t stream;
};

int mems_fp32_forward(const MemsLaunchCtx& ctx, const double* x_dev, int n, float* y_out, double* ll_out,
                      int apply_prior);
int mems_fp32_run(const MemsLaunchCtx& ctx, const MemsRunArgs& ra);
int mems_tc_forward(const MemsLaunchCtx& ctx, const double* x_dev, int n, float* y_out, double* ll_out,
                    int apply_prior);
int mems_tc_run(const MemsLaunchCtx& ctx, const MemsRunArgs& ra);
int mems_tc_choose_nc(int C, int T, int sm_count);
int mems_tc_selftest(int variant, const void* a_f16, const void* b_f16, float* d_out, cudaStream_t st);
int mems_draw_launch(const MemsLaunchCtx& ctx, unsigned long long step0, int n_steps, double* xi_out,
                     double* log_u_out);
