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

// Coarse first-stage surrogate of delayed acceptance: the reference's FFT model (model_VoltageSignal_fft.keras,
// src/utils/preprocessing.py:90-103: P -> 64x6 tanh -> 15 features = [10*mean, |F_1..7|, 100*angle(F_1..7)] of the trace)
// turned back into a band-limited trace by the reference's own reconstruction
// (tests/Xaccelerometer_quality_factor/surrogate_fft.ipynb cell 15):
//     y_c[t] = m + (2/T) sum_k A_k cos(2 pi k t / T + phi_k),  m = 0.1 f_0,  A_k = f_k,  phi_k = 0.01 f_{7+k}
// and scored against the same observation with the same noise variance.  Because the harmonics are orthogonal on the
// uniform grid, ||y_obs - y_c||^2 needs only 2*7+2 projections of y_obs (computed once on the host).
#define MEMS_COARSE_OUT 15
#define MEMS_COARSE_HARM 7
struct MemsCoarse {
    float W1[MEMS_PMAX * MEMS_WIDTH];                  // [P][64]: parameters only, no time column
    float b1[MEMS_WIDTH];
    __align__(16) float Wh[MEMS_NLAYER_H * MEMS_WIDTH * MEMS_WIDTH];
    float bh[MEMS_NLAYER_H * MEMS_WIDTH];
    float Wo[MEMS_WIDTH * 16];                         // [64][15], rows padded to 16
    float bo[16];
    double mean[MEMS_PMAX], sscale[MEMS_PMAX];         // StandardScaler of the coarse model's inputs
    double sum_y2, sum_y;                              // sum y_obs^2, sum y_obs
    double cy[MEMS_COARSE_HARM], sy[MEMS_COARSE_HARM]; // sum_t y_obs[t] cos / sin(2 pi k t / T), k = 1..7
    // tensor-core path: pre-swizzled fp16 operand blocks of the coarse model in global memory (streamed layer by layer into
    // a shared-memory staging area, mems_tc.cu) and the hidden-layer biases times 2 log2(e), preloaded into the accumulator
    const unsigned char* wimg;
    __align__(16) float tcbias[MEMS_NLAYER_H * MEMS_WIDTH];   // float4 loads
};

// Mutable chain state, SoA over chains.
struct MemsChains {
    int C;
    double* x;            // [P][C]
    double* ll;           // [C] log-likelihood core (-0.5*S/sigma2), -inf outside the prior box
    double* scale;        // [C] cuqi self.scale (random-walk MH / delayed acceptance)
    double* lambda;       // [C] cuqi lambd
    int* win_acc;         // [C] accepted updates in the open adaptation window
    int* adapt_i;         // [C] adaptation counter i
    unsigned int* acc_count;  // [C]
    double* scale_cw;     // [P][C] component-wise MH: per-component proposal std
    double* lambda_cw;    // [P][C]
    int* win_cw;          // [P][C]
    double* llc;          // [C] delayed acceptance: coarse log-likelihood of the current state
    unsigned int* prom_count;  // [C] delayed acceptance: proposals promoted to the fine stage
    unsigned long long* rows_eval;  // [1] surrogate rows (chain, time point) evaluated by run launches since init_chains
    unsigned long long seed;
    unsigned long long chain_id0;
};

#define MEMS_MODE_RW 0     // random-walk MH            (cuqi.sampler.MH; the reference's sampler)
#define MEMS_MODE_CW 1     // component-wise MH         (cuqi.sampler.CWMH semantics; unpinned)
#define MEMS_MODE_DA 2     // delayed acceptance        (Christen & Fox 2005; unpinned)

struct MemsRunArgs {
    int n_steps, adapt_window, thin;
    int mode;                          // MEMS_MODE_*
    int n_sub;                         // surrogate evaluations per update: 1 (RW), P (CW), 2 (DA)
    int da_stride;                     // DA: the coarse stage scores every da_stride-th time point
    const MemsCoarse* coarse;          // DA: non-null -> the coarse stage is the FFT surrogate instead (da_stride unused)
    int coarse_tc;                     // tensor-core path: 1 -> the FFT surrogate runs on the tensor cores too (staging fits)
    double target_acc;                 // adaptation target: 0.234 (RW), 0.21/P + 0.23 (CW)
    unsigned long long step0;          // updates already done (global sample index of the state)
    const double* xi;                  // explicit mode: [n_steps][P][C] (RW/DA: chol*z; CW: z), else nullptr
    const double* log_u;               // explicit mode: [n_steps][n_sub][C]
    double* samples_out;               // [n_keep][P][C] or nullptr
    double* loglik_out;                // [n_keep][C] or nullptr
    double* llprop_out;                // explicit mode: [n_steps][n_sub][C] or nullptr
    unsigned char* decisions_out;      // explicit mode: [n_steps][n_sub][C] or nullptr
    int nc;                            // chains per block batch (1..MEMS_NCMAX)
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

// The randomness of update `step` of global chain `gid`: four standard normals (Philox purpose 0)
// and four log-uniforms (purpose 1).  RW/DA use xi = chol*z and log_u[0] (DA: also log_u[1] for
// the second stage); CW uses z[p] and log_u[p] for component p.
// n_lu: how many of the four log-uniforms the caller consumes (the double-precision log is the expensive part).
__device__ __forceinline__ void draw_raw(unsigned long long seed, unsigned long long gid, unsigned long long step,
                                         double* z, double* log_u, int n_lu = 4) {
    const unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
    const unsigned int c0 = (unsigned int)gid, c1 = (unsigned int)(gid >> 32);
    const unsigned int c2 = (unsigned int)step, c3 = (unsigned int)(step >> 32) & 0x00FFFFFFu;
    const Philox4 n = philox4x32_10(c0, c1, c2, c3, k0, k1);
    const Philox4 u = philox4x32_10(c0, c1, c2, c3 | (1u << 24), k0, k1);
    box_muller(n.x, n.y, z[0], z[1]);
    box_muller(n.z, n.w, z[2], z[3]);
    log_u[0] = log(((double)u.x + 1.0) * 2.3283064365386963e-10);   // u in (0, 1]
    if (n_lu > 1) log_u[1] = log(((double)u.y + 1.0) * 2.3283064365386963e-10);
    if (n_lu > 2) log_u[2] = log(((double)u.z + 1.0) * 2.3283064365386963e-10);
    if (n_lu > 3) log_u[3] = log(((double)u.w + 1.0) * 2.3283064365386963e-10);
}

__device__ __forceinline__ void chol_mul(const MemsConsts& pb, const double* z, double* xi) {
    const int P = pb.P;
    for (int p = 0; p < P; ++p) {
        double acc = 0.0;
        for (int q = 0; q <= p; ++q) acc = fma(pb.chol[p * P + q], z[q], acc);
        xi[p] = acc;
    }
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
    double z[MEMS_PMAX][MEMS_NCMAX];      // CW: the step's normals (x_o = x + scale (.) z is drawn once per step)
    double lu[MEMS_PMAX][MEMS_NCMAX];     // log-uniforms of the step (RW: [0]; DA: [0],[1]; CW: [component])
    double scale_cw[MEMS_PMAX][MEMS_NCMAX];
    double lambda_cw[MEMS_PMAX][MEMS_NCMAX];
    double ll[MEMS_NCMAX];
    double scale[MEMS_NCMAX];
    double lambda[MEMS_NCMAX];
    double llc[MEMS_NCMAX];               // DA: coarse log-likelihood of the current state
    double llc_star[MEMS_NCMAX];          // DA: coarse log-likelihood of the proposal
    float xs[MEMS_PMAX][MEMS_NCMAX];      // scaled proposal, f32 network input
    int win_cw[MEMS_PMAX][MEMS_NCMAX];
    int win_acc[MEMS_NCMAX];
    int adapt_i[MEMS_NCMAX];
    unsigned int acc_count[MEMS_NCMAX];
    unsigned int prom_count[MEMS_NCMAX];
    unsigned char inb[MEMS_NCMAX];        // proposal inside the prior box
    unsigned char live[MEMS_NCMAX];       // 1: rows of this chain must be evaluated in the coming pass (2: see DA stage 1)
};

__device__ __forceinline__ void batch_load(ChainBatch& cb, const MemsChains& ch, int P, int mode, int c0, int j, int ncv) {
    if (j < ncv) {
        const int c = c0 + j;
        for (int p = 0; p < P; ++p) cb.x[p][j] = ch.x[(size_t)p * ch.C + c];
        cb.ll[j] = ch.ll[c];
        cb.scale[j] = ch.scale[c];
        cb.lambda[j] = ch.lambda[c];
        cb.win_acc[j] = ch.win_acc[c];
        cb.adapt_i[j] = ch.adapt_i[c];
        cb.acc_count[j] = ch.acc_count[c];
        if (mode == MEMS_MODE_CW) {
            for (int p = 0; p < P; ++p) {
                cb.scale_cw[p][j] = ch.scale_cw[(size_t)p * ch.C + c];
                cb.lambda_cw[p][j] = ch.lambda_cw[(size_t)p * ch.C + c];
                cb.win_cw[p][j] = ch.win_cw[(size_t)p * ch.C + c];
            }
        }
        if (mode == MEMS_MODE_DA) {
            cb.llc[j] = ch.llc[c];
            cb.prom_count[j] = ch.prom_count[c];
        }
    }
}

__device__ __forceinline__ void batch_store(const ChainBatch& cb, const MemsChains& ch, int P, int mode, int c0, int j, int ncv) {
    if (j < ncv) {
        const int c = c0 + j;
        for (int p = 0; p < P; ++p) ch.x[(size_t)p * ch.C + c] = cb.x[p][j];
        ch.ll[c] = cb.ll[j];
        ch.scale[c] = cb.scale[j];
        ch.lambda[c] = cb.lambda[j];
        ch.win_acc[c] = cb.win_acc[j];
        ch.adapt_i[c] = cb.adapt_i[j];
        ch.acc_count[c] = cb.acc_count[j];
        if (mode == MEMS_MODE_CW) {
            for (int p = 0; p < P; ++p) {
                ch.scale_cw[(size_t)p * ch.C + c] = cb.scale_cw[p][j];
                ch.lambda_cw[(size_t)p * ch.C + c] = cb.lambda_cw[p][j];
                ch.win_cw[(size_t)p * ch.C + c] = cb.win_cw[p][j];
            }
        }
        if (mode == MEMS_MODE_DA) {
            ch.llc[c] = cb.llc[j];
            ch.prom_count[c] = cb.prom_count[j];
        }
    }
}

// Scale a parameter vector the way the reference does before the network call:
// StandardScaler.transform in f64 (utils.py:38), then the Keras cast to f32 (utils.py:39).
__device__ __forceinline__ float scale_param(const MemsConsts& pb, int p, double v) {
    return (float)((v - pb.mean[p]) / pb.sscale[p]);
}

// Publish proposal xp[.][j] as network input + prior-box flag.
__device__ __forceinline__ void batch_publish(ChainBatch& cb, const MemsConsts& pb, int j) {
    bool inb = true;
    for (int p = 0; p < pb.P; ++p) {
        const double v = cb.xp[p][j];
        cb.xs[p][j] = scale_param(pb, p, v);
        inb = inb && !(v < pb.low[p]) && !(v > pb.high[p]);
    }
    cb.inb[j] = inb ? 1 : 0;
    cb.live[j] = inb ? 1 : 0;
}

// Before surrogate pass `sub` of update s: thread j prepares the proposal of chain c0+j.
// Time stride of the coming pass is returned by batch_pass_stride().
__device__ __forceinline__ void batch_pre(ChainBatch& cb, const MemsConsts& pb, const MemsChains& ch,
                                          const MemsRunArgs& ra, int c0, int j, int s, int sub) {
    const int P = pb.P;
    const int c = c0 + j;
    if (sub == 0) {                                    // the step's randomness
        double z[4], lu[4];
        if (ra.xi != nullptr) {
            for (int p = 0; p < P; ++p) z[p] = ra.xi[((size_t)s * P + p) * ch.C + c];
            for (int k = 0; k < ra.n_sub; ++k) lu[k] = ra.log_u[((size_t)s * ra.n_sub + k) * ch.C + c];
        } else {
            draw_raw(ch.seed, ch.chain_id0 + (unsigned long long)c, ra.step0 + (unsigned long long)s, z, lu, ra.n_sub);
            if (ra.mode != MEMS_MODE_CW) {
                double xi[MEMS_PMAX];
                chol_mul(pb, z, xi);
                for (int p = 0; p < P; ++p) z[p] = xi[p];
            }
        }
        for (int p = 0; p < P; ++p) cb.z[p][j] = z[p];
        for (int k = 0; k < ra.n_sub; ++k) cb.lu[k][j] = lu[k];
    }
    if (ra.mode == MEMS_MODE_CW) {
        // cuqi CWMH.single_update: x_o = x_t + scale (.) z is fixed at the start of the step; pass `sub`
        // replaces component `sub` of the CURRENT state (earlier components may already have moved).
        for (int p = 0; p < P; ++p) cb.xp[p][j] = cb.x[p][j];
        if (sub == 0)                                     // remember x_o in z[]: x_o[p] = x[p] + scale[p]*z[p]
            for (int p = 0; p < P; ++p) cb.z[p][j] = __dadd_rn(cb.x[p][j], __dmul_rn(cb.scale_cw[p][j], cb.z[p][j]));
        cb.xp[sub][j] = cb.z[sub][j];
        batch_publish(cb, pb, j);
    } else if (sub == 0) {
        // two roundings like numpy's `x_t + self.scale*xi` (no FMA contraction): keeps states bit-exact
        const double sc = cb.scale[j];
        for (int p = 0; p < P; ++p) cb.xp[p][j] = __dadd_rn(cb.x[p][j], __dmul_rn(sc, cb.z[p][j]));
        batch_publish(cb, pb, j);
    }
    // DA, sub == 1: xs unchanged; live[] was set by the first-stage decision
}

// DA with the FFT coarse surrogate: pass `sub` == 0 does not go through the trace surrogate at all
__device__ __forceinline__ bool batch_pass_is_coarse_model(const MemsRunArgs& ra, int sub) {
    return ra.mode == MEMS_MODE_DA && sub == 0 && ra.coarse != nullptr;
}

// Coarse squared misfit S_c[j] = ||y_obs - y_c(x_j)||^2 of `n` points, evaluated cooperatively by `nthr` threads
// (a multiple of 64; thread t owns neuron t%64 of two points per chunk of 2*nthr/64 points).
//   x[p*xstride + idx]: unscaled parameters;  map: optional list of point indices (nullptr = 0..n-1);
//   S[idx] receives the result;  act: scratch of 2*(2*nthr/64)*64 floats;  bar(): barrier over the nthr threads.
template <class Bar>
__device__ __forceinline__ void coarse_fft_pass(const MemsCoarse* __restrict__ cm, int P, int T, const double* x, int xstride,
                                                const unsigned char* map, int n, double* S, float* act, int t, int nthr, Bar bar) {
    const int nn = t & 63, cg = t >> 6, ngrp = nthr >> 6, CH = 2 * ngrp;
    float* bufA = act;
    float* bufB = act + CH * MEMS_WIDTH;
    for (int base = 0; base < n; base += CH) {
        bar();
        if (t < CH * MEMS_PMAX) {                          // scaled inputs (utils.py:38-39 convention: f64 scale, f32 cast)
            const int c = t / MEMS_PMAX, p = t % MEMS_PMAX;
            float v = 0.0f;
            if (base + c < n && p < P) {
                const int idx = map ? (int)map[base + c] : base + c;
                v = (float)((x[(size_t)p * xstride + idx] - cm->mean[p]) / cm->sscale[p]);
            }
            bufB[c * MEMS_WIDTH + p] = v;
        }
        bar();
#pragma unroll
        for (int u = 0; u < 2; ++u) {                      // layer 1: P -> 64
            const int c = cg + u * ngrp;
            float z = bufB[c * MEMS_WIDTH] * __ldg(&cm->W1[nn]);
            for (int p = 1; p < P; ++p) z = fmaf(bufB[c * MEMS_WIDTH + p], __ldg(&cm->W1[p * MEMS_WIDTH + nn]), z);
            bufA[c * MEMS_WIDTH + nn] = tanhf(z + __ldg(&cm->b1[nn]));
        }
        bar();
        float* in = bufA;
        float* out = bufB;
#pragma unroll 1
        for (int l = 0; l < MEMS_NLAYER_H; ++l) {          // 64 -> 64
            const float* W = cm->Wh + l * MEMS_WIDTH * MEMS_WIDTH;
            const float* a0 = in + cg * MEMS_WIDTH;
            const float* a1 = in + (cg + ngrp) * MEMS_WIDTH;
            float z0 = 0.0f, z1 = 0.0f;
#pragma unroll 8
            for (int k = 0; k < MEMS_WIDTH; ++k) {
                const float w = __ldg(&W[k * MEMS_WIDTH + nn]);
                z0 = fmaf(a0[k], w, z0);
                z1 = fmaf(a1[k], w, z1);
            }
            const float b = __ldg(&cm->bh[l * MEMS_WIDTH + nn]);
            out[cg * MEMS_WIDTH + nn] = tanhf(z0 + b);
            out[(cg + ngrp) * MEMS_WIDTH + nn] = tanhf(z1 + b);
            bar();
            float* tmp = in; in = out; out = tmp;
        }
        if (t < CH * 16) {                                 // 64 -> 15, linear
            const int c = t >> 4, o = t & 15;
            float z = 0.0f;
            if (o < MEMS_COARSE_OUT) {
                for (int k = 0; k < MEMS_WIDTH; ++k) z = fmaf(in[c * MEMS_WIDTH + k], __ldg(&cm->Wo[k * 16 + o]), z);
                z += __ldg(&cm->bo[o]);
            }
            out[c * MEMS_WIDTH + o] = z;
        }
        bar();
        if (t < CH && base + t < n) {                      // band-limited misfit from the 15 features
            const float* f = out + t * MEMS_WIDTH;
            const double m = 0.1 * (double)f[0];
            double dot = m * cm->sum_y, nrm = (double)T * m * m;
            const double w2 = 2.0 / (double)T;
            for (int k = 0; k < MEMS_COARSE_HARM; ++k) {
                const double A = (double)f[1 + k], phi = 0.01 * (double)f[1 + MEMS_COARSE_HARM + k];
                double sn, cs;
                sincos(phi, &sn, &cs);
                dot += w2 * A * (cs * cm->cy[k] - sn * cm->sy[k]);
                nrm += w2 * A * A;
            }
            const int idx = map ? (int)map[base + t] : base + t;
            S[idx] = cm->sum_y2 - 2.0 * dot + nrm;
        }
    }
    bar();
}

__device__ __forceinline__ int batch_pass_stride(const MemsRunArgs& ra, int sub) {
    return (ra.mode == MEMS_MODE_DA && sub == 0) ? ra.da_stride : 1;
}

// cuqi vanishing adaptation (MH._sample_adapt / CWMH._sample_adapt): lambda <- exp(log lambda + (hat - star)/sqrt(i+1))
__device__ __forceinline__ double adapt_lambda(double lambda, int win, int Na, int i, double star) {
    const double hat = (double)win / (double)Na;
    const double zeta = 1.0 / sqrt((double)(i + 1));
    return exp(log(lambda) + zeta * (hat - star));
}

__device__ __forceinline__ void batch_outputs(const ChainBatch& cb, const MemsChains& ch, const MemsRunArgs& ra,
                                              int P, int c, int j, int s) {
    if (((s + 1) % ra.thin) == 0) {
        const size_t row = (size_t)((s + 1) / ra.thin - 1);
        if (ra.samples_out)
            for (int p = 0; p < P; ++p) ra.samples_out[(row * P + p) * ch.C + c] = cb.x[p][j];
        if (ra.loglik_out) ra.loglik_out[row * ch.C + c] = cb.ll[j];
    }
}

// After surrogate pass `sub` of update s: thread j consumes S = sum_t (y_obs - f(x*))^2 (over the pass's time points).
__device__ __forceinline__ void batch_post(ChainBatch& cb, const MemsConsts& pb, const MemsChains& ch,
                                           const MemsRunArgs& ra, int c0, int j, int s, int sub, double S) {
    const int P = pb.P;
    const int c = c0 + j;
    const unsigned long long k = ra.step0 + (unsigned long long)s + 1ull;   // global sample index
    const bool adapt_now = ra.adapt_window > 0 && (k % (unsigned long long)ra.adapt_window) == 0ull;
    const size_t orow = ((size_t)s * ra.n_sub + sub) * ch.C + c;
    if (ra.mode == MEMS_MODE_RW) {
        const double ll_star = cb.inb[j] ? pb.neg_half_inv_sigma2 * S : -CUDART_INF;
        const bool acc = mh_accept(cb.lu[0][j], ll_star, cb.ll[j]);
        if (acc) {
            for (int p = 0; p < P; ++p) cb.x[p][j] = cb.xp[p][j];
            cb.ll[j] = ll_star;
            cb.acc_count[j] += 1u;
        }
        // cuqi MH._sample_adapt: the window closed at (s+1) % Na == 0 holds acc[idx:idx+Na], i.e. it
        // excludes the decision just taken, which opens the next window.
        if (adapt_now) {
            const double lam = adapt_lambda(cb.lambda[j], cb.win_acc[j], ra.adapt_window, cb.adapt_i[j], ra.target_acc);
            cb.lambda[j] = lam;
            cb.scale[j] = fmin(lam, 1.0);
            cb.adapt_i[j] += 1;
            cb.win_acc[j] = 0;
        }
        cb.win_acc[j] += acc ? 1 : 0;
        if (ra.xi != nullptr) {                         // parity mode: everything, every update
            if (ra.llprop_out) ra.llprop_out[orow] = ll_star;
            if (ra.decisions_out) ra.decisions_out[orow] = acc ? 1 : 0;
        }
        batch_outputs(cb, ch, ra, P, c, j, s);
    } else if (ra.mode == MEMS_MODE_CW) {
        const double ll_star = cb.inb[j] ? pb.neg_half_inv_sigma2 * S : -CUDART_INF;
        const bool acc = mh_accept(cb.lu[sub][j], ll_star, cb.ll[j]);
        if (acc) {
            cb.x[sub][j] = cb.xp[sub][j];
            cb.ll[j] = ll_star;
            cb.acc_count[j] += 1u;
        }
        if (ra.xi != nullptr) {
            if (ra.llprop_out) ra.llprop_out[orow] = ll_star;
            if (ra.decisions_out) ra.decisions_out[orow] = acc ? 1 : 0;
        }
        // the window of component `sub` closes before this step's decision is added (same rule as MH)
        if (adapt_now) {
            const double lam = adapt_lambda(cb.lambda_cw[sub][j], cb.win_cw[sub][j], ra.adapt_window, cb.adapt_i[j], ra.target_acc);
            cb.lambda_cw[sub][j] = lam;
            cb.scale_cw[sub][j] = fmin(lam, 1.0);
            cb.win_cw[sub][j] = 0;
            if (sub == P - 1) cb.adapt_i[j] += 1;
        }
        cb.win_cw[sub][j] += acc ? 1 : 0;
        if (sub == P - 1) batch_outputs(cb, ch, ra, P, c, j, s);
    } else {                                            // delayed acceptance
        if (sub == 0) {
            // coarse stage: S covers every da_stride-th time point; rescaled to the full trace length
            const int Tc = ra.coarse ? pb.T : (pb.T + ra.da_stride - 1) / ra.da_stride;
            const double llc_star = cb.inb[j] ? pb.neg_half_inv_sigma2 * S * ((double)pb.T / (double)Tc) : -CUDART_INF;
            const bool prom = mh_accept(cb.lu[0][j], llc_star, cb.llc[j]);
            cb.llc_star[j] = llc_star;
            // 1: promoted and to be evaluated; 2: promoted by the min(0, nan) = 0 quirk although the proposal lies outside
            // the prior box (state and proposal both at -inf) -- stage 2 decides without evaluating the surrogate
            cb.live[j] = prom ? (cb.inb[j] ? 1 : 2) : 0;
            if (prom) cb.prom_count[j] += 1u;
            if (ra.xi != nullptr) {
                if (ra.llprop_out) ra.llprop_out[orow] = llc_star;
                if (ra.decisions_out) ra.decisions_out[orow] = prom ? 1 : 0;
            }
        } else {
            bool acc = false;
            double ll_star = -CUDART_INF;
            if (cb.live[j]) {
                ll_star = cb.inb[j] ? pb.neg_half_inv_sigma2 * S : -CUDART_INF;
                // [pi(x*) pi~(x)] / [pi(x) pi~(x*)]
                acc = mh_accept(cb.lu[1][j], (ll_star - cb.ll[j]) - (cb.llc_star[j] - cb.llc[j]), 0.0);
            }
            if (acc) {
                for (int p = 0; p < P; ++p) cb.x[p][j] = cb.xp[p][j];
                cb.ll[j] = ll_star;
                cb.llc[j] = cb.llc_star[j];
                cb.acc_count[j] += 1u;
            }
            if (adapt_now) {
                const double lam = adapt_lambda(cb.lambda[j], cb.win_acc[j], ra.adapt_window, cb.adapt_i[j], ra.target_acc);
                cb.lambda[j] = lam;
                cb.scale[j] = fmin(lam, 1.0);
                cb.adapt_i[j] += 1;
                cb.win_acc[j] = 0;
            }
            cb.win_acc[j] += acc ? 1 : 0;
            if (ra.xi != nullptr) {
                if (ra.llprop_out) ra.llprop_out[orow] = ll_star;
                if (ra.decisions_out) ra.decisions_out[orow] = acc ? 1 : 0;
            }
            batch_outputs(cb, ch, ra, P, c, j, s);
        }
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
                      int apply_prior, int t_stride);
int mems_fp32_run(const MemsLaunchCtx& ctx, const MemsRunArgs& ra);
int mems_tc_forward(const MemsLaunchCtx& ctx, const double* x_dev, int n, float* y_out, double* ll_out,
                    int apply_prior, int t_stride);
int mems_tc_run(const MemsLaunchCtx& ctx, const MemsRunArgs& ra);
int mems_tc_choose_nc(int C, int T, int sm_count);
int mems_tc_max_time_points(size_t smem_optin);
int mems_fp32_max_time_points(size_t smem_optin);
int mems_fp32_choose_nc(int C, int T, int sm_count);
int mems_coarse_ll_launch(const MemsLaunchCtx& ctx, const MemsCoarse* cm_dev, const double* x_dev, int n, double* ll_out);
int mems_draw_launch(const MemsLaunchCtx& ctx, int mode, int n_sub, unsigned long long step0, int n_steps,
                     double* xi_out, double* log_u_out);
