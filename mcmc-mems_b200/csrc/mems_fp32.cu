// FP32 check path: the whole MH update on CUDA cores, FP32 FMAs and tanhf, no tensor cores.
//
// Purpose: pin kernel <-> oracle parity at ~1e-6 relative on surrogate outputs and make
// accept/reject decisions comparable bit for bit before (and beside) the tcgen05 path.
// One thread evaluates one (chain, time point) row of the surrogate
// (reference: src/InverseProblems/utils.py:30-40 builds exactly these rows).
#include "mems_common.cuh"

namespace {

constexpr int NT = 256;   // threads per block

struct Fp32Smem {
    MemsConsts c;
    ChainBatch cb;
    float W1[(MEMS_PMAX + 1) * MEMS_WIDTH];
    float b1[MEMS_WIDTH];
    float bh[MEMS_NLAYER_H * MEMS_WIDTH];
    float w7[MEMS_WIDTH];
    double Spart[NT];
    __align__(16) float Wh[MEMS_NLAYER_H * MEMS_WIDTH * MEMS_WIDTH];
    float act[MEMS_WIDTH * NT];   // per-thread activation column, element n at act[n*NT + tid]
    // followed by: double yobs[T]; float ts[T];
};

__device__ __forceinline__ void load_constants(Fp32Smem& sm, const MemsProblem& pb, double* s_yobs, float* s_ts) {
    const int tid = threadIdx.x;
    if (tid == 0) sm.c = pb.c;
    for (int i = tid; i < (MEMS_PMAX + 1) * MEMS_WIDTH; i += NT) sm.W1[i] = pb.W1[i];
    for (int i = tid; i < MEMS_WIDTH; i += NT) { sm.b1[i] = pb.b1[i]; sm.w7[i] = pb.w7[i]; }
    for (int i = tid; i < MEMS_NLAYER_H * MEMS_WIDTH; i += NT) sm.bh[i] = pb.bh[i];
    const float4* src = reinterpret_cast<const float4*>(pb.Wh);
    float4* dst = reinterpret_cast<float4*>(sm.Wh);
    for (int i = tid; i < MEMS_NLAYER_H * MEMS_WIDTH * MEMS_WIDTH / 4; i += NT) dst[i] = src[i];
    for (int i = tid; i < pb.c.T; i += NT) { s_yobs[i] = pb.yobs[i]; s_ts[i] = pb.ts[i]; }
}

// One surrogate row: scaled inputs (xs[0..P), ts) -> scalar output, everything in FP32.
__device__ __forceinline__ float mlp_row(const Fp32Smem& sm, float* a, int P, const float* xs, float ts, float b7) {
    // layer 1: (P+1) -> 64, matmul first then bias (Keras Dense: matmul + BiasAdd)
#pragma unroll 8
    for (int n = 0; n < MEMS_WIDTH; ++n) {
        float z = xs[0] * sm.W1[n];
        for (int p = 1; p < P; ++p) z = fmaf(xs[p], sm.W1[p * MEMS_WIDTH + n], z);
        z = fmaf(ts, sm.W1[P * MEMS_WIDTH + n], z);
        a[n * NT] = tanhf(z + sm.b1[n]);
    }
    // layers 2..6: 64 -> 64
#pragma unroll 1
    for (int l = 0; l < MEMS_NLAYER_H; ++l) {
        const float4* W = reinterpret_cast<const float4*>(sm.Wh + l * MEMS_WIDTH * MEMS_WIDTH);
        float o[MEMS_WIDTH];
#pragma unroll
        for (int n = 0; n < MEMS_WIDTH; ++n) o[n] = 0.0f;
#pragma unroll 2
        for (int k = 0; k < MEMS_WIDTH; ++k) {
            const float av = a[k * NT];
#pragma unroll
            for (int q = 0; q < MEMS_WIDTH / 4; ++q) {
                const float4 w = W[k * (MEMS_WIDTH / 4) + q];   // warp-uniform address: broadcast
                o[4 * q + 0] = fmaf(av, w.x, o[4 * q + 0]);
                o[4 * q + 1] = fmaf(av, w.y, o[4 * q + 1]);
                o[4 * q + 2] = fmaf(av, w.z, o[4 * q + 2]);
                o[4 * q + 3] = fmaf(av, w.w, o[4 * q + 3]);
            }
        }
#pragma unroll
        for (int n = 0; n < MEMS_WIDTH; ++n) a[n * NT] = tanhf(o[n] + sm.bh[l * MEMS_WIDTH + n]);
    }
    // output layer: 64 -> 1, linear
    float out = 0.0f;
#pragma unroll 8
    for (int k = 0; k < MEMS_WIDTH; ++k) out = fmaf(a[k * NT], sm.w7[k], out);
    return out + b7;
}

// S[j] = sum over the pass's time points t = 0, stride, 2*stride, ... of (y_obs[t] - f(xs_j)[t])^2;
// the result for chain j lands in Spart[j].  y_out (optional) receives the surrogate outputs [chain][T].
__device__ __forceinline__ void eval_batch(Fp32Smem& sm, const double* s_yobs, const float* s_ts, int P, int T,
                                           int stride, float b7, int nc, int ncv, bool skip_dead, float* y_out, int c0) {
    const int tid = threadIdx.x;
    const int j = tid % nc;                    // nc is any value in 1..128; threads >= nts*nc idle
    const int tslot = tid / nc;
    const int nts = NT / nc;
    double Sp = 0.0;
    if (j < ncv && tslot < nts && (!skip_dead || sm.cb.live[j] == 1)) {
        float xs[MEMS_PMAX];
        for (int p = 0; p < P; ++p) xs[p] = sm.cb.xs[p][j];
        float* a = sm.act + tid;
        for (int t = tslot * stride; t < T; t += nts * stride) {
            const float out = mlp_row(sm, a, P, xs, s_ts[t], b7);
            if (y_out) y_out[(size_t)(c0 + j) * T + t] = out;
            const double r = s_yobs[t] - (double)out;
            Sp = fma(r, r, Sp);
        }
    }
    sm.Spart[tid] = Sp;
    __syncthreads();
    if (tid < nc) {
        double S = 0.0;
        for (int k = 0; k < nts; ++k) S += sm.Spart[k * nc + tid];
        sm.Spart[tid] = S;     // only thread tid reads its own slot afterwards
    }
}

__global__ void __launch_bounds__(NT, 1)
fp32_run_kernel(const MemsProblem* __restrict__ pbp, MemsChains ch, MemsRunArgs ra) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Fp32Smem& sm = *reinterpret_cast<Fp32Smem*>(smem_raw);
    const MemsProblem& pb = *pbp;
    double* s_yobs = reinterpret_cast<double*>(smem_raw + sizeof(Fp32Smem));
    float* s_ts = reinterpret_cast<float*>(s_yobs + pb.c.T);
    load_constants(sm, pb, s_yobs, s_ts);
    const int P = pb.c.P, T = pb.c.T, nc = ra.nc;
    const float b7 = pb.b7;
    const int tid = threadIdx.x;
    const int nbatch = (ch.C + nc - 1) / nc;
    for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const int c0 = batch * nc;
        const int ncv = min(nc, ch.C - c0);
        __syncthreads();
        batch_load(sm.cb, ch, P, ra.mode, c0, tid, ncv);
        __syncthreads();
        for (int s = 0; s < ra.n_steps; ++s) {
            for (int sub = 0; sub < ra.n_sub; ++sub) {
                if (tid < ncv) batch_pre(sm.cb, sm.c, ch, ra, c0, tid, s, sub);
                __syncthreads();
                if (batch_pass_is_coarse_model(ra, sub))       // DA first stage on the FFT surrogate: S_c for every chain
                    coarse_fft_pass(ra.coarse, P, T, &sm.cb.xp[0][0], MEMS_NCMAX, nullptr, ncv, sm.Spart, sm.act, tid, NT,
                                    [] { __syncthreads(); });
                else {
                    const int stride = batch_pass_stride(ra, sub);      // (this path evaluates out-of-box proposals too)
                    if (tid == 0) atomicAdd(ch.rows_eval, (unsigned long long)ncv * (unsigned long long)((T + stride - 1) / stride));
                    eval_batch(sm, s_yobs, s_ts, P, T, stride, b7, nc, ncv, true, nullptr, c0);
                }
                if (tid < ncv) batch_post(sm.cb, sm.c, ch, ra, c0, tid, s, sub, sm.Spart[tid]);
                __syncthreads();
            }
        }
        batch_store(sm.cb, ch, P, ra.mode, c0, tid, ncv);
    }
}

__global__ void __launch_bounds__(NT, 1)
fp32_forward_kernel(const MemsProblem* __restrict__ pbp, const double* __restrict__ x, int n, int nc,
                    float* y_out, double* ll_out, int apply_prior, int stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Fp32Smem& sm = *reinterpret_cast<Fp32Smem*>(smem_raw);
    const MemsProblem& pb = *pbp;
    double* s_yobs = reinterpret_cast<double*>(smem_raw + sizeof(Fp32Smem));
    float* s_ts = reinterpret_cast<float*>(s_yobs + pb.c.T);
    load_constants(sm, pb, s_yobs, s_ts);
    const int P = pb.c.P, T = pb.c.T;
    const int tid = threadIdx.x;
    const int nbatch = (n + nc - 1) / nc;
    for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const int c0 = batch * nc;
        const int ncv = min(nc, n - c0);
        __syncthreads();
        if (tid < ncv) {
            bool inb = true;
            for (int p = 0; p < P; ++p) {
                const double v = x[(size_t)p * n + c0 + tid];
                sm.cb.xs[p][tid] = scale_param(sm.c, p, v);
                inb = inb && !(v < sm.c.low[p]) && !(v > sm.c.high[p]);
            }
            sm.cb.inb[tid] = inb ? 1 : 0;
            sm.cb.live[tid] = 1;
        }
        __syncthreads();
        eval_batch(sm, s_yobs, s_ts, P, T, stride, pb.b7, nc, ncv, false, y_out, c0);
        if (tid < ncv && ll_out) {
            // a strided (coarse) pass is rescaled to the full trace length, as in the DA first stage
            const int Tn = (T + stride - 1) / stride;
            const double ll = sm.c.neg_half_inv_sigma2 * sm.Spart[tid] * ((double)T / (double)Tn);
            ll_out[c0 + tid] = (apply_prior && !sm.cb.inb[tid]) ? -CUDART_INF : ll;
        }
    }
}

// Coarse log-likelihood of arbitrary points (the DA start states): ll[i] = -0.5*S_c(x_i)/sigma2, -inf outside the box.
__global__ void __launch_bounds__(NT)
coarse_ll_kernel(const MemsProblem* __restrict__ pbp, const MemsCoarse* __restrict__ cm, const double* __restrict__ x, int n,
                 double* __restrict__ ll_out) {
    __shared__ double xs[MEMS_PMAX][MEMS_NCMAX];
    __shared__ double S[MEMS_NCMAX];
    __shared__ float act[2 * (2 * NT / 64) * MEMS_WIDTH];
    const MemsConsts& c = pbp->c;
    const int tid = threadIdx.x;
    const int nbatch = (n + MEMS_NCMAX - 1) / MEMS_NCMAX;
    for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const int c0 = batch * MEMS_NCMAX, ncv = min(MEMS_NCMAX, n - c0);
        __syncthreads();
        for (int i = tid; i < c.P * ncv; i += NT) xs[i / ncv][i % ncv] = x[(size_t)(i / ncv) * n + c0 + i % ncv];
        __syncthreads();
        coarse_fft_pass(cm, c.P, c.T, &xs[0][0], MEMS_NCMAX, nullptr, ncv, S, act, tid, NT, [] { __syncthreads(); });
        if (tid < ncv) {
            bool inb = true;
            for (int p = 0; p < c.P; ++p) inb = inb && !(xs[p][tid] < c.low[p]) && !(xs[p][tid] > c.high[p]);
            ll_out[c0 + tid] = inb ? c.neg_half_inv_sigma2 * S[tid] : -CUDART_INF;
        }
    }
}

__global__ void draw_kernel(const MemsProblem* __restrict__ pbp, MemsChains ch, int mode, int n_sub,
                            unsigned long long step0, int n_steps, double* xi_out, double* log_u_out) {
    const MemsConsts& pb = pbp->c;
    const size_t total = (size_t)n_steps * ch.C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(i / ch.C), c = (int)(i % ch.C);
        double z[4], lu[4], xi[MEMS_PMAX];
        draw_raw(ch.seed, ch.chain_id0 + (unsigned long long)c, step0 + (unsigned long long)s, z, lu, n_sub);
        if (mode == MEMS_MODE_CW) { for (int p = 0; p < pb.P; ++p) xi[p] = z[p]; }
        else chol_mul(pb, z, xi);
        if (xi_out)
            for (int p = 0; p < pb.P; ++p) xi_out[((size_t)s * pb.P + p) * ch.C + c] = xi[p];
        if (log_u_out)
            for (int k = 0; k < n_sub; ++k) log_u_out[((size_t)s * n_sub + k) * ch.C + c] = lu[k];
    }
}

// Generic Dense stack (NN_Model.predict): one thread per input row, widths <= 64.
constexpr int MLP_NT = 64;
struct MlpDesc {
    int n_layers;
    int widths[10];
    const float* W[9];
    const float* b[9];
};

__global__ void __launch_bounds__(MLP_NT)
mlp_forward_kernel(MlpDesc d, const float* __restrict__ X, long long n, float* __restrict__ Y) {
    __shared__ float bufA[MEMS_WIDTH * MLP_NT];
    __shared__ float bufB[MEMS_WIDTH * MLP_NT];
    const int tid = threadIdx.x;
    for (long long row = blockIdx.x * (long long)MLP_NT + tid; row < n; row += (long long)gridDim.x * MLP_NT) {
        float* in = bufA + tid;
        float* out = bufB + tid;
        for (int k = 0; k < d.widths[0]; ++k) in[k * MLP_NT] = X[row * d.widths[0] + k];
        for (int l = 0; l < d.n_layers; ++l) {
            const int wi = d.widths[l], wo = d.widths[l + 1];
            const float* W = d.W[l];
            const float* b = d.b[l];
            for (int nn = 0; nn < wo; ++nn) {
                float acc = 0.0f;
                for (int k = 0; k < wi; ++k) acc = fmaf(in[k * MLP_NT], __ldg(W + k * wo + nn), acc);
                acc += __ldg(b + nn);
                out[nn * MLP_NT] = (l + 1 < d.n_layers) ? tanhf(acc) : acc;
            }
            float* t = in; in = out; out = t;
        }
        const int wl = d.widths[d.n_layers];
        for (int k = 0; k < wl; ++k) Y[row * wl + k] = in[k * MLP_NT];
    }
}

// Per-chain split-chain moments of the kept samples (the sufficient statistics of split-R-hat):
// out[0] = mean of the first half, out[1] = unbiased variance of the first half, out[2], out[3] = same for
// the last half; each [P][C].  One thread per (parameter, chain), Welford updates, chain-minor reads coalesce.
__global__ void chain_moments_kernel(const double* __restrict__ samples, int n_keep, int P, int C, double* __restrict__ out) {
    const size_t total = (size_t)P * C;
    const int half = n_keep / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        for (int hsel = 0; hsel < 2; ++hsel) {
            const int k0 = hsel == 0 ? 0 : n_keep - half;
            double mean = 0.0, m2 = 0.0;
            for (int k = 0; k < half; ++k) {
                const double v = samples[(size_t)(k0 + k) * total + i];
                const double d = v - mean;
                mean += d / (double)(k + 1);
                m2 = fma(d, v - mean, m2);
            }
            out[(size_t)(2 * hsel) * total + i] = mean;
            out[(size_t)(2 * hsel + 1) * total + i] = half > 1 ? m2 / (double)(half - 1) : 0.0;
        }
    }
}

size_t fp32_smem_bytes(int T) { return sizeof(Fp32Smem) + (size_t)T * (sizeof(double) + sizeof(float)) + 16; }

}  // namespace

int mems_fp32_max_time_points(size_t smem_optin) {
    const size_t fixed = fp32_smem_bytes(0);
    if (smem_optin <= fixed) return 0;
    return (int)((smem_optin - fixed) / (sizeof(double) + sizeof(float)));
}

// Chains per block batch (any value 1..128): minimise (waves of batches) x (time points per thread).
int mems_fp32_choose_nc(int C, int T, int sm_count) {
    int best = 1;
    long long best_cost = -1;
    for (int nc = 1; nc <= MEMS_NCMAX; ++nc) {
        const long long nbatch = (C + nc - 1) / nc;
        const long long waves = (nbatch + sm_count - 1) / sm_count;
        const int nts = NT / nc;
        const long long cost = waves * ((T + nts - 1) / nts);
        if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = nc; }
    }
    return best;
}

int mems_fp32_run(const MemsLaunchCtx& ctx, const MemsRunArgs& ra) {
    const size_t smem = fp32_smem_bytes(ctx.pb_host->c.T);
    cudaError_t e = cudaFuncSetAttribute(fp32_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int nbatch = (ctx.chains.C + ra.nc - 1) / ra.nc;
    const int grid = nbatch < ctx.sm_count ? nbatch : ctx.sm_count;
    fp32_run_kernel<<<grid, NT, smem, ctx.stream>>>(ctx.pb_dev, ctx.chains, ra);
    return (int)cudaGetLastError();
}

int mems_fp32_forward(const MemsLaunchCtx& ctx, const double* x_dev, int n, float* y_out, double* ll_out,
                      int apply_prior, int t_stride) {
    const size_t smem = fp32_smem_bytes(ctx.pb_host->c.T);
    cudaError_t e = cudaFuncSetAttribute(fp32_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    // spread the points over the SMs: smallest power-of-two batch that still fills the grid once
    const int nc = mems_fp32_choose_nc(n, ctx.pb_host->c.T, ctx.sm_count);
    const int nbatch = (n + nc - 1) / nc;
    const int grid = nbatch < 4 * ctx.sm_count ? nbatch : 4 * ctx.sm_count;
    fp32_forward_kernel<<<grid > 0 ? grid : 1, NT, smem, ctx.stream>>>(ctx.pb_dev, x_dev, n, nc, y_out, ll_out,
                                                                   apply_prior, t_stride > 0 ? t_stride : 1);
    return (int)cudaGetLastError();
}

int mems_draw_launch(const MemsLaunchCtx& ctx, int mode, int n_sub, unsigned long long step0, int n_steps,
                     double* xi_out, double* log_u_out) {
    const size_t total = (size_t)n_steps * ctx.chains.C;
    int grid = (int)((total + 255) / 256);
    if (grid > 8 * ctx.sm_count) grid = 8 * ctx.sm_count;
    if (grid < 1) grid = 1;
    draw_kernel<<<grid, 256, 0, ctx.stream>>>(ctx.pb_dev, ctx.chains, mode, n_sub, step0, n_steps, xi_out, log_u_out);
    return (int)cudaGetLastError();
}

int mems_coarse_ll_launch(const MemsLaunchCtx& ctx, const MemsCoarse* cm_dev, const double* x_dev, int n, double* ll_out) {
    int blocks = (n + MEMS_NCMAX - 1) / MEMS_NCMAX;
    if (blocks > 4 * ctx.sm_count) blocks = 4 * ctx.sm_count;
    if (blocks < 1) blocks = 1;
    coarse_ll_kernel<<<blocks, NT, 0, ctx.stream>>>(ctx.pb_dev, cm_dev, x_dev, n, ll_out);
    return (int)cudaGetLastError();
}

int mems_chain_moments_launch(const double* samples, int n_keep, int P, int C, double* out, int sm_count, cudaStream_t st) {
    const size_t total = (size_t)P * C;
    long long blocks = (long long)((total + 255) / 256);
    if (blocks > 16LL * sm_count) blocks = 16LL * sm_count;
    if (blocks < 1) blocks = 1;
    chain_moments_kernel<<<(int)blocks, 256, 0, st>>>(samples, n_keep, P, C, out);
    return (int)cudaGetLastError();
}

int mems_mlp_forward_launch(int n_layers, const int* widths, const float* const* Wd, const float* const* bd,
                            const float* X, long long n, float* Y, int sm_count, cudaStream_t st) {
    MlpDesc d;
    d.n_layers = n_layers;
    for (int i = 0; i <= n_layers; ++i) d.widths[i] = widths[i];
    for (int i = 0; i < n_layers; ++i) { d.W[i] = Wd[i]; d.b[i] = bd[i]; }
    long long blocks = (n + MLP_NT - 1) / MLP_NT;
    if (blocks > 8LL * sm_count) blocks = 8LL * sm_count;
    if (blocks < 1) blocks = 1;
    mlp_forward_kernel<<<(int)blocks, MLP_NT, 0, st>>>(d, X, n, Y);
    return (int)cudaGetLastError();
}
