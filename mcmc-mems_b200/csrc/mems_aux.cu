// Kernels either side of the Metropolis-Hastings path:
//   * batched bounded least-squares initialisation (Levenberg-Marquardt on 3-point finite differences of the
//     surrogate) -- what src/InverseProblems/utils.py:83-93 does with scipy, one start at a time;
//   * lag-autocovariance sums of the kept samples -- the sufficient statistics of the effective sample size that
//     Samples.compute_ess (arviz underneath; tests/Xaccelerometer_geometric/identification.ipynb:308) reports.
// The surrogate itself is evaluated by the path's own forward kernels; nothing here touches the per-update path.
#include <cfloat>

#include "mems_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// Least squares
// ------------------------------------------------------------------------------------------
constexpr int LSQ_NT = 128;

struct LsqState {          // device arrays, one entry per start
    double* x;             // [P][n] accepted point
    double* c;             // [P][n] trial point (centre of the coming finite-difference stencil)
    double* cost;          // [n]    0.5*||r||^2 at x
    double* lam;           // [n]    LM damping
    double* pred;          // [n]    reduction of the cost the quadratic model predicted for the trial step
    double* JtJ;           // [n][P*P] at x
    double* g;             // [n][P]   J^T r at x
    int* done;             // [n]
    int* iters;            // [n]   accepted steps
};

// finite-difference step rel * max(1, |x|); scipy's '3-point' default is rel = eps^(1/3) = 6.0554544523933395e-06
__device__ __forceinline__ double fd_step(double x, double rel) { return rel * fmax(1.0, fabs(x)); }

// stencil of parameter p around c: offsets (in units of h) of the two extra evaluations
//   central: (+1, -1)   forward one-sided: (+1, +2)   backward one-sided: (-1, -2)
__device__ __forceinline__ void fd_scheme(double c, double h, double lb, double ub, double& o1, double& o2) {
    if (c - h >= lb && c + h <= ub) { o1 = 1.0; o2 = -1.0; }
    else if (c + 2.0 * h <= ub) { o1 = 1.0; o2 = 2.0; }
    else { o1 = -1.0; o2 = -2.0; }
}

// evaluation points of the coming iteration: X[p][i*(2P+1) + k], k = 0 centre, 1+2q / 2+2q the stencil of parameter q
__global__ void lsq_points_kernel(LsqState st, int n, int P, const double* __restrict__ lb, const double* __restrict__ ub,
                                  double rel, double* __restrict__ X) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int K = 2 * P + 1;
    const size_t N = (size_t)n * K;
    for (int p = 0; p < P; ++p) {
        const double cp = st.c[(size_t)p * n + i];
        for (int k = 0; k < K; ++k) X[(size_t)p * N + (size_t)i * K + k] = cp;
    }
    for (int q = 0; q < P; ++q) {
        const double cq = st.c[(size_t)q * n + i];
        const double h = fd_step(cq, rel);
        double o1, o2;
        fd_scheme(cq, h, lb[q], ub[q], o1, o2);
        X[(size_t)q * N + (size_t)i * K + 1 + 2 * q] = cq + o1 * h;
        X[(size_t)q * N + (size_t)i * K + 2 + 2 * q] = cq + o2 * h;
    }
}

// Solve (A + lam*diag(A)) d = -g for P <= 4 by Gaussian elimination with partial pivoting (A symmetric PSD).
__device__ bool lsq_solve(const double* A, const double* g, double lam, int P, double* d) {
    double M[MEMS_PMAX][MEMS_PMAX + 1];
    for (int a = 0; a < P; ++a) {
        for (int b = 0; b < P; ++b) M[a][b] = A[a * P + b];
        M[a][a] += lam * fmax(A[a * P + a], DBL_MIN);
        M[a][P] = -g[a];
    }
    for (int a = 0; a < P; ++a) {
        int piv = a;
        for (int r = a + 1; r < P; ++r)
            if (fabs(M[r][a]) > fabs(M[piv][a])) piv = r;
        if (!(fabs(M[piv][a]) > 0.0)) return false;
        if (piv != a)
            for (int b = 0; b <= P; ++b) { const double t = M[a][b]; M[a][b] = M[piv][b]; M[piv][b] = t; }
        for (int r = a + 1; r < P; ++r) {
            const double f = M[r][a] / M[a][a];
            for (int b = a; b <= P; ++b) M[r][b] -= f * M[a][b];
        }
    }
    for (int a = P - 1; a >= 0; --a) {
        double s = M[a][P];
        for (int b = a + 1; b < P; ++b) s -= M[a][b] * d[b];
        d[a] = s / M[a][a];
    }
    return true;
}

// One block per start: residual / cost at the trial point, accept or reject it, finite-difference Jacobian and
// normal equations at an accepted point, damped step to the next trial point.
__global__ void __launch_bounds__(LSQ_NT)
lsq_update_kernel(LsqState st, int n, int P, int T, const float* __restrict__ Y, const double* __restrict__ yobs, int n_obs,
                  const double* __restrict__ lb, const double* __restrict__ ub, double rel, double xtol, double ftol, int first) {
    const int i = blockIdx.x;
    if (st.done[i]) return;
    const int K = 2 * P + 1;
    const float* y0 = Y + ((size_t)i * K) * T;
    const double* yo = yobs + (n_obs > 1 ? (size_t)i * T : 0);
    __shared__ double red[LSQ_NT];
    __shared__ double acc[1 + MEMS_PMAX + MEMS_PMAX * MEMS_PMAX];   // cost, g[P], JtJ[P*P]
    __shared__ int accept_s;
    const int tid = threadIdx.x;

    auto block_sum = [&](double v) -> double {
        red[tid] = v;
        __syncthreads();
        for (int s = LSQ_NT / 2; s > 0; s >>= 1) {
            if (tid < s) red[tid] += red[tid + s];
            __syncthreads();
        }
        const double out = red[0];
        __syncthreads();
        return out;
    };

    double part = 0.0;
    for (int t = tid; t < T; t += LSQ_NT) {
        const double r = (double)y0[t] - yo[t];
        part = fma(r, r, part);
    }
    const double cost_c = 0.5 * block_sum(part);
    if (tid == 0) accept_s = (first || cost_c < st.cost[i]) ? 1 : 0;
    __syncthreads();
    const bool accept = accept_s != 0;

    if (accept) {
        double hh[MEMS_PMAX], o1[MEMS_PMAX], o2[MEMS_PMAX];
        for (int q = 0; q < P; ++q) {
            const double cq = st.c[(size_t)q * n + i];
            hh[q] = fd_step(cq, rel);
            fd_scheme(cq, hh[q], lb[q], ub[q], o1[q], o2[q]);
        }
        double gp[MEMS_PMAX] = {0, 0, 0, 0}, jj[MEMS_PMAX * MEMS_PMAX] = {0};
        for (int t = tid; t < T; t += LSQ_NT) {
            const double f0 = (double)y0[t];
            const double r = f0 - yo[t];
            double J[MEMS_PMAX];
            for (int q = 0; q < P; ++q) {
                const double f1 = (double)y0[(size_t)(1 + 2 * q) * T + t], f2 = (double)y0[(size_t)(2 + 2 * q) * T + t];
                if (o2[q] == -1.0) J[q] = (f1 - f2) / (2.0 * hh[q]);                       // central
                else J[q] = o1[q] * (-3.0 * f0 + 4.0 * f1 - f2) / (2.0 * hh[q]);           // one-sided, 2nd order
            }
            for (int a = 0; a < P; ++a) {
                gp[a] = fma(J[a], r, gp[a]);
                for (int b = 0; b <= a; ++b) jj[a * P + b] = fma(J[a], J[b], jj[a * P + b]);
            }
        }
        for (int a = 0; a < P; ++a) {
            const double s = block_sum(gp[a]);
            if (tid == 0) acc[1 + a] = s;
            for (int b = 0; b <= a; ++b) {
                const double v = block_sum(jj[a * P + b]);
                if (tid == 0) { acc[1 + MEMS_PMAX + a * P + b] = v; acc[1 + MEMS_PMAX + b * P + a] = v; }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        double lam = st.lam[i];
        bool done = false;
        if (accept) {
            const double prev = st.cost[i];
            const double ratio = first ? 1.0 : (prev - cost_c) / fmax(st.pred[i], DBL_MIN);     // gain ratio of the step
            for (int p = 0; p < P; ++p) st.x[(size_t)p * n + i] = st.c[(size_t)p * n + i];
            for (int a = 0; a < P; ++a) st.g[(size_t)i * P + a] = acc[1 + a];
            for (int a = 0; a < P * P; ++a) st.JtJ[(size_t)i * P * P + a] = acc[1 + MEMS_PMAX + a];
            st.cost[i] = cost_c;
            st.iters[i] += 1;
            if (!first) {
                const double t = 2.0 * ratio - 1.0;                        // Nielsen's damping update
                lam = fmax(lam * fmax(1.0 / 3.0, 1.0 - t * t * t), 1e-15);
                if (prev - cost_c < ftol * cost_c && ratio > 0.25) done = true;   // scipy: ftol satisfied on a good step
            }
        } else {
            lam = lam * 4.0;
            if (lam > 1e15) done = true;
        }
        double d[MEMS_PMAX] = {0, 0, 0, 0};
        double xn2 = 0.0, dn2 = 0.0;
        if (!done) {
            if (!lsq_solve(st.JtJ + (size_t)i * P * P, st.g + (size_t)i * P, lam, P, d)) { done = true; }
            for (int p = 0; p < P; ++p) {
                const double xp = st.x[(size_t)p * n + i];
                double cn = xp + d[p];
                cn = fmin(fmax(cn, lb[p]), ub[p]);
                st.c[(size_t)p * n + i] = cn;
                xn2 += xp * xp;
                dn2 += (cn - xp) * (cn - xp);
            }
            if (sqrt(dn2) <= xtol * (xtol + sqrt(xn2))) done = true;      // scipy xtol criterion
            // predicted reduction of the quadratic model for the (clipped) step s: -(g.s + 0.5 s^T JtJ s)
            double pr = 0.0;
            const double* A = st.JtJ + (size_t)i * P * P;
            const double* gg = st.g + (size_t)i * P;
            for (int a = 0; a < P; ++a) {
                const double sa = st.c[(size_t)a * n + i] - st.x[(size_t)a * n + i];
                double As = 0.0;
                for (int b2 = 0; b2 < P; ++b2) As += A[a * P + b2] * (st.c[(size_t)b2 * n + i] - st.x[(size_t)b2 * n + i]);
                pr -= sa * (gg[a] + 0.5 * As);
            }
            st.pred[i] = pr;
        }
        st.lam[i] = lam;
        st.done[i] = done ? 1 : 0;
    }
}

// x_out [P][n], cov_out [n][P][P] = inv(J^T J) at the accepted point (utils.py:88), cost_out [n], iters_out [n]
__global__ void lsq_finish_kernel(LsqState st, int n, int P, double* __restrict__ x_out, double* __restrict__ cov_out,
                                  double* __restrict__ cost_out, int* __restrict__ iters_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int p = 0; p < P; ++p) x_out[(size_t)p * n + i] = st.x[(size_t)p * n + i];
    if (cost_out) cost_out[i] = st.cost[i];
    if (iters_out) iters_out[i] = st.iters[i];
    if (cov_out) {
        const double* A = st.JtJ + (size_t)i * P * P;
        for (int col = 0; col < P; ++col) {                // column `col` of the inverse: solve A d = e_col
            double e[MEMS_PMAX] = {0, 0, 0, 0}, d[MEMS_PMAX] = {0, 0, 0, 0};
            e[col] = -1.0;                                  // lsq_solve solves ... = -g
            const bool ok = lsq_solve(A, e, 0.0, P, d);
            for (int r = 0; r < P; ++r) cov_out[((size_t)i * P + r) * P + col] = ok ? d[r] : CUDART_NAN;
        }
    }
}

__global__ void lsq_init_kernel(LsqState st, int n, int P, const double* __restrict__ x0, const double* __restrict__ lb,
                                const double* __restrict__ ub) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int p = 0; p < P; ++p) {
        const double v = fmin(fmax(x0[(size_t)p * n + i], lb[p]), ub[p]);
        st.x[(size_t)p * n + i] = v;
        st.c[(size_t)p * n + i] = v;
    }
    st.cost[i] = CUDART_INF;
    st.lam[i] = 1e-3;
    st.pred[i] = 0.0;
    st.done[i] = 0;
    st.iters[i] = 0;
}

// ------------------------------------------------------------------------------------------
// Autocovariance sums
// ------------------------------------------------------------------------------------------
constexpr int ACOV_NT = 256;

// Block (bx, p): CB chains of parameter p, centred series in shared memory; thread t owns lags t, t+256, ...
// partial[(bx*P + p)*(L+2) + lag] = sum over the block's chains of the biased autocovariance (1/n) sum_k x_k x_{k+lag};
// slots L, L+1 = sum of chain means, sum of squared chain means.
__global__ void __launch_bounds__(ACOV_NT)
acov_partial_kernel(const double* __restrict__ samples, int n, int P, int C, int L, int CB, double* __restrict__ partial) {
    extern __shared__ double xs[];                         // [CB][n]
    const int p = blockIdx.y, c0 = blockIdx.x * CB;
    const int ncb = min(CB, C - c0);
    const int tid = threadIdx.x;
    __shared__ double mean_s[32];
    for (int cc = tid; cc < ncb; cc += ACOV_NT) {          // chain means (serial per chain, fixed order)
        double m = 0.0;
        for (int k = 0; k < n; ++k) m += samples[((size_t)k * P + p) * C + c0 + cc];
        mean_s[cc] = m / (double)n;
    }
    __syncthreads();
    for (int idx = tid; idx < ncb * n; idx += ACOV_NT) {
        const int cc = idx % ncb, k = idx / ncb;           // consecutive threads -> consecutive chains (coalesced)
        xs[(size_t)cc * n + k] = samples[((size_t)k * P + p) * C + c0 + cc] - mean_s[cc];
    }
    __syncthreads();
    double* out = partial + ((size_t)blockIdx.x * P + p) * (L + 2);
    for (int lag = tid; lag < L; lag += ACOV_NT) {
        double tot = 0.0;
        for (int cc = 0; cc < ncb; ++cc) {
            const double* x = xs + (size_t)cc * n;
            double a = 0.0;
            for (int k = 0; k + lag < n; ++k) a = fma(x[k], x[k + lag], a);
            tot += a / (double)n;
        }
        out[lag] = tot;
    }
    if (tid == 0) {
        double s1 = 0.0, s2 = 0.0;
        for (int cc = 0; cc < ncb; ++cc) { s1 += mean_s[cc]; s2 += mean_s[cc] * mean_s[cc]; }
        out[L] = s1;
        out[L + 1] = s2;
    }
}

// fixed-order reduction over blocks: out[p][slot].  One block per output; thread t adds partials t, t + 256, ... in order,
// then a fixed tree over the 256 thread sums: deterministic, and 65 536 partials (10^6 chains) no longer serialise in one thread
__global__ void __launch_bounds__(256) acov_reduce_kernel(const double* __restrict__ partial, int nblk, int P, int L, double* __restrict__ out) {
    const int W = L + 2;
    __shared__ double red[256];
    for (int idx = blockIdx.x; idx < P * W; idx += gridDim.x) {
        const int p = idx / W, slot = idx % W;
        double s = 0.0;
        for (int b = threadIdx.x; b < nblk; b += 256) s += partial[((size_t)b * P + p) * W + slot];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[idx] = red[0];
        __syncthreads();
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Host-side drivers (called from mems_api.cu)
// ------------------------------------------------------------------------------------------
size_t mems_lsq_workspace_bytes(int n, int P, int T) {
    const size_t K = 2 * (size_t)P + 1;
    size_t b = 0;
    b += sizeof(double) * P * n * K;                 // X
    b += sizeof(float) * n * K * T;                  // Y
    b += sizeof(double) * ((size_t)2 * P * n + 3 * (size_t)n + (size_t)n * P * P + (size_t)n * P);
    b += sizeof(int) * 2 * (size_t)n;
    return b + 1024;
}

int mems_lsq_run(const MemsLaunchCtx& ctx, bool use_tc, const double* x0_dev, int n, const double* yobs_dev, int n_obs,
                 const double* lb_dev, const double* ub_dev, int max_iter, double rel_step, double xtol, double ftol, void* workspace,
                 double* x_out, double* cov_out, double* cost_out, int* iters_out) {
    const int P = ctx.pb_host->c.P, T = ctx.pb_host->c.T;
    const size_t K = 2 * (size_t)P + 1;
    unsigned char* w = static_cast<unsigned char*>(workspace);
    auto take = [&](size_t bytes) { void* p = w; w += (bytes + 255) & ~(size_t)255; return p; };
    double* X = static_cast<double*>(take(sizeof(double) * P * n * K));
    float* Y = static_cast<float*>(take(sizeof(float) * n * K * T));
    LsqState st;
    st.x = static_cast<double*>(take(sizeof(double) * P * n));
    st.c = static_cast<double*>(take(sizeof(double) * P * n));
    st.cost = static_cast<double*>(take(sizeof(double) * n));
    st.lam = static_cast<double*>(take(sizeof(double) * n));
    st.pred = static_cast<double*>(take(sizeof(double) * n));
    st.JtJ = static_cast<double*>(take(sizeof(double) * n * P * P));
    st.g = static_cast<double*>(take(sizeof(double) * n * P));
    st.done = static_cast<int*>(take(sizeof(int) * n));
    st.iters = static_cast<int*>(take(sizeof(int) * n));
    const int nb = (n + 127) / 128;
    lsq_init_kernel<<<nb, 128, 0, ctx.stream>>>(st, n, P, x0_dev, lb_dev, ub_dev);
    for (int it = 0; it < max_iter; ++it) {
        lsq_points_kernel<<<nb, 128, 0, ctx.stream>>>(st, n, P, lb_dev, ub_dev, rel_step, X);
        const int e = use_tc ? mems_tc_forward(ctx, X, (int)(n * K), Y, nullptr, 0, 1)
                             : mems_fp32_forward(ctx, X, (int)(n * K), Y, nullptr, 0, 1);
        if (e != 0) return e;
        lsq_update_kernel<<<n, LSQ_NT, 0, ctx.stream>>>(st, n, P, T, Y, yobs_dev, n_obs, lb_dev, ub_dev, rel_step, xtol, ftol, it == 0 ? 1 : 0);
    }
    lsq_finish_kernel<<<nb, 128, 0, ctx.stream>>>(st, n, P, x_out, cov_out, cost_out, iters_out);
    return (int)cudaGetLastError();
}

int mems_acov_chains_per_block(int n_keep) {
    const size_t budget = 160 * 1024;
    long long cb = (long long)(budget / (sizeof(double) * (size_t)n_keep));
    if (cb > 16) cb = 16;
    return (int)cb;            // 0 -> the series does not fit
}

size_t mems_acov_workspace_bytes(int n_keep, int P, int C, int L) {
    const int cb = mems_acov_chains_per_block(n_keep);
    if (cb < 1) return 0;
    const size_t nblk = ((size_t)C + cb - 1) / cb;
    return sizeof(double) * nblk * P * ((size_t)L + 2) + 256;
}

int mems_acov_launch(const double* samples, int n_keep, int P, int C, int L, void* workspace, double* out, cudaStream_t st) {
    const int cb = mems_acov_chains_per_block(n_keep);
    if (cb < 1) return (int)cudaErrorInvalidValue;
    const int nblk = (C + cb - 1) / cb;
    const size_t smem = sizeof(double) * (size_t)cb * n_keep;
    cudaError_t e = cudaFuncSetAttribute(acov_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    double* partial = static_cast<double*>(workspace);
    acov_partial_kernel<<<dim3(nblk, P), ACOV_NT, smem, st>>>(samples, n_keep, P, C, L, cb, partial);
    acov_reduce_kernel<<<P * (L + 2), 256, 0, st>>>(partial, nblk, P, L, out);
    return (int)cudaGetLastError();
}
