// C ABI of the batched surrogate-MH path (see include/mems_b200.h for the contract and the
// reference interfaces each entry point replaces).  Host-side only: argument checking, constant
// preparation (scaled time column, split-FP16 weight image) and kernel launches.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "mems_common.cuh"

int mems_mlp_forward_launch(int n_layers, const int* widths, const float* const* Wd, const float* const* bd,
                            const float* X, long long n, float* Y, int sm_count, cudaStream_t st);
int mems_chain_moments_launch(const double* samples, int n_keep, int P, int C, double* out, int sm_count, cudaStream_t st);
size_t mems_lsq_workspace_bytes(int n, int P, int T);
int mems_lsq_run(const MemsLaunchCtx& ctx, bool use_tc, const double* x0_dev, int n, const double* yobs_dev, int n_obs,
                 const double* lb_dev, const double* ub_dev, int max_iter, double rel_step, double xtol, double ftol, void* workspace,
                 double* x_out, double* cov_out, double* cost_out, int* iters_out);
size_t mems_acov_workspace_bytes(int n_keep, int P, int C, int L);
int mems_acov_launch(const double* samples, int n_keep, int P, int C, int L, void* workspace, double* out, cudaStream_t st);
size_t mems_sort_padded(size_t n);
int mems_sort_launch(double* keys, size_t n, int sm_count, cudaStream_t st);
int mems_rank_normalize_launch(const double* sorted, size_t M, const double* x, size_t n, double* z, int sm_count, cudaStream_t st);
int mems_geyer_launch(const double* acov, int P, int L, long long n_draws, double n_chains, double* rho_scratch, double* out,
                      int* trunc, cudaStream_t st);
size_t mems_tc_wimg_bytes();
size_t mems_tc_coarse_wimg_bytes();
bool mems_tc_coarse_fits(int T, size_t smem_optin);
void mems_tc_build_coarse_wimg(MemsCoarse& cm, int P, void* host_img);
void mems_tc_build_wimg(const MemsProblem& pb, void* host_img);

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(MEMS_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));           \
    } while (0)

// Every entry point runs on the handle's device and puts the caller's current device back on exit (PyTorch reads the
// current device from the runtime: an engine on cuda:1 must not change what `device="cuda"` means for the caller).
struct DeviceScope {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceScope(int dev) {
        err = cudaGetDevice(&prev);
        if (err != cudaSuccess) { prev = -1; return; }
        if (prev != dev) err = cudaSetDevice(dev);
        else prev = -1;                                // nothing to restore
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};
#define DEVICE_SCOPE(dev)                                                                            \
    DeviceScope device_scope_(dev);                                                                  \
    if (device_scope_.err != cudaSuccess)                                                            \
        return fail(MEMS_E_CUDA, "cudaSetDevice(%d) failed: %s", (int)(dev), cudaGetErrorString(device_scope_.err))

}  // namespace

struct MemsHandle_ {
    MemsConfig cfg;
    int sm_count = 0;
    MemsProblem host_pb;
    MemsProblem* dev_pb = nullptr;
    float* dev_ts = nullptr;
    double* dev_yobs = nullptr;
    void* dev_wimg = nullptr;
    MemsCoarse* dev_coarse = nullptr;      // FFT coarse surrogate of delayed acceptance (device copy)
    MemsCoarse* host_coarse = nullptr;
    void* dev_coarse_wimg = nullptr;       // tensor-core path: operand image of the coarse surrogate (global memory, streamed)
    size_t smem_optin = 0;
    bool have_coarse = false, use_coarse = false;
    std::vector<double> host_yobs;
    MemsChains chains;
    bool have_weights = false, have_problem = false, have_chains = false;
    int sampler = MEMS_MODE_RW;
    int da_stride = 4;
    double cw_scale[MEMS_PMAX] = {1.0, 1.0, 1.0, 1.0};
    unsigned long long steps_done = 0;
    long long launches = 0;
};

namespace {

int choose_nc(const MemsHandle_* h, int C) {
    if (h->cfg.chains_per_block > 0) return h->cfg.chains_per_block;
    return (h->cfg.path == MEMS_PATH_TC) ? mems_tc_choose_nc(C, h->host_pb.c.T, h->sm_count)
                                         : mems_fp32_choose_nc(C, h->host_pb.c.T, h->sm_count);
}

int upload_problem(MemsHandle_* h) {
    h->host_pb.ts = h->dev_ts;
    h->host_pb.yobs = h->dev_yobs;
    h->host_pb.wimg = reinterpret_cast<const uint4*>(h->dev_wimg);
    CUDA_TRY(cudaMemcpy(h->dev_pb, &h->host_pb, sizeof(MemsProblem), cudaMemcpyHostToDevice));
    return MEMS_OK;
}

// projections of y_obs on the first harmonics of the uniform grid (coarse_fft_pass), then the device copy
int upload_coarse(MemsHandle_* h) {
    if (!h->have_coarse || h->host_yobs.empty()) return MEMS_OK;
    MemsCoarse& cm = *h->host_coarse;
    const int T = h->host_pb.c.T;
    const double two_pi = 6.283185307179586476925286766559;
    cm.sum_y2 = cm.sum_y = 0.0;
    for (int t = 0; t < T; ++t) { cm.sum_y2 += h->host_yobs[t] * h->host_yobs[t]; cm.sum_y += h->host_yobs[t]; }
    for (int k = 1; k <= MEMS_COARSE_HARM; ++k) {
        double c = 0.0, sn = 0.0;
        for (int t = 0; t < T; ++t) {
            const double a = two_pi * (double)k * (double)t / (double)T;
            c += h->host_yobs[t] * cos(a);
            sn += h->host_yobs[t] * sin(a);
        }
        cm.cy[k - 1] = c;
        cm.sy[k - 1] = sn;
    }
    CUDA_TRY(cudaMemcpy(h->dev_coarse, h->host_coarse, sizeof(MemsCoarse), cudaMemcpyHostToDevice));
    return MEMS_OK;
}

MemsLaunchCtx make_ctx(MemsHandle_* h, void* stream) {
    MemsLaunchCtx ctx;
    ctx.pb_dev = h->dev_pb;
    ctx.pb_host = &h->host_pb;
    ctx.chains = h->chains;
    ctx.sm_count = h->sm_count;
    ctx.stream = reinterpret_cast<cudaStream_t>(stream);
    return ctx;
}

}  // namespace

extern "C" {

const char* mems_last_error(void) { return g_err; }
int mems_version(void) { return 100; }

int mems_create(const MemsConfig* cfg, mems_handle_t* out) {
    if (!cfg || !out) return fail(MEMS_E_INVALID, "null argument");
    *out = nullptr;
    if (cfg->n_params < 1 || cfg->n_params > MEMS_PMAX)
        return fail(MEMS_E_INVALID, "n_params=%d outside [1,%d]", cfg->n_params, MEMS_PMAX);
    if (cfg->n_time < 1 || cfg->n_time > 4096) return fail(MEMS_E_INVALID, "n_time=%d outside [1,4096]", cfg->n_time);
    if (cfg->n_chains < 1) return fail(MEMS_E_INVALID, "n_chains=%d must be >= 1", cfg->n_chains);
    if (cfg->path != MEMS_PATH_FP32 && cfg->path != MEMS_PATH_TC) return fail(MEMS_E_INVALID, "unknown path %d", cfg->path);
    if (cfg->chains_per_block != 0) {
        const int v = cfg->chains_per_block;
        if (v < 1 || v > MEMS_NCMAX) return fail(MEMS_E_INVALID, "chains_per_block must be in [1,%d]", MEMS_NCMAX);
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(MEMS_E_UNSUPPORTED, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(MEMS_E_INVALID, "device %d out of range", cfg->device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(MEMS_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
    {   // the observation and the time column live in shared memory next to the resident weights
        const int tmax = (cfg->path == MEMS_PATH_TC) ? mems_tc_max_time_points(prop.sharedMemPerBlockOptin)
                                                     : mems_fp32_max_time_points(prop.sharedMemPerBlockOptin);
        if (cfg->n_time > tmax)
            return fail(MEMS_E_INVALID, "n_time=%d does not fit in shared memory on this path (max %d)", cfg->n_time, tmax);
    }
    DEVICE_SCOPE(cfg->device);
    MemsHandle_* h = new (std::nothrow) MemsHandle_();
    if (!h) return fail(MEMS_E_INVALID, "out of host memory");
    h->cfg = *cfg;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    memset(&h->host_pb, 0, sizeof(MemsProblem));
    h->host_pb.c.P = cfg->n_params;
    h->host_pb.c.T = cfg->n_time;
    const size_t C = (size_t)cfg->n_chains, P = (size_t)cfg->n_params;
    memset(&h->chains, 0, sizeof(MemsChains));
    h->chains.C = cfg->n_chains;
    bool ok = cudaMalloc(&h->dev_pb, sizeof(MemsProblem)) == cudaSuccess &&
              cudaMalloc(&h->dev_ts, sizeof(float) * cfg->n_time) == cudaSuccess &&
              cudaMalloc(&h->dev_yobs, sizeof(double) * cfg->n_time) == cudaSuccess &&
              cudaMalloc(&h->dev_wimg, mems_tc_wimg_bytes()) == cudaSuccess &&
              cudaMalloc(&h->chains.x, sizeof(double) * P * C) == cudaSuccess &&
              cudaMalloc(&h->chains.ll, sizeof(double) * C) == cudaSuccess &&
              cudaMalloc(&h->chains.scale, sizeof(double) * C) == cudaSuccess &&
              cudaMalloc(&h->chains.lambda, sizeof(double) * C) == cudaSuccess &&
              cudaMalloc(&h->chains.win_acc, sizeof(int) * C) == cudaSuccess &&
              cudaMalloc(&h->chains.adapt_i, sizeof(int) * C) == cudaSuccess &&
              cudaMalloc(&h->chains.acc_count, sizeof(unsigned int) * C) == cudaSuccess &&
              cudaMalloc(&h->chains.scale_cw, sizeof(double) * P * C) == cudaSuccess &&
              cudaMalloc(&h->chains.lambda_cw, sizeof(double) * P * C) == cudaSuccess &&
              cudaMalloc(&h->chains.win_cw, sizeof(int) * P * C) == cudaSuccess &&
              cudaMalloc(&h->chains.llc, sizeof(double) * C) == cudaSuccess &&
              cudaMalloc(&h->chains.prom_count, sizeof(unsigned int) * C) == cudaSuccess &&
              cudaMalloc(&h->chains.rows_eval, sizeof(unsigned long long)) == cudaSuccess;
    if (!ok) {
        const char* msg = cudaGetErrorString(cudaGetLastError());
        mems_destroy(h);
        return fail(MEMS_E_CUDA, "device allocation failed: %s", msg);
    }
    *out = h;
    return MEMS_OK;
}

void mems_destroy(mems_handle_t h) {
    if (!h) return;
    DeviceScope device_scope_(h->cfg.device);
    cudaFree(h->dev_pb); cudaFree(h->dev_ts); cudaFree(h->dev_yobs); cudaFree(h->dev_wimg);
    cudaFree(h->chains.x); cudaFree(h->chains.ll); cudaFree(h->chains.scale); cudaFree(h->chains.lambda);
    cudaFree(h->chains.win_acc); cudaFree(h->chains.adapt_i); cudaFree(h->chains.acc_count);
    cudaFree(h->chains.scale_cw); cudaFree(h->chains.lambda_cw); cudaFree(h->chains.win_cw);
    cudaFree(h->chains.llc); cudaFree(h->chains.prom_count); cudaFree(h->chains.rows_eval);
    cudaFree(h->dev_coarse);
    cudaFree(h->dev_coarse_wimg);
    delete h->host_coarse;
    delete h;
}

int mems_set_weights(mems_handle_t h, const float* const* W, const float* const* b) {
    if (!h || !W || !b) return fail(MEMS_E_INVALID, "null argument");
    for (int l = 0; l <= MEMS_HIDDEN; ++l)
        if (!W[l] || !b[l]) return fail(MEMS_E_INVALID, "null weight pointer for layer %d", l);
    DEVICE_SCOPE(h->cfg.device);
    MemsProblem& pb = h->host_pb;
    const int P = pb.c.P;
    memcpy(pb.W1, W[0], sizeof(float) * (P + 1) * MEMS_WIDTH);
    memcpy(pb.b1, b[0], sizeof(float) * MEMS_WIDTH);
    for (int l = 0; l < MEMS_NLAYER_H; ++l) {
        memcpy(pb.Wh + l * MEMS_WIDTH * MEMS_WIDTH, W[l + 1], sizeof(float) * MEMS_WIDTH * MEMS_WIDTH);
        memcpy(pb.bh + l * MEMS_WIDTH, b[l + 1], sizeof(float) * MEMS_WIDTH);
    }
    memcpy(pb.w7, W[MEMS_HIDDEN], sizeof(float) * MEMS_WIDTH);
    pb.b7 = b[MEMS_HIDDEN][0];
    std::vector<unsigned char> img(mems_tc_wimg_bytes());
    mems_tc_build_wimg(pb, img.data());
    CUDA_TRY(cudaMemcpy(h->dev_wimg, img.data(), img.size(), cudaMemcpyHostToDevice));
    h->have_weights = true;
    return upload_problem(h);
}

int mems_set_problem(mems_handle_t h, const double* time, const double* scaler_mean, const double* scaler_scale,
                     const double* y_obs, double sigma2, const double* low, const double* high, const double* chol) {
    if (!h || !time || !scaler_mean || !scaler_scale || !y_obs || !low || !high || !chol)
        return fail(MEMS_E_INVALID, "null argument");
    if (!(sigma2 > 0.0)) return fail(MEMS_E_INVALID, "sigma2 must be positive");
    DEVICE_SCOPE(h->cfg.device);
    MemsProblem& pb = h->host_pb;
    const int P = pb.c.P, T = pb.c.T;
    for (int p = 0; p <= P; ++p) {
        if (!(scaler_scale[p] != 0.0)) return fail(MEMS_E_INVALID, "scaler_scale[%d] is zero", p);
        pb.c.mean[p] = scaler_mean[p];
        pb.c.sscale[p] = scaler_scale[p];
    }
    for (int p = 0; p < P; ++p) {
        if (!(low[p] <= high[p])) return fail(MEMS_E_INVALID, "low[%d] > high[%d]", p, p);
        pb.c.low[p] = low[p];
        pb.c.high[p] = high[p];
    }
    for (int i = 0; i < P * P; ++i) pb.c.chol[i] = chol[i];
    pb.c.sigma2 = sigma2;
    pb.c.neg_half_inv_sigma2 = -0.5 / sigma2;
    std::vector<float> ts(T);
    // StandardScaler.transform on the time column in f64, then the Keras cast to f32 (utils.py:38-39)
    for (int t = 0; t < T; ++t) ts[t] = (float)((time[t] - scaler_mean[P]) / scaler_scale[P]);
    CUDA_TRY(cudaMemcpy(h->dev_ts, ts.data(), sizeof(float) * T, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->dev_yobs, y_obs, sizeof(double) * T, cudaMemcpyHostToDevice));
    h->host_yobs.assign(y_obs, y_obs + T);
    h->have_problem = true;
    h->have_chains = false;      // chains.ll / chains.llc were scored against the previous y_obs / sigma2 / bounds
    const int rc = upload_coarse(h);                       // the coarse stage scores the same observation
    if (rc != MEMS_OK) return rc;
    return upload_problem(h);
}

int mems_forward(mems_handle_t h, const double* x_dev, int32_t n, float* y_out, double* loglik_out, void* stream) {
    if (!h) return fail(MEMS_E_INVALID, "null handle");
    if (!h->have_weights || !h->have_problem) return fail(MEMS_E_STATE, "set weights and problem first");
    if (n < 0) return fail(MEMS_E_INVALID, "n < 0");
    if (n == 0) return MEMS_OK;
    if (!x_dev) return fail(MEMS_E_INVALID, "null argument");
    DEVICE_SCOPE(h->cfg.device);
    MemsLaunchCtx ctx = make_ctx(h, stream);
    const int e = (h->cfg.path == MEMS_PATH_TC) ? mems_tc_forward(ctx, x_dev, n, y_out, loglik_out, 0, 1)
                                                : mems_fp32_forward(ctx, x_dev, n, y_out, loglik_out, 0, 1);
    h->launches += 1;
    if (e != 0) return fail(MEMS_E_CUDA, "forward launch failed: %s", cudaGetErrorString((cudaError_t)e));
    return MEMS_OK;
}

namespace {
struct CwInit { double s[MEMS_PMAX]; };
__global__ void init_chains_kernel(MemsChains ch, int P, const double* x0, double scale0, CwInit cw) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ch.C) return;
    for (int p = 0; p < P; ++p) {
        ch.x[(size_t)p * ch.C + c] = x0[(size_t)p * ch.C + c];
        ch.scale_cw[(size_t)p * ch.C + c] = cw.s[p];
        ch.lambda_cw[(size_t)p * ch.C + c] = cw.s[p];
        ch.win_cw[(size_t)p * ch.C + c] = 1;   // cuqi CWMH: acc[:, 0] = 1
    }
    ch.scale[c] = scale0;
    ch.lambda[c] = scale0;
    ch.win_acc[c] = 1;          // cuqi: acc[0] = 1 belongs to the first adaptation window
    ch.adapt_i[c] = 0;
    ch.acc_count[c] = 0u;
    ch.prom_count[c] = 0u;
}
}  // namespace

int mems_set_sampler(mems_handle_t h, int32_t sampler, const double* cw_scale, int32_t da_stride) {
    if (!h) return fail(MEMS_E_INVALID, "null handle");
    if (sampler != MEMS_MODE_RW && sampler != MEMS_MODE_CW && sampler != MEMS_MODE_DA)
        return fail(MEMS_E_INVALID, "unknown sampler %d", sampler);
    if (sampler == MEMS_MODE_CW) {
        if (!cw_scale) return fail(MEMS_E_INVALID, "component-wise MH needs cw_scale[P]");
        for (int p = 0; p < h->host_pb.c.P; ++p) {
            if (!(cw_scale[p] >= 0.0)) return fail(MEMS_E_INVALID, "cw_scale[%d] must be >= 0", p);
            h->cw_scale[p] = cw_scale[p];
        }
    }
    if (sampler == MEMS_MODE_DA) {
        if (da_stride == 0) {                              // coarse stage = the FFT surrogate
            if (!h->have_coarse) return fail(MEMS_E_STATE, "da_stride = 0 selects the FFT coarse surrogate: call mems_set_coarse_fft first");
            h->use_coarse = true;
            h->da_stride = 1;
        } else {
            if (da_stride < 1 || da_stride > h->host_pb.c.T) return fail(MEMS_E_INVALID, "da_stride=%d outside [0,T]", da_stride);
            h->use_coarse = false;
            h->da_stride = da_stride;
        }
    }
    h->sampler = sampler;
    h->have_chains = false;     // chains must be (re)initialised for the new sampler
    return MEMS_OK;
}

int mems_set_coarse_fft(mems_handle_t h, const float* const* W, const float* const* b, const double* scaler_mean,
                        const double* scaler_scale) {
    if (!h || !W || !b || !scaler_mean || !scaler_scale) return fail(MEMS_E_INVALID, "null argument");
    for (int i = 0; i < MEMS_HIDDEN + 1; ++i)
        if (!W[i] || !b[i]) return fail(MEMS_E_INVALID, "null weight pointer for layer %d", i);
    DEVICE_SCOPE(h->cfg.device);
    const int P = h->host_pb.c.P;
    if (!h->host_coarse) {
        h->host_coarse = new (std::nothrow) MemsCoarse();
        if (!h->host_coarse) return fail(MEMS_E_CUDA, "out of host memory");
    }
    if (!h->dev_coarse) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->dev_coarse), sizeof(MemsCoarse)));
    MemsCoarse& cm = *h->host_coarse;
    memset(&cm, 0, sizeof(cm));
    memcpy(cm.W1, W[0], sizeof(float) * P * MEMS_WIDTH);                 // Keras layout [in][out], in = P (no time column)
    memcpy(cm.b1, b[0], sizeof(float) * MEMS_WIDTH);
    for (int l = 0; l < MEMS_NLAYER_H; ++l) {
        memcpy(cm.Wh + l * MEMS_WIDTH * MEMS_WIDTH, W[1 + l], sizeof(float) * MEMS_WIDTH * MEMS_WIDTH);
        memcpy(cm.bh + l * MEMS_WIDTH, b[1 + l], sizeof(float) * MEMS_WIDTH);
    }
    for (int k = 0; k < MEMS_WIDTH; ++k)
        for (int o = 0; o < MEMS_COARSE_OUT; ++o) cm.Wo[k * 16 + o] = W[MEMS_HIDDEN][k * MEMS_COARSE_OUT + o];
    for (int o = 0; o < MEMS_COARSE_OUT; ++o) cm.bo[o] = b[MEMS_HIDDEN][o];
    for (int p = 0; p < P; ++p) {
        if (!(scaler_scale[p] != 0.0)) return fail(MEMS_E_INVALID, "scaler_scale[%d] is zero", p);
        cm.mean[p] = scaler_mean[p];
        cm.sscale[p] = scaler_scale[p];
    }
    if (h->host_pb.c.T < 2 * MEMS_COARSE_HARM + 1)
        return fail(MEMS_E_INVALID, "the FFT coarse surrogate needs n_time >= %d", 2 * MEMS_COARSE_HARM + 1);
    if (h->cfg.path == MEMS_PATH_TC) {                    // operand image for the tensor-core coarse pass
        std::vector<unsigned char> img(mems_tc_coarse_wimg_bytes());
        mems_tc_build_coarse_wimg(cm, P, img.data());
        if (!h->dev_coarse_wimg) CUDA_TRY(cudaMalloc(&h->dev_coarse_wimg, img.size()));
        CUDA_TRY(cudaMemcpy(h->dev_coarse_wimg, img.data(), img.size(), cudaMemcpyHostToDevice));
        cm.wimg = static_cast<const unsigned char*>(h->dev_coarse_wimg);
    }
    h->have_coarse = true;
    h->have_chains = false;
    if (h->host_yobs.empty()) CUDA_TRY(cudaMemcpy(h->dev_coarse, h->host_coarse, sizeof(MemsCoarse), cudaMemcpyHostToDevice));
    return upload_coarse(h);
}

int mems_init_chains(mems_handle_t h, const double* x0_dev, double scale0, uint64_t seed, uint64_t chain_id0,
                     void* stream) {
    if (!h || !x0_dev) return fail(MEMS_E_INVALID, "null argument");
    if (!h->have_weights || !h->have_problem) return fail(MEMS_E_STATE, "set weights and problem first");
    if (!(scale0 > 0.0)) return fail(MEMS_E_INVALID, "scale0 must be positive");
    DEVICE_SCOPE(h->cfg.device);
    h->chains.seed = seed;
    h->chains.chain_id0 = chain_id0;
    h->steps_done = 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int C = h->chains.C;
    CwInit cw;
    for (int p = 0; p < MEMS_PMAX; ++p) cw.s[p] = h->cw_scale[p];
    CUDA_TRY(cudaMemsetAsync(h->chains.rows_eval, 0, sizeof(unsigned long long), st));
    init_chains_kernel<<<(C + 255) / 256, 256, 0, st>>>(h->chains, h->host_pb.c.P, x0_dev, scale0, cw);
    CUDA_TRY(cudaGetLastError());
    MemsLaunchCtx ctx = make_ctx(h, stream);
    const bool tc = h->cfg.path == MEMS_PATH_TC;
    // target_eval[0] = logd(x0): likelihood + prior (-inf outside the box)
    int e = tc ? mems_tc_forward(ctx, h->chains.x, C, nullptr, h->chains.ll, 1, 1)
               : mems_fp32_forward(ctx, h->chains.x, C, nullptr, h->chains.ll, 1, 1);
    h->launches += 2;
    if (e == 0 && h->sampler == MEMS_MODE_DA) {   // coarse log-likelihood of the start state
        if (h->use_coarse) e = mems_coarse_ll_launch(ctx, h->dev_coarse, h->chains.x, C, h->chains.llc);
        else e = tc ? mems_tc_forward(ctx, h->chains.x, C, nullptr, h->chains.llc, 1, h->da_stride)
                    : mems_fp32_forward(ctx, h->chains.x, C, nullptr, h->chains.llc, 1, h->da_stride);
        h->launches += 1;
    }
    if (e != 0) return fail(MEMS_E_CUDA, "init launch failed: %s", cudaGetErrorString((cudaError_t)e));
    h->have_chains = true;
    return MEMS_OK;
}

static int run_common(mems_handle_t h, MemsRunArgs& ra, void* stream) {
    DEVICE_SCOPE(h->cfg.device);
    ra.step0 = h->steps_done;
    ra.nc = choose_nc(h, h->chains.C);
    ra.mode = h->sampler;
    ra.n_sub = (h->sampler == MEMS_MODE_CW) ? h->host_pb.c.P : (h->sampler == MEMS_MODE_DA ? 2 : 1);
    ra.da_stride = h->da_stride;
    ra.coarse = (h->sampler == MEMS_MODE_DA && h->use_coarse) ? h->dev_coarse : nullptr;
    // the coarse stage runs on the tensor cores when its staging blocks fit next to the resident fine weights (MEMS_COARSE_FP32=1
    // in the environment keeps it on the CUDA cores: A/B and fallback)
    ra.coarse_tc = (ra.coarse && h->cfg.path == MEMS_PATH_TC && h->dev_coarse_wimg && mems_tc_coarse_fits(h->host_pb.c.T, h->smem_optin) &&
                    !getenv("MEMS_COARSE_FP32")) ? 1 : 0;
    ra.target_acc = (h->sampler == MEMS_MODE_CW) ? 0.21 / h->host_pb.c.P + 0.23 : 0.234;
    MemsLaunchCtx ctx = make_ctx(h, stream);
    const int e = (h->cfg.path == MEMS_PATH_TC) ? mems_tc_run(ctx, ra) : mems_fp32_run(ctx, ra);
    h->launches += 1;
    if (e != 0) return fail(MEMS_E_CUDA, "run launch failed: %s", cudaGetErrorString((cudaError_t)e));
    h->steps_done += (unsigned long long)ra.n_steps;
    return MEMS_OK;
}

int mems_run(mems_handle_t h, int32_t n_steps, int32_t adapt_window, int32_t thin, double* samples_out,
             double* loglik_out, void* stream) {
    if (!h) return fail(MEMS_E_INVALID, "null handle");
    if (!h->have_chains) return fail(MEMS_E_STATE, "call mems_init_chains first");
    if (n_steps < 0 || adapt_window < 0 || thin < 1) return fail(MEMS_E_INVALID, "bad n_steps/adapt_window/thin");
    if (n_steps == 0) return MEMS_OK;
    MemsRunArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.n_steps = n_steps; ra.adapt_window = adapt_window; ra.thin = thin;
    ra.samples_out = samples_out; ra.loglik_out = loglik_out;
    return run_common(h, ra, stream);
}

int mems_run_explicit(mems_handle_t h, int32_t n_steps, int32_t adapt_window, const double* xi, const double* log_u,
                      double* samples_out, double* loglik_prop_out, uint8_t* decisions_out, void* stream) {
    if (!h || !xi || !log_u) return fail(MEMS_E_INVALID, "null argument");
    if (!h->have_chains) return fail(MEMS_E_STATE, "call mems_init_chains first");
    if (n_steps < 0 || adapt_window < 0) return fail(MEMS_E_INVALID, "bad n_steps/adapt_window");
    if (n_steps == 0) return MEMS_OK;
    MemsRunArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.n_steps = n_steps; ra.adapt_window = adapt_window; ra.thin = 1;
    ra.xi = xi; ra.log_u = log_u;
    ra.samples_out = samples_out; ra.llprop_out = loglik_prop_out; ra.decisions_out = decisions_out;
    return run_common(h, ra, stream);
}

int mems_draw(mems_handle_t h, uint64_t step0, int32_t n_steps, double* xi_out, double* log_u_out, void* stream) {
    const int n_sub = !h ? 1 : (h->sampler == MEMS_MODE_CW) ? h->host_pb.c.P : (h->sampler == MEMS_MODE_DA ? 2 : 1);
    if (!h) return fail(MEMS_E_INVALID, "null handle");
    if (!h->have_chains) return fail(MEMS_E_STATE, "call mems_init_chains first");
    if (n_steps <= 0) return MEMS_OK;
    DEVICE_SCOPE(h->cfg.device);
    MemsLaunchCtx ctx = make_ctx(h, stream);
    const int e = mems_draw_launch(ctx, h->sampler, n_sub, step0, n_steps, xi_out, log_u_out);
    h->launches += 1;
    if (e != 0) return fail(MEMS_E_CUDA, "draw launch failed: %s", cudaGetErrorString((cudaError_t)e));
    return MEMS_OK;
}

int mems_get_state(mems_handle_t h, double* x, double* loglik, double* scale, uint32_t* accept_count,
                   uint64_t* steps_done, void* stream) {
    if (!h) return fail(MEMS_E_INVALID, "null handle");
    if (!h->have_chains) return fail(MEMS_E_STATE, "call mems_init_chains first");
    DEVICE_SCOPE(h->cfg.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t C = (size_t)h->chains.C, P = (size_t)h->host_pb.c.P;
    if (x) CUDA_TRY(cudaMemcpyAsync(x, h->chains.x, sizeof(double) * P * C, cudaMemcpyDeviceToDevice, st));
    if (loglik) CUDA_TRY(cudaMemcpyAsync(loglik, h->chains.ll, sizeof(double) * C, cudaMemcpyDeviceToDevice, st));
    if (scale) CUDA_TRY(cudaMemcpyAsync(scale, h->chains.scale, sizeof(double) * C, cudaMemcpyDeviceToDevice, st));
    if (accept_count)
        CUDA_TRY(cudaMemcpyAsync(accept_count, h->chains.acc_count, sizeof(unsigned int) * C, cudaMemcpyDeviceToDevice, st));
    if (steps_done) {
        CUDA_TRY(cudaStreamSynchronize(st));
        *steps_done = h->steps_done;
    }
    return MEMS_OK;
}

int mems_get_aux(mems_handle_t h, double* scale_cw, double* loglik_coarse, uint32_t* promote_count, void* stream) {
    if (!h) return fail(MEMS_E_INVALID, "null handle");
    if (!h->have_chains) return fail(MEMS_E_STATE, "call mems_init_chains first");
    DEVICE_SCOPE(h->cfg.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t C = (size_t)h->chains.C, P = (size_t)h->host_pb.c.P;
    if (scale_cw) CUDA_TRY(cudaMemcpyAsync(scale_cw, h->chains.scale_cw, sizeof(double) * P * C, cudaMemcpyDeviceToDevice, st));
    if (loglik_coarse) CUDA_TRY(cudaMemcpyAsync(loglik_coarse, h->chains.llc, sizeof(double) * C, cudaMemcpyDeviceToDevice, st));
    if (promote_count)
        CUDA_TRY(cudaMemcpyAsync(promote_count, h->chains.prom_count, sizeof(unsigned int) * C, cudaMemcpyDeviceToDevice, st));
    return MEMS_OK;
}

int mems_mlp_forward(int32_t device, int32_t n_layers, const int32_t* widths, const float* const* W,
                     const float* const* b, const float* X_dev, int64_t n, float* Y_dev, void* stream) {
    if (!widths || !W || !b || !X_dev || !Y_dev) return fail(MEMS_E_INVALID, "null argument");
    if (n_layers < 1 || n_layers > 9) return fail(MEMS_E_INVALID, "n_layers=%d outside [1,9]", n_layers);
    for (int i = 0; i <= n_layers; ++i)
        if (widths[i] < 1 || widths[i] > MEMS_WIDTH) return fail(MEMS_E_INVALID, "widths[%d]=%d outside [1,%d]", i, widths[i], MEMS_WIDTH);
    if (n < 0) return fail(MEMS_E_INVALID, "n < 0");
    if (n == 0) return MEMS_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(MEMS_E_UNSUPPORTED, "no CUDA device; this library has no CPU fallback");
    DEVICE_SCOPE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    float* Wd[9] = {nullptr};
    float* bd[9] = {nullptr};
    int rc = MEMS_OK;
    for (int l = 0; l < n_layers && rc == MEMS_OK; ++l) {
        const size_t wb = sizeof(float) * widths[l] * widths[l + 1], bb = sizeof(float) * widths[l + 1];
        if (cudaMalloc(&Wd[l], wb) != cudaSuccess || cudaMalloc(&bd[l], bb) != cudaSuccess ||
            cudaMemcpyAsync(Wd[l], W[l], wb, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(bd[l], b[l], bb, cudaMemcpyHostToDevice, st) != cudaSuccess)
            rc = fail(MEMS_E_CUDA, "weight upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc == MEMS_OK) {
        const int e = mems_mlp_forward_launch(n_layers, widths, Wd, bd, X_dev, n, Y_dev, prop.multiProcessorCount, st);
        if (e != 0) rc = fail(MEMS_E_CUDA, "mlp launch failed: %s", cudaGetErrorString((cudaError_t)e));
    }
    cudaStreamSynchronize(st);   // weights are temporaries of this call
    for (int l = 0; l < n_layers; ++l) { cudaFree(Wd[l]); cudaFree(bd[l]); }
    return rc;
}

int mems_chain_moments(int32_t device, const double* samples_dev, int32_t n_keep, int32_t n_params, int32_t n_chains,
                       double* out_dev, void* stream) {
    if (!samples_dev || !out_dev) return fail(MEMS_E_INVALID, "null argument");
    if (n_keep < 4 || n_params < 1 || n_chains < 1) return fail(MEMS_E_INVALID, "need n_keep >= 4, n_params >= 1, n_chains >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(MEMS_E_UNSUPPORTED, "no CUDA device; this library has no CPU fallback");
    DEVICE_SCOPE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int e = mems_chain_moments_launch(samples_dev, n_keep, n_params, n_chains, out_dev, prop.multiProcessorCount,
                                            reinterpret_cast<cudaStream_t>(stream));
    if (e != 0) return fail(MEMS_E_CUDA, "chain_moments launch failed: %s", cudaGetErrorString((cudaError_t)e));
    return MEMS_OK;
}

int mems_lsq(mems_handle_t h, const double* x0_dev, int32_t n, const double* yobs_dev, int32_t n_obs,
             const double* bounds_lo, const double* bounds_hi, int32_t max_iter, double rel_step, double xtol, double ftol,
             double* x_out_dev, double* cov_out_dev, double* cost_out_dev, int32_t* iters_out_dev, void* stream) {
    if (!h) return fail(MEMS_E_INVALID, "null handle");
    if (!h->have_weights || !h->have_problem) return fail(MEMS_E_STATE, "set weights and problem first");
    if (n < 0 || max_iter < 1) return fail(MEMS_E_INVALID, "need n >= 0 and max_iter >= 1");
    if (!(rel_step >= 0.0) || rel_step > 0.1) return fail(MEMS_E_INVALID, "rel_step must be in [0, 0.1] (0 = scipy's eps^(1/3))");
    if (rel_step == 0.0) rel_step = 6.0554544523933395e-06;
    if (n == 0) return MEMS_OK;
    if (!x0_dev || !x_out_dev) return fail(MEMS_E_INVALID, "null argument");
    if (yobs_dev && n_obs != 1 && n_obs != n) return fail(MEMS_E_INVALID, "n_obs must be 1 or n");
    if ((bounds_lo == nullptr) != (bounds_hi == nullptr)) return fail(MEMS_E_INVALID, "give both bounds or neither");
    const int P = h->host_pb.c.P, T = h->host_pb.c.T;
    if ((long long)n * (2 * P + 1) > 2000000000LL) return fail(MEMS_E_INVALID, "too many starts");
    double b[2 * MEMS_PMAX];
    for (int p = 0; p < P; ++p) {
        b[p] = bounds_lo ? bounds_lo[p] : h->host_pb.c.low[p];
        b[MEMS_PMAX + p] = bounds_hi ? bounds_hi[p] : h->host_pb.c.high[p];
        if (!(b[p] < b[MEMS_PMAX + p])) return fail(MEMS_E_INVALID, "bounds of parameter %d are empty", p);
    }
    DEVICE_SCOPE(h->cfg.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t wbytes = mems_lsq_workspace_bytes(n, P, T) + 16 * 256;
    void* ws = nullptr;
    double* bdev = nullptr;
    CUDA_TRY(cudaMallocAsync(&ws, wbytes, st));
    cudaError_t e1 = cudaMallocAsync(reinterpret_cast<void**>(&bdev), sizeof(b), st);
    if (e1 != cudaSuccess) { cudaFreeAsync(ws, st); return fail(MEMS_E_CUDA, "cudaMallocAsync failed: %s", cudaGetErrorString(e1)); }
    // the bounds travel through pageable memory: cudaMemcpyAsync stages them before returning
    e1 = cudaMemcpyAsync(bdev, b, sizeof(b), cudaMemcpyHostToDevice, st);
    int rc = MEMS_OK;
    if (e1 != cudaSuccess) rc = fail(MEMS_E_CUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e1));
    if (rc == MEMS_OK) {
        MemsLaunchCtx ctx = make_ctx(h, stream);
        const int e = mems_lsq_run(ctx, h->cfg.path == MEMS_PATH_TC, x0_dev, n, yobs_dev ? yobs_dev : h->dev_yobs,
                                   yobs_dev ? n_obs : 1, bdev, bdev + MEMS_PMAX, max_iter, rel_step, xtol, ftol, ws, x_out_dev, cov_out_dev,
                                   cost_out_dev, iters_out_dev);
        h->launches += 3LL * max_iter + 2;
        if (e != 0) rc = fail(MEMS_E_CUDA, "least-squares launch failed: %s", cudaGetErrorString((cudaError_t)e));
    }
    cudaFreeAsync(bdev, st);
    cudaFreeAsync(ws, st);
    return rc;
}

int mems_chain_autocov(int32_t device, const double* samples_dev, int32_t n_keep, int32_t n_params, int32_t n_chains,
                       int32_t max_lag, double* out_dev, void* stream) {
    if (!samples_dev || !out_dev) return fail(MEMS_E_INVALID, "null argument");
    if (n_keep < 4 || n_params < 1 || n_chains < 1) return fail(MEMS_E_INVALID, "need n_keep >= 4, n_params >= 1, n_chains >= 1");
    if (max_lag < 1 || max_lag > n_keep) return fail(MEMS_E_INVALID, "max_lag must be in [1, n_keep]");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(MEMS_E_UNSUPPORTED, "no CUDA device; this library has no CPU fallback");
    DEVICE_SCOPE(device);
    const size_t wbytes = mems_acov_workspace_bytes(n_keep, n_params, n_chains, max_lag);
    if (wbytes == 0) return fail(MEMS_E_INVALID, "n_keep=%d too long for the shared-memory autocovariance kernel (max 20480)", n_keep);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    void* ws = nullptr;
    CUDA_TRY(cudaMallocAsync(&ws, wbytes, st));
    const int e = mems_acov_launch(samples_dev, n_keep, n_params, n_chains, max_lag, ws, out_dev, st);
    cudaFreeAsync(ws, st);
    if (e != 0) return fail(MEMS_E_CUDA, "autocovariance launch failed: %s", cudaGetErrorString((cudaError_t)e));
    return MEMS_OK;
}

namespace {
int device_sm_count(int device, int* sm_count) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(MEMS_E_UNSUPPORTED, "no CUDA device; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(MEMS_E_INVALID, "device %d out of range", device);
    CUDA_TRY(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device));
    return MEMS_OK;
}
}  // namespace

int64_t mems_sort_capacity(int64_t n) { return n < 0 ? 0 : (int64_t)mems_sort_padded((size_t)n); }

int mems_sort_f64(int32_t device, double* keys_dev, int64_t n, void* stream) {
    if (n < 0) return fail(MEMS_E_INVALID, "n < 0");
    if (n <= 1) return MEMS_OK;
    if (!keys_dev) return fail(MEMS_E_INVALID, "null argument");
    int sms = 0;
    const int rc = device_sm_count(device, &sms);
    if (rc != MEMS_OK) return rc;
    DEVICE_SCOPE(device);
    const int e = mems_sort_launch(keys_dev, (size_t)n, sms, reinterpret_cast<cudaStream_t>(stream));
    if (e != 0) return fail(MEMS_E_CUDA, "sort launch failed: %s", cudaGetErrorString((cudaError_t)e));
    return MEMS_OK;
}

int mems_rank_normalize(int32_t device, const double* sorted_dev, int64_t n_pool, const double* x_dev, int64_t n, double* z_dev,
                        void* stream) {
    if (n < 0 || n_pool < 1) return fail(MEMS_E_INVALID, "need n >= 0 and n_pool >= 1");
    if (n == 0) return MEMS_OK;
    if (!sorted_dev || !x_dev || !z_dev) return fail(MEMS_E_INVALID, "null argument");
    int sms = 0;
    const int rc = device_sm_count(device, &sms);
    if (rc != MEMS_OK) return rc;
    DEVICE_SCOPE(device);
    const int e = mems_rank_normalize_launch(sorted_dev, (size_t)n_pool, x_dev, (size_t)n, z_dev, sms, reinterpret_cast<cudaStream_t>(stream));
    if (e != 0) return fail(MEMS_E_CUDA, "rank_normalize launch failed: %s", cudaGetErrorString((cudaError_t)e));
    return MEMS_OK;
}

int mems_ess_geyer(int32_t device, const double* acov_dev, int32_t n_params, int32_t max_lag, int64_t n_draws, int64_t n_chains,
                   double* ess_out_dev, int32_t* truncated_out_dev, void* stream) {
    if (!acov_dev || !ess_out_dev || !truncated_out_dev) return fail(MEMS_E_INVALID, "null argument");
    if (n_params < 1 || max_lag < 2 || n_draws < 4 || n_chains < 1)
        return fail(MEMS_E_INVALID, "need n_params >= 1, max_lag >= 2, n_draws >= 4, n_chains >= 1");
    int sms = 0;
    const int rc = device_sm_count(device, &sms);
    if (rc != MEMS_OK) return rc;
    DEVICE_SCOPE(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    double* rho = nullptr;
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&rho), sizeof(double) * (size_t)n_params * ((size_t)max_lag + 2), st));
    const int e = mems_geyer_launch(acov_dev, n_params, max_lag, (long long)n_draws, (double)n_chains, rho, ess_out_dev,
                                    truncated_out_dev, st);
    cudaFreeAsync(rho, st);
    if (e != 0) return fail(MEMS_E_CUDA, "geyer launch failed: %s", cudaGetErrorString((cudaError_t)e));
    return MEMS_OK;
}

int64_t mems_launch_count(mems_handle_t h) { return h ? h->launches : 0; }

int mems_rows_evaluated(mems_handle_t h, uint64_t* rows, void* stream) {
    if (!h || !rows) return fail(MEMS_E_INVALID, "null argument");
    if (!h->have_chains) return fail(MEMS_E_STATE, "call mems_init_chains first");
    DEVICE_SCOPE(h->cfg.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned long long v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, h->chains.rows_eval, sizeof(v), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *rows = (uint64_t)v;
    return MEMS_OK;
}

}  // extern "C"
