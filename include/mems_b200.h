/*
 * mems_b200.h -- C ABI of the B200-native batched surrogate Metropolis-Hastings path.
 *
 * The reference (filippozacchei/MCMC-MEMS) is pure Python and has no FFI of its own;
 * the boundary below is what a binding for its hot path has to offer.  Each entry
 * point names the reference interface it replaces (paths relative to the reference
 * tree; "cuqi" = the un-vendored dependency cuqipy==0.8.0, README.md:36).
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types cross the boundary;
 *   - every function returns 0 on success or a negative MEMS_E_* code; the message
 *     of the last failure on the calling thread is mems_last_error();
 *   - "host" arrays are read during the call and copied; "device" arrays are CUDA
 *     device pointers on the handle's device and must stay valid until the stream
 *     reaches the end of the enqueued work;
 *   - all work is enqueued on the cudaStream_t passed as `void* stream`
 *     (NULL = the legacy default stream); calls do not synchronise unless stated;
 *   - one handle per device, not thread-safe per handle; every call runs on the handle's device and
 *     restores the caller's current CUDA device before it returns;
 *   - there is no CPU fallback: without a usable sm_100 device mems_create fails.
 *
 * Per-chain arrays are stored parameter-major, chain-minor ("[P][C]": element
 * (p, c) at p*C + c) so that consecutive chains are consecutive in memory.
 */
#ifndef MEMS_B200_H
#define MEMS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEMS_PMAX 4          /* max number of sampled parameters (reference uses 3 and 4) */
#define MEMS_WIDTH 64        /* hidden width of the surrogate (config_I.json: N_NEURONS)  */
#define MEMS_HIDDEN 6        /* number of tanh layers (config_I.json: N_LAYERS)           */

#define MEMS_PATH_FP32 0     /* CUDA-core FP32 check path (parity reference on the GPU)   */
#define MEMS_PATH_TC   1     /* tcgen05 split-FP16 tensor-core path (the product path)    */

#define MEMS_OK             0
#define MEMS_E_INVALID     -1   /* bad argument / shape                */
#define MEMS_E_CUDA        -2   /* CUDA runtime error                  */
#define MEMS_E_STATE       -3   /* call order (weights/problem/chains) */
#define MEMS_E_UNSUPPORTED -4   /* device or configuration unsupported */

typedef struct MemsConfig {
    int32_t n_params;   /* P: sampled parameters = surrogate inputs minus the time column */
    int32_t n_time;     /* T: time points per signal (len(data_processor.time)); y_obs and the
                           time column live in shared memory: T <= 1961 on the tensor-core
                           path, 3621 on the FP32 path (mems_create reports the limit)      */
    int32_t n_chains;   /* C: independent chains resident on this device                  */
    int32_t path;       /* MEMS_PATH_FP32 or MEMS_PATH_TC                                 */
    int32_t device;     /* CUDA device ordinal                                            */
    int32_t chains_per_block; /* 0 = choose; else 1..128 chains per block batch           */
    int32_t reserved[2];
} MemsConfig;

typedef struct MemsHandle_* mems_handle_t;

const char* mems_last_error(void);
int  mems_version(void);

/* Lifetime.  Replaces: the notebook-level objects built once per run
 * (tests/Xaccelerometer_geometric/identification.ipynb cells 8-14). */
int  mems_create(const MemsConfig* cfg, mems_handle_t* out);
void mems_destroy(mems_handle_t h);

/* Surrogate weights.  Replaces NN_Model.load_model (src/SurrogateModeling/model.py:28-34):
 * the Keras Dense stack [P+1]->64 x6 tanh ->1 in Keras layout kernel[in][out], float32.
 * W[0]: (P+1) x 64, W[1..5]: 64 x 64, W[6]: 64 x 1;  b[0..5]: 64, b[6]: 1.   HOST pointers. */
int  mems_set_weights(mems_handle_t h, const float* const* W, const float* const* b);

/* Problem constants.  Replaces the closure state of create_forward_model_function
 * (src/InverseProblems/utils.py:22-28: time grid, StandardScaler mean_/scale_) and the
 * cuqi objects of identification.ipynb cell 10/14: Gaussian likelihood (y_obs, sigma2 = the
 * iid noise variance), Uniform prior (low, high) and the Gaussian proposal (chol = lower
 * Cholesky factor of its covariance, row-major P x P).                      HOST pointers.
 * time[T], scaler_mean[P+1], scaler_scale[P+1], y_obs[T], low[P], high[P], chol[P*P].
 * Synchronous (blocking copies): no work of this handle may be in flight on any stream.  Chains initialised
 * before the call are invalidated (their stored log-likelihoods belong to the old problem): call
 * mems_init_chains again before mems_run. */
int  mems_set_problem(mems_handle_t h, const double* time, const double* scaler_mean,
                      const double* scaler_scale, const double* y_obs, double sigma2,
                      const double* low, const double* high, const double* chol);

/* Forward model + log-likelihood for n parameter vectors.  Replaces forward_model(x)
 * (src/InverseProblems/utils.py:30-40) and Gaussian(...).logd(y_obs) (identification.ipynb:169).
 * x_dev [P][n] (device, f64).  y_out [n][T] (device, f32) and loglik_out [n] (device, f64) may
 * each be NULL.  loglik = -0.5 * sum_t (y_obs - f)^2 / sigma2, i.e. WITHOUT the additive
 * constant -0.5*T*(log sigma2 + log 2pi) and without the prior term. */
int  mems_forward(mems_handle_t h, const double* x_dev, int32_t n, float* y_out,
                  double* loglik_out, void* stream);

/* Sampler variant (default MEMS_SAMPLER_MH = the reference's sampler).  The two others have NO
 * implementation or artefact in the reference; they follow cuqi.sampler.CWMH
 * semantics and Christen & Fox (2005):
 *   CWMH: per update draw x_o = x + scale (.) z once, then for p = 0..P-1 replace component p,
 *         score, accept with its own uniform -- P surrogate passes per update; cw_scale[P] (HOST)
 *         is the initial per-component proposal std; adaptation target 0.21/P + 0.23.
 *   DA:   stage 1 scores every da_stride-th time point (log-likelihood rescaled by T/Tc) and promotes
 *         with log u1 <= min(0, lc* - lc); stage 2 (promoted only) scores all T points and accepts
 *         with log u2 <= min(0, (l* - l) - (lc* - lc)).  da_stride = 0 selects the FFT coarse surrogate as stage 1
 *         instead (mems_set_coarse_fft).  Every pass lays out only the chains that need it (inside the prior box;
 *         stage 2: promoted), so stage 2 costs in proportion to the promotion rate.
 * Call before mems_init_chains.  With n_sub = 1 (MH), P (CWMH), 2 (DA) the per-update arrays of
 * mems_run_explicit / mems_draw become log_u [n_steps][n_sub][C], loglik_prop_out and
 * decisions_out [n_steps][n_sub][C]; for CWMH xi holds the standard normals z, not chol*z. */
#define MEMS_SAMPLER_MH   0
#define MEMS_SAMPLER_CWMH 1
#define MEMS_SAMPLER_DA   2
int  mems_set_sampler(mems_handle_t h, int32_t sampler, const double* cw_scale, int32_t da_stride);

/* Coarse first-stage model of delayed acceptance: the reference's FFT surrogate (models/model_VoltageSignal_fft.keras,
 * trained on the features of src/utils/preprocessing.py:90-103 -- [10*mean, |F_1..7|, 100*angle(F_1..7)] of the trace),
 * turned back into a band-limited trace with the reference's reconstruction
 * (tests/Xaccelerometer_quality_factor/surrogate_fft.ipynb cell 15) and scored against the problem's y_obs with the
 * problem's sigma2.  One surrogate row per proposal instead of T.
 * W/b: HOST pointer arrays of the 7 Dense layers (Keras layout, [P][64], 5 x [64][64], [64][15]); scaler_mean/scale
 * HOST [P] (the StandardScaler of the FFT model's inputs, parameters only).  Then mems_set_sampler(h, DA, NULL, 0). */
int  mems_set_coarse_fft(mems_handle_t h, const float* const* W, const float* const* b,
                         const double* scaler_mean, const double* scaler_scale);

/* Chain initialisation.  Replaces MH(target, proposal, x0) + the first lines of
 * cuqi MH._sample_adapt (samples[:,0]=x0, target_eval[0]=logd(x0), acc[0]=1, scale=0.1):
 * x0_dev [P][C] (device, f64); scale0 = initial proposal scale; seed = Philox key.
 * chain_id0 = global id of this device's first chain (Philox counters carry global ids so
 * results do not depend on how chains are sharded over GPUs). */
int  mems_init_chains(mems_handle_t h, const double* x0_dev, double scale0, uint64_t seed,
                      uint64_t chain_id0, void* stream);

/* n_steps Metropolis-Hastings updates of every chain with device-side Philox randomness.
 * Replaces the hot loop of cuqi MH._sample_adapt / MH._sample driven from
 * identification.ipynb:302 (single_update: x* = x + scale*xi, accept iff
 * log u <= min(0, l* - l)).  adapt_window = Na (= int(0.1*N) in cuqi; 0 = frozen scale).
 * After every `thin`-th update of this call the state is written:
 *   samples_out [n_steps/thin][P][C] (device, f64, may be NULL),
 *   loglik_out  [n_steps/thin][C]    (device, f64, may be NULL; same convention as mems_forward,
 *                                     -inf never occurs for a stored state unless x0 was outside). */
int  mems_run(mems_handle_t h, int32_t n_steps, int32_t adapt_window, int32_t thin,
              double* samples_out, double* loglik_out, void* stream);

/* Same update with caller-supplied randomness (parity mode, no Philox):
 * xi [n_steps][P][C] (device, f64; already-correlated proposal increments, x* = x + scale*xi),
 * log_u [n_steps][C] (device, f64).  Outputs (device, each may be NULL):
 * samples_out [n_steps][P][C] f64, loglik_prop_out [n_steps][C] f64 (proposal log-likelihood,
 * -inf when the proposal is outside the prior box), decisions_out [n_steps][C] u8. */
int  mems_run_explicit(mems_handle_t h, int32_t n_steps, int32_t adapt_window,
                       const double* xi, const double* log_u, double* samples_out,
                       double* loglik_prop_out, uint8_t* decisions_out, void* stream);

/* The randomness mems_run would consume for updates [step0, step0+n_steps) of every chain:
 * xi_out [n_steps][P][C] = chol * z (f64), log_u_out [n_steps][C] (f64).  Device pointers. */
int  mems_draw(mems_handle_t h, uint64_t step0, int32_t n_steps, double* xi_out,
               double* log_u_out, void* stream);

/* Current chain state (device pointers, each may be NULL): x [P][C] f64, loglik [C] f64,
 * scale [C] f64 (cuqi's self.scale), accept_count [C] u32 (accepted updates since init, the
 * initial acc[0]=1 not included), steps_done (HOST pointer, synchronises the stream). */
int  mems_get_state(mems_handle_t h, double* x, double* loglik, double* scale,
                    uint32_t* accept_count, uint64_t* steps_done, void* stream);

/* Variant-specific state (device pointers, each may be NULL): scale_cw [P][C] f64 (CWMH),
 * loglik_coarse [C] f64 and promote_count [C] u32 (DA). */
int  mems_get_aux(mems_handle_t h, double* scale_cw, double* loglik_coarse, uint32_t* promote_count,
                  void* stream);

/* Generic Dense-stack inference.  Replaces NN_Model.predict / nn_model.model(X)
 * (src/SurrogateModeling/model.py:142-151, src/InverseProblems/utils.py:39) for any of the
 * reference's Sequential models: n_layers Dense layers, widths[0..n_layers] (input width
 * first, all <= 64), tanh on every layer except the last.  W/b are HOST pointer arrays
 * (Keras layout), X_dev [n][widths[0]] f32 and Y_dev [n][widths[n_layers]] f32 device pointers. */
int  mems_mlp_forward(int32_t device, int32_t n_layers, const int32_t* widths,
                      const float* const* W, const float* const* b, const float* X_dev,
                      int64_t n, float* Y_dev, void* stream);

/* Split-chain moments of kept samples: the sufficient statistics of split-R-hat, computed where the
 * samples live.  Replaces the first step of Samples.compute_rhat (arviz underneath; call site
 * tests/Xaccelerometer_geometric/identification.ipynb:311) at scale.
 * samples_dev [n_keep][P][C] f64 -> out_dev [4][P][C] f64: mean and unbiased variance of the first
 * n_keep/2 draws, then of the last n_keep/2 draws, per (parameter, chain). */
int  mems_chain_moments(int32_t device, const double* samples_dev, int32_t n_keep, int32_t n_params,
                        int32_t n_chains, double* out_dev, void* stream);

/* Lag-autocovariance sums of kept samples: the sufficient statistics of the (multi-chain) effective sample
 * size, computed where the samples live.  Replaces the first step of Samples.compute_ess (arviz underneath; call
 * site tests/Xaccelerometer_geometric/identification.ipynb:308) at scale.
 * samples_dev [n_keep][P][C] f64 -> out_dev [P][max_lag + 2] f64: for every parameter the SUM over chains of the
 * biased autocovariance (1/n) sum_k (x_k - m)(x_{k+lag} - m) for lag = 0..max_lag-1, then the sum of the chain
 * means m and the sum of m^2.  Sums (not means) so that shards on several GPUs combine with one all-reduce.
 * Deterministic (fixed reduction order).  n_keep <= 20480. */
int  mems_chain_autocov(int32_t device, const double* samples_dev, int32_t n_keep, int32_t n_params,
                        int32_t n_chains, int32_t max_lag, double* out_dev, void* stream);

/* Rank-normalised diagnostics (arviz defaults behind Samples.compute_ess / compute_rhat: bulk ESS and rank-normalised
 * split R-hat; call sites tests/Xaccelerometer_geometric/identification.ipynb:308,311), three device steps:
 *
 * mems_sort_f64: ascending in-place sort of n f64 keys.  keys_dev must have room for mems_sort_capacity(n) doubles
 * (the network works on a power-of-two length; the tail is scratch).  Deterministic.
 *
 * mems_rank_normalize: for each of the n values x_dev[i], the tie-averaged 1-based rank r among the n_pool sorted
 * values sorted_dev (the pooled draws of every chain -- of every GPU when the pool was gathered), then
 * z_dev[i] = Phi^-1((r - 3/8)/(n_pool + 1/4)).  x_dev and z_dev may alias.
 *
 * mems_ess_geyer: effective sample size from autocovariance sums in the layout mems_chain_autocov writes
 * (acov_dev [n_params][max_lag + 2], SUMS over n_chains chains of n_draws draws, possibly all-reduced over GPUs):
 * Geyer's initial-positive + initial-monotone truncation.  ess_out_dev [n_params] f64; truncated_out_dev [n_params]
 * i32 = 1 when the sequence had not turned negative by max_lag. */
int64_t mems_sort_capacity(int64_t n);
int  mems_sort_f64(int32_t device, double* keys_dev, int64_t n, void* stream);
int  mems_rank_normalize(int32_t device, const double* sorted_dev, int64_t n_pool, const double* x_dev, int64_t n,
                         double* z_dev, void* stream);
int  mems_ess_geyer(int32_t device, const double* acov_dev, int32_t n_params, int32_t max_lag, int64_t n_draws,
                    int64_t n_chains, double* ess_out_dev, int32_t* truncated_out_dev, void* stream);

/* Batched bounded least-squares fit of the surrogate to observed traces: n independent starts advance together.
 * Replaces least_squares_optimization (src/InverseProblems/utils.py:83-93: scipy least_squares with
 * jac='3-point' and bounds, then inv(J^T J)), which the notebooks call once per start
 * (tests/Xaccelerometer_geometric/identification.ipynb:222-227).  Levenberg-Marquardt on the same 3-point
 * finite-difference Jacobian (step rel_step*max(1,|x|), rel_step = 0 selects scipy's eps^(1/3); one-sided next to
 * a bound).  With scipy's step the float32 surrogate makes J, and therefore inv(J^T J), noise-dominated (the
 * reference has the same property); rel_step ~ 3e-4 gives the noise-free covariance.  Every iteration
 * evaluates the surrogate at the 2P+1 stencil points of every start with the handle's forward kernel.
 * x0_dev [P][n] f64; yobs_dev [n_obs][T] f64 with n_obs = 1 (shared) or n (one trace per start), NULL = the
 * handle's y_obs; bounds_lo/hi HOST [P] or NULL = the prior box; stops per start on scipy's xtol / ftol rules or
 * after max_iter iterations.  Outputs (device): x_out [P][n], cov_out [n][P][P] = inv(J^T J) (may be NULL),
 * cost_out [n] = 0.5*||f(x)-y||^2 (may be NULL), iters_out [n] accepted steps (may be NULL). */
int  mems_lsq(mems_handle_t h, const double* x0_dev, int32_t n, const double* yobs_dev, int32_t n_obs,
              const double* bounds_lo, const double* bounds_hi, int32_t max_iter, double rel_step, double xtol,
              double ftol, double* x_out_dev, double* cov_out_dev, double* cost_out_dev, int32_t* iters_out_dev,
              void* stream);

/* Surrogate rows (chain, time point) the run launches of this handle have evaluated since mems_init_chains: the
 * tensor-core path skips proposals outside the prior box and, in delayed acceptance, proposals the first stage did
 * not promote, so this -- not chains x updates x T -- is the work behind a throughput figure (bench bookkeeping).
 * rows: HOST pointer; synchronises the stream. */
int  mems_rows_evaluated(mems_handle_t h, uint64_t* rows, void* stream);

/* Number of kernel launches issued through this handle since creation (bench bookkeeping). */
int64_t mems_launch_count(mems_handle_t h);

#ifdef __cplusplus
}
#endif
#endif /* MEMS_B200_H */
