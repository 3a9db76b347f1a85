// Rank-normalised chain diagnostics where the samples live: the arithmetic behind
// Samples.compute_ess / compute_rhat (arviz defaults underneath; call sites
// tests/Xaccelerometer_geometric/identification.ipynb:308,311) for batches too large to ship to the host.
//
//   sort          in-place ascending sort of f64 keys (bitonic network: shared-memory stages for partner distances
//                 below a 2048-element tile, one compare-exchange kernel per longer distance).  O(n log^2 n), but a
//                 sort happens once per parameter per run and the network is deterministic and branch-free;
//   rank -> z     per draw: tie-averaged rank among ALL draws of all chains (two binary searches in the sorted pool:
//                 a Metropolis chain repeats its state at every rejection, so ties are the rule), then
//                 z = Phi^-1((r - 3/8)/(M + 1/4))  (Blom; arviz `_z_scale`);
//   Geyer         initial-positive + initial-monotone sequence truncation of the chain-averaged autocovariances
//                 (arviz `_ess`), one thread per parameter -- the vector has at most a few thousand entries.
// The pool may be the gathered draws of several GPUs (distributed.py): ranks are then global and the estimators do
// not depend on how the chains are sharded.
#include <cstdint>

#include "mems_common.cuh"

namespace {

constexpr int SORT_NT = 512;
constexpr int SORT_TILE = 2048;            // elements a block sorts / merges in shared memory

__device__ __forceinline__ void cmpx(double& a, double& b, bool asc) {
    if ((a > b) == asc) { const double t = a; a = b; b = t; }
}

// d[n..npad) = +inf (the padding sorts to the end)
__global__ void sort_pad_kernel(double* d, size_t n, size_t npad) {
    for (size_t i = n + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npad; i += (size_t)gridDim.x * blockDim.x)
        d[i] = CUDART_INF;
}

// Stages of the merges k = k_first, 2 k_first, ..., k_last whose partner distance j is below SORT_TILE, in shared memory.
// (k_first = 2, k_last = SORT_TILE: the complete sort of every tile.)
__global__ void __launch_bounds__(SORT_NT) bitonic_tile_kernel(double* d, size_t k_first, size_t k_last) {
    __shared__ double s[SORT_TILE];
    const size_t base = (size_t)blockIdx.x * SORT_TILE;
    for (int i = threadIdx.x; i < SORT_TILE; i += SORT_NT) s[i] = d[base + i];
    __syncthreads();
    for (size_t k = k_first; k <= k_last; k <<= 1) {
        int j = (int)((k >> 1) < (size_t)(SORT_TILE / 2) ? (k >> 1) : (size_t)(SORT_TILE / 2));
        for (; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < SORT_TILE / 2; t += SORT_NT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                cmpx(s[i], s[i + j], ((base + i) & k) == 0);
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < SORT_TILE; i += SORT_NT) d[base + i] = s[i];
}

// One stage (merge size k, partner distance j >= SORT_TILE) in global memory: thread t owns pair (i, i + j).
__global__ void bitonic_step_kernel(double* d, size_t npad, size_t k, size_t j) {
    const size_t half = npad >> 1;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < half; t += (size_t)gridDim.x * blockDim.x) {
        const size_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        double a = d[i], b = d[i + j];
        if ((a > b) == ((i & k) == 0)) { d[i] = b; d[i + j] = a; }
    }
}

// z[i] = Phi^-1((rank(x[i]) - 3/8) / (M + 1/4)), rank = average 1-based position of x[i]'s ties in sorted[0..M)
__global__ void rank_normalize_kernel(const double* __restrict__ sorted, size_t M, const double* __restrict__ x, size_t n,
                                      double* __restrict__ z) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        size_t lo = 0, hi = M;                       // first index with sorted[idx] >= v
        while (lo < hi) {
            const size_t mid = (lo + hi) >> 1;
            if (sorted[mid] < v) lo = mid + 1; else hi = mid;
        }
        const size_t first = lo;
        hi = M;                                      // first index with sorted[idx] > v
        while (lo < hi) {
            const size_t mid = (lo + hi) >> 1;
            if (sorted[mid] <= v) lo = mid + 1; else hi = mid;
        }
        // ties occupy 1-based ranks first+1 .. lo; a value absent from the pool (never the case for the callers) would get
        // the rank it would have between its neighbours
        const double r = 0.5 * ((double)(first + 1) + (double)lo);
        z[i] = normcdfinv((r - 0.375) / ((double)M + 0.25));
    }
}

// Geyer truncation: acov[p][0..L) = SUM over chains of the biased lag autocovariances, acov[p][L], acov[p][L+1] = sum of
// the chain means and of their squares (the layout mems_chain_autocov writes, possibly all-reduced over GPUs).
// out[p] = ESS, trunc[p] = 1 when the sequence had not turned negative by lag L.  rho: scratch [P][L+2].
__global__ void geyer_ess_kernel(const double* __restrict__ acov, int P, int L, long long n_draws, double n_chains,
                                 double* __restrict__ rho_all, double* __restrict__ out, int* __restrict__ trunc) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double* a = acov + (size_t)p * (L + 2);
    double* rho = rho_all + (size_t)p * (L + 2);
    const double m = n_chains, n = (double)n_draws;
    const double s1 = a[L], s2 = a[L + 1];
    const double chain_mean_var = m > 1.0 ? (s2 - s1 * s1 / m) / (m - 1.0) : 0.0;
    const double mean_var = a[0] / m * n / (n - 1.0);
    double var_plus = mean_var * (n - 1.0) / n;
    if (m > 1.0) var_plus += chain_mean_var;
    for (int t = 0; t < L + 2; ++t) rho[t] = 0.0;
    double even = 1.0;
    rho[0] = even;
    double odd = 1.0 - (mean_var - a[1] / m) / var_plus;
    rho[1] = odd;
    long long t = 1;
    const long long limit = (n_draws - 3 < (long long)L - 2) ? n_draws - 3 : (long long)L - 2;
    while (t < limit && even + odd > 0.0) {
        even = 1.0 - (mean_var - a[t + 1] / m) / var_plus;
        odd = 1.0 - (mean_var - a[t + 2] / m) / var_plus;
        if (even + odd >= 0.0) { rho[t + 1] = even; rho[t + 2] = odd; }
        t += 2;
    }
    trunc[p] = (even + odd > 0.0 && limit < n_draws - 3) ? 1 : 0;
    const long long max_t = t - 2;
    if (even > 0.0) rho[max_t + 1] = even;
    t = 1;
    while (t <= max_t - 2) {
        if (rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]) {
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0;
            rho[t + 2] = rho[t + 1];
        }
        t += 2;
    }
    double acc = 0.0;
    for (long long k = 0; k <= max_t; ++k) acc += rho[k];
    const double total = m * n;
    double tau = -1.0 + 2.0 * acc + rho[max_t + 1];
    const double floor_tau = 1.0 / log10(total);
    if (tau < floor_tau) tau = floor_tau;
    out[p] = total / tau;
}

}  // namespace

size_t mems_sort_padded(size_t n) {
    size_t npad = SORT_TILE;
    while (npad < n) npad <<= 1;
    return npad;
}

// keys: device buffer with room for mems_sort_padded(n) doubles; the first n are sorted ascending in place
int mems_sort_launch(double* keys, size_t n, int sm_count, cudaStream_t st) {
    const size_t npad = mems_sort_padded(n);
    const int wide = 8 * sm_count;
    if (npad > n) {
        size_t nb = (npad - n + 255) / 256;
        sort_pad_kernel<<<(int)(nb < (size_t)wide ? nb : (size_t)wide), 256, 0, st>>>(keys, n, npad);
    }
    const size_t ntile = npad / SORT_TILE;
    bitonic_tile_kernel<<<(unsigned)ntile, SORT_NT, 0, st>>>(keys, 2, SORT_TILE);
    for (size_t k = 2 * (size_t)SORT_TILE; k <= npad; k <<= 1) {
        for (size_t j = k >> 1; j >= (size_t)SORT_TILE; j >>= 1) {
            size_t nb = ((npad >> 1) + 255) / 256;
            bitonic_step_kernel<<<(unsigned)(nb < (size_t)(4 * wide) ? nb : (size_t)(4 * wide)), 256, 0, st>>>(keys, npad, k, j);
        }
        bitonic_tile_kernel<<<(unsigned)ntile, SORT_NT, 0, st>>>(keys, k, k);
    }
    return (int)cudaGetLastError();
}

int mems_rank_normalize_launch(const double* sorted, size_t M, const double* x, size_t n, double* z, int sm_count, cudaStream_t st) {
    size_t nb = (n + 255) / 256;
    const size_t cap = (size_t)16 * sm_count;
    rank_normalize_kernel<<<(unsigned)(nb < cap ? (nb ? nb : 1) : cap), 256, 0, st>>>(sorted, M, x, n, z);
    return (int)cudaGetLastError();
}

int mems_geyer_launch(const double* acov, int P, int L, long long n_draws, double n_chains, double* rho_scratch, double* out,
                      int* trunc, cudaStream_t st) {
    geyer_ess_kernel<<<(P + 31) / 32, 32, 0, st>>>(acov, P, L, n_draws, n_chains, rho_scratch, out, trunc);
    return (int)cudaGetLastError();
}
