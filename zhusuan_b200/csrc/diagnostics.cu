// Effective sample size of a chain of vector samples on the device (zhusuan/diagnostics.py:17-64,
// the Stan estimator): per dimension d
//     var+   = biased variance,  var = var+ * n / (n - 1)
//     rho_t  = 1 - (var - acov_t) / var+,   acov_t = mean_{i < n-t} (s_i - mu)(s_{i+t} - mu)
//     ess    = n / (1 + 2 * sum_{t = 0 .. first t with rho_t < 0} rho_t)
// so that long runs never round-trip their samples to the host.
//
// samples: [M, D] row-major (one row per kept iteration).  A block owns 32 consecutive dimensions
// (coalesced across threadIdx.x) and splits the rows over threadIdx.y; the lag loop runs until all
// of the block's dimensions have hit their first negative autocorrelation.  Sums are accumulated
// in double: the reference is float64 NumPy.
#include "common.cuh"

namespace {

constexpr int ESS_TY = 16;

__device__ __forceinline__ double ess_block_sum(double v, double (*red)[33]) {
  // sum over threadIdx.y for each threadIdx.x
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.y == 0) {
    for (int y = 0; y < ESS_TY; ++y) s += red[y][threadIdx.x];
    red[0][threadIdx.x] = s;
  }
  __syncthreads();
  s = red[0][threadIdx.x];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(32 * ESS_TY) ess_kernel(const float* __restrict__ s, int64_t M,
                                                          int64_t D, float* __restrict__ ess) {
  __shared__ double red[ESS_TY][33];
  __shared__ int n_active;
  const int64_t d = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const bool ok = d < D;
  const float* col = s + (ok ? d : 0);
  double acc = 0.0;
  for (int64_t i = threadIdx.y; i < M; i += ESS_TY) acc += ok ? (double)col[i * D] : 0.0;
  const double mu = ess_block_sum(acc, red) / (double)M;
  acc = 0.0;
  for (int64_t i = threadIdx.y; i < M; i += ESS_TY) {
    const double c = ok ? (double)col[i * D] - mu : 0.0;
    acc += c * c;
  }
  const double var_plus = ess_block_sum(acc, red) / (double)M;        // np.var
  const double var = var_plus * (double)M / (double)(M - 1);
  double sum_rho = 0.0;
  bool active = ok;
  for (int64_t t = 0; t < M; ++t) {
    if (threadIdx.x == 0 && threadIdx.y == 0) n_active = 0;
    __syncthreads();
    acc = 0.0;
    if (active)
      for (int64_t i = threadIdx.y; i < M - t; i += ESS_TY)
        acc += ((double)col[i * D] - mu) * ((double)col[(i + t) * D] - mu);
    const double acov = ess_block_sum(acc, red) / (double)(M - t);
    if (active) {
      const double rho = 1.0 - (var - acov) / var_plus;
      if (rho < 0.0) active = false;          // diagnostics.py:36-38 (NaN compares false: kept)
      else sum_rho += rho;
    }
    if (active && threadIdx.y == 0) atomicAdd(&n_active, 1);
    __syncthreads();
    if (n_active == 0) break;
    __syncthreads();
  }
  if (ok && threadIdx.y == 0) ess[d] = (float)((double)M / (1.0 + 2.0 * sum_rho));
}

}  // namespace

extern "C" {

// samples: [M, D] (burn-in already dropped by the caller) -> ess: [D].  M >= 2.
int zsb_effective_sample_size_f32(const float* samples, int64_t M, int64_t D, float* ess,
                                  void* stream) {
  ZSB_REQUIRE(M >= 2 && D >= 0, "zsb_effective_sample_size_f32: need at least 2 samples");
  if (D == 0) return ZSB_OK;
  ZSB_REQUIRE(samples && ess, "zsb_effective_sample_size_f32: null pointer");
  const dim3 block(32, ESS_TY);
  ess_kernel<<<(unsigned)zsb_ceil_div(D, 32), block, 0, (cudaStream_t)stream>>>(samples, M, D, ess);
  return zsb_check_launch("effective_sample_size");
}

}  // extern "C"
