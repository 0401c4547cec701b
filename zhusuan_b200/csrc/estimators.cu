// K6b: score-function / self-normalised estimators on the same [outer, K, inner] log-weight tile
// as the log_mean_exp kernel (reduce.cu).  None of these has a backward pass: the reference wraps
// every one of them in tf.stop_gradient.
//
//   vimco signal      zhusuan/variational/monte_carlo.py:194-223
//       signal[k] = LME_j(l_j) - LME_j(l_j with entry k replaced by mean_{j != k} l_j)
//     The reference materialises an [.., K, K] tensor (tile + matrix_diag).  Here: one thread per
//     column, three passes over the K values of the column (L1/L2 resident after the first), O(K).
//     Leave-one-out sums use  sum_{j != k} e^{l_j - m} = S - e^{l_k - m};  only for k = argmax can
//     that cancel, so that single k is summed explicitly against its own maximum, exactly like
//     the reference's per-row log_mean_exp.
//   normalized weights zhusuan/variational/inclusive_kl.py:139-143
//       w~[k] = e^{l_k - max} / sum_j e^{l_j - max}
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) vimco_signal_kernel(const float* __restrict__ x,
                                                           float* __restrict__ signal,
                                                           float* __restrict__ lme_out,
                                                           int64_t outer, int64_t K,
                                                           int64_t inner) {
  const int64_t ncols = outer * inner;
  const float fK = (float)K, fK1 = (float)(K - 1);
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncols;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = c / inner, i = c % inner;
    const float* p = x + o * K * inner + i;
    float* s_out = signal + o * K * inner + i;
    // pass 1: sum, max (first argmax), max of the rest
    float sum = 0.f, m1 = -INFINITY, m2 = -INFINITY;
    int64_t top = 0;
    for (int64_t k = 0; k < K; ++k) {
      const float v = p[k * inner];
      sum += v;
      if (v > m1) { m2 = m1; m1 = v; top = k; }
      else m2 = fmaxf(m2, v);
    }
    const float sub_top = (sum - m1) / fK1;               // mean of the others, for k = top
    const float M_top = fmaxf(m2, sub_top);
    // pass 2: S = sum_j e^{l_j - m1};  S_top = sum_{j != top} e^{l_j - M_top}
    float S = 0.f, S_top = 0.f;
    for (int64_t k = 0; k < K; ++k) {
      const float v = p[k * inner];
      S += expf(v - m1);
      if (k != top) S_top += expf(v - M_top);
    }
    const float lme = logf(S / fK) + m1;
    if (lme_out) lme_out[c] = lme;
    // pass 3: control variate per k
    for (int64_t k = 0; k < K; ++k) {
      const float v = p[k * inner];
      const float sub = (sum - v) / fK1;
      float cv;
      if (k == top) {
        cv = logf((S_top + expf(sub - M_top)) / fK) + M_top;
      } else {
        // max_{j != k} l_j = m1 >= mean_{j != k} l_j
        cv = logf((S - expf(v - m1) + expf(sub - m1)) / fK) + m1;
      }
      s_out[k * inner] = lme - cv;
    }
  }
}

__global__ void __launch_bounds__(256) normalized_weights_kernel(const float* __restrict__ x,
                                                                 float* __restrict__ w,
                                                                 int64_t outer, int64_t K,
                                                                 int64_t inner) {
  const int64_t ncols = outer * inner;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncols;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = c / inner, i = c % inner;
    const float* p = x + o * K * inner + i;
    float* q = w + o * K * inner + i;
    float m = -INFINITY;
    for (int64_t k = 0; k < K; ++k) m = fmaxf(m, p[k * inner]);
    float S = 0.f;
    for (int64_t k = 0; k < K; ++k) S += expf(p[k * inner] - m);
    for (int64_t k = 0; k < K; ++k) q[k * inner] = expf(p[k * inner] - m) / S;
  }
}

unsigned col_grid(int64_t ncols) {
  int64_t blocks = zsb_ceil_div(ncols, 256);
  if (blocks > ZSB_NUM_SMS * 16) blocks = ZSB_NUM_SMS * 16;
  return (unsigned)blocks;
}

}  // namespace

extern "C" {

// x: [outer, K, inner] -> signal: same shape; lme (optional): [outer, inner].  K >= 2.
int zsb_vimco_signal_f32(const float* x, float* signal, float* lme, int64_t outer, int64_t K,
                         int64_t inner, void* stream) {
  ZSB_REQUIRE(outer >= 0 && inner >= 0, "zsb_vimco_signal_f32: bad sizes");
  ZSB_REQUIRE(K >= 2, "zsb_vimco_signal_f32: size along the sample axis must be at least 2");
  if (outer * inner == 0) return ZSB_OK;
  ZSB_REQUIRE(x && signal, "zsb_vimco_signal_f32: null pointer");
  vimco_signal_kernel<<<col_grid(outer * inner), 256, 0, (cudaStream_t)stream>>>(
      x, signal, lme, outer, K, inner);
  return zsb_check_launch("vimco_signal");
}

// x: [outer, K, inner] -> w: same shape, softmax over K in the reference's operation order.
int zsb_normalized_weights_f32(const float* x, float* w, int64_t outer, int64_t K, int64_t inner,
                               void* stream) {
  ZSB_REQUIRE(outer >= 0 && K > 0 && inner >= 0, "zsb_normalized_weights_f32: bad sizes");
  if (outer * inner == 0) return ZSB_OK;
  ZSB_REQUIRE(x && w, "zsb_normalized_weights_f32: null pointer");
  normalized_weights_kernel<<<col_grid(outer * inner), 256, 0, (cudaStream_t)stream>>>(
      x, w, outer, K, inner);
  return zsb_check_launch("normalized_weights");
}

}  // extern "C"
