// Device samplers of the discrete / gamma-family registry distributions.
//
//   Categorical._sample  zhusuan/distributions/univariate.py:478-494  (tf.random.categorical)
//   Dirichlet._sample    zhusuan/distributions/multivariate.py:660-663 (Gamma(alpha, 1) normalised)
//   Gamma._sample        zhusuan/distributions/univariate.py (tf.random_gamma)
//
// TensorFlow's own streams cannot be reproduced without TensorFlow, so -- as for every other draw
// of this library -- the algorithm is fixed here and restated in oracle/samplers.py:
//   * Categorical: inverse CDF of softmax(logits) with ONE uniform per draw (injected, or word 0
//     of Philox block (0, draw index, iter, stream)); warp per draw, three passes over the row
//     (max, sum of exp, prefix scan), ties and round-off resolved towards the LAST category with
//     non-zero mass so an index is always valid;
//   * Gamma(alpha, 1): Marsaglia-Tsang (2000) squeeze with Philox normals / uniforms, attempt k of
//     element e using Philox block (k, e, iter, stream); alpha < 1 boosted by u^(1/alpha);
//     Dirichlet = one warp per row, gammas normalised by their row sum.  An injected-noise mode
//     takes the gamma variates themselves (the parity surface of the normalisation).
#include "common.cuh"

namespace {

#define ZSB_STREAM_CATEGORICAL 6u
#define ZSB_STREAM_GAMMA 7u

// one warp per draw; draw d = sample * rows + row reads logits row (row % logit_rows)
__global__ void __launch_bounds__(256) categorical_sample_kernel(
    const float* __restrict__ logits, int64_t logit_rows, int64_t rows, int C,
    const float* __restrict__ u_in, uint64_t seed, uint32_t iter, int32_t* __restrict__ out,
    int64_t n_draws, const uint32_t* __restrict__ epoch) {
  if (epoch) iter += *epoch;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int chunk = (C + 31) / 32;                 // contiguous categories per lane
  for (int64_t d = warp0; d < n_draws; d += nwarps) {
    const float* __restrict__ l = logits + ((d % rows) % logit_rows) * (int64_t)C;
    float u;
    if (u_in) {
      u = u_in[d];
    } else {
      const Philox4 r = philox4x32_10(0u, (uint32_t)d, iter ^ (uint32_t)((uint64_t)d >> 32),
                                      ZSB_STREAM_CATEGORICAL, (uint32_t)seed,
                                      (uint32_t)(seed >> 32));
      u = u32_to_uniform(r.x);
    }
    const int c0 = lane * chunk, c1 = min(C, c0 + chunk);
    float m = -INFINITY;
    for (int c = c0; c < c1; ++c) m = fmaxf(m, l[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = c0; c < c1; ++c) s += expf(l[c] - m);
    // inclusive prefix of the per-lane sums
    float pre = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const float total = __shfl_sync(0xffffffffu, pre, 31);
    const float target = u * total;
    // first lane whose inclusive prefix exceeds the target
    const unsigned hit = __ballot_sync(0xffffffffu, pre > target && s > 0.f);
    int pick;
    if (hit) {
      const int src = __ffs(hit) - 1;
      const float base = __shfl_sync(0xffffffffu, pre - s, src);
      pick = -1;
      if (lane == src) {
        float acc = base;
        int last = c0;
        for (int c = c0; c < c1; ++c) {
          const float e = expf(l[c] - m);
          if (e > 0.f) last = c;
          acc += e;
          if (acc > target) { pick = c; break; }
        }
        if (pick < 0) pick = last;               // round-off at the chunk's end
      }
      pick = __shfl_sync(0xffffffffu, pick, src);
    } else {
      // u * total rounded up to the total: last category with non-zero mass
      int last = -1;
      for (int c = c0; c < c1; ++c)
        if (expf(l[c] - m) > 0.f) last = c;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
      pick = last < 0 ? 0 : last;
    }
    if (lane == 0) out[d] = pick;
  }
}

// Marsaglia-Tsang for alpha >= 1 (d = alpha - 1/3, c = 1 / sqrt(9 d)); attempt k draws block
// (k, elem, iter, stream): words (x, y) -> one normal (Box-Muller, first output), word z -> u.
__device__ __forceinline__ float gamma_mt(float alpha, uint32_t elem_lo, uint32_t elem_hi,
                                          uint64_t seed, uint32_t iter) {
  const bool boost = alpha < 1.0f;
  const float a = boost ? alpha + 1.0f : alpha;
  const float dd = a - (1.0f / 3.0f);
  const float cc = 1.0f / sqrtf(9.0f * dd);
  float g = dd;                                    // fallback after 64 rejections (p < 1e-80)
  uint32_t w_boost = 0u;
  for (uint32_t k = 0; k < 64u; ++k) {
    const Philox4 r = philox4x32_10(k, elem_lo, iter ^ elem_hi, ZSB_STREAM_GAMMA, (uint32_t)seed,
                                    (uint32_t)(seed >> 32));
    if (k == 0) w_boost = r.w;
    float z0, z1;
    box_muller(r.x, r.y, z0, z1);
    const float v1 = 1.0f + cc * z0;
    if (v1 <= 0.f) continue;
    const float v = v1 * v1 * v1;
    const float u = u32_to_uniform_open(r.z);
    if (logf(u) < 0.5f * z0 * z0 + dd - dd * v + dd * logf(v)) { g = dd * v; break; }
  }
  if (boost) g *= powf(u32_to_uniform_open(w_boost), 1.0f / alpha);
  return g;
}

// One warp per output row: row r of [n_rows, C]; alpha row (r % alpha_rows); optional injected
// gammas [n_rows, C].  normalise = 1: Dirichlet (divide by the row sum); 0: plain Gamma(alpha, 1)
// scaled by 1 / beta[(r % beta_rows), c] when beta != nullptr.
__global__ void __launch_bounds__(256) gamma_rows_kernel(
    const float* __restrict__ alpha, int64_t alpha_rows, const float* __restrict__ beta,
    int64_t beta_rows, const float* __restrict__ gam_in, int64_t n_rows, int C, int normalise,
    uint64_t seed, uint32_t iter, float* __restrict__ out, const uint32_t* __restrict__ epoch) {
  if (epoch) iter += *epoch;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rows; r += nwarps) {
    const float* __restrict__ a = alpha + (r % alpha_rows) * (int64_t)C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) {
      const int64_t e = r * C + c;
      float g = gam_in ? gam_in[e]
                       : gamma_mt(a[c], (uint32_t)e, (uint32_t)((uint64_t)e >> 32), seed, iter);
      if (!normalise && beta) g = g / beta[(r % beta_rows) * (int64_t)C + c];
      out[e] = g;
      s += g;
    }
    if (normalise) {
      s = warp_sum(s);
      __syncwarp();
      for (int c = lane; c < C; c += 32) out[r * C + c] = out[r * C + c] / s;
    }
  }
}

#define ZSB_STREAM_COUNT 8u

// Inverse transform of a unimodal integer distribution with ONE uniform, visiting the outcomes in
// the order mode, mode+1, mode-1, mode+2, ... (any fixed enumeration of the support with running
// sums is a valid inverse CDF; this one needs O(std) steps).  pm = pmf(mode); up(k) = pmf(k+1) /
// pmf(k); down(k) = pmf(k-1) / pmf(k); support [0, kmax].
template <class Up, class Down>
__device__ __forceinline__ int invert_from_mode(float u, int mode, float pm, int kmax, int max_steps,
                                                Up up, Down down) {
  float s = pm;
  if (u < s) return mode;
  int lo = mode, hi = mode;
  float plo = pm, phi = pm;
  for (int it = 0; it < max_steps; ++it) {
    if (hi < kmax) {
      phi *= up(hi); ++hi; s += phi;
      if (u < s) return hi;
    }
    if (lo > 0) {
      plo *= down(lo); --lo; s += plo;
      if (u < s) return lo;
    }
    if (hi >= kmax && lo <= 0) break;
  }
  return hi;          // u within float round-off of 1: the far tail
}

// kind 0: Poisson(rate = a[i % a_n]);  kind 1: Binomial(n, p = sigmoid(a[i % a_n])).
__global__ void __launch_bounds__(256) count_sample_kernel(
    int kind, const float* __restrict__ a, int64_t a_n, int n_exp, const float* __restrict__ u_in,
    uint64_t seed, uint32_t iter, int32_t* __restrict__ out, int64_t n,
    const uint32_t* __restrict__ epoch) {
  if (epoch) iter += *epoch;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float u;
    if (u_in) {
      u = u_in[i];
    } else {
      const Philox4 r = philox4x32_10((uint32_t)(i >> 2), (uint32_t)((uint64_t)i >> 34), iter,
                                      ZSB_STREAM_COUNT, (uint32_t)seed, (uint32_t)(seed >> 32));
      const uint32_t w = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
      u = u32_to_uniform(w);
    }
    const float par = a[i % a_n];
    int k;
    if (kind == 0) {
      const float lam = par;
      if (!(lam > 0.f)) { out[i] = 0; continue; }
      const int m = (int)floorf(lam);
      const float pm = expf((float)m * logf(lam) - lam - lgammaf((float)m + 1.f));
      const int steps = (int)(12.f * sqrtf(lam) + 64.f);
      k = invert_from_mode(u, m, pm, 0x7fffffff, steps,
                           [=](int j) { return lam / (float)(j + 1); },
                           [=](int j) { return (float)j / lam; });
    } else {
      const float p = 1.f / (1.f + expf(-par));
      const float q = 1.f - p;
      if (p <= 0.f) { out[i] = 0; continue; }
      if (q <= 0.f) { out[i] = n_exp; continue; }
      int m = (int)floorf((float)(n_exp + 1) * p);
      m = m > n_exp ? n_exp : m;
      const float pm = expf(lgammaf((float)n_exp + 1.f) - lgammaf((float)m + 1.f) -
                            lgammaf((float)(n_exp - m) + 1.f) + (float)m * logf(p) +
                            (float)(n_exp - m) * log1pf(-p));
      const float odds = p / q;
      k = invert_from_mode(u, m, pm, n_exp, n_exp + 1,
                           [=](int j) { return (float)(n_exp - j) / (float)(j + 1) * odds; },
                           [=](int j) { return (float)j / (float)(n_exp - j + 1) / odds; });
    }
    out[i] = k;
  }
}

#define ZSB_STREAM_BASE 9u

// out[i] = uniform [0, 1) (kind 0) or standard normal (kind 1): the base noise of the
// reparameterised samplers; one Philox block per 4 consecutive elements.
__global__ void __launch_bounds__(256) base_noise_kernel(int kind, float* __restrict__ out,
                                                         int64_t n, uint64_t seed, uint32_t iter,
                                                         const uint32_t* __restrict__ epoch) {
  if (epoch) iter += *epoch;
  const int64_t n4 = (n + 3) / 4;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n4;
       b += (int64_t)gridDim.x * blockDim.x) {
    const Philox4 r = philox4x32_10((uint32_t)b, (uint32_t)((uint64_t)b >> 32), iter,
                                    ZSB_STREAM_BASE, (uint32_t)seed, (uint32_t)(seed >> 32));
    float v[4];
    if (kind == 0) {
      v[0] = u32_to_uniform(r.x); v[1] = u32_to_uniform(r.y);
      v[2] = u32_to_uniform(r.z); v[3] = u32_to_uniform(r.w);
    } else {
      box_muller(r.x, r.y, v[0], v[1]);
      box_muller(r.z, r.w, v[2], v[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * b + j < n) out[4 * b + j] = v[j];
  }
}

inline unsigned warp_grid(int64_t n_warps) {
  int64_t blocks = zsb_ceil_div(n_warps, 8);
  if (blocks > ZSB_NUM_SMS * 16) blocks = ZSB_NUM_SMS * 16;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace

extern "C" {

// Categorical._sample (univariate.py:478-494): out[s, r] for s < n_samples, r < rows; logits
// [logit_rows, n_categories] broadcast by r % logit_rows; u: optional injected uniforms
// [n_samples * rows] in [0, 1).
int zsb_sample_categorical_i32(const float* logits, int64_t logit_rows, int64_t rows,
                               int64_t n_categories, int64_t n_samples, const float* u,
                               uint64_t seed, uint32_t iter, int32_t* out, void* stream) {
  ZSB_REQUIRE(logits && out && logit_rows > 0 && rows > 0 && n_categories > 0 &&
                  n_categories < (1 << 30) && n_samples >= 0,
              "zsb_sample_categorical_i32: bad args");
  const int64_t n = n_samples * rows;
  if (n == 0) return ZSB_OK;
  categorical_sample_kernel<<<warp_grid(n), 256, 0, (cudaStream_t)stream>>>(
      logits, logit_rows, rows, (int)n_categories, u, seed, iter, out, n, zsb_epoch_ptr());
  return zsb_check_launch("sample_categorical");
}

// Dirichlet._sample (multivariate.py:660-663): out[n_rows, n_categories], alpha broadcast by
// row % alpha_rows; gammas: optional injected Gamma(alpha, 1) variates [n_rows, n_categories].
int zsb_sample_dirichlet_f32(const float* alpha, int64_t alpha_rows, int64_t n_rows,
                             int64_t n_categories, const float* gammas, uint64_t seed,
                             uint32_t iter, float* out, void* stream) {
  ZSB_REQUIRE(alpha && out && alpha_rows > 0 && n_rows >= 0 && n_categories > 0 &&
                  n_categories < (1 << 30),
              "zsb_sample_dirichlet_f32: bad args");
  if (n_rows == 0) return ZSB_OK;
  gamma_rows_kernel<<<warp_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(
      alpha, alpha_rows, nullptr, 1, gammas, n_rows, (int)n_categories, 1, seed, iter, out,
      zsb_epoch_ptr());
  return zsb_check_launch("sample_dirichlet");
}

// Gamma._sample: out[n_rows, row_len] = Gamma(alpha, 1) / beta (beta may be NULL = 1).
int zsb_sample_gamma_f32(const float* alpha, int64_t alpha_rows, const float* beta,
                         int64_t beta_rows, int64_t n_rows, int64_t row_len, uint64_t seed,
                         uint32_t iter, float* out, void* stream) {
  ZSB_REQUIRE(alpha && out && alpha_rows > 0 && n_rows >= 0 && row_len > 0 &&
                  row_len < (1 << 30) && (!beta || beta_rows > 0),
              "zsb_sample_gamma_f32: bad args");
  if (n_rows == 0) return ZSB_OK;
  gamma_rows_kernel<<<warp_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(
      alpha, alpha_rows, beta, beta ? beta_rows : 1, nullptr, n_rows, (int)row_len, 0, seed, iter,
      out, zsb_epoch_ptr());
  return zsb_check_launch("sample_gamma");
}

// Base noise of the reparameterised samplers whose transform is composed on the host side
// (Uniform / Laplace / FoldNormal / (Bin)Concrete / MatrixVariateNormal: tf.random_uniform /
// tf.random_normal in univariate.py:306-317, 622-640, 1246-1265, 1363-1379): kind 0 = U[0, 1),
// kind 1 = N(0, 1); element i = word i % 4 of Philox block (i / 4, 0, iter, 9).
int zsb_sample_base_noise_f32(int kind, float* out, int64_t n, uint64_t seed, uint32_t iter,
                              void* stream) {
  ZSB_REQUIRE((kind == 0 || kind == 1) && out && n >= 0, "zsb_sample_base_noise_f32: bad args");
  if (n == 0) return ZSB_OK;
  int64_t blocks = zsb_ceil_div((n + 3) / 4, 256);
  if (blocks > ZSB_NUM_SMS * 16) blocks = ZSB_NUM_SMS * 16;
  base_noise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(kind, out, n, seed, iter,
                                                                        zsb_epoch_ptr());
  return zsb_check_launch("sample_base_noise");
}

// Poisson._sample (univariate.py:915-920, tf.random_poisson) / Binomial._sample (univariate.py:
// 1025-1045: n_experiments categorical draws summed): out[i] for i < n, parameter broadcast by
// i % param_n; kind 0 = Poisson(rate), 1 = Binomial(n_experiments, sigmoid(logits)); one uniform per
// draw (injected `u` [n] or Philox), inverse transform from the mode.
int zsb_sample_count_i32(int kind, const float* param, int64_t param_n, int64_t n_experiments,
                         const float* u, uint64_t seed, uint32_t iter, int32_t* out, int64_t n,
                         void* stream) {
  ZSB_REQUIRE((kind == 0 || kind == 1) && param && out && param_n > 0 && n >= 0 &&
                  (kind == 0 || (n_experiments > 0 && n_experiments < (1LL << 30))),
              "zsb_sample_count_i32: bad args");
  if (n == 0) return ZSB_OK;
  int64_t blocks = zsb_ceil_div(n, 256);
  if (blocks > ZSB_NUM_SMS * 16) blocks = ZSB_NUM_SMS * 16;
  count_sample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      kind, param, param_n, (int)n_experiments, u, seed, iter, out, n, zsb_epoch_ptr());
  return zsb_check_launch("sample_count");
}

}  // extern "C"
