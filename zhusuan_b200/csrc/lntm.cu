// K8 for config 5: the E-step log-joint of the Logistic-Normal Topic Model and its gradient in one
// fused, sparsity-aware kernel.
//
// Reference (examples/topic_models/lntm_mcem.py:33-48, 97-99; UnnormalizedMultinomial._log_prob,
// zhusuan/distributions/multivariate.py:435-443 with normalize_logits=False):
//     theta = softmax(eta)                         eta  [chains, docs, K]
//     phi   = softmax(beta)                        beta [K, V]
//     log p = sum_k Normal(eta_k; mean_k, exp(logstd_k)).log_prob            (cond_log_prob('eta'))
//           + sum_v x[d, v] * log(theta @ phi)[v]                            (cond_log_prob('x'))
// TensorFlow materialises doc_word = theta @ phi as a [chains * docs, V] matrix (at config 5:
// 1024 x 10 000 x 8192 floats = 335 TB -- it cannot run) and differentiates through it.  x is a bag of
// words: a document touches a few hundred of the V columns, so here only those are ever formed:
//     S_j   = sum_k theta_k * phi[k, w_j]          for the document's words w_j (CSR)
//     log p += c_j * log S_j
//     dtheta_k += (c_j / S_j) * phi[k, w_j]
//     deta  = theta * (dtheta - <theta, dtheta>) - (eta - mean) * exp(-2 logstd)
// = 4 K flops per (chain, word occurrence) instead of 4 K V per chain-document; the [rows, V] matrix
// never exists anywhere.  phi is kept transposed ([V, K], 4 MB at config 5: L2 resident) so that a
// word's topic vector is one contiguous 512-byte row.
//
// Mapping: one block = one document x 64 chains; a quad of threads owns a chain (each thread K/4
// topics, as float4 groups interleaved across the quad: conflict-free LDS.128 of the phi tile, the
// 8 chains of a warp read the same words by broadcast); 32 words of the document at a time are
// staged in shared memory.  Bound by the fp32 FMA pipe (2 FMAs + 1/16 LDS.128 per topic-word).
#include "common.cuh"

namespace {

constexpr int LN_CHAINS = 64;        // chains per block
constexpr int LN_WORDS = 32;         // words staged per round

// phi_t[v, k] = softmax_v(beta[k, :])[v]: one block per topic row, two passes
__global__ void __launch_bounds__(256) lntm_phi_t_kernel(const float* __restrict__ beta, int K,
                                                         int64_t V, float* __restrict__ phi_t) {
  __shared__ float red[32];
  const int k = blockIdx.x;
  const float* __restrict__ b = beta + (int64_t)k * V;
  float m = -INFINITY;
  for (int64_t v = threadIdx.x; v < V; v += blockDim.x) m = fmaxf(m, b[v]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int64_t v = threadIdx.x; v < V; v += blockDim.x) s += expf(b[v] - m);
  s = block_sum(s, red);
  const float inv = 1.f / s;
  for (int64_t v = threadIdx.x; v < V; v += blockDim.x) phi_t[v * K + k] = expf(b[v] - m) * inv;
}

template <int G>                     // G = float4 groups per thread = K / 16
__global__ void __launch_bounds__(256, 2) lntm_logjoint_kernel(
    const float* __restrict__ eta, const float* __restrict__ eta_mean,
    const float* __restrict__ eta_logstd, const float* __restrict__ phi_t,
    const int64_t* __restrict__ doc_ptr, const int32_t* __restrict__ word_idx,
    const float* __restrict__ word_cnt, float* __restrict__ lp_out, float* __restrict__ grad_out,
    int64_t chains, int64_t docs) {
  constexpr int K = 16 * G;
  __shared__ float4 tile[LN_WORDS][K / 4];       // phi_t rows of the staged words
  __shared__ float cnt[LN_WORDS];
  const int q = threadIdx.x & 3;                 // thread inside the chain's quad
  const int64_t d = blockIdx.x;
  const int64_t c = (int64_t)blockIdx.y * LN_CHAINS + (threadIdx.x >> 2);
  const bool live = c < chains;
  const float* __restrict__ e = eta + ((live ? c : 0) * docs + d) * K;

  // this thread's topics: float4 groups g*4 + q, g < G
  float4 th[G], dth[G];
  float mx = -INFINITY;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    th[g] = *reinterpret_cast<const float4*>(e + 4 * (g * 4 + q));
    dth[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    mx = fmaxf(mx, fmaxf(fmaxf(th[g].x, th[g].y), fmaxf(th[g].z, th[g].w)));
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  // prior (Normal, group_ndims = 1) on eta before it is overwritten by theta
  float lp = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int k0 = 4 * (g * 4 + q);
    const float4 mu = *reinterpret_cast<const float4*>(eta_mean + k0);
    const float4 ls = *reinterpret_cast<const float4*>(eta_logstd + k0);
    const float ev[4] = {th[g].x, th[g].y, th[g].z, th[g].w};
    const float mv[4] = {mu.x, mu.y, mu.z, mu.w};
    const float lv[4] = {ls.x, ls.y, ls.z, ls.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float prec = expf(-2.f * lv[i]);
      const float dd = ev[i] - mv[i];
      lp += -0.9189385332046727f - lv[i] - 0.5f * prec * dd * dd;     // univariate.py:174-181
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    th[g].x = expf(th[g].x - mx); th[g].y = expf(th[g].y - mx);
    th[g].z = expf(th[g].z - mx); th[g].w = expf(th[g].w - mx);
    sum += (th[g].x + th[g].y) + (th[g].z + th[g].w);
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  const float inv = 1.f / sum;
#pragma unroll
  for (int g = 0; g < G; ++g) { th[g].x *= inv; th[g].y *= inv; th[g].z *= inv; th[g].w *= inv; }

  const int64_t w0 = doc_ptr[d], w1 = doc_ptr[d + 1];
  for (int64_t wb = w0; wb < w1; wb += LN_WORDS) {
    const int nw = (int)((w1 - wb < LN_WORDS) ? (w1 - wb) : LN_WORDS);
    __syncthreads();                                   // previous round consumed
    for (int i = threadIdx.x; i < nw * (K / 4); i += blockDim.x) {
      const int w = i / (K / 4), kk = i % (K / 4);
      tile[w][kk] = *reinterpret_cast<const float4*>(
          phi_t + (int64_t)word_idx[wb + w] * K + 4 * kk);
    }
    if ((int)threadIdx.x < nw) cnt[threadIdx.x] = word_cnt[wb + threadIdx.x];
    __syncthreads();
    for (int w = 0; w < nw; ++w) {
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 ph = tile[w][g * 4 + q];
        s = fmaf(th[g].x, ph.x, s); s = fmaf(th[g].y, ph.y, s);
        s = fmaf(th[g].z, ph.z, s); s = fmaf(th[g].w, ph.w, s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      const float cw = cnt[w];
      const float r = cw / s;                          // multivariate.py:435-443 differentiated
      if (q == 0) lp += cw * logf(s);
      if (grad_out == nullptr) continue;               // value only (MH test): skip the axpy
#pragma unroll
      for (int g = 0; g < G; ++g) {                    // (re-read: keeps the kernel at 2 blocks/SM)
        const float4 ph = tile[w][g * 4 + q];
        dth[g].x = fmaf(r, ph.x, dth[g].x); dth[g].y = fmaf(r, ph.y, dth[g].y);
        dth[g].z = fmaf(r, ph.z, dth[g].z); dth[g].w = fmaf(r, ph.w, dth[g].w);
      }
    }
  }
  // softmax backward + prior gradient
  float dot = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g)
    dot += (th[g].x * dth[g].x + th[g].y * dth[g].y) + (th[g].z * dth[g].z + th[g].w * dth[g].w);
  dot += __shfl_xor_sync(0xffffffffu, dot, 1);
  dot += __shfl_xor_sync(0xffffffffu, dot, 2);
  lp += __shfl_xor_sync(0xffffffffu, lp, 1);
  lp += __shfl_xor_sync(0xffffffffu, lp, 2);
  if (!live) return;
  if (grad_out) {
    float* __restrict__ go = grad_out + (c * docs + d) * K;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int k0 = 4 * (g * 4 + q);
      const float4 ev = *reinterpret_cast<const float4*>(e + k0);       // prior gradient
      const float4 mu = *reinterpret_cast<const float4*>(eta_mean + k0);
      const float4 ls = *reinterpret_cast<const float4*>(eta_logstd + k0);
      float4 o;
      o.x = fmaf(th[g].x, dth[g].x - dot, -expf(-2.f * ls.x) * (ev.x - mu.x));
      o.y = fmaf(th[g].y, dth[g].y - dot, -expf(-2.f * ls.y) * (ev.y - mu.y));
      o.z = fmaf(th[g].z, dth[g].z - dot, -expf(-2.f * ls.z) * (ev.z - mu.z));
      o.w = fmaf(th[g].w, dth[g].w - dot, -expf(-2.f * ls.w) * (ev.w - mu.w));
      *reinterpret_cast<float4*>(go + 4 * (g * 4 + q)) = o;
    }
  }
  if (lp_out && q == 0) lp_out[c * docs + d] = lp;
}

}  // namespace

extern "C" {

// phi_t [V, K] = softmax(beta [K, V], axis = vocabulary) transposed (lntm_mcem.py:41).
int zsb_lntm_phi_t_f32(const float* beta, int64_t n_topics, int64_t n_vocab, float* phi_t,
                       void* stream) {
  ZSB_REQUIRE(beta && phi_t && n_topics > 0 && n_vocab > 0, "zsb_lntm_phi_t_f32: bad args");
  lntm_phi_t_kernel<<<(unsigned)n_topics, 256, 0, (cudaStream_t)stream>>>(beta, (int)n_topics,
                                                                          n_vocab, phi_t);
  return zsb_check_launch("lntm_phi_t");
}

// E-step log-joint of the LNTM and its gradient w.r.t. eta (lntm_mcem.py:33-48, 97-99).
//   eta [chains, docs, n_topics]; eta_mean / eta_logstd [n_topics]; phi_t [n_vocab, n_topics];
//   corpus in CSR: doc_ptr [docs + 1] (int64), word_idx [nnz] (int32), word_cnt [nnz] (float);
//   lp_out [chains, docs] and / or grad_out like eta.  n_topics in {16, 32, 64, 128}.
int zsb_lntm_logjoint_f32(const float* eta, const float* eta_mean, const float* eta_logstd,
                          const float* phi_t, const int64_t* doc_ptr, const int32_t* word_idx,
                          const float* word_cnt, float* lp_out, float* grad_out, int64_t chains,
                          int64_t docs, int64_t n_topics, void* stream) {
  ZSB_REQUIRE(eta && eta_mean && eta_logstd && phi_t && doc_ptr && (lp_out || grad_out) &&
                  chains > 0 && docs > 0 && docs < (1LL << 31),
              "zsb_lntm_logjoint_f32: bad args");
  const dim3 grid((unsigned)docs, (unsigned)zsb_ceil_div(chains, LN_CHAINS));
  ZSB_REQUIRE(grid.y < 65536, "zsb_lntm_logjoint_f32: too many chains");
  cudaStream_t st = (cudaStream_t)stream;
#define ZSB_LN(G)                                                                              \
  lntm_logjoint_kernel<G><<<grid, 256, 0, st>>>(eta, eta_mean, eta_logstd, phi_t, doc_ptr,     \
                                                word_idx, word_cnt, lp_out, grad_out, chains, docs)
  switch (n_topics) {
    case 16: ZSB_LN(1); break;
    case 32: ZSB_LN(2); break;
    case 64: ZSB_LN(4); break;
    case 128: ZSB_LN(8); break;
    default:
      zsb_set_error("zsb_lntm_logjoint_f32: n_topics must be 16, 32, 64 or 128");
      return ZSB_ERR_INVALID;
  }
#undef ZSB_LN
  return zsb_check_launch("lntm_logjoint");
}

}  // extern "C"
