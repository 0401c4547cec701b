// K5 + K8 for BASELINE config 4: SGHMC over per-chain Bayesian-NN weights, ONE launch per step.
//
// Model (examples/bayesian_neural_nets/bnn_sgmcmc.py:19-35, 74-77), layer sizes [n_in, H, 1]:
//   w0 [chains, H, n_in+1] ~ N(0, exp(logstd0)),  w1 [chains, 1, H+1] ~ N(0, exp(logstd1))
//   h0 = [x, 1];  a1 = h0 w0^T / sqrt(n_in+1);  r1 = relu(a1);  h1 = [r1, 1]
//   y_mean = h1 w1^T / sqrt(H+1);   y ~ N(y_mean, exp(y_logstd))
//   log_joint = sum log p(w) + mean_b log p(y_b | x_b, w) * n_train
// Update (zhusuan/sgmcmc.py:326-371), per chain, fused:
//   [resample v ~ N(0, sqrt(lr))]                                    sgmcmc.py:327-336
//   2nd order: q1 = q + v/2;  g = grad log_joint(q1);  v = dh (dh v + lr g + xi);  q = q1 + v/2
//   1st order: g = grad log_joint(q);  v = (1-alpha) v + lr g + xi;  q += v
//   xi ~ N(0, sqrt(2 (alpha-beta) lr));  partial sums of v^2 for mean_k            sgmcmc.py:358
//
// One warp per chain.  Lane l owns hidden units l and l+32: their w0 rows, gradient accumulators
// and w1 entries stay in registers across the whole minibatch; the minibatch (x, y) is staged once
// per block in shared memory.  The gradient (tf.gradients in the reference, sgmcmc.py:96-98) is the
// hand-derived backward of the two-layer net.  HBM traffic = read+write of q and v only
// (16 * 601 B per chain-step at [10, 50, 1]); ~0.36 MFLOP per chain-step on the fp32 pipes.
#include "common.cuh"

namespace {

constexpr int MAX_IN1 = 16;   // n_in + 1 <= 16
constexpr int MAX_B = 512;    // minibatch rows staged in shared memory

struct BnnArgs {
  float* w0; float* w1; float* v0; float* v1;
  const float* x; const float* y;
  const float* logstd0; int64_t logstd0_n; const float* logstd1; int64_t logstd1_n;
  const float* noise0; const float* noise1; const float* rs0; const float* rs1;
  float* part0; float* part1;
  int64_t chains; int B, n_in, H;
  float y_logstd, n_train, lr, alpha, beta;
  int second_order, resample;
  uint64_t seed; uint32_t iter; int64_t row0;
};

__device__ __forceinline__ float noise_at(const float* injected, int64_t idx, uint64_t seed,
                                          uint32_t stream_id, uint32_t iter, int64_t row,
                                          int64_t col) {
  if (injected) return injected[idx];
  float z[4];
  philox_normal4(seed, stream_id, iter, (uint32_t)row, (uint32_t)(col >> 2), z);
  return z[col & 3];
}

constexpr int PB = 4;   // data points processed together (independent FMA / shuffle chains)

// Same numbers as noise_at(), but one Philox block (4 normals) is generated once and reused for
// the up-to-4 consecutive columns that share it.
struct NoiseCache {
  const float* injected; uint64_t seed; uint32_t stream_id, iter; int64_t row;
  int64_t blk; float z[4];
  __device__ __forceinline__ NoiseCache(const float* inj, uint64_t s, uint32_t st, uint32_t it,
                                        int64_t r)
      : injected(inj), seed(s), stream_id(st), iter(it), row(r), blk(-1) {}
  __device__ __forceinline__ float at(int64_t flat_idx, int64_t col) {
    if (injected) return injected[flat_idx];
    if ((col >> 2) != blk) {
      blk = col >> 2;
      philox_normal4(seed, stream_id, iter, (uint32_t)row, (uint32_t)blk, z);
    }
    const int w = (int)(col & 3);
    return w == 0 ? z[0] : w == 1 ? z[1] : w == 2 ? z[2] : z[3];
  }
};

template <int IN1>
__global__ void __launch_bounds__(256, 2) sghmc_bnn_kernel(BnnArgs a) {
  extern __shared__ float sh[];
  constexpr int in1 = IN1;
  const int H1 = a.H + 1;
  const int Bp = (a.B + PB - 1) / PB * PB;   // padded with zero-weight rows
  float* xs = sh;                       // [Bp][in1]  (bias column appended)
  float* ys = sh + Bp * in1;            // [Bp]
  float* wt = ys + Bp;                  // [Bp] 1 for real rows, 0 for padding
  __shared__ float red[32];
  for (int i = threadIdx.x; i < Bp * in1; i += blockDim.x) {
    const int b = i / in1, k = i % in1;
    xs[i] = (b < a.B) ? ((k < a.n_in) ? a.x[b * a.n_in + k] : 1.f) : 0.f;
  }
  for (int i = threadIdx.x; i < Bp; i += blockDim.x) {
    ys[i] = (i < a.B) ? a.y[i] : 0.f;
    wt[i] = (i < a.B) ? 1.f : 0.f;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const float inv_s0 = rsqrtf((float)in1), inv_s1 = rsqrtf((float)H1);
  const float prec_y = expf(-2.f * a.y_logstd);
  const float lik_scale = a.n_train / (float)a.B;
  const float sd_xi = sqrtf(mul(mul(2.f, sub(a.alpha, a.beta)), a.lr));
  const float sd_v = sqrtf(a.lr);
  const float dh = expf(mul(-0.5f, a.alpha)), oma = sub(1.f, a.alpha);
  float ksum0 = 0.f, ksum1 = 0.f;
  const int ls0_n = (int)a.logstd0_n, ls1_n = (int)a.logstd1_n;     // 32-bit index math only
  const bool ls0_full = ls0_n == a.H * in1, ls1_full = ls1_n == H1;

  for (int64_t c = (int64_t)blockIdx.x * 8 + wib; c < a.chains; c += (int64_t)gridDim.x * 8) {
    float* w0c = a.w0 + c * a.H * in1;
    float* v0c = a.v0 + c * a.H * in1;
    float* w1c = a.w1 + c * H1;
    float* v1c = a.v1 + c * H1;
    const int64_t grow = a.row0 + c;
    const int64_t c0off = c * a.H * in1;
    // ---- momentum resample (sgmcmc.py:327-336): written back so phase 3 can re-read it
    if (a.resample) {
      NoiseCache rc0(a.rs0, a.seed, ZSB_STREAM_SGMCMC_RESAMPLE, a.iter, grow);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int m = lane + 32 * u;
        if (m < a.H) {
#pragma unroll
          for (int k = 0; k < in1; ++k) {
            const int idx = m * in1 + k;
            v0c[idx] = mul(rc0.at(c0off + idx, idx), sd_v);
          }
          v1c[m] = mul(noise_at(a.rs1, c * H1 + m, a.seed + 1, ZSB_STREAM_SGMCMC_RESAMPLE,
                                a.iter, grow, m), sd_v);
        }
      }
      if (lane == 0)
        v1c[a.H] = mul(noise_at(a.rs1, c * H1 + a.H, a.seed + 1, ZSB_STREAM_SGMCMC_RESAMPLE,
                                a.iter, grow, a.H), sd_v);
      __syncwarp();
    }
    // ---- this lane's parameters (hidden units m = lane, lane + 32; lane 0 also the h1 bias)
    float W[2][IN1], G[2][IN1];
    float w1r[2], g1r[2];
    float w1b = 0.f, g1b = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int m = lane + 32 * u;
      const bool mv = m < a.H;
#pragma unroll
      for (int k = 0; k < in1; ++k) {
        const int idx = m * in1 + k;
        float w = mv ? w0c[idx] : 0.f;
        if (a.second_order && mv) w = add(w, mul(0.5f, v0c[idx]));   // q1 = q + v/2
        W[u][k] = w; G[u][k] = 0.f;
      }
      float w = mv ? w1c[m] : 0.f;
      if (a.second_order && mv) w = add(w, mul(0.5f, v1c[m]));
      w1r[u] = w; g1r[u] = 0.f;
    }
    if (lane == 0) {
      w1b = w1c[a.H];
      if (a.second_order) w1b = add(w1b, mul(0.5f, v1c[a.H]));
    }
    // ---- forward + backward over the minibatch, PB points at a time
    for (int b0 = 0; b0 < Bp; b0 += PB) {
      float a1[2][PB], part[PB];
#pragma unroll
      for (int p = 0; p < PB; ++p) part[p] = (lane == 0) ? w1b : 0.f;   // bias unit of h1
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int p = 0; p < PB; ++p) {
          const float* xb = xs + (b0 + p) * in1;
          float sacc = 0.f;
#pragma unroll
          for (int k = 0; k < in1; ++k) sacc = fmaf(W[u][k], xb[k], sacc);
          a1[u][p] = sacc * inv_s0;
          part[p] = fmaf(w1r[u], fmaxf(a1[u][p], 0.f), part[p]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int p = 0; p < PB; ++p) part[p] += __shfl_xor_sync(0xffffffffu, part[p], o);
      }
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        const float* xb = xs + (b0 + p) * in1;
        const float ym = part[p] * inv_s1;
        // d log_joint / d y_mean (zero for padding rows), then back through the output layer
        const float dout = prec_y * (ys[b0 + p] - ym) * lik_scale * wt[b0 + p] * inv_s1;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float r1 = fmaxf(a1[u][p], 0.f);
          g1r[u] = fmaf(dout, r1, g1r[u]);
          const float da1 = (a1[u][p] > 0.f) ? dout * w1r[u] * inv_s0 : 0.f;
#pragma unroll
          for (int k = 0; k < in1; ++k) G[u][k] = fmaf(da1, xb[k], G[u][k]);
        }
        if (lane == 0) g1b += dout;
      }
    }
    // ---- prior gradient, SGHMC update, write back
    NoiseCache nc0(a.noise0, a.seed, ZSB_STREAM_SGMCMC_NOISE, a.iter, grow);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int m = lane + 32 * u;
      if (m < a.H) {
#pragma unroll
        for (int k = 0; k < in1; ++k) {
          const int idx = m * in1 + k;
          const float ls = a.logstd0[ls0_full ? (int)idx : ((int)idx % ls0_n)];
          const float g = G[u][k] - expf(-2.f * ls) * W[u][k];
          const float xi = mul(nc0.at(c0off + idx, idx), sd_xi);
          const float vold = v0c[idx];
          float nv, nq;
          if (a.second_order) {
            nv = mul(dh, add(add(mul(dh, vold), mul(a.lr, g)), xi));
            nq = add(W[u][k], mul(0.5f, nv));
          } else {
            nv = add(add(mul(oma, vold), mul(a.lr, g)), xi);
            nq = add(W[u][k], nv);
          }
          w0c[idx] = nq; v0c[idx] = nv;
          ksum0 += nv * nv;
        }
        const float ls = a.logstd1[ls1_full ? m : (m % ls1_n)];
        const float g = g1r[u] - expf(-2.f * ls) * w1r[u];
        const float xi = mul(noise_at(a.noise1, c * H1 + m, a.seed + 1, ZSB_STREAM_SGMCMC_NOISE,
                                      a.iter, grow, m), sd_xi);
        const float vold = v1c[m];
        float nv, nq;
        if (a.second_order) {
          nv = mul(dh, add(add(mul(dh, vold), mul(a.lr, g)), xi));
          nq = add(w1r[u], mul(0.5f, nv));
        } else {
          nv = add(add(mul(oma, vold), mul(a.lr, g)), xi);
          nq = add(w1r[u], nv);
        }
        w1c[m] = nq; v1c[m] = nv;
        ksum1 += nv * nv;
      }
    }
    if (lane == 0) {
      const float ls = a.logstd1[ls1_full ? a.H : (a.H % ls1_n)];
      const float g = g1b - expf(-2.f * ls) * w1b;
      const float xi = mul(noise_at(a.noise1, c * H1 + a.H, a.seed + 1, ZSB_STREAM_SGMCMC_NOISE,
                                    a.iter, grow, a.H), sd_xi);
      const float vold = v1c[a.H];
      float nv, nq;
      if (a.second_order) {
        nv = mul(dh, add(add(mul(dh, vold), mul(a.lr, g)), xi));
        nq = add(w1b, mul(0.5f, nv));
      } else {
        nv = add(add(mul(oma, vold), mul(a.lr, g)), xi);
        nq = add(w1b, nv);
      }
      w1c[a.H] = nq; v1c[a.H] = nv;
      ksum1 += nv * nv;
    }
  }
  ksum0 = block_sum(ksum0, red);
  if (threadIdx.x == 0) a.part0[blockIdx.x] = ksum0;
  ksum1 = block_sum(ksum1, red);
  if (threadIdx.x == 0) a.part1[blockIdx.x] = ksum1;
}

__global__ void bnn_mean_k_kernel(const float* part0, const float* part1, int n_part, float n0,
                                  float n1, float* mean_k) {
  __shared__ float red[32];
  float s0 = 0.f, s1 = 0.f;
  for (int i = threadIdx.x; i < n_part; i += blockDim.x) { s0 += part0[i]; s1 += part1[i]; }
  s0 = block_sum(s0, red);
  s1 = block_sum(s1, red);
  if (threadIdx.x == 0) { mean_k[0] = s0 / n0; mean_k[1] = s1 / n1; }
}

}  // namespace

extern "C" {

// One fused SGHMC step for the two-layer BNN regression log-joint (bnn_sgmcmc.py:19-35, 74-91).
// w0/v0: [chains, H, n_in+1]; w1/v1: [chains, 1, H+1]; x: [B, n_in]; y: [B]; logstd0/1: prior
// log-stddevs broadcast modularly over one chain's weights; noise*/resample*: injected standard
// normals shaped like w0 / w1 (NULL -> in-kernel Philox); part: 2 * zsb_sgmcmc_parts() floats;
// mean_k: 2 floats out (sgmcmc.py:358, per latent).
int zsb_sgmcmc_sghmc_bnn_f32(float* w0, float* w1, float* v0, float* v1, const float* x,
                             const float* y, int B, int n_in, int H, const float* logstd0,
                             int64_t logstd0_n, const float* logstd1, int64_t logstd1_n,
                             float y_logstd, float n_train, float lr, float alpha, float beta,
                             int second_order, int resample, const float* noise0,
                             const float* noise1, const float* resample0, const float* resample1,
                             uint64_t seed, uint32_t iter, int64_t row0, float* part,
                             float* mean_k, int64_t chains, void* stream) {
  ZSB_REQUIRE(w0 && w1 && v0 && v1 && x && y && logstd0 && logstd1 && part && mean_k,
              "zsb_sgmcmc_sghmc_bnn_f32: null arg");
  ZSB_REQUIRE(chains > 0 && B > 0 && B <= MAX_B && n_in > 0 && n_in + 1 <= MAX_IN1 && H > 0 &&
                  H <= 64 && logstd0_n > 0 && logstd1_n > 0,
              "zsb_sgmcmc_sghmc_bnn_f32: need 0 < B <= 512, n_in <= 15, H <= 64");
  BnnArgs a;
  a.w0 = w0; a.w1 = w1; a.v0 = v0; a.v1 = v1; a.x = x; a.y = y;
  a.logstd0 = logstd0; a.logstd0_n = logstd0_n; a.logstd1 = logstd1; a.logstd1_n = logstd1_n;
  a.noise0 = noise0; a.noise1 = noise1; a.rs0 = resample0; a.rs1 = resample1;
  const int cap = ZSB_NUM_SMS * 8;
  a.part0 = part; a.part1 = part + cap;
  a.chains = chains; a.B = B; a.n_in = n_in; a.H = H;
  a.y_logstd = y_logstd; a.n_train = n_train; a.lr = lr; a.alpha = alpha; a.beta = beta;
  a.second_order = second_order; a.resample = resample;
  a.seed = seed; a.iter = iter; a.row0 = row0;
  int64_t blocks = zsb_ceil_div(chains, 8);
  if (blocks > cap) blocks = cap;
  const int Bp = (B + PB - 1) / PB * PB;
  const size_t smem = (size_t)(Bp * (n_in + 1) + 2 * Bp) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  switch (n_in + 1) {
#define ZSB_BNN_CASE(N) case N: sghmc_bnn_kernel<N><<<(unsigned)blocks, 256, smem, st>>>(a); break;
    ZSB_BNN_CASE(2) ZSB_BNN_CASE(3) ZSB_BNN_CASE(4) ZSB_BNN_CASE(5) ZSB_BNN_CASE(6)
    ZSB_BNN_CASE(7) ZSB_BNN_CASE(8) ZSB_BNN_CASE(9) ZSB_BNN_CASE(10) ZSB_BNN_CASE(11)
    ZSB_BNN_CASE(12) ZSB_BNN_CASE(13) ZSB_BNN_CASE(14) ZSB_BNN_CASE(15) ZSB_BNN_CASE(16)
#undef ZSB_BNN_CASE
  }
  int rc = zsb_check_launch("sgmcmc_sghmc_bnn");
  if (rc) return rc;
  bnn_mean_k_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(
      a.part0, a.part1, (int)blocks, (float)(chains * H * (n_in + 1)), (float)(chains * (H + 1)),
      mean_k);
  return zsb_check_launch("sgmcmc_sghmc_bnn_mean_k");
}

}  // extern "C"
