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
// One warp per chain (persistent grid).  Lane l owns hidden units l and l+32: their w0 rows, gradient accumulators
// and w1 entries stay in registers across the whole minibatch; the minibatch (x, y) is staged once
// per block in shared memory, rows padded to a multiple of 4 floats so a row is read with 128-bit
// broadcast loads.  Round 1 spent 17.6 k warp instructions per chain and step, half of them outside
// the minibatch loop: every lane regenerated (under divergence) each Philox block it touched, and
// each weight paid an integer modulo + expf for its prior precision.  Now the warp generates each
// noise block once into shared memory and the prior precisions are tabulated once per block.  The gradient (tf.gradients in the reference, sgmcmc.py:96-98) is the
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

constexpr int PB = 4;   // data points processed together (independent FMA / shuffle chains)

// `n` standard normals of (row, elements 0..n-1) of a Philox stream into a per-warp shared buffer
// (n rounded up to 4 floats), or the injected ones.  Element e is component e & 3 of block e >> 2
// -- the same numbers the element-wise kernels draw (sgmcmc.cu), so the fused and the generic path
// agree draw for draw.  Each block is generated ONCE per warp (round 1 generated a block in every
// lane that touched it, under divergence: 22 Philox evaluations per lane and step instead of 5).
__device__ __forceinline__ void warp_fill_normals(float* buf, int n, const float* injected,
                                                  int64_t flat0, uint64_t seed, uint32_t stream,
                                                  uint32_t iter, int64_t row, int lane) {
  if (injected) {
    for (int i = lane; i < n; i += 32) buf[i] = injected[flat0 + i];
  } else {
    const int nblk = (n + 3) >> 2;
    for (int b = lane; b < nblk; b += 32) {
      float z[4];
      philox_normal4(seed, stream, iter, (uint32_t)row, (uint32_t)b, z);
      reinterpret_cast<float4*>(buf)[b] = make_float4(z[0], z[1], z[2], z[3]);
    }
  }
  __syncwarp();
}

template <int IN1>
__global__ void __launch_bounds__(256, 2) sghmc_bnn_kernel(BnnArgs a) {
  extern __shared__ float4 sh4[];
  constexpr int in1 = IN1;
  constexpr int X4 = (IN1 + 3) / 4;     // a staged minibatch row = X4 float4 (bias column, 0 pad)
  constexpr int XP = 4 * X4;
  const int H1 = a.H + 1;
  const int n0 = a.H * in1;                       // weights of layer 0 per chain
  const int n0p = (n0 + 3) & ~3, n1p = (H1 + 3) & ~3;
  const int Bp = (a.B + PB - 1) / PB * PB;        // padded with zero-weight rows
  float* xs = reinterpret_cast<float*>(sh4);                 // [Bp + PB][XP]
  float2* yc = reinterpret_cast<float2*>(xs + (Bp + PB) * XP);      // [Bp] {y, dout coefficient or 0}
  float* pr0 = reinterpret_cast<float*>(yc + Bp);            // [n0p] prior precision exp(-2 ls)
  float* pr1 = pr0 + n0p;                                    // [n1p]
  float* nzb = pr1 + n1p;                                    // [8 warps][2][n0p + n1p] staging
  __shared__ float red[32];
  const float inv_s0 = rsqrtf((float)in1), inv_s1 = rsqrtf((float)H1);
  {
    const float prec_y = expf(-2.f * a.y_logstd);
    const float lik_scale = a.n_train / (float)a.B;
    // d log_joint / d (h1 . w1) = prec_y (y - y_mean) * (n_train / B) / sqrt(H + 1)
    const float cf = prec_y * lik_scale * inv_s1;
    for (int i = threadIdx.x; i < (Bp + PB) * XP; i += blockDim.x) {
      const int b = i / XP, k = i % XP;
      xs[i] = (b < a.B) ? ((k < a.n_in) ? a.x[b * a.n_in + k] : (k == a.n_in ? 1.f : 0.f)) : 0.f;
    }
    for (int i = threadIdx.x; i < Bp; i += blockDim.x)
      yc[i] = (i < a.B) ? make_float2(a.y[i], cf) : make_float2(0.f, 0.f);
    const int ls0_n = (int)a.logstd0_n, ls1_n = (int)a.logstd1_n;
    for (int i = threadIdx.x; i < n0p; i += blockDim.x)
      pr0[i] = (i < n0) ? expf(-2.f * a.logstd0[i % ls0_n]) : 0.f;
    for (int i = threadIdx.x; i < n1p; i += blockDim.x)
      pr1[i] = (i < H1) ? expf(-2.f * a.logstd1[i % ls1_n]) : 0.f;
  }
  __syncthreads();

  // warp index through a shuffle: provably warp-uniform, so the chain loop's exit is uniform and the
  // shuffles inside it compile to plain SHFL instead of WARPSYNC.COLLECTIVE sequences
  const int lane = threadIdx.x & 31, nwb = blockDim.x >> 5;
  const int wib = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  // per-warp staging: A = weights in / noise / new weights out, B = momentum in / new momentum out.
  // Every global access of a chain's state is a coalesced, independent copy through these buffers
  // (round 1 read v element by element between dependent stores: 22 exposed global latencies).
  float* A0 = nzb + wib * 2 * (n0p + n1p);
  float* A1 = A0 + n0p;
  float* B0 = A1 + n1p;
  float* B1 = B0 + n0p;
  const float sd_xi = sqrtf(mul(mul(2.f, sub(a.alpha, a.beta)), a.lr));
  const float sd_v = sqrtf(a.lr);
  const float dh = expf(mul(-0.5f, a.alpha)), oma = sub(1.f, a.alpha);
  float ksum0 = 0.f, ksum1 = 0.f;

  for (int64_t c = (int64_t)blockIdx.x * nwb + wib; c < a.chains;
       c += (int64_t)gridDim.x * nwb) {
    float* w0c = a.w0 + c * n0;
    float* v0c = a.v0 + c * n0;
    float* w1c = a.w1 + c * H1;
    float* v1c = a.v1 + c * H1;
    const int64_t grow = a.row0 + c;
    for (int i = lane; i < n0; i += 32) A0[i] = w0c[i];
    for (int i = lane; i < H1; i += 32) A1[i] = w1c[i];
    if (a.resample) {   // momentum resample v ~ N(0, sqrt(lr)) (sgmcmc.py:327-336)
      warp_fill_normals(B0, n0, a.rs0, c * n0, a.seed, ZSB_STREAM_SGMCMC_RESAMPLE, a.iter, grow,
                        lane);
      warp_fill_normals(B1, H1, a.rs1, c * H1, a.seed + 1, ZSB_STREAM_SGMCMC_RESAMPLE, a.iter,
                        grow, lane);
      for (int i = lane; i < n0; i += 32) B0[i] = mul(B0[i], sd_v);
      for (int i = lane; i < H1; i += 32) B1[i] = mul(B1[i], sd_v);
    } else {
      for (int i = lane; i < n0; i += 32) B0[i] = v0c[i];
      for (int i = lane; i < H1; i += 32) B1[i] = v1c[i];
    }
    __syncwarp();
    // ---- this lane's parameters (hidden units m = lane, lane + 32; lane 0 also the h1 bias)
    float W[2][IN1], G[2][IN1];
    float w1r[2], w1s[2], g1r[2];
    float w1b = 0.f, g1b = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int m = lane + 32 * u;
      const bool mv = m < a.H;
#pragma unroll
      for (int k = 0; k < in1; ++k) {
        const int idx = m * in1 + k;
        float w = mv ? A0[idx] : 0.f;
        if (a.second_order && mv) w = add(w, mul(0.5f, B0[idx]));   // q1 = q + v/2
        W[u][k] = w; G[u][k] = 0.f;
      }
      float w = mv ? A1[m] : 0.f;
      if (a.second_order && mv) w = add(w, mul(0.5f, B1[m]));
      w1r[u] = w; w1s[u] = w * inv_s0; g1r[u] = 0.f;
    }
    if (lane == 0) {
      w1b = A1[a.H];
      if (a.second_order) w1b = add(w1b, mul(0.5f, B1[a.H]));
    }
    __syncwarp();                          // A is overwritten with the update noise below
    // ---- forward + backward over the minibatch, PB points at a time.  A staged row is read as
    // X4 128-bit shared loads (every lane the same address: broadcast).  With s = W x (the
    // pre-activation before the 1/sqrt(n_in+1) scale), r = max(s, 0):
    //   h1 . w1 = sum_m (w1_m / sqrt(n_in+1)) r_m + bias;   dout = cf_b (y_b - (h1 . w1)/sqrt(H+1))
    //   d/dw1_m += dout r_m / sqrt(n_in+1) (scale applied once after the loop)
    //   d/dW_mk += [s_m > 0] dout (w1_m / sqrt(n_in+1)) x_k
    const float bias_l = (lane == 0) ? w1b : 0.f;                        // bias unit of h1
    auto load_row = [&](int b, float* xv) {
      const float4* xr = reinterpret_cast<const float4*>(xs + b * XP);
#pragma unroll
      for (int j = 0; j < X4; ++j) {
        const float4 t = xr[j];
        xv[4 * j] = t.x; xv[4 * j + 1] = t.y; xv[4 * j + 2] = t.z; xv[4 * j + 3] = t.w;
      }
    };
    auto forward1 = [&](int b, float& s0, float& s1, float& pt) {
      float xv[XP];
      load_row(b, xv);
      float sacc0 = 0.f, sacc1 = 0.f;
#pragma unroll
      for (int k = 0; k < in1; ++k) {
        sacc0 = fmaf(W[0][k], xv[k], sacc0);
        sacc1 = fmaf(W[1][k], xv[k], sacc1);
      }
      s0 = sacc0; s1 = sacc1;
      pt = fmaf(w1s[1], fmaxf(sacc1, 0.f), fmaf(w1s[0], fmaxf(sacc0, 0.f), bias_l));
    };
    // software pipeline: the forward pass of block i+1 (one data point per slot) is issued between
    // the butterfly rounds of block i's h1 . w1 reduction -- five dependent shuffles on which the
    // warp would otherwise idle.  xs holds PB zero rows past Bp, so the look-ahead needs no branch.
    static_assert(PB == 4, "the interleave below is written for 4 points per block");
    float sa[2][PB], part[PB];
#pragma unroll
    for (int p = 0; p < PB; ++p) forward1(p, sa[0][p], sa[1][p], part[p]);
    for (int b0 = 0; b0 < Bp; b0 += PB) {
      float sn[2][PB], pn[PB];
#define ZSB_BFLY(o)                                                                   \
  _Pragma("unroll") for (int p = 0; p < PB; ++p)                                      \
      part[p] += __shfl_xor_sync(0xffffffffu, part[p], o);
      ZSB_BFLY(16)
      forward1(b0 + PB + 0, sn[0][0], sn[1][0], pn[0]);
      ZSB_BFLY(8)
      forward1(b0 + PB + 1, sn[0][1], sn[1][1], pn[1]);
      ZSB_BFLY(4)
      forward1(b0 + PB + 2, sn[0][2], sn[1][2], pn[2]);
      ZSB_BFLY(2)
      forward1(b0 + PB + 3, sn[0][3], sn[1][3], pn[3]);
      ZSB_BFLY(1)
#undef ZSB_BFLY
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        float xv[XP];
        load_row(b0 + p, xv);
        const float2 yw = yc[b0 + p];
        const float dout = fmaf(-inv_s1, part[p], yw.x) * yw.y;    // 0 for padding rows
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          g1r[u] = fmaf(dout, fmaxf(sa[u][p], 0.f), g1r[u]);
          const float da = (sa[u][p] > 0.f) ? dout * w1s[u] : 0.f;
#pragma unroll
          for (int k = 0; k < in1; ++k) G[u][k] = fmaf(da, xv[k], G[u][k]);
        }
        g1b += dout;                       // every lane accumulates; only lane 0's copy is used
      }
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        part[p] = pn[p];
        sa[0][p] = sn[0][p]; sa[1][p] = sn[1][p];
      }
    }
    // ---- prior gradient, SGHMC update (noise in A, old momentum in B; results overwrite them)
    warp_fill_normals(A0, n0, a.noise0, c * n0, a.seed, ZSB_STREAM_SGMCMC_NOISE, a.iter, grow,
                      lane);
    warp_fill_normals(A1, H1, a.noise1, c * H1, a.seed + 1, ZSB_STREAM_SGMCMC_NOISE, a.iter, grow,
                      lane);
    auto update = [&](float q1, float g, float xi, float vold, float& nq, float& nv) {
      if (a.second_order) {
        nv = mul(dh, add(add(mul(dh, vold), mul(a.lr, g)), xi));
        nq = add(q1, mul(0.5f, nv));
      } else {
        nv = add(add(mul(oma, vold), mul(a.lr, g)), xi);
        nq = add(q1, nv);
      }
    };
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int m = lane + 32 * u;
      if (m < a.H) {
#pragma unroll
        for (int k = 0; k < in1; ++k) {
          const int idx = m * in1 + k;
          const float g = G[u][k] - pr0[idx] * W[u][k];
          float nv, nq;
          update(W[u][k], g, mul(A0[idx], sd_xi), B0[idx], nq, nv);
          A0[idx] = nq; B0[idx] = nv;
          ksum0 += nv * nv;
        }
        const float g = g1r[u] * inv_s0 - pr1[m] * w1r[u];
        float nv, nq;
        update(w1r[u], g, mul(A1[m], sd_xi), B1[m], nq, nv);
        A1[m] = nq; B1[m] = nv;
        ksum1 += nv * nv;
      }
    }
    if (lane == 0) {
      const float g = g1b - pr1[a.H] * w1b;
      float nv, nq;
      update(w1b, g, mul(A1[a.H], sd_xi), B1[a.H], nq, nv);
      A1[a.H] = nq; B1[a.H] = nv;
      ksum1 += nv * nv;
    }
    __syncwarp();
    for (int i = lane; i < n0; i += 32) { w0c[i] = A0[i]; v0c[i] = B0[i]; }
    for (int i = lane; i < H1; i += 32) { w1c[i] = A1[i]; v1c[i] = B1[i]; }
    __syncwarp();                          // the buffers are restaged for the next chain
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
  // persistent grid: two resident blocks per SM, each warp walks its chains.  (7 warps per block
  // would fill the last round of 8192 chains better -- 3.95 instead of 3.46 rounds -- but measured
  // the same 0.13 ms: the kernel is bound by per-warp latency x warps in flight, not by the tail.)
  const int nw = 8;
  int64_t blocks = zsb_ceil_div(chains, nw);
  if (blocks > 2 * ZSB_NUM_SMS) blocks = 2 * ZSB_NUM_SMS;
  const int Bp = (B + PB - 1) / PB * PB;
  const int n0p = (H * (n_in + 1) + 3) & ~3, n1p = (H + 1 + 3) & ~3;
  const size_t smem = (size_t)((Bp + PB) * ((n_in + 1 + 3) / 4 * 4) + 2 * Bp + 17 * (n0p + n1p)) *
                      sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  switch (n_in + 1) {
#define ZSB_BNN_CASE(N)                                                                          \
  case N:                                                                                        \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(sghmc_bnn_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                           (int)smem);                                                           \
    sghmc_bnn_kernel<N><<<(unsigned)blocks, 32 * nw, smem, st>>>(a);                                 \
    break;
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
