// K2 / K3 / K4: parallel-chain HMC kernels (reference: zhusuan/hmc.py).
//
// Layout: every latent is canonicalised by the host to a row-major [chains, row_len] float32
// matrix (chain axes flattened to rows, data axes flattened to row_len); per-dimension vectors
// (mass, EWMV mean/var) are [row_len].  Sampler scalars live in a 16-float device state block
// (common.cuh: ZSB_ST_*), so an iteration never needs a host read-back.
//
// Elementwise arithmetic uses explicit round-to-nearest mul/add (no FMA contraction) in the
// reference's operation order so that the CPU oracle (NumPy float32) reproduces it bit-for-bit;
// only reductions (different summation tree) and libm calls (exp/log) can differ in the last ulp.
// These kernels are HBM-bound, so forgoing FMA costs nothing.
#include "common.cuh"

namespace {

// Device-driven iterations (CUDA-graph replay): when the host passes iter == ZSB_ITER_FROM_STATE the
// Philox iteration counter is read from the sampler state block, which zsb_hmc_begin_f32 advances
// on the device, so a captured graph draws fresh numbers on every replay.
#define ZSB_ITER_FROM_STATE 0xFFFFFFFFu
__device__ __forceinline__ uint32_t resolve_iter(uint32_t iter, const float* state) {
  return (iter == ZSB_ITER_FROM_STATE && state) ? (uint32_t)state[ZSB_ST_T] : iter;
}


// ---------------------------------------------------------------------------------------------
// a1  random_momentum (hmc.py:21-23): p = N(0,1) * sqrt(mass);   optional kinetic 0.5*sum p^2/m.
// LANES threads per chain; each lane walks 4-element Philox blocks.
template <int LANES>
__global__ void __launch_bounds__(256) momentum_kernel(float* __restrict__ p,
                                                       const float* __restrict__ noise,
                                                       const float* __restrict__ mass,
                                                       int64_t mass_n, int64_t chains,
                                                       int64_t row_len, uint64_t seed,
                                                       uint32_t iter_in, uint32_t stream_id,
                                                       int64_t row0, float* __restrict__ k_out,
                                                       int accumulate,
                                                       const float* __restrict__ state) {
  const uint32_t iter = resolve_iter(iter_in, state);
  const int rows_per_block = 256 / LANES;
  const int lane = threadIdx.x % LANES;
  const int64_t nblk = (row_len + 3) / 4;
  for (int64_t row = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / LANES; row < chains;
       row += (int64_t)gridDim.x * rows_per_block) {
    float kin = 0.f;
    for (int64_t b = lane; b < nblk; b += LANES) {
      float z[4];
      if (!noise) philox_normal4(seed, stream_id, iter, (uint32_t)(row0 + row), (uint32_t)b, z);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t c = b * 4 + e;
        if (c < row_len) {
          const float zz = noise ? noise[row * row_len + c] : z[e];
          const float m = mass[c % mass_n];
          const float pv = mul(zz, sqrtf(m));
          p[row * row_len + c] = pv;
          kin += fdiv(mul(pv, pv), m);
        }
      }
    }
    kin = sub_warp_sum<LANES>(kin);
    if (lane == 0 && k_out) {
      const float v = mul(0.5f, kin);
      k_out[row] = accumulate ? add(k_out[row], v) : v;
    }
  }
}

// Vectorised form of the same draw (row_len % 4 == 0, one mass entry per column, in-kernel Philox,
// 16-byte aligned p / mass): the CTA stages mass and sqrt(mass) in shared memory once, a lane turns
// one Philox block into one 128-bit store.  Same counters, same operation order per element and
// per kinetic sum as momentum_kernel: the two paths are bit-identical (tests/test_gpu_hmc.py).
template <int LANES>
__global__ void __launch_bounds__(256) momentum_vec4_kernel(float* __restrict__ p,
                                                            const float* __restrict__ mass,
                                                            int64_t chains, int row_len,
                                                            uint64_t seed, uint32_t iter_in,
                                                            uint32_t stream_id, int64_t row0,
                                                            float* __restrict__ k_out,
                                                            int accumulate,
                                                            const float* __restrict__ state) {
  extern __shared__ float4 mom_sm[];                 // [nblk] mass, [nblk] sqrt(mass)
  const uint32_t iter = resolve_iter(iter_in, state);
  const int nblk = row_len >> 2;
  for (int i = threadIdx.x; i < nblk; i += 256) {
    const float4 m = __ldg(reinterpret_cast<const float4*>(mass) + i);
    mom_sm[i] = m;
    mom_sm[nblk + i] = make_float4(sqrtf(m.x), sqrtf(m.y), sqrtf(m.z), sqrtf(m.w));
  }
  __syncthreads();
  constexpr int rows_per_block = 256 / LANES;
  const int lane = threadIdx.x % LANES;
  for (int64_t row = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / LANES; row < chains;
       row += (int64_t)gridDim.x * rows_per_block) {
    float4* __restrict__ pr = reinterpret_cast<float4*>(p + row * row_len);
    const uint32_t grow = (uint32_t)(row0 + row);
    float kin = 0.f;
#pragma unroll 2
    for (int b = lane; b < nblk; b += LANES) {
      float z[4];
      philox_normal4(seed, stream_id, iter, grow, (uint32_t)b, z);
      const float4 m = mom_sm[b], sm = mom_sm[nblk + b];
      float4 pv;
      pv.x = mul(z[0], sm.x); pv.y = mul(z[1], sm.y);
      pv.z = mul(z[2], sm.z); pv.w = mul(z[3], sm.w);
      pr[b] = pv;
      kin += fdiv(mul(pv.x, pv.x), m.x);
      kin += fdiv(mul(pv.y, pv.y), m.y);
      kin += fdiv(mul(pv.z, pv.z), m.z);
      kin += fdiv(mul(pv.w, pv.w), m.w);
    }
    kin = sub_warp_sum<LANES>(kin);
    if (lane == 0 && k_out) {
      const float v = mul(0.5f, kin);
      k_out[row] = accumulate ? add(k_out[row], v) : v;
    }
  }
}

// a5 kinetic (hmc.py:32-34): k[c] (+)= 0.5 * sum_d p^2 / mass
template <int LANES>
__global__ void __launch_bounds__(256) kinetic_kernel(const float* __restrict__ p,
                                                      const float* __restrict__ mass,
                                                      int64_t mass_n, int64_t chains,
                                                      int64_t row_len, float* __restrict__ k_out,
                                                      int accumulate) {
  const int rows_per_block = 256 / LANES;
  const int lane = threadIdx.x % LANES;
  for (int64_t row = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / LANES; row < chains;
       row += (int64_t)gridDim.x * rows_per_block) {
    float kin = 0.f;
    for (int64_t c = lane; c < row_len; c += LANES) {
      const float pv = p[row * row_len + c];
      kin += fdiv(mul(pv, pv), mass[c % mass_n]);
    }
    kin = sub_warp_sum<LANES>(kin);
    if (lane == 0) {
      const float v = mul(0.5f, kin);
      k_out[row] = accumulate ? add(k_out[row], v) : v;
    }
  }
}
// 128-bit form (row_len % 4 == 0, mass_n == row_len, 16-byte aligned).  The IEEE division has a
// slow-path branch per element that keeps the compiler from hoisting loads across it, so four
// float4 loads per lane are issued explicitly before any of them is consumed; the out-of-range
// slots of the last batch contribute p = 0 / m = 1, i.e. exactly +0 to a non-negative sum.
template <int LANES>
__global__ void __launch_bounds__(256) kinetic_vec4_kernel(const float* __restrict__ p,
                                                           const float* __restrict__ mass,
                                                           int64_t chains, int row_len,
                                                           float* __restrict__ k_out,
                                                           int accumulate) {
  constexpr int rows_per_block = 256 / LANES;
  const int lane = threadIdx.x % LANES;
  const int n4 = row_len >> 2;
  const float4* __restrict__ mr = reinterpret_cast<const float4*>(mass);
  for (int64_t row = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / LANES; row < chains;
       row += (int64_t)gridDim.x * rows_per_block) {
    const float4* __restrict__ pr = reinterpret_cast<const float4*>(p + row * row_len);
    float kin = 0.f;
    for (int c4 = lane; c4 < n4; c4 += 4 * LANES) {
      float4 pv[4], mv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = c4 + u * LANES;
        const bool ok = i < n4;
        pv[u] = ok ? pr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        mv[u] = ok ? __ldg(mr + i) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        kin += fdiv(mul(pv[u].x, pv[u].x), mv[u].x);
        kin += fdiv(mul(pv[u].y, pv[u].y), mv[u].y);
        kin += fdiv(mul(pv[u].z, pv[u].z), mv[u].z);
        kin += fdiv(mul(pv[u].w, pv[u].w), mv[u].w);
      }
    }
    kin = sub_warp_sum<LANES>(kin);
    if (lane == 0) {
      const float v = mul(0.5f, kin);
      k_out[row] = accumulate ? add(k_out[row], v) : v;
    }
  }
}

// a3 leapfrog_integrator, position half (hmc.py:39, 26-27): q += s1 * (p / mass)
__global__ void __launch_bounds__(256) leapfrog_q_kernel(float* __restrict__ q,
                                                         const float* __restrict__ p,
                                                         const float* __restrict__ mass,
                                                         int64_t mass_n, int64_t row_len,
                                                         const float* __restrict__ eps_dev,
                                                         float scale, int64_t n) {
  const float s1 = mul(*eps_dev, scale);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    q[i] = add(q[i], mul(s1, fdiv(p[i], mass[(i % row_len) % mass_n])));
}
// float4 variants (row_len % 4 == 0, mass_n in {1, row_len})
__global__ void __launch_bounds__(256) leapfrog_q4_kernel(float* __restrict__ q,
                                                          const float* __restrict__ p,
                                                          const float* __restrict__ mass,
                                                          int mass_scalar, uint32_t q4,
                                                          const float* __restrict__ eps_dev,
                                                          float scale, uint32_t n4) {
  const float s1 = mul(*eps_dev, scale);
  for (uint32_t i4 = blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += gridDim.x * blockDim.x) {
    const float4 qv = ld4(q, i4), pv = ld4(p, i4);
    float4 mv;
    if (mass_scalar) { const float m = mass[0]; mv = make_float4(m, m, m, m); }
    else mv = ld4(mass, i4 % q4);
    st4(q, i4, make_float4(add(qv.x, mul(s1, fdiv(pv.x, mv.x))), add(qv.y, mul(s1, fdiv(pv.y, mv.y))),
                           add(qv.z, mul(s1, fdiv(pv.z, mv.z))), add(qv.w, mul(s1, fdiv(pv.w, mv.w)))));
  }
}
__global__ void __launch_bounds__(256) leapfrog_p4_kernel(float* __restrict__ p,
                                                          const float* __restrict__ g,
                                                          const float* __restrict__ eps_dev,
                                                          float scale, uint32_t n4) {
  const float s2 = mul(*eps_dev, scale);
  for (uint32_t i4 = blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += gridDim.x * blockDim.x) {
    const float4 pv = ld4(p, i4), gv = ld4(g, i4);
    st4(p, i4, make_float4(add(pv.x, mul(s2, gv.x)), add(pv.y, mul(s2, gv.y)),
                           add(pv.z, mul(s2, gv.z)), add(pv.w, mul(s2, gv.w))));
  }
}
__global__ void __launch_bounds__(256) select4_kernel(float* __restrict__ q,
                                                      const float* __restrict__ qn,
                                                      const int32_t* __restrict__ accept,
                                                      uint32_t q4, uint32_t n4) {
  for (uint32_t i4 = blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += gridDim.x * blockDim.x)
    if (accept[i4 / q4]) st4(q, i4, ld4(qn, i4));
}

// momentum half (hmc.py:42): p += s2 * grad
__global__ void __launch_bounds__(256) leapfrog_p_kernel(float* __restrict__ p,
                                                         const float* __restrict__ g,
                                                         const float* __restrict__ eps_dev,
                                                         float scale, int64_t n) {
  const float s2 = mul(*eps_dev, scale);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = add(p[i], mul(s2, g[i]));
}

// a6 + a7 get_acceptance_rate / MH decision (hmc.py:46-61, 485-486, 498).
__device__ __forceinline__ void mh_decide(float lp0, float lp1, float k0, float k1, float u,
                                          float& h0, float& h1, float& acc, int& accept,
                                          float& lp_sel, bool& bad_old) {
  h0 = add(-lp0, k0);                                  // hmc.py:31-35 potential + kinetic
  h1 = add(-lp1, k1);
  bad_old = !isfinite(lp0);                            // hmc.py:51-53 check_numerics
  float a = expf(fminf(add(-h1, h0), 0.f));            // hmc.py:54-55
  if (!(isfinite(a) && isfinite(lp1))) a = 0.f;        // hmc.py:56-59
  acc = a;
  accept = (u < a) ? 1 : 0;                            // hmc.py:486
  lp_sel = accept ? lp1 : lp0;                         // hmc.py:498
}

__global__ void __launch_bounds__(256) mh_kernel(const float* __restrict__ lp0,
                                                 const float* __restrict__ lp1,
                                                 const float* __restrict__ k0,
                                                 const float* __restrict__ k1,
                                                 const float* __restrict__ u, uint64_t seed,
                                                 uint32_t iter_in, int64_t row0, int64_t chains,
                                                 float* __restrict__ h0o, float* __restrict__ h1o,
                                                 float* __restrict__ acco,
                                                 int32_t* __restrict__ accepto,
                                                 float* __restrict__ lpselo,
                                                 float* __restrict__ acc_part,
                                                 float* __restrict__ state) {
  __shared__ float red[32];
  const uint32_t iter = resolve_iter(iter_in, state);
  float local = 0.f;
  bool any_bad = false;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < chains;
       c += (int64_t)gridDim.x * blockDim.x) {
    const float uu = u ? u[c] : philox_uniform_row(seed, ZSB_STREAM_UNIFORM, iter,
                                                   (uint32_t)(row0 + c));
    float h0, h1, acc, lps; int accept; bool bad;
    mh_decide(lp0[c], lp1[c], k0[c], k1[c], uu, h0, h1, acc, accept, lps, bad);
    if (h0o) h0o[c] = h0;
    if (h1o) h1o[c] = h1;
    acco[c] = acc;
    if (accepto) accepto[c] = accept;
    if (lpselo) lpselo[c] = lps;
    local += acc;
    any_bad |= bad;
  }
  local = block_sum(local, red);
  if (threadIdx.x == 0 && acc_part) acc_part[blockIdx.x] = local;
  if (any_bad && state) atomicOr(reinterpret_cast<unsigned int*>(state) + ZSB_ST_FLAGS, 1u);
}

// q <- accept ? q_new : q   (hmc.py:488-497), broadcasting accept over the data axes.
__global__ void __launch_bounds__(256) select_kernel(float* __restrict__ q,
                                                     const float* __restrict__ qn,
                                                     const int32_t* __restrict__ accept,
                                                     int64_t row_len, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    if (accept[i / row_len]) q[i] = qn[i];
}

// fixed-order sum of the per-block partials -> stats[0] = sum(acc), stats[1] = n_chains (local)
__global__ void acc_sum_kernel(const float* __restrict__ part, int n_part, int64_t chains,
                               float* __restrict__ stats) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n_part; i += blockDim.x) s += part[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) { stats[0] = s; stats[1] = (float)chains; }
}

// a8 StepsizeTuner.tune (hmc.py:89-112) + assign step_size (hmc.py:379).  One thread.
__global__ void tune_kernel(float* __restrict__ st, const float* __restrict__ stats, int has_tuner,
                            int adapt, float fresh, float gamma, float t0, float kappa,
                            float delta, float t_now) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float abar = fdiv(stats[0], stats[1]);              // hmc.py:377 reduce_mean (global)
  st[ZSB_ST_ACC_MEAN] = abar;
  if (t_now >= 0.f) st[ZSB_ST_T] = t_now;                   // < 0: the device already advanced t
  if (!has_tuner) return;                                   // hmc.py:504-505
  if (adapt) {
    const float step = add(mul(sub(1.f, fresh), st[ZSB_ST_TUNER_STEP]), 1.f);      // :92
    const float rate1 = fdiv(1.f, add(step, t0));                                  // :93
    const float hbar = add(mul(mul(sub(1.f, fresh), sub(1.f, rate1)), st[ZSB_ST_H_BAR]),
                           mul(rate1, sub(delta, abar)));                          // :94-96
    const float log_eps = sub(st[ZSB_ST_MU], mul(fdiv(sqrtf(step), gamma), hbar)); // :97
    const float rate = powf(step, -kappa);                                         // :98
    const float leb = add(mul(rate, log_eps),
                          mul(mul(sub(1.f, fresh), sub(1.f, rate)),
                              st[ZSB_ST_LOG_EPS_BAR]));                            // :99-102
    st[ZSB_ST_TUNER_STEP] = step;
    st[ZSB_ST_H_BAR] = hbar;
    st[ZSB_ST_LOG_EPS_BAR] = leb;
    st[ZSB_ST_STEP_SIZE] = expf(log_eps);                                          // :106, 379
  } else {
    st[ZSB_ST_STEP_SIZE] = expf(st[ZSB_ST_LOG_EPS_BAR]);                           // :110
  }
}

// eps_used <- step_size (no search this iteration, hmc.py:472 else-branch / 464)
__global__ void begin_kernel(float* __restrict__ st, int start_search) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (start_search < 0) {                  // device-driven iteration: t += 1 here (hmc.py:418)
    st[ZSB_ST_T] = st[ZSB_ST_T] + 1.0f;
    start_search = 0;
  }
  st[ZSB_ST_EPS_USED] = st[ZSB_ST_STEP_SIZE];
  if (start_search) { st[ZSB_ST_SEARCH_LAST] = 1.0f; st[ZSB_ST_SEARCH_COND] = 1.0f; }  // :343
}
// a9 one pass of the _init_step_size loop body's bookkeeping (hmc.py:326-338).
__global__ void search_update_kernel(float* __restrict__ st, const float* __restrict__ stats,
                                     float target) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float a = fdiv(stats[0], stats[1]);
  const float eps = st[ZSB_ST_EPS_USED];
  const float factor = 1.5f;
  st[ZSB_ST_EPS_USED] = (a < target) ? mul(eps, fdiv(1.0f, factor)) : mul(eps, factor);
  const bool last_lt = st[ZSB_ST_SEARCH_LAST] < target, a_lt = a < target;
  st[ZSB_ST_SEARCH_COND] = (last_lt == a_lt) ? 1.0f : 0.0f;   // not xor
  st[ZSB_ST_SEARCH_LAST] = a;
}

// a10 EWMV statistics (hmc.py:130-145), one pass over q:
//   part[b][0][d] = sum_c (q[c,d] - mean[d]),  part[b][1][d] = sum_c (q[c,d] - mean[d])^2
// for the rows owned by block-row b.  Coalesced along d; fixed-order final sum in stage 2.
__global__ void __launch_bounds__(256) mass_stats_kernel(const float* __restrict__ q,
                                                         const float* __restrict__ mean,
                                                         int64_t chains, int64_t D,
                                                         float* __restrict__ part) {
  const int64_t d = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float m = mean[d];
  float s1 = 0.f, s2 = 0.f;
  int64_t c = blockIdx.x;
  const int64_t st = gridDim.x;
  for (; c + 3 * st < chains; c += 4 * st) {            // 4 independent loads in flight
    const float x0 = q[c * D + d] - m, x1 = q[(c + st) * D + d] - m;
    const float x2 = q[(c + 2 * st) * D + d] - m, x3 = q[(c + 3 * st) * D + d] - m;
    s1 += x0; s2 += x0 * x0;
    s1 += x1; s2 += x1 * x1;
    s1 += x2; s2 += x2 * x2;
    s1 += x3; s2 += x3 * x3;
  }
  for (; c < chains; c += st) {
    const float x = q[c * D + d] - m;
    s1 += x;
    s2 += x * x;
  }
  part[((int64_t)blockIdx.x * 2 + 0) * D + d] = s1;
  part[((int64_t)blockIdx.x * 2 + 1) * D + d] = s2;
}
// Same sums, same order per dimension, four dimensions per thread (128-bit loads; D % 4 == 0).
__global__ void __launch_bounds__(256) mass_stats4_kernel(const float* __restrict__ q,
                                                          const float* __restrict__ mean,
                                                          int64_t chains, int64_t D,
                                                          float* __restrict__ part) {
  const int64_t D4 = D >> 2;
  const int64_t d4 = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (d4 >= D4) return;
  const float4 m = reinterpret_cast<const float4*>(mean)[d4];
  const float4* __restrict__ q4 = reinterpret_cast<const float4*>(q);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  auto acc = [&](const float4 v) {
    const float x0 = v.x - m.x, x1 = v.y - m.y, x2 = v.z - m.z, x3 = v.w - m.w;
    s1.x += x0; s2.x += x0 * x0;
    s1.y += x1; s2.y += x1 * x1;
    s1.z += x2; s2.z += x2 * x2;
    s1.w += x3; s2.w += x3 * x3;
  };
  int64_t c = blockIdx.x;
  const int64_t st = gridDim.x;
  for (; c + 3 * st < chains; c += 4 * st) {            // 4 independent 128-bit loads in flight
    const float4 v0 = q4[c * D4 + d4], v1 = q4[(c + st) * D4 + d4];
    const float4 v2 = q4[(c + 2 * st) * D4 + d4], v3 = q4[(c + 3 * st) * D4 + d4];
    acc(v0); acc(v1); acc(v2); acc(v3);
  }
  for (; c < chains; c += st) acc(q4[c * D4 + d4]);
  reinterpret_cast<float4*>(part + ((int64_t)blockIdx.x * 2 + 0) * D)[d4] = s1;
  reinterpret_cast<float4*>(part + ((int64_t)blockIdx.x * 2 + 1) * D)[d4] = s2;
}
// Stage 2: stats[which][d] = sum_b part[b][which][d] in a fixed order.  A 256-thread block owns 32
// consecutive columns of the [2*D] output; its 8 warps take the partials b = w, w + 8, ... (four
// independent 128-byte-coalesced loads in flight per thread), then warp 0 adds the 8 sub-sums in
// warp order.  (One thread per column walking all n_part partials took as long as stage 1.)
__global__ void __launch_bounds__(256) mass_stats_final_kernel(const float* __restrict__ part,
                                                               int n_part, int64_t D,
                                                               float* __restrict__ stats) {
  __shared__ float sub[8][32];
  const int tx = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * 32 + tx;               // over 2*D
  const bool ok = j < 2 * D;
  const int64_t which = ok ? j / D : 0, d = ok ? j % D : 0;
  const float* __restrict__ src = part + which * D + d;
  const int64_t stride = 2 * D;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (ok) {
    int b = w;
    for (; b + 24 < n_part; b += 32) {
      const float v0 = src[(int64_t)b * stride], v1 = src[(int64_t)(b + 8) * stride];
      const float v2 = src[(int64_t)(b + 16) * stride], v3 = src[(int64_t)(b + 24) * stride];
      s0 += v0; s1 += v1; s2 += v2; s3 += v3;
    }
    for (; b < n_part; b += 8) s0 += src[(int64_t)b * stride];
  }
  sub[w][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && ok) {
    float s = sub[0][tx];
#pragma unroll
    for (int i = 1; i < 8; ++i) s += sub[i][tx];
    stats[j] = s;
  }
}
// EWMV.update + get_precision + mass gating (hmc.py:130-159, 283-305), per dimension.
//   w = (1-decay)/(1-decay^tt); delta = S1/C; mean += w*delta;
//   var = (1-w)*var + w*(S2/C - w*delta^2)   [= reduce_mean(incr*(q-mean_new)), algebraically]
__global__ void __launch_bounds__(256) mass_update_kernel(float* __restrict__ mean,
                                                          float* __restrict__ var,
                                                          float* __restrict__ mass,
                                                          const float* __restrict__ stats,
                                                          float n_chains_global, int64_t D,
                                                          float decay, float tt, int adapt,
                                                          int use_ones, float* __restrict__ st) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (use_ones < 0) {
    // device-driven iteration: use_ones = -(mass_collect_iters + 1); the EWMV count was advanced
    // by zsb_hmc_begin/tune on the device (nobody writes the state block in this kernel)
    const int mci = -use_ones - 1;
    use_ones = ((int)st[ZSB_ST_T] < mci) ? 1 : 0;
    tt = st[ZSB_ST_EWMV_T] + 1.0f;
  } else if (d == 0 && adapt && st) {
    st[ZSB_ST_EWMV_T] = tt;
  }
  if (d >= D) return;
  float v = var[d];
  if (adapt) {
    const float w = (1.f - decay) / (1.f - powf(decay, tt));
    const float delta = stats[d] / n_chains_global;
    const float s2 = stats[D + d] / n_chains_global;
    mean[d] = mean[d] + w * delta;
    v = (1.f - w) * v + w * (s2 - w * delta * delta);
    var[d] = v;
  }
  mass[d] = use_ones ? 1.0f : 1.0f / v;
}

// ---------------------------------------------------------------------------------------------
// Fused whole-iteration kernel for a diagonal-Gaussian target (config 1: Normal(mean, std) with
// group_ndims=1, examples/toy_examples/gaussian.py:15-20).  One warp per chain, E elements per
// lane held in registers for the entire trajectory: momentum draw, L+1 gradient evaluations,
// Hamiltonians, MH decision and the in-place select -- q is read once and written once
// (algorithmic HBM traffic 8*D bytes per chain-ITERATION instead of 16*D per leapfrog step).
template <int E>
__global__ void __launch_bounds__(256) diag_normal_traj_kernel(
    float* __restrict__ q, const float* __restrict__ noise, const float* __restrict__ u,
    const float* __restrict__ mean, int64_t mean_n, const float* __restrict__ logstd,
    int64_t logstd_n, const float* __restrict__ mass, int64_t mass_n,
    const float* __restrict__ state, int n_leapfrogs, int64_t chains, int64_t D, uint64_t seed,
    uint32_t iter_in, int64_t row0, int search_mode, float* __restrict__ p0_out,
    float* __restrict__ h0o, float* __restrict__ h1o, float* __restrict__ lp0o,
    float* __restrict__ lpselo, float* __restrict__ acco, int32_t* __restrict__ accepto,
    float* __restrict__ acc_part, float* __restrict__ state_flags) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31;
  const uint32_t iter = resolve_iter(iter_in, state);
  const float eps = state[ZSB_ST_EPS_USED];
  float local_acc = 0.f;
  bool any_bad = false;
  // warp index through a shuffle: provably warp-uniform, so the row loop's exit is uniform and the
  // shuffles inside it compile to plain SHFL (not WARPSYNC.COLLECTIVE sequences)
  const int warp_id = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  // parameter vectors are either scalars, one entry per column, or (rarely) a shorter period:
  // resolve the period once instead of three 64-bit modulos per element
  const int Di = (int)D;
  const int mean_p = mean_n >= D ? Di : (int)mean_n, ls_p = logstd_n >= D ? Di : (int)logstd_n,
            mass_p = mass_n >= D ? Di : (int)mass_n;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp_id; row < chains;
       row += (int64_t)gridDim.x * 8) {
    float q0[E], qc[E], pc[E], mu[E], prec[E], ls[E], ms[E];
    // Philox blocks are 4 consecutive columns; lane owns columns c = j*32 + lane (coalesced
    // rows), i.e. word c%4 of block c/4.  Each block is generated ONCE per row -- lane l draws
    // blocks l, l+32, ... -- and its four normals are handed to the four lanes that own its
    // columns with warp shuffles (round 1 drew the full block in every one of them: 4x the
    // Philox + Box-Muller work in the hot loop of config 1').
    constexpr int NBLK = (E + 3) / 4;
    float zb[NBLK][4];
    if (!noise) {
#pragma unroll
      for (int t = 0; t < NBLK; ++t) {
        const int64_t blk = (int64_t)t * 32 + lane;
        if (blk * 4 < D)
          philox_normal4(seed, ZSB_STREAM_MOMENTUM, iter, (uint32_t)(row0 + row), (uint32_t)blk,
                         zb[t]);
        else
          zb[t][0] = zb[t][1] = zb[t][2] = zb[t][3] = 0.f;
      }
    }
    float lp0 = 0.f, k0 = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int64_t c = (int64_t)j * 32 + lane;
      float zsh = 0.f;
      if (!noise) {           // all lanes take part in the shuffles (no divergence on c < D)
        const int src = (j & 3) * 8 + (lane >> 2);          // lane holding block c / 4
        const float a0 = __shfl_sync(0xffffffffu, zb[j >> 2][0], src);
        const float a1 = __shfl_sync(0xffffffffu, zb[j >> 2][1], src);
        const float a2 = __shfl_sync(0xffffffffu, zb[j >> 2][2], src);
        const float a3 = __shfl_sync(0xffffffffu, zb[j >> 2][3], src);
        const int w = lane & 3;
        zsh = w == 0 ? a0 : w == 1 ? a1 : w == 2 ? a2 : a3;
      }
      if (c < D) {
        const int ci = (int)c;
        q0[j] = q[row * D + c];
        mu[j] = mean[mean_p == Di ? ci : (mean_p == 1 ? 0 : ci % mean_p)];
        ls[j] = logstd[ls_p == Di ? ci : (ls_p == 1 ? 0 : ci % ls_p)];
        ms[j] = mass[mass_p == Di ? ci : (mass_p == 1 ? 0 : ci % mass_p)];
        prec[j] = expf(mul(-2.f, ls[j]));                       // univariate.py:177
        const float z = noise ? noise[row * D + c] : zsh;
        pc[j] = mul(z, sqrtf(ms[j]));                           // hmc.py:22
        if (p0_out) p0_out[row * D + c] = pc[j];
        qc[j] = q0[j];
        const float d = sub(q0[j], mu[j]);
        lp0 += sub(sub(-0.9189385332046727f, ls[j]), mul(mul(0.5f, prec[j]), mul(d, d)));
        k0 += fdiv(mul(pc[j], pc[j]), ms[j]);
      }
    }
    // trajectory (hmc.py:347-372): i = 0..L ; search_mode: the 1-step probe of hmc.py:316-321
    const int L = search_mode ? 1 : n_leapfrogs;
    bool unit = true;                      // p / 1.0f == p exactly: skip the IEEE divide
#pragma unroll
    for (int j = 0; j < E; ++j)
      if ((int64_t)j * 32 + lane < D) unit = unit && (ms[j] == 1.0f);
    unit = __all_sync(0xffffffffu, unit);
    for (int i = 0; i <= L; ++i) {
      const float s1 = (i > 0) ? eps : 0.f;
      const float s2 = (i > 0 && i < L) ? eps : fdiv(eps, 2.f);
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const int64_t c = (int64_t)j * 32 + lane;
        if (c < D) {
          qc[j] = add(qc[j], mul(s1, unit ? pc[j] : fdiv(pc[j], ms[j])));
          const float g = -mul(prec[j], sub(qc[j], mu[j]));     // d/dx Normal log_prob
          pc[j] = add(pc[j], mul(s2, g));
        }
      }
    }
    float lp1 = 0.f, k1 = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int64_t c = (int64_t)j * 32 + lane;
      if (c < D) {
        const float d = sub(qc[j], mu[j]);
        lp1 += sub(sub(-0.9189385332046727f, ls[j]), mul(mul(0.5f, prec[j]), mul(d, d)));
        k1 += fdiv(mul(pc[j], pc[j]), ms[j]);
      }
    }
    lp0 = warp_sum(lp0); lp1 = warp_sum(lp1);
    k0 = mul(0.5f, warp_sum(k0)); k1 = mul(0.5f, warp_sum(k1));
    const float uu = u ? u[row] : philox_uniform_row(seed, ZSB_STREAM_UNIFORM, iter,
                                                     (uint32_t)(row0 + row));
    float h0, h1, acc, lps; int accept; bool bad;
    mh_decide(lp0, lp1, k0, k1, uu, h0, h1, acc, accept, lps, bad);
    any_bad |= bad;
    if (lane == 0) {
      local_acc += acc;
      acco[row] = acc;
      if (!search_mode) {
        if (h0o) h0o[row] = h0;
        if (h1o) h1o[row] = h1;
        if (lp0o) lp0o[row] = lp0;
        if (lpselo) lpselo[row] = lps;
        if (accepto) accepto[row] = accept;
      }
    }
    if (!search_mode && accept) {
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const int64_t c = (int64_t)j * 32 + lane;
        if (c < D) q[row * D + c] = qc[j];
      }
    }
  }
  local_acc = block_sum(local_acc, red);
  if (threadIdx.x == 0) acc_part[blockIdx.x] = local_acc;
  if (any_bad && state_flags)
    atomicOr(reinterpret_cast<unsigned int*>(state_flags) + ZSB_ST_FLAGS, 1u);
}

inline int pick_lanes(int64_t work_items) {
  int lanes = 1;
  while (lanes < 32 && lanes < work_items) lanes <<= 1;
  return lanes;
}
inline unsigned rows_grid(int64_t chains, int lanes) {
  int64_t blocks = zsb_ceil_div(chains, 256 / lanes);
  const int64_t cap = (int64_t)ZSB_NUM_SMS * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}
inline unsigned flat_grid(int64_t n) {
  int64_t blocks = zsb_ceil_div(n, 256);
  const int64_t cap = (int64_t)ZSB_NUM_SMS * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace

extern "C" {

int zsb_hmc_acc_parts(void) { return ZSB_NUM_SMS * 16; }  // capacity the caller must give acc_part

int zsb_hmc_momentum_f32(float* p, const float* noise, const float* mass, int64_t mass_n,
                         int64_t chains, int64_t row_len, uint64_t seed, uint32_t iter,
                         uint32_t stream_id, int64_t row0, float* k_out, int accumulate,
                         const float* iter_state, void* stream) {
  ZSB_REQUIRE(chains >= 0 && row_len > 0 && mass_n > 0, "zsb_hmc_momentum_f32: bad sizes");
  if (chains == 0) return ZSB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int lanes = pick_lanes((row_len + 3) / 4);
  const unsigned g = rows_grid(chains, lanes);
  if (!noise && mass_n == row_len && row_len <= 6144 && zsb_vec4_ok(chains, row_len, {p, mass})) {
    const size_t sm = (size_t)(row_len / 4) * 2 * sizeof(float4);
#define ZSB_L(LN) momentum_vec4_kernel<LN><<<g, 256, sm, st>>>(p, mass, chains, (int)row_len, seed,  \
                                                              iter, stream_id, row0, k_out,        \
                                                              accumulate, iter_state)
    switch (lanes) {
      case 1: ZSB_L(1); break; case 2: ZSB_L(2); break; case 4: ZSB_L(4); break;
      case 8: ZSB_L(8); break; case 16: ZSB_L(16); break; default: ZSB_L(32); break;
    }
#undef ZSB_L
    return zsb_check_launch("hmc_momentum");
  }
#define ZSB_L(LN) momentum_kernel<LN><<<g, 256, 0, st>>>(p, noise, mass, mass_n, chains, row_len, \
                                                        seed, iter, stream_id, row0, k_out,      \
                                                        accumulate, iter_state)
  switch (lanes) {
    case 1: ZSB_L(1); break; case 2: ZSB_L(2); break; case 4: ZSB_L(4); break;
    case 8: ZSB_L(8); break; case 16: ZSB_L(16); break; default: ZSB_L(32); break;
  }
#undef ZSB_L
  return zsb_check_launch("hmc_momentum");
}

int zsb_hmc_kinetic_f32(const float* p, const float* mass, int64_t mass_n, int64_t chains,
                        int64_t row_len, float* k_out, int accumulate, void* stream) {
  ZSB_REQUIRE(chains >= 0 && row_len > 0 && mass_n > 0 && k_out, "zsb_hmc_kinetic_f32: bad sizes");
  if (chains == 0) return ZSB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int lanes = pick_lanes(row_len);
  const unsigned g = rows_grid(chains, lanes);
  if (mass_n == row_len && zsb_vec4_ok(chains, row_len, {p, mass})) {
    const int lanes4 = pick_lanes(row_len / 4);
    const unsigned g4 = rows_grid(chains, lanes4);
#define ZSB_L(LN) kinetic_vec4_kernel<LN><<<g4, 256, 0, st>>>(p, mass, chains, (int)row_len, k_out, \
                                                             accumulate)
    switch (lanes4) {
      case 1: ZSB_L(1); break; case 2: ZSB_L(2); break; case 4: ZSB_L(4); break;
      case 8: ZSB_L(8); break; case 16: ZSB_L(16); break; default: ZSB_L(32); break;
    }
#undef ZSB_L
    return zsb_check_launch("hmc_kinetic");
  }
#define ZSB_L(LN) kinetic_kernel<LN><<<g, 256, 0, st>>>(p, mass, mass_n, chains, row_len, k_out, \
                                                       accumulate)
  switch (lanes) {
    case 1: ZSB_L(1); break; case 2: ZSB_L(2); break; case 4: ZSB_L(4); break;
    case 8: ZSB_L(8); break; case 16: ZSB_L(16); break; default: ZSB_L(32); break;
  }
#undef ZSB_L
  return zsb_check_launch("hmc_kinetic");
}

int zsb_hmc_leapfrog_q_f32(float* q, const float* p, const float* mass, int64_t mass_n,
                           int64_t row_len, const float* eps_dev, float scale, int64_t n,
                           void* stream) {
  ZSB_REQUIRE(n >= 0 && row_len > 0 && mass_n > 0 && eps_dev, "zsb_hmc_leapfrog_q_f32: bad args");
  if (n == 0) return ZSB_OK;
  if ((mass_n == 1 || mass_n == row_len) && n % row_len == 0 &&
      zsb_vec4_ok(n / row_len, row_len, {q, p, mass_n == 1 ? nullptr : mass})) {
    leapfrog_q4_kernel<<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(
        q, p, mass, mass_n == 1, (uint32_t)(row_len / 4), eps_dev, scale, (uint32_t)(n / 4));
    return zsb_check_launch("hmc_leapfrog_q4");
  }
  leapfrog_q_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(q, p, mass, mass_n, row_len,
                                                                    eps_dev, scale, n);
  return zsb_check_launch("hmc_leapfrog_q");
}
int zsb_hmc_leapfrog_p_f32(float* p, const float* g, const float* eps_dev, float scale, int64_t n,
                           void* stream) {
  ZSB_REQUIRE(n >= 0 && eps_dev, "zsb_hmc_leapfrog_p_f32: bad args");
  if (n == 0) return ZSB_OK;
  if (zsb_vec4_ok(1, n, {p, g})) {
    leapfrog_p4_kernel<<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(p, g, eps_dev, scale,
                                                                            (uint32_t)(n / 4));
    return zsb_check_launch("hmc_leapfrog_p4");
  }
  leapfrog_p_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, eps_dev, scale, n);
  return zsb_check_launch("hmc_leapfrog_p");
}

int zsb_hmc_mh_f32(const float* lp0, const float* lp1, const float* k0, const float* k1,
                   const float* u, uint64_t seed, uint32_t iter, int64_t row0, int64_t chains,
                   float* h0, float* h1, float* acc, int32_t* accept, float* lp_sel,
                   float* acc_part, int* n_part_out, float* state, void* stream) {
  ZSB_REQUIRE(chains > 0 && acc && acc_part && n_part_out, "zsb_hmc_mh_f32: bad args");
  const unsigned g = flat_grid(chains);
  *n_part_out = (int)g;
  mh_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(lp0, lp1, k0, k1, u, seed, iter, row0, chains,
                                                 h0, h1, acc, accept, lp_sel, acc_part, state);
  return zsb_check_launch("hmc_mh");
}

int zsb_hmc_select_f32(float* q, const float* q_new, const int32_t* accept, int64_t chains,
                       int64_t row_len, void* stream) {
  ZSB_REQUIRE(chains >= 0 && row_len > 0, "zsb_hmc_select_f32: bad sizes");
  const int64_t n = chains * row_len;
  if (n == 0) return ZSB_OK;
  if (zsb_vec4_ok(chains, row_len, {q, q_new})) {
    select4_kernel<<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(
        q, q_new, accept, (uint32_t)(row_len / 4), (uint32_t)(n / 4));
    return zsb_check_launch("hmc_select4");
  }
  select_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(q, q_new, accept, row_len, n);
  return zsb_check_launch("hmc_select");
}

int zsb_hmc_acc_sum_f32(const float* acc_part, int n_part, int64_t chains, float* stats,
                        void* stream) {
  ZSB_REQUIRE(n_part > 0 && stats, "zsb_hmc_acc_sum_f32: bad args");
  acc_sum_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(acc_part, n_part, chains, stats);
  return zsb_check_launch("hmc_acc_sum");
}

int zsb_hmc_begin_f32(float* state, int start_search, void* stream) {
  begin_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state, start_search);
  return zsb_check_launch("hmc_begin");
}
int zsb_hmc_search_update_f32(float* state, const float* stats, float target, void* stream) {
  search_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state, stats, target);
  return zsb_check_launch("hmc_search_update");
}
int zsb_hmc_tune_f32(float* state, const float* stats, int has_tuner, int adapt, float fresh_start,
                     float gamma, float t0, float kappa, float delta, float t_now, void* stream) {
  tune_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state, stats, has_tuner, adapt, fresh_start,
                                                  gamma, t0, kappa, delta, t_now);
  return zsb_check_launch("hmc_tune");
}

namespace {
__global__ void ewmv_bump_kernel(float* __restrict__ st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) st[ZSB_ST_EWMV_T] = st[ZSB_ST_EWMV_T] + 1.0f;
}
}  // namespace
// device-driven iterations: advance the EWMV update count after an adaptive mass update
int zsb_hmc_ewmv_bump_f32(float* state, void* stream) {
  ewmv_bump_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state);
  return zsb_check_launch("hmc_ewmv_bump");
}

// part: scratch of zsb_hmc_mass_parts()*2*D floats; stats: [2*D] local sums out.
int zsb_hmc_mass_parts(void) { return ZSB_NUM_SMS * 4; }
int zsb_hmc_mass_stats_f32(const float* q, const float* ewmv_mean, int64_t chains, int64_t D,
                           float* part, float* stats, void* stream) {
  ZSB_REQUIRE(chains > 0 && D > 0 && part && stats, "zsb_hmc_mass_stats_f32: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int nb = (int)(chains < ZSB_NUM_SMS * 4 ? chains : ZSB_NUM_SMS * 4);
  const bool vec = D % 4 == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(part) |
                                   reinterpret_cast<uintptr_t>(ewmv_mean)) & 15u) == 0;
  if (vec) {
    const int threads = D / 4 >= 256 ? 256 : (int)((D / 4 + 31) / 32 * 32);
    dim3 grid(nb, (unsigned)zsb_ceil_div(D / 4, threads));
    mass_stats4_kernel<<<grid, threads, 0, st>>>(q, ewmv_mean, chains, D, part);
  } else {
    dim3 grid(nb, (unsigned)zsb_ceil_div(D, 256));
    mass_stats_kernel<<<grid, 256, 0, st>>>(q, ewmv_mean, chains, D, part);
  }
  int rc = zsb_check_launch("hmc_mass_stats");
  if (rc) return rc;
  mass_stats_final_kernel<<<(unsigned)zsb_ceil_div(2 * D, 32), 256, 0, st>>>(part, nb, D, stats);
  return zsb_check_launch("hmc_mass_stats_final");
}
int zsb_hmc_mass_update_f32(float* ewmv_mean, float* ewmv_var, float* mass, const float* stats,
                            float n_chains_global, int64_t D, float decay, float ewmv_t_new,
                            int adapt, int use_ones, float* state, void* stream) {
  ZSB_REQUIRE(D > 0 && ewmv_mean && ewmv_var && mass, "zsb_hmc_mass_update_f32: bad args");
  mass_update_kernel<<<(unsigned)zsb_ceil_div(D, 256), 256, 0, (cudaStream_t)stream>>>(
      ewmv_mean, ewmv_var, mass, stats, n_chains_global, D, decay, ewmv_t_new, adapt, use_ones,
      state);
  return zsb_check_launch("hmc_mass_update");
}

// Fused diagonal-Gaussian HMC iteration (or, with search_mode=1, the one-step acceptance probe of
// _init_step_size, hmc.py:314-326, which leaves q untouched and only fills acc / acc_part).
int zsb_hmc_diag_normal_step_f32(float* q, const float* noise, const float* u, const float* mean,
                                 int64_t mean_n, const float* logstd, int64_t logstd_n,
                                 const float* mass, int64_t mass_n, float* state, int n_leapfrogs,
                                 int64_t chains, int64_t D, uint64_t seed, uint32_t iter,
                                 int64_t row0, int search_mode, float* p0_out, float* h0, float* h1,
                                 float* lp0, float* lp_sel, float* acc, int32_t* accept,
                                 float* acc_part, int* n_part_out, void* stream) {
  ZSB_REQUIRE(chains > 0 && D > 0 && D <= 1024 && mean_n > 0 && logstd_n > 0 && mass_n > 0 &&
                  n_leapfrogs >= 0 && state && acc && acc_part && n_part_out,
              "zsb_hmc_diag_normal_step_f32: bad args (D must be in 1..1024)");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = zsb_ceil_div(chains, 8);
  if (blocks > ZSB_NUM_SMS * 8) blocks = ZSB_NUM_SMS * 8;
  *n_part_out = (int)blocks;
#define ZSB_L(EE)                                                                               \
  diag_normal_traj_kernel<EE><<<(unsigned)blocks, 256, 0, st>>>(                                 \
      q, noise, u, mean, mean_n, logstd, logstd_n, mass, mass_n, state, n_leapfrogs, chains, D, \
      seed, iter, row0, search_mode, p0_out, h0, h1, lp0, lp_sel, acc, accept, acc_part, state)
  if (D <= 128) ZSB_L(4);
  else if (D <= 256) ZSB_L(8);
  else if (D <= 512) ZSB_L(16);
  else ZSB_L(32);
#undef ZSB_L
  return zsb_check_launch("hmc_diag_normal_step");
}

}  // extern "C"
