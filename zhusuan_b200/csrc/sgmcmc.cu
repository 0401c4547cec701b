// K5: SG-MCMC update kernels (reference: zhusuan/sgmcmc.py:170-523).
//
// Each update is one fused elementwise pass over [chains, row_len] state (q, v, alpha/aux) with the
// Gaussian noise either injected (parity runs) or drawn in-kernel (Philox4x32-10 keyed by
// (seed; stream, iteration, global chain, 4-element block) so results do not depend on the GPU
// count).  HBM-bound: algorithmic traffic 8*D (SGLD), 16*D (PSGLD, SGHMC), 24*D (SGNHT) bytes per
// chain-step plus the gradient read.  Arithmetic uses explicit RN mul/add in the reference's
// operation order (see hmc.cu) so the NumPy oracle matches bit-for-bit on injected noise.
#include "common.cuh"

namespace {


struct Noise {
  const float* injected;  // [n] standard normals or NULL
  uint64_t seed;
  uint32_t iter, stream_id;
  int64_t row0, row_len;
  __device__ __forceinline__ float at(int64_t i) const {
    if (injected) return injected[i];
    const int64_t row = i / row_len, c = i % row_len;
    float z[4];
    philox_normal4(seed, stream_id, iter, (uint32_t)(row0 + row), (uint32_t)(c >> 2), z);
    return z[c & 3];
  }
};

// float4 flavour: one Philox block per thread-iteration (same numbers as Noise::at)
struct Noise4 {
  const float* injected; uint64_t seed; uint32_t iter, stream_id; int64_t row0;
  __device__ __forceinline__ float4 at(uint32_t i4, uint32_t row, uint32_t c4) const {
    if (injected) return ld4(injected, i4);
    float z[4];
    philox_normal4(seed, stream_id, iter, (uint32_t)(row0 + row), c4, z);
    return make_float4(z[0], z[1], z[2], z[3]);
  }
};
#define ZSB_V4(expr_x, expr_y, expr_z, expr_w) make_float4(expr_x, expr_y, expr_z, expr_w)

template <class F>
__global__ void __launch_bounds__(256) ew_kernel(int64_t n, F f) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    f(i);
}
// elementwise + block partial sum of the returned value (for mean_k = mean(v^2))
template <class F>
__global__ void __launch_bounds__(256) ew_sum_kernel(int64_t n, float* __restrict__ part, F f) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    s += f(i);
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void final_mean_kernel(const float* __restrict__ part, int n_part, float n_elems,
                                  float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n_part; i += blockDim.x) s += part[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s / n_elems;
}

inline unsigned flat_grid(int64_t n) {
  int64_t blocks = zsb_ceil_div(n, 256);
  const int64_t cap = (int64_t)ZSB_NUM_SMS * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace

extern "C" {

int zsb_sgmcmc_parts(void) { return ZSB_NUM_SMS * 8; }

// SGLD._update_single (sgmcmc.py:195-200): q += 0.5*lr*g + N(0, sqrt(lr))
int zsb_sgmcmc_sgld_f32(float* q, const float* g, const float* noise, float lr, int64_t chains,
                        int64_t row_len, uint64_t seed, uint32_t iter, int64_t row0,
                        void* stream) {
  ZSB_REQUIRE(chains >= 0 && row_len > 0, "zsb_sgmcmc_sgld_f32: bad sizes");
  const int64_t n = chains * row_len;
  if (n == 0) return ZSB_OK;
  Noise nz{noise, seed, iter, ZSB_STREAM_SGMCMC_NOISE, row0, row_len};
  const float sd = sqrtf(lr), hl = mul(0.5f, lr);
  if (zsb_vec4_ok(chains, row_len, {q, g, noise})) {
    Noise4 n4{noise, seed, iter, ZSB_STREAM_SGMCMC_NOISE, row0};
    auto f4 = [=] __device__(uint32_t i4, uint32_t row, uint32_t c4) {
      const float4 qv = ld4(q, i4), gv = ld4(g, i4), z = n4.at(i4, row, c4);
      st4(q, i4, ZSB_V4(add(add(qv.x, mul(hl, gv.x)), mul(z.x, sd)),
                        add(add(qv.y, mul(hl, gv.y)), mul(z.y, sd)),
                        add(add(qv.z, mul(hl, gv.z)), mul(z.z, sd)),
                        add(add(qv.w, mul(hl, gv.w)), mul(z.w, sd))));
    };
    ew4_kernel<<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((uint32_t)(n / 4),
                                                                  (uint32_t)(row_len / 4), f4);
    return zsb_check_launch("sgmcmc_sgld4");
  }
  auto f = [=] __device__(int64_t i) {
    q[i] = add(add(q[i], mul(hl, g[i])), mul(nz.at(i), sd));
  };
  ew_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, f);
  return zsb_check_launch("sgmcmc_sgld");
}

// PSGLD (sgmcmc.py:225-257): aux = decay*aux + (1-decay)*g^2; G = 1/(eps+sqrt(aux));
//   q += 0.5*lr*G*g + N(0, sqrt(lr*G))
int zsb_sgmcmc_psgld_f32(float* q, float* aux, const float* g, const float* noise, float lr,
                         float decay, float epsilon, int64_t chains, int64_t row_len,
                         uint64_t seed, uint32_t iter, int64_t row0, void* stream) {
  ZSB_REQUIRE(chains >= 0 && row_len > 0, "zsb_sgmcmc_psgld_f32: bad sizes");
  const int64_t n = chains * row_len;
  if (n == 0) return ZSB_OK;
  Noise nz{noise, seed, iter, ZSB_STREAM_SGMCMC_NOISE, row0, row_len};
  const float hl = mul(0.5f, lr), omd = sub(1.f, decay);
  auto f = [=] __device__(int64_t i) {
    const float gi = g[i];
    const float a = add(mul(decay, aux[i]), mul(omd, mul(gi, gi)));
    aux[i] = a;
    const float G = fdiv(1.f, add(epsilon, sqrtf(a)));
    q[i] = add(add(q[i], mul(mul(hl, G), gi)), mul(nz.at(i), sqrtf(mul(lr, G))));
  };
  ew_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, f);
  return zsb_check_launch("sgmcmc_psgld");
}

// v <- N(0, sqrt(lr))  (momentum creation sgmcmc.py:320-324 and resampling :327-336)
int zsb_sgmcmc_resample_v_f32(float* v, const float* noise, float lr, int64_t chains,
                              int64_t row_len, uint64_t seed, uint32_t iter, int64_t row0,
                              void* stream) {
  ZSB_REQUIRE(chains >= 0 && row_len > 0, "zsb_sgmcmc_resample_v_f32: bad sizes");
  const int64_t n = chains * row_len;
  if (n == 0) return ZSB_OK;
  Noise nz{noise, seed, iter, ZSB_STREAM_SGMCMC_RESAMPLE, row0, row_len};
  const float sd = sqrtf(lr);
  if (zsb_vec4_ok(chains, row_len, {v, noise})) {
    Noise4 n4{noise, seed, iter, ZSB_STREAM_SGMCMC_RESAMPLE, row0};
    auto f4 = [=] __device__(uint32_t i4, uint32_t row, uint32_t c4) {
      const float4 z = n4.at(i4, row, c4);
      st4(v, i4, ZSB_V4(mul(z.x, sd), mul(z.y, sd), mul(z.z, sd), mul(z.w, sd)));
    };
    ew4_kernel<<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((uint32_t)(n / 4),
                                                                  (uint32_t)(row_len / 4), f4);
    return zsb_check_launch("sgmcmc_resample_v4");
  }
  auto f = [=] __device__(int64_t i) { v[i] = mul(nz.at(i), sd); };
  ew_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, f);
  return zsb_check_launch("sgmcmc_resample_v");
}

// 2nd-order integrator, first half (sgmcmc.py:351 / 493): q1 = q + 0.5*v  (in place)
int zsb_sgmcmc_half_q_f32(float* q, const float* v, int64_t n, void* stream) {
  ZSB_REQUIRE(n >= 0, "zsb_sgmcmc_half_q_f32: bad size");
  if (n == 0) return ZSB_OK;
  if (zsb_vec4_ok(1, n, {q, v})) {
    auto f4 = [=] __device__(uint32_t i4, uint32_t, uint32_t) {
      const float4 qv = ld4(q, i4), vv = ld4(v, i4);
      st4(q, i4, ZSB_V4(add(qv.x, mul(0.5f, vv.x)), add(qv.y, mul(0.5f, vv.y)),
                        add(qv.z, mul(0.5f, vv.z)), add(qv.w, mul(0.5f, vv.w))));
    };
    ew4_kernel<<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((uint32_t)(n / 4),
                                                                  (uint32_t)(n / 4), f4);
    return zsb_check_launch("sgmcmc_half_q4");
  }
  auto f = [=] __device__(int64_t i) { q[i] = add(q[i], mul(0.5f, v[i])); };
  ew_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, f);
  return zsb_check_launch("sgmcmc_half_q");
}

// SGHMC velocity/position update (sgmcmc.py:338-356).  second_order: q holds q1 on entry.
//   1st: v = (1-alpha)*v + lr*g + xi ; q += v
//   2nd: v = dh*(dh*v + lr*g + xi)   ; q = q1 + 0.5*v      dh = exp(-0.5*alpha)
//   xi ~ N(0, sqrt(2*(alpha-beta)*lr));  part[] receives block sums of v^2 (mean_k, :358).
int zsb_sgmcmc_sghmc_f32(float* q, float* v, const float* g, const float* noise, float lr,
                         float alpha, float beta, int second_order, int64_t chains,
                         int64_t row_len, uint64_t seed, uint32_t iter, int64_t row0,
                         float* part, float* mean_k, void* stream) {
  ZSB_REQUIRE(chains >= 0 && row_len > 0 && part && mean_k, "zsb_sgmcmc_sghmc_f32: bad args");
  const int64_t n = chains * row_len;
  if (n == 0) return ZSB_OK;
  Noise nz{noise, seed, iter, ZSB_STREAM_SGMCMC_NOISE, row0, row_len};
  const float sd = sqrtf(mul(mul(2.f, sub(alpha, beta)), lr));
  const float dh = expf(mul(-0.5f, alpha)), oma = sub(1.f, alpha);
  if (zsb_vec4_ok(chains, row_len, {q, v, g, noise})) {
    Noise4 n4{noise, seed, iter, ZSB_STREAM_SGMCMC_NOISE, row0};
    auto upd = [=] __device__(float qe, float ve, float ge, float ze, float& nq) -> float {
      const float xi = mul(ze, sd);
      float nv;
      if (second_order) {
        nv = mul(dh, add(add(mul(dh, ve), mul(lr, ge)), xi));
        nq = add(qe, mul(0.5f, nv));
      } else {
        nv = add(add(mul(oma, ve), mul(lr, ge)), xi);
        nq = add(qe, nv);
      }
      return nv;
    };
    auto f4 = [=] __device__(uint32_t i4, uint32_t row, uint32_t c4) -> float {
      const float4 qv = ld4(q, i4), vv = ld4(v, i4), gv = ld4(g, i4), z = n4.at(i4, row, c4);
      float4 nq, nv;
      nv.x = upd(qv.x, vv.x, gv.x, z.x, nq.x);
      nv.y = upd(qv.y, vv.y, gv.y, z.y, nq.y);
      nv.z = upd(qv.z, vv.z, gv.z, z.z, nq.z);
      nv.w = upd(qv.w, vv.w, gv.w, z.w, nq.w);
      st4(q, i4, nq);
      st4(v, i4, nv);
      return nv.x * nv.x + nv.y * nv.y + nv.z * nv.z + nv.w * nv.w;
    };
    const unsigned grid4 = flat_grid(n / 4);
    ew4_sum_kernel<<<grid4, 256, 0, (cudaStream_t)stream>>>((uint32_t)(n / 4),
                                                           (uint32_t)(row_len / 4), part, f4);
    int rc4 = zsb_check_launch("sgmcmc_sghmc4");
    if (rc4) return rc4;
    final_mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(part, (int)grid4, (float)n, mean_k);
    return zsb_check_launch("sgmcmc_sghmc_mean_k");
  }
  auto f = [=] __device__(int64_t i) -> float {
    const float xi = mul(nz.at(i), sd);
    float nv;
    if (second_order) {
      nv = mul(dh, add(add(mul(dh, v[i]), mul(lr, g[i])), xi));
      q[i] = add(q[i], mul(0.5f, nv));
    } else {
      nv = add(add(mul(oma, v[i]), mul(lr, g[i])), xi);
      q[i] = add(q[i], nv);
    }
    v[i] = nv;
    return nv * nv;
  };
  const unsigned grid = flat_grid(n);
  ew_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, part, f);
  int rc = zsb_check_launch("sgmcmc_sghmc");
  if (rc) return rc;
  final_mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(part, (int)grid, (float)n, mean_k);
  return zsb_check_launch("sgmcmc_sghmc_mean_k");
}

// SGNHT with vector alpha (sgmcmc.py:460-523, use_vector_alpha=True).  second_order: q holds q1.
//   1st: v=(1-al)*v+lr*g+xi; q+=v; k=v^2; al += tune*(k-lr)
//   2nd: k1=v_old^2; a1=al+0.5*tune*(k1-lr); dh=exp(-0.5*a1); v=dh*(dh*v+lr*g+xi); q=q1+0.5*v;
//        k=v^2; al=a1+0.5*tune*(k-lr)
//   xi ~ N(0, sqrt(2*a*lr)).  mean_k_out (nullable, [n]) receives k.
int zsb_sgmcmc_sgnht_vec_f32(float* q, float* v, float* alpha, const float* g, const float* noise,
                             float lr, float a, float tune_rate, int second_order, int64_t chains,
                             int64_t row_len, uint64_t seed, uint32_t iter, int64_t row0,
                             float* mean_k_out, void* stream) {
  ZSB_REQUIRE(chains >= 0 && row_len > 0, "zsb_sgmcmc_sgnht_vec_f32: bad sizes");
  const int64_t n = chains * row_len;
  if (n == 0) return ZSB_OK;
  Noise nz{noise, seed, iter, ZSB_STREAM_SGMCMC_NOISE, row0, row_len};
  const float sd = sqrtf(mul(mul(2.f, a), lr));
  const float ht = mul(0.5f, tune_rate);
  auto f = [=] __device__(int64_t i) {
    const float xi = mul(nz.at(i), sd);
    const float ov = v[i], al = alpha[i];
    float nv, na, k;
    if (second_order) {
      const float a1 = add(al, mul(ht, sub(mul(ov, ov), lr)));
      const float dh = expf(mul(-0.5f, a1));
      nv = mul(dh, add(add(mul(dh, ov), mul(lr, g[i])), xi));
      q[i] = add(q[i], mul(0.5f, nv));
      k = mul(nv, nv);
      na = add(a1, mul(ht, sub(k, lr)));
    } else {
      nv = add(add(mul(sub(1.f, al), ov), mul(lr, g[i])), xi);
      q[i] = add(q[i], nv);
      k = mul(nv, nv);
      na = add(al, mul(tune_rate, sub(k, lr)));
    }
    v[i] = nv;
    alpha[i] = na;
    if (mean_k_out) mean_k_out[i] = k;
  };
  ew_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, f);
  return zsb_check_launch("sgmcmc_sgnht_vec");
}

// mean(v^2) over all elements -> out[0]   (maybe_reduce_mean, sgmcmc.py:464-468; scalar alpha)
int zsb_sgmcmc_mean_sq_f32(const float* v, int64_t n, float* part, float* out, void* stream) {
  ZSB_REQUIRE(n > 0 && part && out, "zsb_sgmcmc_mean_sq_f32: bad args");
  auto f = [=] __device__(int64_t i) -> float { return v[i] * v[i]; };
  const unsigned grid = flat_grid(n);
  ew_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, part, f);
  int rc = zsb_check_launch("sgmcmc_mean_sq");
  if (rc) return rc;
  final_mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(part, (int)grid, (float)n, out);
  return zsb_check_launch("sgmcmc_mean_sq_final");
}

// SGNHT with scalar alpha, velocity/position part.  alpha_eff is a DEVICE scalar:
//   1st order: alpha_eff = alpha;  v=(1-alpha)*v+lr*g+xi; q+=v
//   2nd order: alpha_eff = alpha1 (computed by the caller from mean(v_old^2));
//              v=dh*(dh*v+lr*g+xi), dh=exp(-0.5*alpha1); q=q1+0.5*v
// part/mean_k: mean(v_new^2) for the alpha update done by zsb_sgmcmc_sgnht_alpha_f32.
int zsb_sgmcmc_sgnht_scalar_f32(float* q, float* v, const float* alpha_eff, const float* g,
                                const float* noise, float lr, float a, int second_order,
                                int64_t chains, int64_t row_len, uint64_t seed, uint32_t iter,
                                int64_t row0, float* part, float* mean_k, void* stream) {
  ZSB_REQUIRE(chains >= 0 && row_len > 0 && alpha_eff && part && mean_k,
              "zsb_sgmcmc_sgnht_scalar_f32: bad args");
  const int64_t n = chains * row_len;
  if (n == 0) return ZSB_OK;
  Noise nz{noise, seed, iter, ZSB_STREAM_SGMCMC_NOISE, row0, row_len};
  const float sd = sqrtf(mul(mul(2.f, a), lr));
  auto f = [=] __device__(int64_t i) -> float {
    const float xi = mul(nz.at(i), sd);
    const float al = *alpha_eff;
    float nv;
    if (second_order) {
      const float dh = expf(mul(-0.5f, al));
      nv = mul(dh, add(add(mul(dh, v[i]), mul(lr, g[i])), xi));
      q[i] = add(q[i], mul(0.5f, nv));
    } else {
      nv = add(add(mul(sub(1.f, al), v[i]), mul(lr, g[i])), xi);
      q[i] = add(q[i], nv);
    }
    v[i] = nv;
    return nv * nv;
  };
  const unsigned grid = flat_grid(n);
  ew_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, part, f);
  int rc = zsb_check_launch("sgmcmc_sgnht_scalar");
  if (rc) return rc;
  final_mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(part, (int)grid, (float)n, mean_k);
  return zsb_check_launch("sgmcmc_sgnht_scalar_mean_k");
}

}  // extern "C"

namespace {
__global__ void alpha_axpy_kernel(float* out, const float* in, const float* mean_k, float coef,
                                  float lr) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    out[0] = __fadd_rn(in[0], __fmul_rn(coef, __fsub_rn(mean_k[0], lr)));
}
}  // namespace

extern "C" {
// out = in + coef * (mean_k - lr)  on device scalars (alpha thermostat updates, sgmcmc.py:490, 495, 506)
int zsb_sgmcmc_sgnht_alpha_f32(float* out, const float* in, const float* mean_k, float coef,
                               float lr, void* stream) {
  ZSB_REQUIRE(out && in && mean_k, "zsb_sgmcmc_sgnht_alpha_f32: bad args");
  alpha_axpy_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(out, in, mean_k, coef, lr);
  return zsb_check_launch("sgmcmc_sgnht_alpha");
}
}  // extern "C"
