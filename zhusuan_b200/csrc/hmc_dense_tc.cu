// Dense-Gaussian leapfrog pass on 5th-gen tensor cores (impl 1): tcgen05.mma kind::tf32 with a
// 3xTF32 split so the fp32 gradient  g = b - q P  keeps ~fp32 accuracy:
//     q = q_hi + q_lo,  P = P_hi + P_lo   (hi = top 19 bits, exactly what the TF32 datapath reads)
//     q P ~= q_hi P_hi + q_hi P_lo + q_lo P_hi          (dropped term ~2^-22 relative)
// accumulated in fp32 in TMEM.  Same fused leapfrog epilogue as the SIMT kernel (hmc_dense.cu).
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled tiles of q_hi(=q), q_lo, P_hi, P_lo
//               into a 2-stage shared-memory ring (96 KB / stage), mbarrier complete_tx signalling
//   warp 1      MMA issuer: one elected lane issues 12 tcgen05.mma (128x256x8, 3 per k-step) per
//               stage; tcgen05.commit frees the smem slot / publishes the accumulator
//   warps 2-5   epilogue: tcgen05.ld (32x32b.x32) the fp32 accumulator (thread == chain row),
//               p += s2*g, q_next = q + eps*p/m, q_next_lo, row partials of lp and K -> HBM
//   TMEM        2 x 256 columns: accumulator double buffer (epilogue of tile i overlaps MMA of i+1)
// Tiles (128 chains x 256 dims) are assigned round-robin with the N index fastest, so the CTAs
// running concurrently share their A rows and the whole (8 MB hi+lo) P through the 126 MB L2.
#include "common.cuh"
#include <cuda.h>

namespace {

constexpr int BM = 128, BN = 256, BK = 32, STAGES = 2;
constexpr int A_TILE_BYTES = BM * BK * 4;            // 16 KB
constexpr int B_TILE_BYTES = BN * BK * 4;            // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;   // 96 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr long long WAIT_TIMEOUT_CYCLES = 4000000000LL;   // ~2 s: trap instead of hanging the box

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) {
      printf("zsb dense_tc: mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n",
             blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, 128B-swizzled operand tile (rows x 32 fp32 = 128 B per row, 8-row atoms of 1024 B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address      bits [0,14)
  d |= (uint64_t)0 << 16;                             // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // SBO = 1024 B       bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor: D=F32, A=B=TF32, both K-major, M=128, N=256
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
dense_leapfrog_tc_kernel(const __grid_constant__ CUtensorMap map_qhi,
                         const __grid_constant__ CUtensorMap map_qlo,
                         const __grid_constant__ CUtensorMap map_phi,
                         const __grid_constant__ CUtensorMap map_plo,
                         const float* __restrict__ q_cur, float* __restrict__ q_next,
                         float* __restrict__ q_next_lo, const float* __restrict__ p_in,
                         float* __restrict__ p_out, const float* __restrict__ bvec,
                         const float* __restrict__ mu, const float* __restrict__ mass,
                         const float* __restrict__ state, float p_scale,
                         float* __restrict__ lp_part, float* __restrict__ k_part, int64_t chains,
                         int D) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment required by the 128B swizzle atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
  const uint32_t full_bar = bars;                       // [STAGES]
  const uint32_t empty_bar = bars + 8 * STAGES;         // [STAGES]
  const uint32_t tfull_bar = bars + 16 * STAGES;        // [2]
  const uint32_t tempty_bar = bars + 16 * STAGES + 16;  // [2]
  const uint32_t tmem_slot = bars + 16 * STAGES + 32;   // u32
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(
      smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_n = (D + BN - 1) / BN;
  const int64_t n_tiles_m = (chains + BM - 1) / BM;
  const int64_t n_tiles = n_tiles_m * n_tiles_n;
  const int n_kb = D / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qhi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qlo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_phi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_plo) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (int)((t / n_tiles_n) * BM);
        const int n0 = (int)((t % n_tiles_n) * BN);
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          mbar_expect_tx(fb, STAGE_BYTES);
          tma_load_2d(sa, &map_qhi, fb, kb * BK, m0);
          tma_load_2d(sa + A_TILE_BYTES, &map_qlo, fb, kb * BK, m0);
          tma_load_2d(sa + 2 * A_TILE_BYTES, &map_phi, fb, kb * BK, n0);
          tma_load_2d(sa + 2 * A_TILE_BYTES + B_TILE_BYTES, &map_plo, fb, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);         // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t a_hi = make_smem_desc(sa);
          const uint64_t a_lo = make_smem_desc(sa + A_TILE_BYTES);
          const uint64_t b_hi = make_smem_desc(sa + 2 * A_TILE_BYTES);
          const uint64_t b_lo = make_smem_desc(sa + 2 * A_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t ko = (uint64_t)((k * 8 * 4) >> 4);   // +32 B along K per UMMA_K=8
            // small cross terms first, hi*hi last
            umma_tf32(d_tmem, a_lo + ko, b_hi + ko, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_tf32(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
            umma_tf32(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
          }
          umma_commit(empty_bar + 8 * stage);             // frees the smem slot when MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar + 8 * acc);                 // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row_in_tile = quarter * 32 + lane;
    const float eps = state[ZSB_ST_EPS_USED];
    const float s2 = mul(eps, p_scale);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int64_t tile_m = t / n_tiles_n;
      const int tile_n = (int)(t % n_tiles_n);
      const int64_t m = tile_m * BM + row_in_tile;
      const int n0 = tile_n * BN;
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      float lp_acc = 0.f, k_acc = 0.f;
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(trow + (uint32_t)c, v);      // all 32 lanes participate (sync.aligned)
        tmem_ld_wait();
        const int n = n0 + c;
        if (m < chains && n < D) {
          const float* pin = p_in + m * D + n;
          const float* qc = q_cur + m * D + n;
          float* po = p_out + m * D + n;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 pv = *reinterpret_cast<const float4*>(pin + j);
            const float4 qv = *reinterpret_cast<const float4*>(qc + j);
            const float4 ms = *reinterpret_cast<const float4*>(mass + n + j);
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), mv = bv;
            if (bvec) bv = *reinterpret_cast<const float4*>(bvec + n + j);
            if (mu) mv = *reinterpret_cast<const float4*>(mu + n + j);
            const float pe[4] = {pv.x, pv.y, pv.z, pv.w}, qe[4] = {qv.x, qv.y, qv.z, qv.w};
            const float me[4] = {ms.x, ms.y, ms.z, ms.w}, be[4] = {bv.x, bv.y, bv.z, bv.w};
            const float ue[4] = {mv.x, mv.y, mv.z, mv.w};
            float pn[4], qn[4], ql[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float g = sub(be[e], __uint_as_float(v[j + e]));
              pn[e] = add(pe[e], mul(s2, g));
              qn[e] = add(qe[e], mul(eps, fdiv(pn[e], me[e])));
              ql[e] = qn[e] - __uint_as_float(__float_as_uint(qn[e]) & 0xFFFFE000u);
              lp_acc += (qe[e] - ue[e]) * g;
              k_acc += fdiv(mul(pn[e], pn[e]), me[e]);
            }
            *reinterpret_cast<float4*>(po + j) = make_float4(pn[0], pn[1], pn[2], pn[3]);
            if (q_next) {
              *reinterpret_cast<float4*>(q_next + m * D + n + j) =
                  make_float4(qn[0], qn[1], qn[2], qn[3]);
              *reinterpret_cast<float4*>(q_next_lo + m * D + n + j) =
                  make_float4(ql[0], ql[1], ql[2], ql[3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + 8 * acc);               // 128 arrivals free the accumulator
      if (m < chains) {
        if (lp_part) lp_part[(int64_t)tile_n * chains + m] = lp_acc;
        if (k_part) k_part[(int64_t)tile_n * chains + m] = k_acc;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// q_lo = q - (q with the low 13 mantissa bits cleared): the residual the TF32 datapath drops.
__global__ void __launch_bounds__(256) split_lo_kernel(const float* __restrict__ q,
                                                       float* __restrict__ lo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(q)[i];
    float4 r;
    r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    reinterpret_cast<float4*>(lo)[i] = r;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
            cudaSuccess || qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D row-major [rows, cols] fp32 tensor, box = [box_rows, 32 cols], 128B swizzle, zero OOB fill.
int make_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    zsb_set_error("dense_tc: cuTensorMapEncodeTiled unavailable");
    return ZSB_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    zsb_set_error("dense_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return ZSB_ERR_CUDA;
  }
  return ZSB_OK;
}

}  // namespace

int zsb_dense_tc_ntiles(int D) { return (D + BN - 1) / BN; }

// q_lo scratch: the TC path needs the TF32 residual of every A operand.  `q_cur_lo` must hold the
// residual of q_cur on entry; the kernel writes q_next's residual to `q_next_lo`.
int zsb_dense_leapfrog_tc_launch(const float* q_cur, const float* q_cur_lo, float* q_next,
                                 float* q_next_lo, const float* p_in, float* p_out,
                                 const float* P_hi, const float* P_lo, const float* bvec,
                                 const float* mu, const float* mass, const float* state,
                                 float p_scale, float* lp_part, float* k_part, int64_t chains,
                                 int D, cudaStream_t st) {
  if (D % BK != 0 || D < BK) {
    zsb_set_error("dense_tc: D must be a multiple of %d", BK);
    return ZSB_ERR_INVALID;
  }
  if (chains >= (1LL << 31) || (q_next && !q_next_lo) || !q_cur_lo) {
    zsb_set_error("dense_tc: bad arguments");
    return ZSB_ERR_INVALID;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dense_leapfrog_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      zsb_set_error("dense_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return ZSB_ERR_CUDA;
    }
    attr_set = true;
  }
  CUtensorMap m_qhi, m_qlo, m_phi, m_plo;
  int rc;
  if ((rc = make_map(&m_qhi, q_cur, (uint64_t)chains, (uint64_t)D, BM))) return rc;
  if ((rc = make_map(&m_qlo, q_cur_lo, (uint64_t)chains, (uint64_t)D, BM))) return rc;
  if ((rc = make_map(&m_phi, P_hi, (uint64_t)D, (uint64_t)D, BN))) return rc;
  if ((rc = make_map(&m_plo, P_lo, (uint64_t)D, (uint64_t)D, BN))) return rc;
  const int64_t n_tiles = ((chains + BM - 1) / BM) * ((D + BN - 1) / BN);
  int dev = 0, sms = ZSB_NUM_SMS;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  dense_leapfrog_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(
      m_qhi, m_qlo, m_phi, m_plo, q_cur, q_next, q_next_lo, p_in, p_out, bvec, mu, mass, state,
      p_scale, lp_part, k_part, chains, D);
  return zsb_check_launch("hmc_dense_leapfrog_tc");
}

int zsb_dense_split_lo_launch(const float* q, float* lo, int64_t n, cudaStream_t st) {
  if (n % 4 != 0) {
    zsb_set_error("dense split: element count must be a multiple of 4");
    return ZSB_ERR_INVALID;
  }
  const int64_t n4 = n / 4;
  int64_t blocks = zsb_ceil_div(n4, 256);
  if (blocks > ZSB_NUM_SMS * 16) blocks = ZSB_NUM_SMS * 16;
  if (blocks < 1) blocks = 1;
  split_lo_kernel<<<(unsigned)blocks, 256, 0, st>>>(q, lo, n4);
  return zsb_check_launch("hmc_dense_split_lo");
}
