// Shared tcgen05 / TMA / mbarrier building blocks of the tensor-core kernels (hmc_dense_tc.cu,
// gemm_logjoint_tc.cu): tile constants, pipeline barriers, TMA loads, UMMA issue / commit,
// TMEM loads, shared-memory and instruction descriptors, the CTA-pair (cta_group::2) variants,
// the warp-transpose reduction of the epilogues and the host-side tensor-map encoder.
// (No reference counterpart: ZhuSuan has no kernels; these serve the GEMMs inside the log-joints of
// zhusuan/hmc.py:347-372 and examples/variational_autoencoders/iwae.py:23-32.)
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace {

constexpr int BM = 128;   // dimensions per tile (TMEM lanes)
constexpr int BN = 256;   // chains per tile (TMEM columns)
constexpr int NUM_EPI_WARPS = 8;                         // 2 per TMEM lane quarter
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;      // TMA warp + MMA warp + epilogue
constexpr int TMEM_COLS = 512;
constexpr long long WAIT_TIMEOUT_CYCLES = 4000000000LL;   // ~2 s: trap instead of hanging the box

template <int BK>
struct Cfg {
  static constexpr int A_TILE = BM * BK * 4;
  static constexpr int B_TILE = BN * BK * 4;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;      // 96 KB (BK=32) / 48 KB (BK=16)
  static constexpr int STAGES = (192 * 1024) / STAGE;        // 2 / 4
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
  static constexpr int ROW_BYTES = BK * 4;                   // 128 / 64 -> swizzle width
  static constexpr int SBO = 8 * ROW_BYTES;                  // 8-row swizzle atom
  static constexpr uint64_t LAYOUT = (BK == 32) ? 2 : 4;     // SWIZZLE_128B / SWIZZLE_64B
};

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
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
// K-major swizzled operand tile: rows of BK fp32 (128 B or 64 B), 8-row swizzle atoms.
template <int BK>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address      bits [0,14)
  d |= (uint64_t)(Cfg<BK>::SBO >> 4) << 32;           // stride byte offset bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= Cfg<BK>::LAYOUT << 61;                         // swizzle mode
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
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

// In: every lane holds v[0..15] (value of "column" j at this lane's row).  Out: returns, on lanes
// l < 16 (and mirrored on l+16), the sum over the 32 lanes of column l & 15 (31 shuffles).
__device__ __forceinline__ float warp_transpose_sum16(float v[16], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int o = 8, cnt = 16; o >= 1; o >>= 1, cnt >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < cnt / 2; ++i) {
      const float keep = upper ? v[i + cnt / 2] : v[i];
      const float send = upper ? v[i] : v[i + cnt / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}


constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // shared::cluster address of CTA 0's copy

template <int BK>
struct Cfg2 {
  static constexpr int A_TILE = BM * BK * 4;                 // own 128 dimension rows
  static constexpr int B_TILE = (BN / 2) * BK * 4;           // own half of the chain rows
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;      // 64 KB (BK=32) / 32 KB (BK=16)
  static constexpr int STAGES = (192 * 1024) / STAGE;        // 3 / 6
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;"
               ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map,
                                                uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrives in BOTH CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_2sm() {   // TF32 x TF32 -> F32, M=256 (pair), N=256
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ uint32_t make_idesc_2sm_f16() {   // F16 x F16 -> F32 (formats 0)
  return (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
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

// 2-D row-major [rows, cols] fp32 tensor, box = [box_rows, bk cols], swizzle = row bytes.
int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
             int bk, int half_elems = 0) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    zsb_set_error("dense_tc: cuTensorMapEncodeTiled unavailable");
    return ZSB_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * (half_elems ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(half_elems ? 2 * bk : bk), box_rows};   // same bytes per row
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, half_elems ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                   2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    zsb_set_error("dense_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return ZSB_ERR_CUDA;
  }
  return ZSB_OK;
}

// 2-D row-major [rows, cols] fp32 tensor, box = [box_rows, box_cols], NO swizzle: a staging tile
// that ordinary threads read back (the in-kernel fp16 split of hmc_dense_tc.cu).
int make_map_plain(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                   uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    zsb_set_error("dense_tc: cuTensorMapEncodeTiled unavailable");
    return ZSB_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    zsb_set_error("dense_tc: cuTensorMapEncodeTiled (plain) failed (%d)", (int)r);
    return ZSB_ERR_CUDA;
  }
  return ZSB_OK;
}

}  // namespace
