// tcgen05 3xTF32 implementation of the dense-Gaussian leapfrog pass (impl 1) -- placeholder until
// the tensor-core kernel lands; reports "unsupported" so callers fail loudly instead of silently
// taking another path.
#include "common.cuh"

int zsb_dense_tc_ntiles(int D) { return (D + 127) / 128; }

int zsb_dense_leapfrog_tc_launch(const float*, float*, const float*, float*, const float*,
                                 const float*, const float*, const float*, const float*,
                                 const float*, float, float*, float*, int64_t, int, cudaStream_t) {
  zsb_set_error("zsb_hmc_dense_leapfrog_f32: impl 1 (tcgen05) is not built in this version");
  return ZSB_ERR_UNSUPPORTED;
}
