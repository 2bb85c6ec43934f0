// tcgen05 3xTF32 gather-GEMM (placeholder until the tensor-core kernel lands in this file).
#include "common.cuh"

extern "C" int p3d_sparse_conv_gather_gemm_tf32x3(const float *, const int32_t *, const int32_t *, int64_t, int, int,
                                                  int, const float *, const float *, const float *, const float *,
                                                  int, float *, p3d_stream_t) {
  return P3D_ERR_UNSUPPORTED;
}
