// C-ABI housekeeping + precision dispatch for libp3d_b200.
#include "common.cuh"

#include <stdlib.h>

namespace p3d {
thread_local int g_last_cuda_error = 0;
bool pdl_enabled() {
  static const bool on = !(getenv("P3D_PDL") && atoi(getenv("P3D_PDL")) == 0);
  return on;
}
}  // namespace p3d

extern "C" int p3d_sparse_conv_gather_gemm_fp32(const float *in, const int32_t *nbr, const int32_t *n_out_dev,
                                                int64_t n_out_cap, int K, int Cin, int Cout, const float *weight,
                                                const float *scale, const float *shift, const float *residual,
                                                int relu, float *out, p3d_stream_t stream);
extern "C" int p3d_sparse_conv_gather_gemm_tf32x3(const float *in, const int32_t *nbr, const int32_t *n_out_dev,
                                                  int64_t n_out_cap, int K, int Cin, int Cout, const float *weight,
                                                  const float *scale, const float *shift, const float *residual,
                                                  int relu, float *out, p3d_stream_t stream);

extern "C" const char *p3d_status_string(int status) {
  switch (status) {
    case P3D_OK: return "ok";
    case P3D_ERR_INVALID_ARG: return "invalid argument (shape, null pointer, alignment or attribute)";
    case P3D_ERR_WORKSPACE: return "workspace smaller than p3d_*_workspace_bytes()";
    case P3D_ERR_CUDA: return "CUDA runtime error (see p3d_last_cuda_error)";
    case P3D_ERR_UNSUPPORTED: return "valid in the reference but outside this build's limits";
    default: return "unknown status";
  }
}

extern "C" int p3d_last_cuda_error(void) { return p3d::g_last_cuda_error; }
extern "C" int p3d_abi_version(void) { return 1; }

extern "C" int p3d_sparse_conv_gather_gemm(const float *in, const int32_t *nbr, const int32_t *n_out_dev,
                                           int64_t n_out_cap, int K, int Cin, int Cout, const float *weight,
                                           const float *scale, const float *shift, const float *residual, int relu,
                                           int precision, float *out, p3d_stream_t stream) {
  if (precision == P3D_CONV_FP32)
    return p3d_sparse_conv_gather_gemm_fp32(in, nbr, n_out_dev, n_out_cap, K, Cin, Cout, weight, scale, shift,
                                            residual, relu, out, stream);
  if (precision == P3D_CONV_TF32X3)
    return p3d_sparse_conv_gather_gemm_tf32x3(in, nbr, n_out_dev, n_out_cap, K, Cin, Cout, weight, scale, shift,
                                              residual, relu, out, stream);
  return P3D_ERR_INVALID_ARG;
}
