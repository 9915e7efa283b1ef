// TF32 tensor-core GEMMs (tcgen05 + TMEM + TMA) -- placeholder until the kernels land.
#include "common.cuh"

namespace rslrl {
int tc_linear_forward(int, const rslrl_linear_fwd_t*, cudaStream_t) {
  set_error("rslrl_linear_forward: RSLRL_PREC_TF32 kernels are not built into this library yet");
  return 3;
}
int tc_linear_backward_data(int, const rslrl_linear_bwd_data_t*, cudaStream_t) {
  set_error("rslrl_linear_backward_data: RSLRL_PREC_TF32 kernels are not built into this library yet");
  return 3;
}
int tc_linear_backward_weight(int, const rslrl_linear_bwd_weight_t*, cudaStream_t) {
  set_error("rslrl_linear_backward_weight: RSLRL_PREC_TF32 kernels are not built into this library yet");
  return 3;
}
}  // namespace rslrl
