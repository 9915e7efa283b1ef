// C-ABI entry points for the dense layers: dispatch on the arithmetic mode only (one device target,
// no fallbacks).  fp32 -> gemm_simt.cu, tf32 / tf32x3 (3xTF32, fp32-class) -> gemm_tc.cu (tcgen05 / TMEM / TMA).
#include "common.cuh"

namespace rslrl {
int simt_linear_forward(int n, const rslrl_linear_fwd_t* pr, cudaStream_t st);
int simt_linear_backward_data(int n, const rslrl_linear_bwd_data_t* pr, cudaStream_t st);
int simt_linear_backward_weight(int n, const rslrl_linear_bwd_weight_t* pr, cudaStream_t st);
int tc_linear_forward(int n, const rslrl_linear_fwd_t* pr, cudaStream_t st, bool x3 = false);
int head_linear_backward(int n, const rslrl_linear_bwd_head_t* pr, cudaStream_t st);
int tc_linear_backward_data(int n, const rslrl_linear_bwd_data_t* pr, cudaStream_t st, bool x3 = false);
int tc_linear_backward_weight(int n, const rslrl_linear_bwd_weight_t* pr, cudaStream_t st, bool x3 = false);
int tc16_linear_backward_data(int n, const rslrl_linear_bwd_data_t* pr, cudaStream_t st);
int tc16_linear_backward_weight(int n, const rslrl_linear_bwd_weight_t* pr, cudaStream_t st);
}  // namespace rslrl

extern "C" {

int rslrl_linear_forward(int n, const rslrl_linear_fwd_t* probs_host, int precision, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_linear_forward";
  RSLRL_CHECK_ARG(n > 0 && n <= RSLRL_MAX_PROBLEMS && probs_host, what);
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_fwd_t& p = probs_host[i];
    RSLRL_CHECK_ARG(p.x && p.w && p.y && p.M > 0 && p.N > 0 && p.K > 0 && p.ldx >= p.K && p.ldy >= p.N, what);
  }
  if (precision == RSLRL_PREC_FP32) return simt_linear_forward(n, probs_host, (cudaStream_t)stream);
  if (precision == RSLRL_PREC_TF32 || precision == RSLRL_PREC_TF32X3) {
    const int rc = tc_linear_forward(n, probs_host, (cudaStream_t)stream, precision == RSLRL_PREC_TF32X3);
    // -1: an operand cannot be described to TMA (row pitch not a multiple of 16 bytes) -> exact fp32 kernel
    return rc >= 0 ? rc : simt_linear_forward(n, probs_host, (cudaStream_t)stream);
  }
  set_error("%s: unknown precision %d", what, precision);
  return 2;
}

int rslrl_linear_backward_data(int n, const rslrl_linear_bwd_data_t* probs_host, int precision,
                               rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_linear_backward_data";
  RSLRL_CHECK_ARG(n > 0 && n <= RSLRL_MAX_PROBLEMS && probs_host, what);
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_bwd_data_t& p = probs_host[i];
    RSLRL_CHECK_ARG(p.dy && p.w && p.dx && p.M > 0 && p.N > 0 && p.K > 0 && p.ldy >= p.N && p.ldx >= p.K, what);
    RSLRL_CHECK_ARG(p.act == RSLRL_ACT_IDENTITY || (p.h && p.ldh >= p.K), what);
  }
  if (precision == RSLRL_PREC_FP32) return simt_linear_backward_data(n, probs_host, (cudaStream_t)stream);
  if (precision == RSLRL_PREC_TF32 || precision == RSLRL_PREC_TF32X3) {
    const int rc = tc_linear_backward_data(n, probs_host, (cudaStream_t)stream, precision == RSLRL_PREC_TF32X3);
    return rc >= 0 ? rc : simt_linear_backward_data(n, probs_host, (cudaStream_t)stream);
  }
  if (precision == RSLRL_PREC_FP16) {
    const int rc = tc16_linear_backward_data(n, probs_host, (cudaStream_t)stream);
    if (rc < 0) {
      set_error("%s(fp16): shape not eligible (needs K %% 64 == 0, an activation with saved fp16 h, 16-byte aligned "
                "rows, at most 2 problems)", what);
      return 4;
    }
    return rc;
  }
  set_error("%s: unknown precision %d", what, precision);
  return 2;
}

int rslrl_linear_backward_weight(int n, const rslrl_linear_bwd_weight_t* probs_host, int precision,
                                 rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_linear_backward_weight";
  RSLRL_CHECK_ARG(n > 0 && n <= RSLRL_MAX_PROBLEMS && probs_host, what);
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_bwd_weight_t& p = probs_host[i];
    RSLRL_CHECK_ARG(p.dy && p.x && p.dw && p.M > 0 && p.N > 0 && p.K > 0 && p.ldy >= p.N && p.ldx >= p.K, what);
    RSLRL_CHECK_ARG(p.scale_dev == nullptr || precision == RSLRL_PREC_FP16, what);  // dynamic scale: fp16 chain only
  }
  if (precision == RSLRL_PREC_FP32) return simt_linear_backward_weight(n, probs_host, (cudaStream_t)stream);
  if (precision == RSLRL_PREC_TF32 || precision == RSLRL_PREC_TF32X3) {
    const int rc = tc_linear_backward_weight(n, probs_host, (cudaStream_t)stream, precision == RSLRL_PREC_TF32X3);
    if (rc < 0) return simt_linear_backward_weight(n, probs_host, (cudaStream_t)stream);
    return rc;
  }
  if (precision == RSLRL_PREC_FP16) {
    const int rc = tc16_linear_backward_weight(n, probs_host, (cudaStream_t)stream);
    if (rc < 0) {
      set_error("%s(fp16): shape not eligible (needs 16-byte aligned rows of dy and x, at most 8 problems)", what);
      return 4;
    }
    return rc;
  }
  set_error("%s: unknown precision %d", what, precision);
  return 2;
}

int rslrl_linear_backward_head(int n, const rslrl_linear_bwd_head_t* probs_host, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_linear_backward_head";
  RSLRL_CHECK_ARG(n > 0 && n <= RSLRL_MAX_PROBLEMS && probs_host, what);
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_bwd_head_t& p = probs_host[i];
    RSLRL_CHECK_ARG(p.dy && p.w && p.h && p.dw && p.M > 0 && p.N > 0 && p.K > 0 && p.ldy >= p.N && p.ldh >= p.K, what);
  }
  const int rc = head_linear_backward(n, probs_host, (cudaStream_t)stream);
  if (rc < 0) {
    set_error("%s: problem does not qualify (needs N <= 12, K %% 4 == 0, 16-byte aligned rows)", what);
    return 4;
  }
  return rc;
}

}  // extern "C"
