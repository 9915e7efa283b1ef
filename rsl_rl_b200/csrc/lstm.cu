// LSTM cell, forward and backward, as fused elementwise kernels around the library's GEMMs.
// Replaces the recurrence inside rsl_rl/networks/memory.py:27-37 (`nn.LSTM`: cuDNN on the reference's CUDA path).
//
// The dense parts of an LSTM layer are ordinary products and run on the GEMM kernels of this library (host side:
// rsl_rl_b200/networks/lstm.py):
//   * input projection  Xg = X W_ih^T + b_ih  HOISTED over all T steps of all trajectories: one [T*n, D] x [D, 4H] launch;
//   * per step          Hg = h_{t-1} W_hh^T   ([n, H] x [H, 4H]);
//   * backward          dh_{t-1} = dG_t W_hh per step; dW_ih, dW_hh, db as TWO products over all T*n rows after the loop
//                       (every dG_t is kept), instead of 2 T small ones.
// What remains per step is elementwise and lives here, one launch each:
//   forward   gates = Xg_t + Hg + b_hh -> i, f, o = sigmoid, g = tanh (PyTorch gate order i, f, g, o);
//             c_t = f c_{t-1} + i g;  h_t = o tanh(c_t);  activated gates saved for the backward pass
//   backward  dh = dh_out + dh_rec;  do = dh tanh(c_t);  dc = dh o (1 - tanh^2 c_t) + dc_next;
//             dG = (dc g i(1-i), dc c_{t-1} f(1-f), dc i (1-g^2), do o(1-o));  dc_{t-1} = dc f
// Thread <-> (row, 4 consecutive hidden units): 16-byte accesses, HBM-bound (forward 2*4H + H in, 4H + 2H out per row).
#include "common.cuh"

namespace rslrl {

constexpr int kLstmThreads = 256;

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(kLstmThreads)
lstm_cell_forward_kernel(const float* __restrict__ xg, const float* __restrict__ hg, const float* __restrict__ b_hh,
                         const float* __restrict__ c_prev, float* __restrict__ gates_act, float* __restrict__ c_out,
                         float* __restrict__ h_out, long long n, int H) {
  const int H4 = H >> 2;
  const long long total = n * H4;
  for (long long idx = (long long)blockIdx.x * kLstmThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kLstmThreads) {
    const long long r = idx / H4;
    const int j = (int)(idx - r * H4) * 4;
    const float* xr = xg + r * 4 * H;
    const float* hr = hg + r * 4 * H;
    float4 pre[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 a = *(const float4*)(xr + g * H + j), b = *(const float4*)(hr + g * H + j);
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b_hh != nullptr) c = __ldg((const float4*)(b_hh + g * H + j));
      pre[g] = make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
    }
    const float4 cp = *(const float4*)(c_prev + r * H + j);
    float4 gi, gf, gg, go, cn, hn;
#define RSLRL_LSTM_LANE(m)                                      \
    gi.m = sigmoidf_acc(pre[0].m); gf.m = sigmoidf_acc(pre[1].m); \
    gg.m = tanhf(pre[2].m); go.m = sigmoidf_acc(pre[3].m);       \
    cn.m = gf.m * cp.m + gi.m * gg.m; hn.m = go.m * tanhf(cn.m);
    RSLRL_LSTM_LANE(x) RSLRL_LSTM_LANE(y) RSLRL_LSTM_LANE(z) RSLRL_LSTM_LANE(w)
#undef RSLRL_LSTM_LANE
    if (gates_act != nullptr) {
      float* ga = gates_act + r * 4 * H;
      *(float4*)(ga + j) = gi;
      *(float4*)(ga + H + j) = gf;
      *(float4*)(ga + 2 * H + j) = gg;
      *(float4*)(ga + 3 * H + j) = go;
    }
    *(float4*)(c_out + r * H + j) = cn;
    *(float4*)(h_out + r * H + j) = hn;
  }
}

__global__ void __launch_bounds__(kLstmThreads)
lstm_cell_backward_kernel(const float* __restrict__ dh_out, const float* __restrict__ dh_rec,
                          const float* __restrict__ dc_next, const float* __restrict__ gates_act,
                          const float* __restrict__ c_prev, const float* __restrict__ c_cur,
                          float* __restrict__ dgates, float* __restrict__ dc_prev, long long n, int H) {
  const int H4 = H >> 2;
  const long long total = n * H4;
  for (long long idx = (long long)blockIdx.x * kLstmThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kLstmThreads) {
    const long long r = idx / H4;
    const int j = (int)(idx - r * H4) * 4;
    const float* ga = gates_act + r * 4 * H;
    const float4 gi = *(const float4*)(ga + j), gf = *(const float4*)(ga + H + j), gg = *(const float4*)(ga + 2 * H + j),
                 go = *(const float4*)(ga + 3 * H + j);
    const float4 cp = *(const float4*)(c_prev + r * H + j), cc = *(const float4*)(c_cur + r * H + j);
    float4 dh = make_float4(0.f, 0.f, 0.f, 0.f), dcn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dh_out != nullptr) dh = *(const float4*)(dh_out + r * H + j);
    if (dh_rec != nullptr) {
      const float4 d2 = *(const float4*)(dh_rec + r * H + j);
      dh.x += d2.x; dh.y += d2.y; dh.z += d2.z; dh.w += d2.w;
    }
    if (dc_next != nullptr) dcn = *(const float4*)(dc_next + r * H + j);
    float4 di, df, dg, dO, dcp;
#define RSLRL_LSTM_LANE(m)                                                        \
    {                                                                             \
      const float tc = tanhf(cc.m);                                               \
      const float d_o = dh.m * tc;                                                \
      const float dc = dh.m * go.m * (1.f - tc * tc) + dcn.m;                     \
      di.m = dc * gg.m * gi.m * (1.f - gi.m);                                     \
      df.m = dc * cp.m * gf.m * (1.f - gf.m);                                     \
      dg.m = dc * gi.m * (1.f - gg.m * gg.m);                                     \
      dO.m = d_o * go.m * (1.f - go.m);                                           \
      dcp.m = dc * gf.m;                                                          \
    }
    RSLRL_LSTM_LANE(x) RSLRL_LSTM_LANE(y) RSLRL_LSTM_LANE(z) RSLRL_LSTM_LANE(w)
#undef RSLRL_LSTM_LANE
    float* dgr = dgates + r * 4 * H;
    *(float4*)(dgr + j) = di;
    *(float4*)(dgr + H + j) = df;
    *(float4*)(dgr + 2 * H + j) = dg;
    *(float4*)(dgr + 3 * H + j) = dO;
    *(float4*)(dc_prev + r * H + j) = dcp;
  }
}

static int lstm_blocks(long long work) {
  long long b = (work + kLstmThreads - 1) / kLstmThreads;
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  return b < 1 ? 1 : (int)b;
}

}  // namespace rslrl

extern "C" {

int rslrl_lstm_cell_forward(const float* xg, const float* hg, const float* b_hh, const float* c_prev, float* gates_act,
                            float* c_out, float* h_out, int64_t n, int H, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_lstm_cell_forward";
  RSLRL_CHECK_ARG(xg && hg && c_prev && c_out && h_out && n > 0 && H > 0 && (H % 4) == 0, what);
  RSLRL_CHECK_ARG((((uintptr_t)xg | (uintptr_t)hg | (uintptr_t)b_hh | (uintptr_t)c_prev | (uintptr_t)gates_act |
                    (uintptr_t)c_out | (uintptr_t)h_out) & 15) == 0, what);
  lstm_cell_forward_kernel<<<lstm_blocks(n * (H / 4)), kLstmThreads, 0, (cudaStream_t)stream>>>(
      xg, hg, b_hh, c_prev, gates_act, c_out, h_out, n, H);
  return finish_launch(what);
}

int rslrl_lstm_cell_backward(const float* dh_out, const float* dh_rec, const float* dc_next, const float* gates_act,
                             const float* c_prev, const float* c_cur, float* dgates, float* dc_prev, int64_t n, int H,
                             rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_lstm_cell_backward";
  RSLRL_CHECK_ARG(gates_act && c_prev && c_cur && dgates && dc_prev && n > 0 && H > 0 && (H % 4) == 0, what);
  RSLRL_CHECK_ARG((((uintptr_t)dh_out | (uintptr_t)dh_rec | (uintptr_t)dc_next | (uintptr_t)gates_act |
                    (uintptr_t)c_prev | (uintptr_t)c_cur | (uintptr_t)dgates | (uintptr_t)dc_prev) & 15) == 0, what);
  lstm_cell_backward_kernel<<<lstm_blocks(n * (H / 4)), kLstmThreads, 0, (cudaStream_t)stream>>>(
      dh_out, dh_rec, dc_next, gates_act, c_prev, c_cur, dgates, dc_prev, n, H);
  return finish_launch(what);
}

// ---- whole-layer drivers: the time loop runs in C (one GEMM + one cell launch per step), not in the caller -----------
// xg [T*n,4H] is the hoisted input projection; h_all / c_all are [(T+1),n,H] with slot 0 = initial state (the caller
// fills it), slot t+1 = state after step t; gates [T,n,4H] (NULL: inference); hg [n,4H] scratch.
int rslrl_lstm_layer_forward(const float* xg, const float* w_hh, const float* b_hh, float* h_all, float* c_all,
                             float* gates, float* hg, int T, int64_t n, int H, int precision, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_lstm_layer_forward";
  RSLRL_CHECK_ARG(xg && w_hh && h_all && c_all && hg && T > 0 && n > 0 && H > 0 && (H % 4) == 0, what);
  const size_t nH = (size_t)n * H, nG = (size_t)n * 4 * H;
  for (int t = 0; t < T; ++t) {
    rslrl_linear_fwd_t p{};
    p.x = h_all + t * nH; p.w = w_hh; p.b = nullptr; p.y = hg;
    p.M = (int)n; p.N = 4 * H; p.K = H; p.ldx = H; p.ldy = 4 * H; p.act = RSLRL_ACT_IDENTITY;
    int rc = rslrl_linear_forward(1, &p, precision, stream);
    if (rc != 0) return rc;
    rc = rslrl_lstm_cell_forward(xg + t * nG, hg, b_hh, c_all + t * nH, gates ? gates + t * nG : nullptr,
                                 c_all + (t + 1) * nH, h_all + (t + 1) * nH, n, H, stream);
    if (rc != 0) return rc;
  }
  return 0;
}

// Backward through time: dout [T,n,H] (may be NULL), dhT / dcT gradients of the final state (may be NULL), w_hh_t [H,4H] =
// W_hh transposed ->
// dgates [T,n,4H]; dh_scratch / dc_scratch are [2][n,H]; on return slot 0 of each holds d h_0 / d c_0
// (d h_0 only when need_dh0).
int rslrl_lstm_layer_backward(const float* dout, const float* dhT, const float* dcT, const float* gates,
                              const float* c_all, const float* w_hh_t, float* dgates, float* dh_scratch,
                              float* dc_scratch, int need_dh0, int T, int64_t n, int H, int precision,
                              rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_lstm_layer_backward";
  RSLRL_CHECK_ARG(gates && c_all && w_hh_t && dgates && dh_scratch && dc_scratch && T > 0 && n > 0 && H > 0 && (H % 4) == 0,
                  what);
  const size_t nH = (size_t)n * H, nG = (size_t)n * 4 * H;
  const float* dh_rec = dhT;
  const float* dc = dcT;
  for (int t = T - 1; t >= 0; --t) {
    float* dc_prev = dc_scratch + (size_t)(t & 1) * nH;
    int rc = rslrl_lstm_cell_backward(dout ? dout + t * nH : nullptr, dh_rec, dc, gates + t * nG, c_all + t * nH,
                                      c_all + (t + 1) * nH, dgates + t * nG, dc_prev, n, H, stream);
    if (rc != 0) return rc;
    dc = dc_prev;
    if (t > 0 || need_dh0) {
      // dh_{t-1} = dG_t W_hh, issued as the FORWARD product dG_t (W_hh^T)^T against the transposed copy of W_hh the caller
      // passes ([H,4H] row-major): a plain K-major GEMM without activation, which the tensor-core forward kernel covers
      float* dh = dh_scratch + (size_t)(t & 1) * nH;
      rslrl_linear_fwd_t q{};
      q.x = dgates + t * nG; q.w = w_hh_t; q.b = nullptr; q.y = dh;
      q.M = (int)n; q.N = H; q.K = 4 * H; q.ldx = 4 * H; q.ldy = H; q.act = RSLRL_ACT_IDENTITY;
      rc = rslrl_linear_forward(1, &q, precision, stream);
      if (rc != 0) return rc;
      dh_rec = dh;
    }
  }
  return 0;
}

}  // extern "C"
