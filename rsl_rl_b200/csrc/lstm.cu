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

struct LstmFwdArgs {
  const float *xg, *hg, *b_hh, *c_prev;
  float *gates_act, *c_out, *h_out;
};
struct LstmFwdBatch { LstmFwdArgs net[2]; };
struct LstmBwdArgs {
  const float *dh_out, *dh_rec, *dc_next, *gates_act, *c_prev, *c_cur;
  float *dgates, *dc_prev;
};
struct LstmBwdBatch { LstmBwdArgs net[2]; };

// blockIdx.y = net (actor / critic memories of equal shape run in one launch)
__global__ void __launch_bounds__(kLstmThreads)
lstm_cell_forward_kernel(const __grid_constant__ LstmFwdBatch batch, long long n, int H) {
  const LstmFwdArgs& a = batch.net[blockIdx.y];
  const float* __restrict__ xg = a.xg; const float* __restrict__ hg = a.hg; const float* __restrict__ b_hh = a.b_hh;
  const float* __restrict__ c_prev = a.c_prev; float* __restrict__ gates_act = a.gates_act;
  float* __restrict__ c_out = a.c_out; float* __restrict__ h_out = a.h_out;
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
lstm_cell_backward_kernel(const __grid_constant__ LstmBwdBatch batch, long long n, int H) {
  const LstmBwdArgs& a = batch.net[blockIdx.y];
  const float* __restrict__ dh_out = a.dh_out; const float* __restrict__ dh_rec = a.dh_rec;
  const float* __restrict__ dc_next = a.dc_next; const float* __restrict__ gates_act = a.gates_act;
  const float* __restrict__ c_prev = a.c_prev; const float* __restrict__ c_cur = a.c_cur;
  float* __restrict__ dgates = a.dgates; float* __restrict__ dc_prev = a.dc_prev;
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

static int launch_cell_forward(const LstmFwdBatch& b, int n_nets, long long n, int H, cudaStream_t st) {
  prefer_max_smem(lstm_cell_forward_kernel, true);
  lstm_cell_forward_kernel<<<dim3(lstm_blocks(n * (H / 4)), n_nets), kLstmThreads, 0, st>>>(b, n, H);
  return finish_launch("rslrl_lstm_cell_forward");
}
static int launch_cell_backward(const LstmBwdBatch& b, int n_nets, long long n, int H, cudaStream_t st) {
  prefer_max_smem(lstm_cell_backward_kernel, true);
  lstm_cell_backward_kernel<<<dim3(lstm_blocks(n * (H / 4)), n_nets), kLstmThreads, 0, st>>>(b, n, H);
  return finish_launch("rslrl_lstm_cell_backward");
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
  LstmFwdBatch b{};
  b.net[0] = LstmFwdArgs{xg, hg, b_hh, c_prev, gates_act, c_out, h_out};
  return launch_cell_forward(b, 1, n, H, (cudaStream_t)stream);
}

int rslrl_lstm_cell_backward(const float* dh_out, const float* dh_rec, const float* dc_next, const float* gates_act,
                             const float* c_prev, const float* c_cur, float* dgates, float* dc_prev, int64_t n, int H,
                             rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_lstm_cell_backward";
  RSLRL_CHECK_ARG(gates_act && c_prev && c_cur && dgates && dc_prev && n > 0 && H > 0 && (H % 4) == 0, what);
  RSLRL_CHECK_ARG((((uintptr_t)dh_out | (uintptr_t)dh_rec | (uintptr_t)dc_next | (uintptr_t)gates_act |
                    (uintptr_t)c_prev | (uintptr_t)c_cur | (uintptr_t)dgates | (uintptr_t)dc_prev) & 15) == 0, what);
  LstmBwdBatch b{};
  b.net[0] = LstmBwdArgs{dh_out, dh_rec, dc_next, gates_act, c_prev, c_cur, dgates, dc_prev};
  return launch_cell_backward(b, 1, n, H, (cudaStream_t)stream);
}

// ---- whole-layer drivers: the time loop runs in C (one GEMM launch + one cell launch per step for ALL nets), not in the
// caller.  `nets` describes 1 or 2 LSTM layers of equal shape (the actor's and the critic's memory) that advance together:
// their recurrent projections are one grouped GEMM launch and their cells one launch per step.
//   forward:  xg [T*n,4H] hoisted input projection; h_all / c_all [(T+1),n,H] with slot 0 = initial state (filled by the
//             caller), slot t+1 = state after step t; gates [T,n,4H] (NULL: inference); hg [n,4H] scratch.
//   backward: dout [T,n,H] / dhT / dcT (each may be NULL), w_hh_t [H,4H] = W_hh transposed (dh_{t-1} = dG_t W_hh runs as
//             a forward product) -> dgates [T,n,4H]; dh_scratch / dc_scratch [2][n,H], slot 0 of each = d h_0 / d c_0 on
//             return (d h_0 only when need_dh0).
int rslrl_lstm_layers_forward(int n_nets, const rslrl_lstm_fwd_t* nets, int T, int64_t n, int H, int precision,
                              rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_lstm_layers_forward";
  RSLRL_CHECK_ARG(nets && n_nets >= 1 && n_nets <= 2 && T > 0 && n > 0 && H > 0 && (H % 4) == 0, what);
  for (int i = 0; i < n_nets; ++i)
    RSLRL_CHECK_ARG(nets[i].xg && nets[i].w_hh && nets[i].h_all && nets[i].c_all && nets[i].hg, what);
  const size_t nH = (size_t)n * H, nG = (size_t)n * 4 * H;
  for (int t = 0; t < T; ++t) {
    rslrl_linear_fwd_t p[2] = {};
    LstmFwdBatch cb{};
    for (int i = 0; i < n_nets; ++i) {
      const rslrl_lstm_fwd_t& q = nets[i];
      p[i].x = q.h_all + t * nH; p[i].w = q.w_hh; p[i].b = nullptr; p[i].y = q.hg;
      p[i].M = (int)n; p[i].N = 4 * H; p[i].K = H; p[i].ldx = H; p[i].ldy = 4 * H; p[i].act = RSLRL_ACT_IDENTITY;
      cb.net[i] = LstmFwdArgs{q.xg + t * nG, q.hg, q.b_hh, q.c_all + t * nH, q.gates ? q.gates + t * nG : nullptr,
                              q.c_all + (t + 1) * nH, q.h_all + (t + 1) * nH};
    }
    int rc = rslrl_linear_forward(n_nets, p, precision, stream);
    if (rc != 0) return rc;
    rc = launch_cell_forward(cb, n_nets, n, H, (cudaStream_t)stream);
    if (rc != 0) return rc;
  }
  return 0;
}

int rslrl_lstm_layers_backward(int n_nets, const rslrl_lstm_bwd_t* nets, int need_dh0, int T, int64_t n, int H,
                               int precision, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_lstm_layers_backward";
  RSLRL_CHECK_ARG(nets && n_nets >= 1 && n_nets <= 2 && T > 0 && n > 0 && H > 0 && (H % 4) == 0, what);
  for (int i = 0; i < n_nets; ++i)
    RSLRL_CHECK_ARG(nets[i].gates && nets[i].c_all && nets[i].w_hh_t && nets[i].dgates && nets[i].dh_scratch &&
                        nets[i].dc_scratch, what);
  const size_t nH = (size_t)n * H, nG = (size_t)n * 4 * H;
  const float* dh_rec[2] = {nets[0].dhT, n_nets > 1 ? nets[1].dhT : nullptr};
  const float* dc[2] = {nets[0].dcT, n_nets > 1 ? nets[1].dcT : nullptr};
  for (int t = T - 1; t >= 0; --t) {
    LstmBwdBatch cb{};
    for (int i = 0; i < n_nets; ++i) {
      const rslrl_lstm_bwd_t& q = nets[i];
      float* dc_prev = q.dc_scratch + (size_t)(t & 1) * nH;
      cb.net[i] = LstmBwdArgs{q.dout ? q.dout + t * nH : nullptr, dh_rec[i], dc[i], q.gates + t * nG, q.c_all + t * nH,
                              q.c_all + (t + 1) * nH, q.dgates + t * nG, dc_prev};
      dc[i] = dc_prev;
    }
    int rc = launch_cell_backward(cb, n_nets, n, H, (cudaStream_t)stream);
    if (rc != 0) return rc;
    if (t > 0 || need_dh0) {
      // dh_{t-1} = dG_t W_hh as the FORWARD product dG_t (W_hh^T)^T: a plain K-major GEMM without activation, which the
      // tensor-core forward kernel covers (the data-gradient kernel is specialised to activation derivatives)
      rslrl_linear_fwd_t p[2] = {};
      for (int i = 0; i < n_nets; ++i) {
        const rslrl_lstm_bwd_t& q = nets[i];
        float* dh = q.dh_scratch + (size_t)(t & 1) * nH;
        p[i].x = q.dgates + t * nG; p[i].w = q.w_hh_t; p[i].b = nullptr; p[i].y = dh;
        p[i].M = (int)n; p[i].N = H; p[i].K = 4 * H; p[i].ldx = 4 * H; p[i].ldy = H; p[i].act = RSLRL_ACT_IDENTITY;
        dh_rec[i] = dh;
      }
      rc = rslrl_linear_forward(n_nets, p, precision, stream);
      if (rc != 0) return rc;
    }
  }
  return 0;
}

}  // extern "C"
