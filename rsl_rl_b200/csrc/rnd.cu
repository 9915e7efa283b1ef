// Intrinsic reward of Random Network Distillation in ONE launch per env step:
//   reward[i] = scale * || target(x_i) - predictor(x_i) ||_2
// replacing, per env step, two small MLP forwards (4-6 GEMM launches), a subtraction, a norm and a scale
// (reference rsl_rl/modules/rnd.py:113-135: get_intrinsic_reward).  The two embedding networks are tiny (the
// reference's defaults: rnd_state -> hidden -> num_outputs with widths of a few dozen), so their weights live in shared
// memory for the whole launch and the launch is bound by reading x once: 4 * D bytes per env.
//
// A block walks tiles of 32 rows.  Thread t <-> row t % 32 of the tile and the output units (t / 32) + 4 j: the input
// activations of a row sit in a shared-memory row of odd pitch (lanes = rows: conflict-free), the weight of a unit is
// warp-uniform (broadcast), four units are accumulated together so that one activation load feeds four FMAs.
// fp32 FMA arithmetic throughout (the fp32 class of the reference).
#include "common.cuh"

namespace rslrl {

constexpr int kRndThreads = 128;
constexpr int kRndRows = 32;
constexpr int kRndMaxLayers = 4;
constexpr int kRndMaxWidth = 128;

struct RndNet {
  const float* w[kRndMaxLayers];   // [dims[l+1], dims[l]] row-major
  const float* b[kRndMaxLayers];   // [dims[l+1]] or null
  int act[kRndMaxLayers];          // activation after layer l
  int dims[kRndMaxLayers + 1];
  int L;
};

struct RndParams {
  RndNet net[2];   // 0: target, 1: predictor
  const float* x;  // [n, D] with row pitch ldx
  long long ldx, n;
  float* out;      // [n]
  float scale;
  int pitch;       // shared-memory row pitch of the activation tiles (odd)
};

__global__ void __launch_bounds__(kRndThreads) rnd_reward_kernel(const RndParams p) {
  extern __shared__ __align__(16) float rnd_smem[];
  // layout: weights + biases of both nets | tile A | tile B | target output tile
  float* wbase = rnd_smem;
  int off = 0;
  int woff[2][kRndMaxLayers], boff[2][kRndMaxLayers];
  for (int k = 0; k < 2; ++k)
    for (int l = 0; l < p.net[k].L; ++l) {
      woff[k][l] = off;
      off += p.net[k].dims[l + 1] * p.net[k].dims[l];
      boff[k][l] = off;
      off += p.net[k].dims[l + 1];
    }
  float* const tileA = wbase + ((off + 3) & ~3);
  float* const tileB = tileA + kRndRows * p.pitch;
  float* const tileT = tileB + kRndRows * p.pitch;
  for (int k = 0; k < 2; ++k)
    for (int l = 0; l < p.net[k].L; ++l) {
      const int nw = p.net[k].dims[l + 1] * p.net[k].dims[l], nb = p.net[k].dims[l + 1];
      for (int i = threadIdx.x; i < nw; i += kRndThreads) wbase[woff[k][l] + i] = __ldg(p.net[k].w[l] + i);
      for (int i = threadIdx.x; i < nb; i += kRndThreads)
        wbase[boff[k][l] + i] = p.net[k].b[l] ? __ldg(p.net[k].b[l] + i) : 0.f;
    }
  __syncthreads();

  const int r = threadIdx.x & 31, u0 = threadIdx.x >> 5;
  const int D = p.net[0].dims[0];
  const int Dout = p.net[0].dims[p.net[0].L];
  for (long long row0 = (long long)blockIdx.x * kRndRows; row0 < p.n; row0 += (long long)gridDim.x * kRndRows) {
    for (int k = 0; k < 2; ++k) {
      // input tile -> tileA (coalesced along the columns of a row)
      for (int i = threadIdx.x; i < kRndRows * D; i += kRndThreads) {
        const int rr = i / D, c = i - rr * D;
        tileA[rr * p.pitch + c] = row0 + rr < p.n ? __ldg(p.x + (row0 + rr) * p.ldx + c) : 0.f;
      }
      __syncthreads();
      const float* in = tileA;
      float* outp = tileB;
      const RndNet& net = p.net[k];
      for (int l = 0; l < net.L; ++l) {
        const int K = net.dims[l], Nn = net.dims[l + 1];
        const float* W = wbase + woff[k][l];
        const float* B = wbase + boff[k][l];
        const bool last = l == net.L - 1;
        float* dst = last ? (k == 0 ? tileT : outp) : outp;
        for (int u = u0; u < Nn; u += 16) {   // units u, u + 4, u + 8, u + 12
          float acc[4];
          const float* wr[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int uu = u + 4 * j < Nn ? u + 4 * j : u;
            acc[j] = B[uu];
            wr[j] = W + uu * K;
          }
          const float* xr = in + r * p.pitch;
          for (int kk = 0; kk < K; ++kk) {
            const float xv = xr[kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fmaf(xv, wr[j][kk], acc[j]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (u + 4 * j < Nn) dst[r * p.pitch + u + 4 * j] = act_forward(net.act[l], acc[j]);
        }
        __syncthreads();
        in = dst;
        outp = (dst == tileB) ? tileA : tileB;
      }
      if (k == 1) {
        // `in` = predictor output tile, tileT = target output tile
        if (threadIdx.x < kRndRows && row0 + threadIdx.x < p.n) {
          float s = 0.f;
          for (int c = 0; c < Dout; ++c) {
            const float g = tileT[threadIdx.x * p.pitch + c] - in[threadIdx.x * p.pitch + c];
            s = fmaf(g, g, s);
          }
          p.out[row0 + threadIdx.x] = sqrtf(s) * p.scale;
        }
        __syncthreads();
      }
    }
  }
}

}  // namespace rslrl

extern "C" int rslrl_rnd_reward(const rslrl_rnd_net_t* target, const rslrl_rnd_net_t* predictor, const float* x,
                                int64_t ldx, int64_t n, float* out, double scale, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_rnd_reward";
  RSLRL_CHECK_ARG(target && predictor && x && out && n > 0, what);
  RndParams p{};
  const rslrl_rnd_net_t* nets[2] = {target, predictor};
  int maxw = 0;
  long long floats = 0;
  for (int k = 0; k < 2; ++k) {
    const rslrl_rnd_net_t& q = *nets[k];
    RSLRL_CHECK_ARG(q.num_layers >= 1 && q.num_layers <= kRndMaxLayers, what);
    p.net[k].L = q.num_layers;
    for (int l = 0; l <= q.num_layers; ++l) {
      RSLRL_CHECK_ARG(q.dims[l] >= 1 && q.dims[l] <= kRndMaxWidth, what);
      p.net[k].dims[l] = q.dims[l];
      if (q.dims[l] > maxw) maxw = q.dims[l];
    }
    for (int l = 0; l < q.num_layers; ++l) {
      RSLRL_CHECK_ARG(q.w[l] != nullptr, what);
      p.net[k].w[l] = q.w[l];
      p.net[k].b[l] = q.b[l];
      p.net[k].act[l] = q.act[l];
      floats += (long long)q.dims[l + 1] * q.dims[l] + q.dims[l + 1];
    }
  }
  RSLRL_CHECK_ARG(target->dims[0] == predictor->dims[0] && ldx >= target->dims[0], what);
  RSLRL_CHECK_ARG(target->dims[target->num_layers] == predictor->dims[predictor->num_layers], what);
  p.x = x; p.ldx = ldx; p.n = n; p.out = out; p.scale = (float)scale;
  p.pitch = maxw | 1;
  const size_t smem = (size_t)(((floats + 3) & ~3LL) + 3LL * kRndRows * p.pitch) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("%s: the two networks (%lld parameters) do not fit in shared memory", what, floats);
    return 4;
  }
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    if (cudaFuncSetAttribute(rnd_reward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute failed", what);
      (void)cudaGetLastError();
      return 1;
    }
    smem_set = 200 * 1024;
  }
  long long tiles = (n + kRndRows - 1) / kRndRows;
  const long long cap = (long long)kNumSMs * 4;
  const int grid = (int)(tiles < cap ? tiles : cap);
  rnd_reward_kernel<<<grid, kRndThreads, smem, (cudaStream_t)stream>>>(p);
  return finish_launch(what);
}
