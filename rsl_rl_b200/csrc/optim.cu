// Gradient-norm clip + KL-adaptive learning rate + Adam on flat buffers, scalars resident on device.
//
// Reference: rsl_rl/algorithms/ppo.py:260-294 (KL rule; two host syncs per minibatch), :380
// (clip_grad_norm_: _foreach_norm + stack + vector_norm + _foreach_mul_), :381 (Adam, ~8 foreach
// launches over 17 tensors).  Here all parameters / gradients / Adam moments live in four flat fp32
// buffers (parameters are views), so the step is two launches with no host round trip:
//   pass 1  rslrl_grad_norm_lr : sum g^2 (fp64 partials, last-block finalise) -> clip scale, new lr,
//                                t += 1, bias corrections  (rslrl_opt_state_t in HBM)
//   pass 2  rslrl_adam_step    : 16-byte vectorised elementwise update, optional gradient clear.
// HBM-bound: pass 1 reads 4P bytes, pass 2 reads 16P and writes 12P (+4P clear) -> 32P bytes/step.
#include "common.cuh"

namespace rslrl {

constexpr int kOptThreads = 256;

constexpr int kOptMaxBlocks = kNumSMs * 4;

// per-block partials summed by the last block in index order: every rank derives bit-identical clip_scale / lr from
// identical (all-reduced) gradients, whatever order its blocks retire in
struct NormWorkspace {
  double partial[kOptMaxBlocks];
  unsigned int ticket;
  unsigned int pad_;
};

__global__ void __launch_bounds__(kOptThreads)
grad_norm_lr_kernel(const float* __restrict__ g, long long P, const float* __restrict__ kl_sum,
                    rslrl_opt_state_t* state, NormWorkspace* ws, int world, double max_norm, int adaptive,
                    double desired_kl, int round_lr_fp32, double beta1, double beta2) {
  __shared__ double s_red[32];
  __shared__ bool s_last;
  __shared__ double s_total;
  double acc[1] = {0.0};
  const long long tid = (long long)blockIdx.x * kOptThreads + threadIdx.x;
  const long long stride = (long long)gridDim.x * kOptThreads;
  const long long P4 = P >> 2;
  const float4* g4 = (const float4*)g;
  for (long long i = tid; i < P4; i += stride) {
    const float4 v = __ldg(g4 + i);
    acc[0] += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  for (long long i = (P4 << 2) + tid; i < P; i += stride) {
    const float v = __ldg(g + i);
    acc[0] += (double)v * v;
  }
  block_sum<1>(acc, s_red);
  if (threadIdx.x == 0) {
    ws->partial[blockIdx.x] = acc[0];
    __threadfence();
    s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 32) {  // fixed-shape tree over the partials in index order (deterministic)
    double s = 0.0;
    for (int j = threadIdx.x; j < (int)gridDim.x; j += 32) s += *((volatile double*)&ws->partial[j]);
    s = warp_sum(s);
    if (threadIdx.x == 0) s_total = s;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double sumsq = s_total;
  const double w = (double)world;
  const double norm = sqrt(sumsq) / w;  // ||mean over ranks of g||
  double coef = 1.0;
  if (max_norm > 0.0) {
    coef = max_norm / (norm + 1e-6);
    coef = coef < 1.0 ? coef : 1.0;
  }
  double lr = state->lr;
  if (adaptive) {
    // ppo.py:272-284: kl is an fp32 0-dim tensor; comparisons happen in fp32
    const float kl = kl_sum ? (*kl_sum) / (float)world : 0.f;
    if (kl > (float)(desired_kl * 2.0)) lr = fmax(1e-5, lr / 1.5);
    else if (kl < (float)(desired_kl / 2.0) && kl > 0.0f) lr = fmin(1e-2, lr * 1.5);
    if (round_lr_fp32) lr = (double)(float)lr;  // ppo.py:288-290 broadcasts lr through an fp32 tensor
    state->kl_mean = (double)kl;
  }
  const long long t = state->step + 1;
  const double bc1 = 1.0 - pow(beta1, (double)t);
  const double bc2 = 1.0 - pow(beta2, (double)t);
  state->lr = lr;
  state->clip_scale = coef / w;
  state->total_norm = norm;
  state->step_size = lr / bc1;
  state->bc2_sqrt = sqrt(bc2);
  state->step = t;
  __threadfence();
  ws->ticket = 0u;
}

__device__ __forceinline__ void adam_elem(float& p, float& g, float& m, float& v, float scale, float w1, float b2,
                                          float w2, float bc2_sqrt, float eps, float neg_step) {
  const float gc = g * scale;
  m = m + (gc - m) * w1;
  v = v * b2 + w2 * gc * gc;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p + neg_step * (m / denom);
}

__global__ void __launch_bounds__(kOptThreads)
adam_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long P, const rslrl_opt_state_t* __restrict__ state, double beta1, double beta2, double eps_d,
                 int zero_grad) {
  const float scale = (float)state->clip_scale;
  const float neg_step = (float)(-state->step_size);
  const float bc2_sqrt = (float)state->bc2_sqrt;
  const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), eps = (float)eps_d;
  const long long tid = (long long)blockIdx.x * kOptThreads + threadIdx.x;
  const long long stride = (long long)gridDim.x * kOptThreads;
  const long long P4 = P >> 2;
  for (long long i = tid; i < P4; i += stride) {
    float4 pp = ((float4*)p)[i], gg = ((float4*)g)[i], mm = ((float4*)m)[i], vv = ((float4*)v)[i];
    adam_elem(pp.x, gg.x, mm.x, vv.x, scale, w1, b2, w2, bc2_sqrt, eps, neg_step);
    adam_elem(pp.y, gg.y, mm.y, vv.y, scale, w1, b2, w2, bc2_sqrt, eps, neg_step);
    adam_elem(pp.z, gg.z, mm.z, vv.z, scale, w1, b2, w2, bc2_sqrt, eps, neg_step);
    adam_elem(pp.w, gg.w, mm.w, vv.w, scale, w1, b2, w2, bc2_sqrt, eps, neg_step);
    ((float4*)p)[i] = pp;
    ((float4*)m)[i] = mm;
    ((float4*)v)[i] = vv;
    if (zero_grad) ((float4*)g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long i = (P4 << 2) + tid; i < P; i += stride) {
    float pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    adam_elem(pp, gg, mm, vv, scale, w1, b2, w2, bc2_sqrt, eps, neg_step);
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (zero_grad) g[i] = 0.f;
  }
}

static int opt_blocks(long long P) {
  long long b = (P / 4 + kOptThreads - 1) / kOptThreads;
  if (b > kOptMaxBlocks) b = kOptMaxBlocks;
  return b < 1 ? 1 : (int)b;
}

}  // namespace rslrl

extern "C" {

int64_t rslrl_optim_workspace_bytes(int64_t P) {
  (void)P;
  return (int64_t)sizeof(rslrl::NormWorkspace);
}

int rslrl_grad_norm_lr(const float* g, int64_t P, const float* kl_sum, rslrl_opt_state_t* state, void* workspace,
                       int world_size, double max_grad_norm, int adaptive, double desired_kl, int round_lr_fp32,
                       double beta1, double beta2, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_grad_norm_lr";
  RSLRL_CHECK_ARG(g && state && workspace && P > 0 && world_size >= 1, what);
  RSLRL_CHECK_ARG(((uintptr_t)g & 15) == 0, what);
  grad_norm_lr_kernel<<<opt_blocks(P), kOptThreads, 0, (cudaStream_t)stream>>>(
      g, P, kl_sum, state, (NormWorkspace*)workspace, world_size, max_grad_norm, adaptive, desired_kl, round_lr_fp32,
      beta1, beta2);
  return finish_launch(what);
}

int rslrl_adam_step(float* p, float* g, float* m, float* v, int64_t P, const rslrl_opt_state_t* state,
                    double beta1, double beta2, double eps, int zero_grad, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_adam_step";
  RSLRL_CHECK_ARG(p && g && m && v && state && P > 0, what);
  RSLRL_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, what);
  adam_step_kernel<<<opt_blocks(P), kOptThreads, 0, (cudaStream_t)stream>>>(p, g, m, v, P, state, beta1, beta2, eps,
                                                                            zero_grad);
  return finish_launch(what);
}

}  // extern "C"
