// Shared helpers for the sm_100a kernels behind include/rslrl_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rslrl_b200.h"

namespace rslrl {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);
void count_launch(unsigned long long n = 1);

// Records the launch and converts launch errors into the library's error channel.
inline int finish_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  count_launch();
  return 0;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// The update() chain is ~7 back-to-back launches per minibatch, each 10-60 us long, so launch latency and prologues
// (barrier init, TMEM allocation, descriptor prefetch) are a visible share.  Kernels of the chain call pdl_trigger()
// first thing (the NEXT kernel's CTAs may then be scheduled as soon as SM resources free up) and pdl_wait() before
// their first access to global memory (returns once the PREVIOUS kernel has completed and flushed).  Every kernel
// launched with the attribute executes pdl_wait(), so completion stays transitive along the stream.
// Opt-in (RSLRL_PDL=1); by default everything is launched with plain stream order.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();

// Shared-memory carve-out: the tcgen05 kernels of the chain run with the maximum shared-memory configuration (227 KB);
// a kernel without a preference gets a small carve-out, and switching the L1 / shared split between two consecutive
// kernels needs idle SMs.  With the preference set on the small kernels of the chain (loss, optimiser, LSTM cells, ...)
// every launch keeps the same split.  RSLRL_SMEM_CARVEOUT=0 / 1 (A/B switch).
// Measured (profiles/experiments/notes/r2_carveout.md): the recurrent update alternates a 227 KB GEMM and a cell kernel 48
// times per time loop -- with the preference on the cell kernels C4 drops from 56 to 46 ms per iteration (always on for
// them); on the feed-forward chain (7 launches per minibatch) it is neutral to slightly negative (5.08 -> 5.20 ms), so the
// other kernels only follow RSLRL_SMEM_CARVEOUT=1.
bool carveout_enabled();
template <typename K>
inline void prefer_max_smem(K kernel, bool always = false) {
  static bool done = false;   // one flag per kernel type (template instantiation)
  if (!done) {
    if (always || carveout_enabled())
      (void)cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    (void)cudaGetLastError();
    done = true;
  }
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                       Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  prefer_max_smem(kernel);
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define RSLRL_CHECK_ARG(cond, what)                      \
  do {                                                   \
    if (!(cond)) {                                       \
      rslrl::set_error("%s: bad argument: %s", what, #cond); \
      return 2;                                          \
    }                                                    \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NV doubles per thread; result valid in thread 0.  smem: NV * 32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = lane < nwarps ? smem[i * 32 + lane] : 0.0;
      v[i] = warp_sum(x);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float act_forward(int act, float z) {
  switch (act) {
    case RSLRL_ACT_ELU: return z > 0.f ? z : expm1f(z);
    case RSLRL_ACT_RELU: return fmaxf(z, 0.f);
    case RSLRL_ACT_TANH: return tanhf(z);
    case RSLRL_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    case RSLRL_ACT_LRELU: return z > 0.f ? z : 0.01f * z;
    case RSLRL_ACT_SELU: return 1.0507009873554804934f * (z > 0.f ? z : 1.6732632423543772848f * expm1f(z));
    case RSLRL_ACT_SOFTPLUS: return z > 20.f ? z : log1pf(expf(z));
    default: return z;
  }
}

// d act(z)/dz expressed through the activation OUTPUT h = act(z) (no pre-activation is stored).
__device__ __forceinline__ float act_grad_from_output(int act, float h) {
  switch (act) {
    case RSLRL_ACT_ELU: return h > 0.f ? 1.f : h + 1.f;  // exp(z) = h + 1 for z <= 0
    case RSLRL_ACT_RELU: return h > 0.f ? 1.f : 0.f;
    case RSLRL_ACT_TANH: return 1.f - h * h;
    case RSLRL_ACT_SIGMOID: return h * (1.f - h);
    case RSLRL_ACT_LRELU: return h > 0.f ? 1.f : 0.01f;
    case RSLRL_ACT_SELU:
      return h > 0.f ? 1.0507009873554804934f : h + 1.0507009873554804934f * 1.6732632423543772848f;
    case RSLRL_ACT_SOFTPLUS: return 1.f - expf(-h);  // sigmoid(z) = 1 - exp(-softplus(z))
    default: return 1.f;
  }
}

}  // namespace rslrl
