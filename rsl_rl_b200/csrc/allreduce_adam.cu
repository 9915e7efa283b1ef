// One kernel per optimiser step for the data-parallel PPO update: gradient all-reduce over NVLink peer memory +
// gradient-norm clip + KL-adaptive learning rate + Adam, replacing
//   rsl_rl/algorithms/ppo.py:441-469 (reduce_parameters: torch.cat + all_reduce + 17 copy_),
//   :272-290 (KL all-reduce + lr broadcast, two host syncs), :380 (clip_grad_norm_), :381 (Adam.step).
//
// Every rank keeps its flat gradient buffer [P | KL tail] in symmetric (peer-mapped) memory.  Block b of every rank
// owns slice b of the parameters:
//   A  cross-rank flag barrier per block ("my backward is complete" <-> "so is every peer's"), epoch stamped
//   1  one-shot reduce: slice b of every rank is read straight from its HBM over NVLink (fixed rank order, so all
//      ranks compute bit-identical sums) into shared memory; sum of squares in fp64
//   B  "done reading" flags to the peers (non-blocking)
//   2  grid barrier (blocks are co-resident: grid <= #SMs), total norm from the per-block partials in fixed order,
//      clip coefficient, KL rule, Adam bias corrections - computed redundantly per block from the OLD state
//   3  Adam on slice b from shared memory (p, m, v touched once)
//   4  wait for the peers' "done reading" flags of block b, then clear slice b of the local gradient buffer
// With world == 1 the flag phases vanish and the kernel is clip + KL-LR + Adam in one launch instead of two.
// All scalars stay on the device (rslrl_opt_state_t); the kernel is CUDA-graph capturable (epoch in device memory).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace rslrl {

constexpr int kArThreads = 1024;  // one float4 per thread at the PPO sizes (P ~ 0.4-0.6 M): one NVLink latency per phase
constexpr int kArMaxBlocks = 148;
constexpr int kArMaxWorld = 8;

struct ArWorkspace {
  double partial[kArMaxBlocks];
  unsigned long long arrivals;  // grid barrier: monotonically increasing arrival count
  unsigned int epoch;           // launches so far (identical on every rank)
  unsigned int pad_;
};

struct ArParams {
  const float* peer_grads[kArMaxWorld];  // gradient buffer [P + 4] of every rank (own rank included)
  unsigned int* peer_flags[kArMaxWorld]; // flag buffer [3][kArMaxBlocks][kArMaxWorld] of every rank
  float* peer_results[kArMaxWorld];      // two-step mode: reduced-gradient buffer [P] of every rank (else null)
  float* peer_stage[kArMaxWorld];        // push mode: staging buffer [world][P + 4] of every rank (else null)
  float* p;
  __half* p16;  // optional fp16 shadow of p (same element offsets), refreshed by the Adam phase
  float* g;  // own gradient buffer (= peer_grads[rank])
  float* m;
  float* v;
  long long P;
  rslrl_opt_state_t* state;
  ArWorkspace* ws;
  int world, rank, adaptive, round_lr_fp32, use_kl, zero_grad;
  double max_norm, desired_kl, beta1, beta2, eps;
  long long timeout_cycles;  // spin limit of the peer / grid waits (RSLRL_PEER_TIMEOUT_S, default 30 min)
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_volatile_f(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// Spin on a peer's flag.  Ordinary rank skew (rank 0 saving a checkpoint, logging) must simply be waited out, as NCCL
// does; the limit only turns a DEAD peer into a reported trap instead of a hung GPU.  It is configurable
// (RSLRL_PEER_TIMEOUT_S, default 1800 s) and PPO.update() additionally lines the ranks up with one NCCL collective
// before the first peer step, so the spin never sees more than kernel-level skew.
__device__ __forceinline__ void wait_flag(const unsigned int* flag, unsigned int stamp, long long limit) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - stamp) < 0) {
    if (clock64() - t0 > limit) {
      printf("rslrl allreduce_clip_adam: peer flag timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// Optional in-kernel timeline (build with -DRSLRL_AR_TRACE, read with tools/mgpu_phase_trace.py): per launch and block,
// globaltimer stamps at 0 entry, 1 peers' backward complete (phase A), 2 slice reduced (phase 1), 3 grid barrier passed,
// 4 Adam done, 5 peers done reading / exit.  Ring of kArTraceLaunches launches.
#ifdef RSLRL_AR_TRACE
constexpr int kArTraceLaunches = 64;
__device__ unsigned long long g_ar_trace[kArTraceLaunches * kArMaxBlocks * 8];
__device__ __forceinline__ void ar_trace(unsigned int epoch, int slot) {
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_ar_trace[((epoch % kArTraceLaunches) * kArMaxBlocks + blockIdx.x) * 8 + slot] = t;
  }
}
#define AR_TRACE(slot) ar_trace(epoch, slot)
#else
#define AR_TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ void adam_elem1(float& p, float g, float& m, float& v, float scale, float w1, float b2,
                                           float w2, float bc2_sqrt, float eps, float neg_step) {
  const float gc = g * scale;
  m = m + (gc - m) * w1;
  v = v * b2 + w2 * gc * gc;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p + neg_step * (m / denom);
}

__global__ void __launch_bounds__(kArThreads, 1) allreduce_clip_adam_kernel(const __grid_constant__ ArParams prm) {
  extern __shared__ __align__(16) float s_slice[];
  __shared__ double s_red[32];
  __shared__ double s_scal[4];  // clip_scale, step_size, bc2_sqrt

  pdl_trigger();
  pdl_wait();
  const int W = prm.world, rank = prm.rank, b = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
  ArWorkspace* const ws = prm.ws;
  const unsigned int epoch = *((volatile unsigned int*)&ws->epoch);
  const unsigned int stamp = epoch + 1u;
  const long long P = prm.P, P4 = P >> 2;  // P is a multiple of 4 (flat buffers are 16-byte padded)
  const long long per = (P4 + G - 1) / G;
  const long long i0 = (long long)b * per, i1 = (i0 + per < P4) ? i0 + per : P4;

  // old optimiser state (read before the grid barrier; block 0 rewrites it after)
  const double lr_old = prm.state->lr;
  const long long step_old = prm.state->step;

  AR_TRACE(0);
  const bool push = W > 1 && prm.peer_stage[0] != nullptr;
  const bool two_step = W > 1 && prm.peer_results[0] != nullptr;
  const long long RS = P + 4;   // row stride of a staging buffer (floats): [parity][src rank][P gradients | KL tail]
  // (double buffered by launch parity: a rank can start pushing launch n+1 while a peer still reads launch n's copy;
  // it cannot reach launch n+2 before that peer has entered n+1, i.e. finished reading n)
  const size_t stage_par = (size_t)(epoch & 1u) * (size_t)W * (size_t)RS;
  // ---- PUSH mode, step 1: this rank's share of slice b goes to its owner(s) -- no "is your backward complete" round
  // trip: a rank pushes its own finished gradients and the receiver waits for "data landed" flags (one NVLink hop per
  // step; the pull protocol below needs two for the same step).  One-hop variant (small worlds): every rank owns
  // every element.  Two-step variant: rank r owns sub-slice r of slice b.
  if (push) {
    const long long len = i1 - i0, per_r = (len + W - 1) / W;
    for (long long i = i0 + tid; i < i1; i += kArThreads) {
      const float4 x = *(const float4*)(prm.g + 4 * i);
      if (two_step) {
        const int o = (int)((i - i0) / per_r);
        if (o != rank) *(float4*)(prm.peer_stage[o] + stage_par + (size_t)rank * RS + 4 * i) = x;
      } else {
        for (int k = 0; k < W; ++k)
          if (k != rank) *(float4*)(prm.peer_stage[k] + stage_par + (size_t)rank * RS + 4 * i) = x;
      }
    }
    if (prm.use_kl && tid < W && tid != rank) prm.peer_stage[tid][stage_par + (size_t)rank * RS + P] = prm.g[P];
    __threadfence_system();
    __syncthreads();
    if (tid < W && tid != rank) {
      st_release_sys(prm.peer_flags[tid] + (0 * kArMaxBlocks + b) * kArMaxWorld + rank, stamp);
      wait_flag(prm.peer_flags[rank] + (0 * kArMaxBlocks + b) * kArMaxWorld + tid, stamp, prm.timeout_cycles);
    }
    __syncthreads();
  }
  // ---- A (pull mode): every rank's backward is complete
  if (W > 1 && !push) {
    if (tid < W && tid != rank) {
      __threadfence_system();
      st_release_sys(prm.peer_flags[tid] + (0 * kArMaxBlocks + b) * kArMaxWorld + rank, stamp);
      wait_flag(prm.peer_flags[rank] + (0 * kArMaxBlocks + b) * kArMaxWorld + tid, stamp, prm.timeout_cycles);
    }
    __syncthreads();
  }

  AR_TRACE(1);
  // ---- 1: reduce slice b over the ranks (fixed order), keep it in shared memory, sum of squares
  double acc[1] = {0.0};
  const float* const stage = push ? prm.peer_stage[rank] + stage_par : nullptr;   // local staging buffer (written by the peers)
  // contribution of rank k to element i: local gradients for k == rank, else the pushed copy / the peer's buffer
  auto contrib = [&](int k, long long i) -> float4 {
    if (k == rank) return *(const float4*)(prm.g + 4 * i);
    return push ? ld_volatile_f4(stage + (size_t)k * RS + 4 * i) : ld_volatile_f4(prm.peer_grads[k] + 4 * i);
  };
  if (!two_step) {
    // one step: slice b of EVERY rank is summed here in fixed rank order ((W-1) x slice bytes inbound per rank)
    for (long long i = i0 + tid; i < i1; i += kArThreads) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < W; ++k) {
        const float4 x = contrib(k, i);
        s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
      }
      *(float4*)(s_slice + 4 * (i - i0)) = s;
      acc[0] += (double)s.x * s.x + (double)s.y * s.y + (double)s.z * s.z + (double)s.w * s.w;
    }
  } else {
    // two steps (reduce-scatter + all-gather inside the slice): rank r reduces sub-slice r of slice b over all ranks
    // (fixed order) and WRITES the sums into every rank's result buffer; after a cross-rank flag barrier every rank
    // reads the whole reduced slice from its LOCAL result buffer.  Inbound + outbound traffic per rank is
    // 2 (W-1)/W x slice instead of (W-1) x slice, and all ranks hold bit-identical sums by construction.
    const long long len = i1 - i0, per_r = (len + W - 1) / W;
    const long long s0 = i0 + (long long)rank * per_r < i1 ? i0 + (long long)rank * per_r : i1;
    const long long s1 = s0 + per_r < i1 ? s0 + per_r : i1;
    for (long long i = s0 + tid; i < s1; i += kArThreads) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < W; ++k) {
        const float4 x = contrib(k, i);
        s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
      }
      for (int k = 0; k < W; ++k) *(float4*)(prm.peer_results[k] + 4 * i) = s;
    }
    __threadfence_system();
    __syncthreads();
    // B (early, pull mode): this rank is done reading slice b of its peers' gradients
    if (!push && tid < W && tid != rank)
      st_release_sys(prm.peer_flags[tid] + (1 * kArMaxBlocks + b) * kArMaxWorld + rank, stamp);
    // C: the reduced sub-slices of every rank have landed in this rank's result buffer
    if (tid < W && tid != rank) {
      st_release_sys(prm.peer_flags[tid] + (2 * kArMaxBlocks + b) * kArMaxWorld + rank, stamp);
      wait_flag(prm.peer_flags[rank] + (2 * kArMaxBlocks + b) * kArMaxWorld + tid, stamp, prm.timeout_cycles);
    }
    __syncthreads();
    const float* res = prm.peer_results[rank];
    for (long long i = i0 + tid; i < i1; i += kArThreads) {
      const float4 s = ld_volatile_f4(res + 4 * i);
      *(float4*)(s_slice + 4 * (i - i0)) = s;
      acc[0] += (double)s.x * s.x + (double)s.y * s.y + (double)s.z * s.z + (double)s.w * s.w;
    }
  }
  // Adam operands of this thread's first (at the PPO sizes: only) element: fetched BEFORE the grid barrier, so that the
  // barrier wait hides their DRAM latency and the phase after it is arithmetic + stores
  const long long ipre = i0 + tid;
  const bool has_pre = ipre < i1;
  float4 pre_p = make_float4(0.f, 0.f, 0.f, 0.f), pre_m = pre_p, pre_v = pre_p;
  if (has_pre) {
    pre_p = ((const float4*)prm.p)[ipre];
    pre_m = ((const float4*)prm.m)[ipre];
    pre_v = ((const float4*)prm.v)[ipre];
  }
  float kl_sum = 0.f;
  if (prm.use_kl && tid == 0) {  // the minibatch KL rides in the tail of the gradient buffer (4 bytes per rank)
    for (int k = 0; k < W; ++k)
      kl_sum += (k == rank) ? prm.g[P] : (push ? ld_volatile_f(stage + (size_t)k * RS + P) : ld_volatile_f(prm.peer_grads[k] + P));
  }
  block_sum<1>(acc, s_red);

  // ---- B: tell the peers that their slice b has been read (they may clear it)
  if (W > 1 && !two_step && !push) {
    __syncthreads();
    if (tid < W && tid != rank)
      st_release_sys(prm.peer_flags[tid] + (1 * kArMaxBlocks + b) * kArMaxWorld + rank, stamp);
  }

  AR_TRACE(2);
  // ---- 2: grid barrier, then norm / clip / lr / bias corrections (redundantly per block, deterministic order)
  // Adam bias corrections of step t (fp64 pow): computed while the other blocks arrive
  double bc1 = 1.0, bc2 = 1.0;
  if (tid == 0) {
    ws->partial[b] = acc[0];
    __threadfence();
    atomicAdd(&ws->arrivals, 1ull);
    bc1 = 1.0 - pow(prm.beta1, (double)(step_old + 1));
    bc2 = 1.0 - pow(prm.beta2, (double)(step_old + 1));
    const unsigned long long target = (unsigned long long)stamp * (unsigned long long)G;
    const long long t0 = clock64();
    while (ld_acquire_gpu(&ws->arrivals) < target) {
      if (clock64() - t0 > prm.timeout_cycles) {
        printf("rslrl allreduce_clip_adam: grid barrier timeout (block %d)\n", b);
        __trap();
      }
    }
  }
  __syncthreads();
  if (tid < 32) {
    double s = 0.0;
    for (int j = tid; j < G; j += 32) s += *((volatile double*)&ws->partial[j]);
    s = warp_sum(s);
    if (tid == 0) {
      const double w = (double)W;
      const double norm = sqrt(s) / w;  // || mean over ranks of g ||
      double coef = 1.0;
      if (prm.max_norm > 0.0) {
        coef = prm.max_norm / (norm + 1e-6);
        coef = coef < 1.0 ? coef : 1.0;
      }
      double lr = lr_old;
      float kl = 0.f;
      if (prm.adaptive) {  // ppo.py:272-284: fp32 comparisons
        kl = prm.use_kl ? kl_sum / (float)W : 0.f;
        if (kl > (float)(prm.desired_kl * 2.0)) lr = fmax(1e-5, lr / 1.5);
        else if (kl < (float)(prm.desired_kl / 2.0) && kl > 0.0f) lr = fmin(1e-2, lr * 1.5);
        if (prm.round_lr_fp32) lr = (double)(float)lr;
      }
      const long long t = step_old + 1;
      s_scal[0] = coef / w;
      s_scal[1] = lr / bc1;
      s_scal[2] = sqrt(bc2);
      if (b == 0) {
        rslrl_opt_state_t* st = prm.state;
        st->lr = lr;
        st->clip_scale = coef / w;
        st->total_norm = norm;
        st->step_size = lr / bc1;
        st->bc2_sqrt = sqrt(bc2);
        if (prm.adaptive) st->kl_mean = (double)kl;
        st->step = t;
        ws->epoch = stamp;
      }
    }
  }
  __syncthreads();

  AR_TRACE(3);
  // ---- 3: Adam on slice b
  {
    const float scale = (float)s_scal[0], neg_step = (float)(-s_scal[1]), bc2_sqrt = (float)s_scal[2];
    const float w1 = (float)(1.0 - prm.beta1), b2 = (float)prm.beta2, w2 = (float)(1.0 - prm.beta2);
    const float eps = (float)prm.eps;
    for (long long i = i0 + tid; i < i1; i += kArThreads) {
      float4 pp, mm, vv;
      if (i == ipre) { pp = pre_p; mm = pre_m; vv = pre_v; }
      else { pp = ((float4*)prm.p)[i]; mm = ((float4*)prm.m)[i]; vv = ((float4*)prm.v)[i]; }
      const float4 gg = *(const float4*)(s_slice + 4 * (i - i0));
      adam_elem1(pp.x, gg.x, mm.x, vv.x, scale, w1, b2, w2, bc2_sqrt, eps, neg_step);
      adam_elem1(pp.y, gg.y, mm.y, vv.y, scale, w1, b2, w2, bc2_sqrt, eps, neg_step);
      adam_elem1(pp.z, gg.z, mm.z, vv.z, scale, w1, b2, w2, bc2_sqrt, eps, neg_step);
      adam_elem1(pp.w, gg.w, mm.w, vv.w, scale, w1, b2, w2, bc2_sqrt, eps, neg_step);
      ((float4*)prm.p)[i] = pp;
      ((float4*)prm.m)[i] = mm;
      ((float4*)prm.v)[i] = vv;
      if (prm.p16 != nullptr) {
        const __half2 lo = __floats2half2_rn(pp.x, pp.y), hi = __floats2half2_rn(pp.z, pp.w);
        uint2 o;
        o.x = *(const unsigned int*)&lo;
        o.y = *(const unsigned int*)&hi;
        ((uint2*)prm.p16)[i] = o;
      }
    }
  }

  AR_TRACE(4);
  // ---- 4: the peers have read slice b of this rank's gradients -> clear it for the next backward pass
  if (W > 1 && !push) {   // (push mode: nobody reads this rank's gradient buffer remotely)
    if (tid < W && tid != rank) wait_flag(prm.peer_flags[rank] + (1 * kArMaxBlocks + b) * kArMaxWorld + tid, stamp, prm.timeout_cycles);
    __syncthreads();
  }
  if (prm.zero_grad) {
    for (long long i = i0 + tid; i < i1; i += kArThreads) ((float4*)prm.g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  AR_TRACE(5);
}

}  // namespace rslrl

extern "C" {

int64_t rslrl_allreduce_adam_workspace_bytes(void) { return (int64_t)sizeof(rslrl::ArWorkspace); }
int64_t rslrl_allreduce_adam_flag_bytes(void) {
  return (int64_t)(3 * rslrl::kArMaxBlocks * rslrl::kArMaxWorld * sizeof(unsigned int));
}

int rslrl_allreduce_clip_adam(int world, int rank, const float* const* peer_grads_host,
                              unsigned int* const* peer_flags_host, float* const* peer_results_host,
                              float* const* peer_stage_host, float* p, float* m, float* v, int64_t P,
                              rslrl_opt_state_t* state, void* workspace, double max_grad_norm, int adaptive,
                              double desired_kl, int round_lr_fp32, int use_kl, double beta1, double beta2, double eps,
                              int zero_grad, void* p16_shadow, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_allreduce_clip_adam";
  RSLRL_CHECK_ARG(((uintptr_t)p16_shadow & 7) == 0, what);
  RSLRL_CHECK_ARG(world >= 1 && world <= kArMaxWorld && rank >= 0 && rank < world, what);
  RSLRL_CHECK_ARG(peer_grads_host && p && m && v && state && workspace && P > 0 && (P % 4) == 0, what);
  RSLRL_CHECK_ARG(world == 1 || peer_flags_host, what);
  ArParams prm{};
  for (int k = 0; k < world; ++k) {
    RSLRL_CHECK_ARG(peer_grads_host[k] && ((uintptr_t)peer_grads_host[k] & 15) == 0, what);
    prm.peer_grads[k] = peer_grads_host[k];
    if (world > 1) {
      RSLRL_CHECK_ARG(peer_flags_host[k] != nullptr, what);
      prm.peer_flags[k] = peer_flags_host[k];
      if (peer_results_host != nullptr) {
        RSLRL_CHECK_ARG(peer_results_host[k] != nullptr && ((uintptr_t)peer_results_host[k] & 15) == 0, what);
        prm.peer_results[k] = peer_results_host[k];
      }
      if (peer_stage_host != nullptr) {
        RSLRL_CHECK_ARG(peer_stage_host[k] != nullptr && ((uintptr_t)peer_stage_host[k] & 15) == 0, what);
        prm.peer_stage[k] = peer_stage_host[k];
      }
    }
  }
  RSLRL_CHECK_ARG((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) == 0, what);
  prm.p = p; prm.p16 = (__half*)p16_shadow; prm.g = const_cast<float*>(peer_grads_host[rank]); prm.m = m; prm.v = v; prm.P = P;
  prm.state = state; prm.ws = (ArWorkspace*)workspace;
  prm.world = world; prm.rank = rank; prm.adaptive = adaptive; prm.round_lr_fp32 = round_lr_fp32;
  prm.use_kl = use_kl; prm.zero_grad = zero_grad;
  {
    static long long cycles = 0;
    if (cycles == 0) {
      const char* env = getenv("RSLRL_PEER_TIMEOUT_S");
      double sec = env ? atof(env) : 1800.0;
      if (!(sec > 0.0)) sec = 1800.0;
      cycles = (long long)(sec * 2.0e9);  // clock64 ticks at <= ~2 GHz
    }
    prm.timeout_cycles = cycles;
  }
  prm.max_norm = max_grad_norm; prm.desired_kl = desired_kl; prm.beta1 = beta1; prm.beta2 = beta2; prm.eps = eps;

  int sms = kNumSMs;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // every rank must use the same grid (block b pairs with block b of the peers): a function of P alone
  const long long P4 = P / 4;
  long long grid = (P4 + 255) / 256;  // >= 256 float4 (4 KB) per block
  if (grid > kArMaxBlocks) grid = kArMaxBlocks;
  if (grid < 1) grid = 1;
  if (grid > sms) {
    set_error("%s: needs %lld co-resident blocks, the device has %d SMs", what, grid, sms);
    return 3;
  }
  const long long per = (P4 + grid - 1) / grid;
  const size_t smem = (size_t)per * 16;
  if (smem > 200 * 1024) {
    set_error("%s: %lld parameters exceed the shared-memory slice capacity", what, (long long)P);
    return 3;
  }
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(allreduce_clip_adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return 1;
    }
    smem_set = smem;
  }
  launch_chain_kernel(allreduce_clip_adam_kernel, dim3((int)grid), dim3(kArThreads), smem, (cudaStream_t)stream, prm);
  return finish_launch(what);
}

#ifdef RSLRL_AR_TRACE
int rslrl_debug_read_trace_allreduce(unsigned long long* host) {
  return (int)cudaMemcpyFromSymbol(host, rslrl::g_ar_trace,
                                   sizeof(unsigned long long) * rslrl::kArTraceLaunches * rslrl::kArMaxBlocks * 8);
}
#endif

}  // extern "C"
