// GAE reverse scan + returns + advantage normalisation in one kernel (grid-wide barrier inside).
// Replaces RolloutStorage.compute_returns (reference rsl_rl/storage/rollout_storage.py:127-149),
// which issues ~400 elementwise launches over 24 [N,1] slices.
//
// Layout: rewards/values/returns/advantages are [T,N] fp32, dones [T,N] u8 (the reference's [T,N,1]
// buffers, same bytes).  One thread owns one env: for fixed t a warp touches 32 consecutive floats
// (one 128 B line), so every access is fully coalesced.  All 3*T loads of an env are issued before
// the scan starts (T is a template bound), which is what hides the HBM latency: the scan itself is
// only T dependent FMUL/FADD triples.  The fp32 operation order of the reference is kept with
// explicit round-to-nearest intrinsics (no FMA contraction), so `returns` and the un-normalised
// advantages are bit-identical to the reference.  The normalisation needs the global mean and
// unbiased std of all T*N advantages: per-thread fp64 partials -> warp shuffles -> one fp64 pair per
// block in a workspace -> grid barrier -> every block reduces the per-block pairs in the same order
// (deterministic) -> returns and advantages, still in registers, are written once.
// Algorithmic traffic: 17*T*N + 4*N bytes.
//
// Round-2 shape of the kernel (the round-1 one ran at 0.29 of the HBM peak at 65 536 envs):
//   * NOTHING is stored before the barrier: the scan keeps `returns` in the registers that held the rewards and both
//     outputs leave after the barrier, so the arrival fence has no store queue to drain and the barrier sits directly
//     behind the last load;
//   * the grid is at most one block per SM (block size grows with N), so the barrier is <= 148 arrivals polled with
//     ld.acquire, and the per-block partial pairs are reduced by one warp;
//   * the launch is COOPERATIVE (cudaLaunchAttributeCooperative): co-residency of the spinning blocks is guaranteed by
//     the driver (or the launch fails) whatever else shares the device; the wait is unbounded like any grid.sync().
// The last block to leave re-arms the counters, so the workspace only has to be zero ONCE (at allocation).
//
// TMA-staged kernel (gae_norm_tma_kernel, the default whenever N % 16 == 0 and T <= 32): a block owns a contiguous span
// of envs; for a fixed step that span is one contiguous run of bytes in every [T,N] array, so the block's whole input
// arrives as 3 T + 1 bulk copies (cp.async.bulk -> shared memory, one mbarrier) issued by one warp instead of 3 T scalar
// loads per thread, and the outputs leave as 2 T bulk stores from shared memory (the returns rows while the statistics
// are still being reduced).  The scan itself is unchanged (one env per thread, registers, the reference's fp32 op order).
// Up to 8 192 envs the blocks form ONE thread-block cluster (<= 16 CTAs): the per-block fp64 partials are exchanged
// through distributed shared memory behind the hardware cluster barrier - no cooperative launch, no global-memory
// barrier; larger rollouts keep the cooperative grid barrier below.
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rslrl {

constexpr int kGaeThreads = 128;      // grid-stride fallback kernel
constexpr int kGaeMaxBlocks = 592;    // workspace capacity (pairs of doubles): 4 blocks per SM on 148 SMs

// Optional in-kernel timeline (build with -DRSLRL_GAE_TRACE; profiles/experiments/gae_trace.py): per block globaltimer at
// 0 entry, 1 loads landed + scan done, 2 partials published (barrier arrival), 3 barrier passed, 4 statistics known, 5 exit.
#ifdef RSLRL_GAE_TRACE
__device__ unsigned long long g_gae_trace[148 * 8];
__device__ __forceinline__ void gae_trace(int slot) {
  if (threadIdx.x == 0 && blockIdx.x < 148) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_gae_trace[blockIdx.x * 8 + slot] = t;
  }
}
#define GAE_TRACE(slot) gae_trace(slot)
#else
#define GAE_TRACE(slot) do { } while (0)
#endif

struct GaeBarrier {
  unsigned int count;  // arrivals of the current generation (atomicInc wraps it back to 0 at gridDim.x: self re-arming)
  unsigned int gen;    // generation: flipped by the last arriver, polled by everybody else
};

__device__ __forceinline__ unsigned int gae_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void gae_st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Sense-reversing grid barrier (co-residency guaranteed by the cooperative launch): no departure phase, nothing to
// reset between launches, grid size may differ from launch to launch.
__device__ __forceinline__ void gae_grid_barrier(GaeBarrier* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int gen = gae_ld_acquire(&bar->gen);   // cannot change before this block has arrived
    __threadfence();
    if (atomicInc(&bar->count, gridDim.x - 1) == gridDim.x - 1) {
      gae_st_release(&bar->gen, gen + 1u);
    } else {
      while (gae_ld_acquire(&bar->gen) == gen) {
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void gae_finalize_stats(double s, double ss, double* partials, double* smem,
                                                   GaeBarrier* bar, int64_t B, float& mean_f, float& denom_f) {
  double v[2] = {s, ss};
  block_sum<2>(v, smem);
  if (threadIdx.x == 0) *(double2*)(partials + 2 * blockIdx.x) = make_double2(v[0], v[1]);
  GAE_TRACE(2);
  gae_grid_barrier(bar);
  GAE_TRACE(3);
  if (threadIdx.x < 32) {   // one warp, fixed order (every block derives the same bits); all loads in flight at once
    constexpr int kPer = (kGaeMaxBlocks + 31) / 32;
    double2 p[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = threadIdx.x + 32 * k;
      p[k] = i < (int)gridDim.x ? __ldcg((const double2*)(partials + 2 * i)) : make_double2(0.0, 0.0);
    }
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      t0 += p[k].x;
      t1 += p[k].y;
    }
    t0 = warp_sum(t0);
    t1 = warp_sum(t1);
    if (threadIdx.x == 0) {
      const double n = (double)B;
      const double mean = t0 / n;
      double var = (t1 - t0 * t0 / n) / (n - 1.0);  // unbiased, as torch.Tensor.std()
      var = var > 0.0 ? var : 0.0;
      smem[0] = mean;
      smem[1] = sqrt(var);
    }
  }
  __syncthreads();
  GAE_TRACE(4);
  mean_f = (float)smem[0];
  denom_f = __fadd_rn((float)smem[1], 1e-8f);
}

// One env per thread, whole trajectory in registers.
template <int TMAX>
__global__ void __launch_bounds__(512, 1)
gae_norm_reg_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                    const uint8_t* __restrict__ dones, const float* __restrict__ last_values,
                    float* __restrict__ returns, float* __restrict__ advantages, double* __restrict__ partials,
                    GaeBarrier* bar, int T, int N, float gamma, float lam, int normalize) {
  __shared__ double smem[64];
  GAE_TRACE(0);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float r[TMAX], v[TMAX];   // r[i] is overwritten by returns[i] during the scan
  double s = 0.0, ss = 0.0;
  if (n < N) {
    uint8_t d[TMAX];
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
      if (i < T) {
        const size_t o = (size_t)i * N + n;
        r[i] = __ldg(rewards + o);
        v[i] = __ldg(values + o);
        d[i] = __ldg(dones + o);
      }
    }
    float next_v = __ldg(last_values + n);
    float adv = 0.f;
#pragma unroll
    for (int i = TMAX - 1; i >= 0; --i) {
      if (i < T) {
        const float nnt = __fsub_rn(1.0f, (float)d[i]);
        const float ng = __fmul_rn(nnt, gamma);
        const float delta = __fsub_rn(__fadd_rn(r[i], __fmul_rn(ng, next_v)), v[i]);
        adv = __fadd_rn(delta, __fmul_rn(__fmul_rn(ng, lam), adv));
        const float ret = __fadd_rn(adv, v[i]);
        r[i] = ret;
        const float x = __fsub_rn(ret, v[i]);
        s += (double)x;
        ss += (double)x * (double)x;
        next_v = v[i];
      }
    }
  }
  float mean_f = 0.f, denom_f = 1.f;
  GAE_TRACE(1);
  if (normalize) gae_finalize_stats(s, ss, partials, smem, bar, (int64_t)T * N, mean_f, denom_f);
  if (n < N) {
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
      if (i < T) {
        const float x = __fsub_rn(r[i], v[i]);
        returns[(size_t)i * N + n] = r[i];
        advantages[(size_t)i * N + n] = normalize ? __fdiv_rn(__fsub_rn(x, mean_f), denom_f) : x;
      }
    }
  }
  GAE_TRACE(5);
}

// Any T, any N: grid-stride over envs, raw advantages staged through `advantages` itself.
__global__ void __launch_bounds__(kGaeThreads)
gae_norm_generic_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                        const uint8_t* __restrict__ dones, const float* __restrict__ last_values,
                        float* __restrict__ returns, float* advantages, double* __restrict__ partials, GaeBarrier* bar,
                        int T, int N, float gamma, float lam, int normalize) {
  __shared__ double smem[64];
  double s = 0.0, ss = 0.0;
  const int stride = gridDim.x * kGaeThreads;
  for (int n = blockIdx.x * kGaeThreads + threadIdx.x; n < N; n += stride) {
    float next_v = __ldg(last_values + n);
    float adv = 0.f;
    for (int i = T - 1; i >= 0; --i) {
      const size_t o = (size_t)i * N + n;
      const float r = __ldg(rewards + o), v = __ldg(values + o);
      const float nnt = __fsub_rn(1.0f, (float)__ldg(dones + o));
      const float ng = __fmul_rn(nnt, gamma);
      const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(ng, next_v)), v);
      adv = __fadd_rn(delta, __fmul_rn(__fmul_rn(ng, lam), adv));
      const float ret = __fadd_rn(adv, v);
      returns[o] = ret;
      const float x = __fsub_rn(ret, v);
      advantages[o] = x;
      s += (double)x;
      ss += (double)x * (double)x;
      next_v = v;
    }
  }
  if (!normalize) return;
  float mean_f, denom_f;
  gae_finalize_stats(s, ss, partials, smem, bar, (int64_t)T * N, mean_f, denom_f);
  for (int n = blockIdx.x * kGaeThreads + threadIdx.x; n < N; n += stride)
    for (int i = 0; i < T; ++i) {
      const size_t o = (size_t)i * N + n;
      advantages[o] = __fdiv_rn(__fsub_rn(advantages[o], mean_f), denom_f);  // own writes: same thread
    }
}


// ---------------------------------------------------------------------------------------------- TMA-staged kernel
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"((uint64_t)dst), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ double ld_cluster_f64(const double* local_ptr, uint32_t cta_rank) {
  uint32_t remote;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(cta_rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
  return v;
}

// nb = envs per block (multiple of 16, <= blockDim.x); CLUSTER: the whole grid is one cluster of gridDim.x CTAs.
template <int TMAX, bool CLUSTER>
__global__ void __launch_bounds__(512, 1)
gae_norm_tma_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                    const uint8_t* __restrict__ dones, const float* __restrict__ last_values,
                    float* __restrict__ returns, float* __restrict__ advantages, double* __restrict__ partials,
                    GaeBarrier* bar, int T, int N, int nb, float gamma, float lam, int normalize) {
  extern __shared__ __align__(128) uint8_t gae_smem[];
  __shared__ double s_red[64];
  __shared__ __align__(16) double s_part[2];
  __shared__ __align__(8) uint64_t s_mbar;
  float* const s_r = (float*)gae_smem;                      // [T][nb] rewards, overwritten by the returns
  float* const s_v = s_r + (size_t)T * nb;                  // [T][nb] values, overwritten by the advantages
  uint8_t* const s_d = (uint8_t*)(s_v + (size_t)T * nb);    // [T][nb]
  float* const s_last = (float*)(s_d + (size_t)T * nb);     // [nb]
  GAE_TRACE(0);
  const int tid = threadIdx.x, lane = tid & 31;
  const int n0 = blockIdx.x * nb;
  const int cnt = min(nb, N - n0);                          // envs of this block (multiple of 16)
  if (tid == 0) {
    mbar_init(&s_mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32) {   // one warp issues the bulk copies: lane <-> steps lane, lane + 32, ...
    if (lane == 0) mbar_expect_tx(&s_mbar, (uint32_t)cnt * (9u * (uint32_t)T + 4u));
    __syncwarp();
    for (int t = lane; t < T; t += 32) {
      const size_t o = (size_t)t * N + n0;
      bulk_load(s_r + (size_t)t * nb, rewards + o, (uint32_t)cnt * 4u, &s_mbar);
      bulk_load(s_v + (size_t)t * nb, values + o, (uint32_t)cnt * 4u, &s_mbar);
      bulk_load(s_d + (size_t)t * nb, dones + o, (uint32_t)cnt, &s_mbar);
    }
    if (lane == 0) bulk_load(s_last, last_values + n0, (uint32_t)cnt * 4u, &s_mbar);
  }
  mbar_wait(&s_mbar, 0);

  float x[TMAX];   // raw advantages of this thread's env
  double s = 0.0, ss = 0.0;
  const bool live = tid < cnt;
  if (live) {
    float r[TMAX], v[TMAX];
    uint8_t d[TMAX];
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
      if (i < T) {
        r[i] = s_r[(size_t)i * nb + tid];
        v[i] = s_v[(size_t)i * nb + tid];
        d[i] = s_d[(size_t)i * nb + tid];
      }
    }
    float next_v = s_last[tid];
    float adv = 0.f;
#pragma unroll
    for (int i = TMAX - 1; i >= 0; --i) {
      if (i < T) {
        const float nnt = __fsub_rn(1.0f, (float)d[i]);
        const float ng = __fmul_rn(nnt, gamma);
        const float delta = __fsub_rn(__fadd_rn(r[i], __fmul_rn(ng, next_v)), v[i]);
        adv = __fadd_rn(delta, __fmul_rn(__fmul_rn(ng, lam), adv));
        const float ret = __fadd_rn(adv, v[i]);
        s_r[(size_t)i * nb + tid] = ret;
        x[i] = __fsub_rn(ret, v[i]);
        s += (double)x[i];
        ss += (double)x[i] * (double)x[i];
        next_v = v[i];
      }
    }
  }
  GAE_TRACE(1);
  // returns rows leave now (bulk stores from shared memory), overlapping the reduction of the statistics
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid < 32) {
    for (int t = lane; t < T; t += 32) bulk_store(returns + (size_t)t * N + n0, s_r + (size_t)t * nb, (uint32_t)cnt * 4u);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }

  float mean_f = 0.f, denom_f = 1.f;
  if (normalize) {
    if constexpr (CLUSTER) {
      double v2[2] = {s, ss};
      block_sum<2>(v2, s_red);
      if (tid == 0) { s_part[0] = v2[0]; s_part[1] = v2[1]; }
      GAE_TRACE(2);
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
      GAE_TRACE(3);
      if (tid < 32) {   // every CTA sums the partials of all CTAs in the same order: identical bits everywhere
        double t0 = 0.0, t1 = 0.0;
        if (lane < (int)gridDim.x) {
          t0 = ld_cluster_f64(&s_part[0], (uint32_t)lane);
          t1 = ld_cluster_f64(&s_part[1], (uint32_t)lane);
        }
        t0 = warp_sum(t0);
        t1 = warp_sum(t1);
        if (lane == 0) {
          const double n = (double)T * (double)N;
          const double mean = t0 / n;
          double var = (t1 - t0 * t0 / n) / (n - 1.0);  // unbiased, as torch.Tensor.std()
          var = var > 0.0 ? var : 0.0;
          s_red[0] = mean;
          s_red[1] = sqrt(var);
        }
      }
      // nobody leaves (and frees its shared memory) before every CTA has read the partials
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      __syncthreads();
      GAE_TRACE(4);
      mean_f = (float)s_red[0];
      denom_f = __fadd_rn((float)s_red[1], 1e-8f);
    } else {
      gae_finalize_stats(s, ss, partials, s_red, bar, (int64_t)T * N, mean_f, denom_f);
    }
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < TMAX; ++i)
      if (i < T) s_v[(size_t)i * nb + tid] = normalize ? __fdiv_rn(__fsub_rn(x[i], mean_f), denom_f) : x[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid < 32) {
    for (int t = lane; t < T; t += 32) bulk_store(advantages + (size_t)t * N + n0, s_v + (size_t)t * nb, (uint32_t)cnt * 4u);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  if constexpr (CLUSTER) {
    if (normalize) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  GAE_TRACE(5);
}

static int device_sms() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// cooperative launch: the driver guarantees that all blocks are co-resident (or rejects the launch)
template <typename K>
static int launch_coop(K kernel, int blocks, int threads, void** args, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)kernel, args);
  if (e != cudaSuccess) {
    set_error("rslrl_gae_norm: %s", cudaGetErrorString(e));
    return 1;
  }
  count_launch();
  return 0;
}

template <int TMAX>
static int launch_reg(void** args, int N, cudaStream_t st) {
  // at most one block per SM: the block size grows with the env count (128 .. 512 threads)
  const int sms = device_sms();
  int threads = ((N + sms - 1) / sms + 31) / 32 * 32;   // one block per SM, every SM used
  if (threads < 128) threads = 128;
  if (threads > 512) return -1;  // more envs than one wave of register-resident threads: grid-stride kernel
  const int blocks = (N + threads - 1) / threads;
  return launch_coop(gae_norm_reg_kernel<TMAX>, blocks, threads, args, st);
}


// TMA-staged path: 0 launched, 1 error, -1 not eligible
template <int TMAX>
static int launch_tma(void** args_in, const void* const* ptrs, int T, int N, int normalize, cudaStream_t st) {
  for (int i = 0; i < 6; ++i)
    if ((uintptr_t)ptrs[i] & 15) return -1;
  if (N % 16 != 0 || T > TMAX) return -1;
  static const bool off = [] { const char* e = getenv("RSLRL_GAE_TMA"); return e && e[0] == '0'; }();
  if (off) return -1;
  const int sms = device_sms();
  static const int cluster_max = [] { const char* e = getenv("RSLRL_GAE_CLUSTER_MAX"); return e ? atoi(e) : 16; }();
  const bool cluster = normalize && cluster_max >= 1 && N <= cluster_max * 512;
  int blocks, nb;
  if (cluster) {
    blocks = cluster_max;
    while (blocks > 1 && blocks * 32 > N) blocks >>= 1;          // at least 32 envs per CTA
    nb = ((N + blocks - 1) / blocks + 15) / 16 * 16;
    blocks = (N + nb - 1) / nb;
    if (blocks & (blocks - 1)) {                                  // keep the cluster size a power of two
      int p2 = 1;
      while (p2 < blocks) p2 <<= 1;
      blocks = p2 >> 1;
      nb = ((N + blocks - 1) / blocks + 15) / 16 * 16;
      blocks = (N + nb - 1) / nb;
      if (blocks & (blocks - 1)) return -1;
    }
  } else {
    nb = ((N + sms - 1) / sms + 15) / 16 * 16;
    if (nb < 32) nb = 32;
    blocks = (N + nb - 1) / nb;
  }
  if (nb > 512 || blocks > kGaeMaxBlocks) return -1;
  const int threads = (nb + 31) / 32 * 32;
  const size_t smem = (size_t)nb * (9 * (size_t)T + 4);
  void* args[14];
  for (int i = 0; i < 10; ++i) args[i] = args_in[i];
  args[10] = &nb;
  args[11] = args_in[10];
  args[12] = args_in[11];
  args[13] = args_in[12];
  const void* kern = cluster ? (const void*)gae_norm_tma_kernel<TMAX, true> : (const void*)gae_norm_tma_kernel<TMAX, false>;
  static size_t smem_set[2] = {0, 0};
  static int cluster_ok = -1;   // 16-CTA clusters need the non-portable attribute and a GPC that can host them
  if (smem > smem_set[cluster]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)) != cudaSuccess) {
      (void)cudaGetLastError();
      return -1;
    }
    smem_set[cluster] = 200 * 1024;
  }
  if (smem > 200 * 1024) return -1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cluster) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = blocks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    if (cluster_ok < 0) {
      cluster_ok = 0;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
        cudaLaunchConfig_t probe = cfg;
        cudaLaunchAttribute pa[1];
        pa[0] = attr[0];
        pa[0].val.clusterDim.x = 16;
        probe.attrs = pa;
        probe.gridDim = dim3(16);
        probe.blockDim = dim3(512);
        probe.dynamicSmemBytes = 200 * 1024;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, kern, &probe) == cudaSuccess && nc >= 1) cluster_ok = 1;
      }
      (void)cudaGetLastError();
    }
    if (!cluster_ok && blocks > 8) return -1;
  } else {
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = normalize ? 1 : 0;
  }
  cudaError_t e = cudaLaunchKernelExC(&cfg, kern, args);
  if (e != cudaSuccess) {
    set_error("rslrl_gae_norm: %s", cudaGetErrorString(e));
    return 1;
  }
  count_launch();
  return 0;
}

}  // namespace rslrl

extern "C" {

int64_t rslrl_gae_workspace_bytes(int N) {
  (void)N;
  return (int64_t)rslrl::kGaeMaxBlocks * 2 * sizeof(double) + (int64_t)sizeof(rslrl::GaeBarrier) + 8;
}

int rslrl_gae_norm(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                   float* returns, float* advantages, void* workspace, int T, int N, double gamma_d, double lam_d,
                   int normalize, rslrl_stream_t stream) {
  using namespace rslrl;
  float gamma = (float)gamma_d, lam = (float)lam_d;  // python scalars are cast to fp32 at each tensor op
  RSLRL_CHECK_ARG(rewards && values && dones && last_values && returns && advantages && workspace, "rslrl_gae_norm");
  RSLRL_CHECK_ARG(T > 0 && N > 0 && (int64_t)T * N > 1, "rslrl_gae_norm");
  cudaStream_t st = (cudaStream_t)stream;
  double* partials = (double*)workspace;
  GaeBarrier* bar = (GaeBarrier*)(partials + 2 * kGaeMaxBlocks);
  void* args[] = {&rewards, &values, &dones, &last_values, &returns, &advantages, &partials, &bar,
                  &T,       &N,      &gamma, &lam,         &normalize};
  int rc = -1;
  const void* ptrs[6] = {rewards, values, dones, last_values, returns, advantages};
  if (T <= 8) rc = launch_tma<8>(args, ptrs, T, N, normalize, st);
  else if (T <= 16) rc = launch_tma<16>(args, ptrs, T, N, normalize, st);
  else if (T <= 24) rc = launch_tma<24>(args, ptrs, T, N, normalize, st);
  else if (T <= 32) rc = launch_tma<32>(args, ptrs, T, N, normalize, st);
  if (rc >= 0) return rc;
  if (T <= 8) rc = launch_reg<8>(args, N, st);
  else if (T <= 16) rc = launch_reg<16>(args, N, st);
  else if (T <= 24) rc = launch_reg<24>(args, N, st);
  else if (T <= 32) rc = launch_reg<32>(args, N, st);
  if (rc >= 0) return rc;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gae_norm_generic_kernel, kGaeThreads, 0);
  int cap = device_sms() * (per_sm > 0 ? per_sm : 1);
  if (cap > kGaeMaxBlocks) cap = kGaeMaxBlocks;
  const int blocks = (N + kGaeThreads - 1) / kGaeThreads;
  return launch_coop(gae_norm_generic_kernel, blocks < cap ? blocks : cap, kGaeThreads, args, st);
}

#ifdef RSLRL_GAE_TRACE
int rslrl_debug_read_trace_gae(unsigned long long* host) {
  return (int)cudaMemcpyFromSymbol(host, rslrl::g_gae_trace, sizeof(unsigned long long) * 148 * 8);
}
#endif

}  // extern "C"
