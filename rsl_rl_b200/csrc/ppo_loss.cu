// Gaussian policy head (rollout) and the fused PPO objective forward+backward (update).
//
// rslrl_ppo_loss_fwd_bwd replaces ~60 forward and ~100 autograd-backward elementwise/reduction
// launches per minibatch of the reference (rsl_rl/algorithms/ppo.py:255-315 + loss.backward()) by one
// kernel.  One thread owns one sample row: a warp reads 32 consecutive rows = one contiguous
// 32*A*4-byte span of mu / actions / old_mu / old_sigma, so all sectors fetched are used (rows are
// walked in 16-byte chunks when A % 4 == 0).  Reductions (KL, the three loss means, dL/dstd[A]) go
// per-thread fp64 -> warp shuffle -> shared -> one fp64 atomicAdd per block into a workspace; the
// last block to finish (threadfence + ticket) publishes the results and clears the workspace, so
// there is no extra reduction launch, no host involvement and nothing to zero between launches.
// Algorithmic traffic: (20*A + 24) bytes per sample.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace rslrl {

constexpr int kLossThreads = 128;
constexpr int kMaxActions = 64;
constexpr float kLogSqrt2Pi = 0.91893853320467274178f;   // log(sqrt(2 pi))
constexpr float kEntropyConst = 1.41893853320467274178f;  // 0.5 + 0.5 log(2 pi)

// Block partials are added into one of kLossSlots slot sets (slot = blockIdx.x % kLossSlots) instead of one set of
// addresses: same-address fp64 atomics serialise in L2 (~27 clk each), which at 3 000 - 6 000 blocks (C3 minibatch)
// was most of the kernel's time; the last block sums the slots in index order.
constexpr int kLossSlots = 64;
constexpr int kLossAcc = 4 + 2 * kMaxActions;   // kl, surrogate, value loss, sum(dvalue), dstd[A], sum_rows(dmu)[A]
struct LossWorkspace {
  double acc[kLossSlots * kLossAcc];
  unsigned int ticket;
  unsigned int pad_;
};

// last block, all threads: totals over the slots (fixed order) -> s_tot[kLossAcc], slots cleared for the next launch
__device__ __forceinline__ void loss_collect_slots(LossWorkspace* ws, double* s_tot) {
  for (int t = threadIdx.x; t < kLossAcc; t += blockDim.x) {
    double a = 0.0;
    for (int sl = 0; sl < kLossSlots; ++sl) {
      a += __ldcg(&ws->acc[sl * kLossAcc + t]);
      ws->acc[sl * kLossAcc + t] = 0.0;
    }
    s_tot[t] = a;
  }
  __syncthreads();
}

struct SumWorkspace {
  double acc[2];
  unsigned int ticket;
  unsigned int pad_;
};

// ---- dynamic power-of-two scale of the fp16 gradient copies (S * dmu, S * dvalue) ---------------------------------------
// scale_state (device float[8], optional): [0] S of the most recent launch, [1] 1 / S (read by the weight-gradient
// kernel of the same minibatch), [2] S for the NEXT launch, [3] running max |gradient| of the current launch (uint bits),
// [4] saturation events (uint): launches in which S * max|gradient| left fp16's finite range.  Every launch scales with
// the value the previous one left in [2] (chosen so that S * max lands in [2^11, 2^12): 16x headroom for growth between
// consecutive minibatches, 2^-26 of the maximum still representable); the last block of a launch publishes [0], [1],
// counts a saturation if there was one and derives [2] from the maximum it saw.
#define RSLRL_LOSS_SCALE_PROLOGUE                                                                     \
  const float scale16 = scale_state != nullptr ? __ldcg(scale_state + 2) : scale16_arg;               \
  __shared__ unsigned int s_gmax;                                                                     \
  if (threadIdx.x == 0) s_gmax = 0u;

// saturating conversion guard: a scaled gradient beyond fp16's range becomes +-65504 (counted by the caller), never inf
__device__ __forceinline__ float sat16(float x) { return fminf(fmaxf(x, -65504.f), 65504.f); }

__device__ __forceinline__ void loss_scale_track(float* scale_state, unsigned int* s_gmax, float local_max) {
  if (scale_state == nullptr) return;
  float m = local_max;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(s_gmax, __float_as_uint(m));
}
__device__ __forceinline__ void loss_scale_block_publish(float* scale_state, const unsigned int* s_gmax) {
  // after a __syncthreads: one global atomic per block
  if (scale_state != nullptr && threadIdx.x == 0 && *s_gmax != 0u) atomicMax((unsigned int*)(scale_state + 3), *s_gmax);
}
__device__ __forceinline__ void loss_scale_finalize(float* scale_state, float S) {
  // last block, one thread, after the ticket fence
  if (scale_state == nullptr) return;
  const float mx = __uint_as_float(*((volatile unsigned int*)(scale_state + 3)));
  scale_state[0] = S;
  scale_state[1] = 1.0f / S;
  if (mx * S > 60000.0f) atomicAdd((unsigned int*)(scale_state + 4), 1u);
  float next = S;
  if (mx > 0.f && isfinite(mx)) {
    next = exp2f(floorf(log2f(4096.0f / mx)));
    next = fminf(fmaxf(next, 9.5367431640625e-07f), 1073741824.0f);   // [2^-20, 2^30]
  }
  scale_state[2] = next;
  *((unsigned int*)(scale_state + 3)) = 0u;
}

__global__ void __launch_bounds__(kLossThreads)
gaussian_act_kernel(const float* __restrict__ mu, const float* __restrict__ std_param, int std_type,
                    const float* __restrict__ z, float* __restrict__ actions, float* __restrict__ log_prob,
                    float* __restrict__ sigma_out, long long n, int A) {
  __shared__ float s_sigma[kMaxActions], s_logs[kMaxActions], s_i2v[kMaxActions];
  if (threadIdx.x < A) {
    const float p = std_param[threadIdx.x];
    const float s = std_type == RSLRL_STD_LOG ? expf(p) : p;
    s_sigma[threadIdx.x] = s;
    s_logs[threadIdx.x] = logf(s);
    s_i2v[threadIdx.x] = 2.f * (s * s);
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * kLossThreads + threadIdx.x;
  if (i >= n) return;
  const float* m = mu + i * A;
  const float* zz = z + i * A;
  float* out = actions + i * A;
  float lp = 0.f;
  for (int a = 0; a < A; ++a) {
    const float mean = m[a], s = s_sigma[a];
    const float act = __fadd_rn(__fmul_rn(zz[a], s), mean);  // normal_().mul_(std).add_(mean)
    out[a] = act;
    const float d = act - mean;
    lp += -(d * d) / s_i2v[a] - s_logs[a] - kLogSqrt2Pi;
    if (sigma_out) sigma_out[i * A + a] = s;
  }
  log_prob[i] = lp;
}

template <bool VEC>
__global__ void __launch_bounds__(kLossThreads)
ppo_loss_kernel(const float* __restrict__ mu, const float* __restrict__ std_param, int std_type,
                const float* __restrict__ value, const float* __restrict__ actions,
                const float* __restrict__ old_logp, const float* __restrict__ advantages,
                const float* __restrict__ returns, const float* __restrict__ old_values,
                const float* __restrict__ old_mu, const float* __restrict__ old_sigma,
                const double* __restrict__ adv_stats, float* __restrict__ dmu, int dmu_stride,
                float* __restrict__ dvalue, int dvalue_stride, float* dstd, float* db_mu, float* db_value,
                float* kl_out, double* loss_sums, LossWorkspace* ws, long long rows, long long mb_orig,
                int A, float clip, float c_v, float c_e, int use_clipped, __half* __restrict__ dmu16, int dmu16_stride,
                __half* __restrict__ dvalue16, int dvalue16_stride, float scale16_arg, float* scale_state) {
  pdl_trigger();
  pdl_wait();
  RSLRL_LOSS_SCALE_PROLOGUE
  __shared__ float s_sigma[kMaxActions], s_logs[kMaxActions], s_2var[kMaxActions];
  __shared__ double s_red[4 * 32];
  __shared__ float s_dstd[kMaxActions], s_dbmu[kMaxActions];
  __shared__ bool s_last;
  __shared__ double s_tot[kLossAcc];
  if (threadIdx.x < A) {
    const float p = std_param[threadIdx.x];
    const float s = std_type == RSLRL_STD_LOG ? expf(p) : p;
    s_sigma[threadIdx.x] = s;
    s_logs[threadIdx.x] = logf(s);
    s_2var[threadIdx.x] = 2.f * (s * s);
    s_dstd[threadIdx.x] = 0.f;
    s_dbmu[threadIdx.x] = 0.f;
  }
  __syncthreads();

  const long long i = (long long)blockIdx.x * kLossThreads + threadIdx.x;
  const bool valid = i < rows;
  const long long j = valid ? i % mb_orig : 0;
  const float lo = 1.0f - clip, hi = 1.0f + clip;
  const float g = 1.0f / (float)rows;

  float lp = 0.f, kl = 0.f;
  if (valid) {
    const float* m = mu + i * A;
    const float* ac = actions + i * A;
    const bool do_kl = i < mb_orig;
    const float* om = old_mu + j * A;
    const float* os = old_sigma + j * A;
    if (VEC) {
      for (int c = 0; c < A; c += 4) {
        const float4 m4 = *(const float4*)(m + c), a4 = *(const float4*)(ac + c);
        const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, aa[4] = {a4.x, a4.y, a4.z, a4.w};
        float omv[4] = {0, 0, 0, 0}, osv[4] = {1, 1, 1, 1};
        if (do_kl) {
          const float4 o4 = __ldg((const float4*)(om + c)), s4 = __ldg((const float4*)(os + c));
          omv[0] = o4.x; omv[1] = o4.y; omv[2] = o4.z; omv[3] = o4.w;
          osv[0] = s4.x; osv[1] = s4.y; osv[2] = s4.z; osv[3] = s4.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float d = aa[q] - mm[q];
          lp += -(d * d) / s_2var[c + q] - s_logs[c + q] - kLogSqrt2Pi;
          if (do_kl) {
            const float dm = omv[q] - mm[q];
            kl += logf(s_sigma[c + q] / osv[q] + 1.0e-5f) + (osv[q] * osv[q] + dm * dm) / s_2var[c + q] - 0.5f;
          }
        }
      }
    } else {
      for (int a = 0; a < A; ++a) {
        const float d = ac[a] - m[a];
        lp += -(d * d) / s_2var[a] - s_logs[a] - kLogSqrt2Pi;
        if (do_kl) {
          const float dm = om[a] - m[a], so = os[a];
          kl += logf(s_sigma[a] / so + 1.0e-5f) + (so * so + dm * dm) / s_2var[a] - 0.5f;
        }
      }
    }
  }

  float g_lp = 0.f, surr = 0.f, vl = 0.f, dv_sum = 0.f, gmax = 0.f;
  if (valid) {
    float adv = __ldg(advantages + j);
    if (adv_stats) {  // per-minibatch normalisation (ppo.py:221-223), unbiased std
      const double n = (double)mb_orig;
      const double mean = adv_stats[0] / n;
      double var = (adv_stats[1] - adv_stats[0] * adv_stats[0] / n) / (n - 1.0);
      var = var > 0.0 ? var : 0.0;
      adv = (adv - (float)mean) / ((float)sqrt(var) + 1e-8f);
    }
    const float ratio = expf(lp - __ldg(old_logp + j));
    const float s1 = -adv * ratio;
    const float clamped = fminf(fmaxf(ratio, lo), hi);
    const float s2 = -adv * clamped;
    surr = fmaxf(s1, s2);
    // aten maximum backward: grad to the larger input, grad/2 to both on ties
    const float g1 = s1 > s2 ? g : (s1 == s2 ? 0.5f * g : 0.f);
    const float g2 = s2 > s1 ? g : (s1 == s2 ? 0.5f * g : 0.f);
    const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;  // clamp backward mask (inclusive)
    const float g_ratio = g1 * (-adv) + g2 * (-adv) * in_range;
    g_lp = g_ratio * ratio;

    const float v = value[i], R = __ldg(returns + j);
    float dv_out;
    if (use_clipped) {
      const float vo = __ldg(old_values + j);
      const float dv = v - vo;
      const float vclip = vo + fminf(fmaxf(dv, -clip), clip);
      const float e1 = v - R, e2 = vclip - R;
      const float l1 = e1 * e1, l2 = e2 * e2;
      vl = fmaxf(l1, l2);
      const float ga = l1 > l2 ? g : (l1 == l2 ? 0.5f * g : 0.f);
      const float gb = l2 > l1 ? g : (l1 == l2 ? 0.5f * g : 0.f);
      const float in_v = (dv >= -clip && dv <= clip) ? 1.f : 0.f;
      dv_out = c_v * (ga * 2.f * e1 + gb * 2.f * e2 * in_v);
    } else {
      const float e = R - v;
      vl = e * e;
      dv_out = c_v * g * (-2.f) * e;
    }
    dvalue[i * dvalue_stride] = dv_out;
    if (dvalue16 != nullptr) dvalue16[i * dvalue16_stride] = __float2half_rn(sat16(dv_out * scale16));
    dv_sum = dv_out;
    gmax = fabsf(dv_out);
  }

  // second pass over the row: dL/dmu and the per-action dL/dsigma column sums
  {
    const float* m = mu + (valid ? i : 0) * A;
    const float* ac = actions + (valid ? i : 0) * A;
    float* dm = dmu + (valid ? i : 0) * dmu_stride;
    const int lane = threadIdx.x & 31;
    for (int c = 0; c < A; c += 4) {
      float d[4] = {0, 0, 0, 0};
      const int nq = A - c < 4 ? A - c : 4;
      if (valid) {
        if (VEC) {
          const float4 m4 = *(const float4*)(m + c), a4 = *(const float4*)(ac + c);
          d[0] = a4.x - m4.x; d[1] = a4.y - m4.y; d[2] = a4.z - m4.z; d[3] = a4.w - m4.w;
        } else {
          for (int q = 0; q < nq; ++q) d[q] = ac[c + q] - m[c + q];
        }
      }
      float o[4] = {0, 0, 0, 0};
      for (int q = 0; q < nq; ++q) {
        const float s = s_sigma[c + q];
        const float var = s * s;
        o[q] = g_lp * d[q] / var;
        float ds = valid ? g_lp * (d[q] * d[q] / (var * s) - 1.f / s) : 0.f;
        ds = warp_sum(ds);
        if (lane == 0) atomicAdd(&s_dstd[c + q], ds);
        if (db_mu != nullptr) {
          const float sm = warp_sum(o[q]);
          if (lane == 0) atomicAdd(&s_dbmu[c + q], sm);
        }
      }
      if (valid) {
        if (VEC) *(float4*)(dm + c) = make_float4(o[0], o[1], o[2], o[3]);
        else for (int q = 0; q < nq; ++q) dm[c + q] = o[q];
        gmax = fmaxf(fmaxf(gmax, fmaxf(fabsf(o[0]), fabsf(o[1]))), fmaxf(fabsf(o[2]), fabsf(o[3])));
        if (VEC && dmu16 != nullptr) {  // scaled fp16 copy (A % 4 == 0): the A operand of the fp16 output-layer backward
          const __half2 lo2 = __floats2half2_rn(sat16(o[0] * scale16), sat16(o[1] * scale16)), hi2 = __floats2half2_rn(sat16(o[2] * scale16), sat16(o[3] * scale16));
          uint2 w;
          w.x = *(const unsigned int*)&lo2;
          w.y = *(const unsigned int*)&hi2;
          *(uint2*)(dmu16 + i * dmu16_stride + c) = w;
        }
      }
    }
  }

  loss_scale_track(scale_state, &s_gmax, gmax);
  double red[4] = {(double)kl, (double)surr, (double)vl, (double)dv_sum};
  block_sum<4>(red, s_red);   // (its barriers also order the shared-memory maximum)
  loss_scale_block_publish(scale_state, &s_gmax);
  double* const slot = ws->acc + (size_t)(blockIdx.x & (kLossSlots - 1)) * kLossAcc;
  if (threadIdx.x == 0) {
    atomicAdd(&slot[0], red[0]);
    atomicAdd(&slot[1], red[1]);
    atomicAdd(&slot[2], red[2]);
    atomicAdd(&slot[3], red[3]);
  }
  __syncthreads();
  if (threadIdx.x < A) {
    atomicAdd(&slot[4 + threadIdx.x], (double)s_dstd[threadIdx.x]);
    if (db_mu != nullptr) atomicAdd(&slot[4 + kMaxActions + threadIdx.x], (double)s_dbmu[threadIdx.x]);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last block: publish and clear
  loss_collect_slots(ws, s_tot);
  if (threadIdx.x < A) {
    const int a = threadIdx.x;
    const float s = s_sigma[a];
    // entropy bonus -c_e * mean_i sum_a log sigma_a  ->  d/dsigma_a = -c_e / sigma_a (x sigma for log-std)
    double gsum = s_tot[4 + a];
    if (std_type == RSLRL_STD_LOG) gsum = gsum * (double)s - (double)c_e;
    else gsum = gsum - (double)c_e / (double)s;
    dstd[a] += (float)gsum;
    if (db_mu != nullptr) db_mu[a] += (float)s_tot[4 + kMaxActions + a];
  }
  if (threadIdx.x == 0) {
    const double klm = s_tot[0] / (double)mb_orig;
    if (kl_out) *kl_out = (float)klm;
    if (db_value) *db_value += (float)s_tot[3];
    float ent = 0.f;
    for (int a = 0; a < A; ++a) ent += kEntropyConst + s_logs[a];
    if (loss_sums) {
      loss_sums[0] += s_tot[1] / (double)rows;
      loss_sums[1] += s_tot[2] / (double)rows;
      loss_sums[2] += (double)ent;
    }
    loss_scale_finalize(scale_state, scale16);
    __threadfence();
    ws->ticket = 0u;
  }
}

// Same objective, four threads per sample row (thread <-> 4 consecutive actions, A <= 16, 16-byte aligned rows).
// The row-per-thread kernel above leaves ~5 warps on an SM at the C2 minibatch (24 576 rows), i.e. one warp per
// scheduler with a ~2 800-instruction dependent chain each; here a row's chain is ~4x shorter and there are 4x more
// warps to hide its latencies.  Row sums (log-prob) are quad shuffles, the per-action column sums (dL/dsigma) are
// strided warp shuffles; everything else is identical, including the tie rules and the last-block publish.
// Optional fused data gradient of the two OUTPUT layers (fp16-storage path): with `head.dz_a != nullptr` the kernel also
// writes  dZ3 = S * (dL/dmu W4_actor) * act'(H3_actor)  and  S * (dL/dV W4_critic) * act'(H3_critic)  as fp16 -- the
// gradients at the output of the last hidden layer, i.e. what a separate K = 12 / K = 1 tensor-core launch produced
// before (12 us per minibatch at 4 % tensor activity, for 25 MB of traffic).  The block's 64 rows of S * dL/dmu (fp16, the
// same values the dmu16 copy holds) and S * dL/dV are staged in shared memory; warps 0-3 then run the actor's
// [64 x 16] x [16 x K3] product as warp-level m16n8k16 MMAs (fp16 operands, fp32 accumulation: the arithmetic of the
// tensor-core launch it replaces) with the column order of the B fragments chosen so that every lane ends up with 8
// CONSECUTIVE hidden units of its two rows (one 16-byte load of H3 and one 16-byte store of dZ3 per 32-column group);
// warps 4-7 do the critic's rank-1 product elementwise, 16 lanes per 256-byte row.
struct LossHead {
  const float* w_a;    // [A, K3] actor output-layer weight (fp32 master)
  const float* w_c;    // [1, K3] critic output-layer weight
  const __half* h_a;   // [rows, K3] saved activations of the last hidden layer (row stride ldh)
  const __half* h_c;
  __half* dz_a;        // [rows, K3] outputs (row stride ldh)
  __half* dz_c;
  int K3, ldh, act;
};
constexpr int kHeadMaxK3 = 128;
constexpr int kHeadPitch = 24;   // halfs per shared-memory row of the head operands (48 bytes)

constexpr int kQuadThreads = 256;
__global__ void __launch_bounds__(kQuadThreads)
ppo_loss_quad_kernel(const float* __restrict__ mu, const float* __restrict__ std_param, int std_type,
                     const float* __restrict__ value, const float* __restrict__ actions,
                     const float* __restrict__ old_logp, const float* __restrict__ advantages,
                     const float* __restrict__ returns, const float* __restrict__ old_values,
                     const float* __restrict__ old_mu, const float* __restrict__ old_sigma,
                     const double* __restrict__ adv_stats, float* __restrict__ dmu, int dmu_stride,
                     float* __restrict__ dvalue, int dvalue_stride, float* dstd, float* db_mu, float* db_value,
                     float* kl_out, double* loss_sums, LossWorkspace* ws, long long rows, long long mb_orig,
                     int A, float clip, float c_v, float c_e, int use_clipped, __half* __restrict__ dmu16,
                     int dmu16_stride, __half* __restrict__ dvalue16, int dvalue16_stride, float scale16_arg,
                     float* scale_state, const LossHead head) {
  pdl_trigger();
  pdl_wait();
  RSLRL_LOSS_SCALE_PROLOGUE
  // fused output-layer data gradient: W4^T of the actor as fp16 [K3][16 actions] and the block's S * dL/dmu rows, both
  // with a 48-byte row pitch (conflict-free 4-byte fragment loads); critic weight (fp16-rounded) and S * dL/dV as floats
  __shared__ __align__(16) __half s_wt[kHeadMaxK3 * kHeadPitch];
  __shared__ __align__(16) __half s_dm16[(kQuadThreads / 4) * kHeadPitch];
  __shared__ __align__(16) float s_wc[kHeadMaxK3];
  __shared__ float s_dv[kQuadThreads / 4];
  if (head.dz_a != nullptr) {
    for (int t = threadIdx.x; t < 16 * head.K3; t += kQuadThreads) {
      const int col = t >> 4, a = t & 15;
      s_wt[col * kHeadPitch + a] = __float2half_rn(a < A ? __ldg(head.w_a + a * head.K3 + col) : 0.f);
    }
    for (int t = threadIdx.x; t < head.K3; t += kQuadThreads) s_wc[t] = __half2float(__float2half_rn(__ldg(head.w_c + t)));
  }
  __shared__ float s_sigma[kMaxActions], s_logs[kMaxActions], s_2var[kMaxActions];
  __shared__ double s_red[4 * 32];
  __shared__ float s_dstd[kMaxActions], s_dbmu[kMaxActions];
  __shared__ bool s_last;
  __shared__ double s_tot[kLossAcc];
  if (threadIdx.x < A) {
    const float p = std_param[threadIdx.x];
    const float s = std_type == RSLRL_STD_LOG ? expf(p) : p;
    s_sigma[threadIdx.x] = s;
    s_logs[threadIdx.x] = logf(s);
    s_2var[threadIdx.x] = 2.f * (s * s);
    s_dstd[threadIdx.x] = 0.f;
    s_dbmu[threadIdx.x] = 0.f;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, qd = threadIdx.x & 3, c = qd * 4;
  const long long i = (long long)blockIdx.x * (kQuadThreads / 4) + (threadIdx.x >> 2);
  const bool valid = i < rows;
  const bool mine = valid && c < A;          // this thread owns actions [c, c + 4) of row i
  const bool lead = valid && qd == 0;        // the thread of the quad that accounts for the per-row scalars
  const long long j = valid ? i % mb_orig : 0;
  const float lo = 1.0f - clip, hi = 1.0f + clip;
  const float g = 1.0f / (float)rows;

  float lp = 0.f, kl = 0.f;
  float d[4] = {0.f, 0.f, 0.f, 0.f};
  if (mine) {
    const float4 m4 = *(const float4*)(mu + i * A + c), a4 = *(const float4*)(actions + i * A + c);
    const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, aa[4] = {a4.x, a4.y, a4.z, a4.w};
    const bool do_kl = i < mb_orig;
    float omv[4] = {0, 0, 0, 0}, osv[4] = {1, 1, 1, 1};
    if (do_kl) {
      const float4 o4 = __ldg((const float4*)(old_mu + j * A + c)), s4 = __ldg((const float4*)(old_sigma + j * A + c));
      omv[0] = o4.x; omv[1] = o4.y; omv[2] = o4.z; omv[3] = o4.w;
      osv[0] = s4.x; osv[1] = s4.y; osv[2] = s4.z; osv[3] = s4.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      d[q] = aa[q] - mm[q];
      lp += -(d[q] * d[q]) / s_2var[c + q] - s_logs[c + q] - kLogSqrt2Pi;
      if (do_kl) {
        const float dm = omv[q] - mm[q];
        kl += logf(s_sigma[c + q] / osv[q] + 1.0e-5f) + (osv[q] * osv[q] + dm * dm) / s_2var[c + q] - 0.5f;
      }
    }
  }
  // log-prob of the row: sum over the quad (every thread of the quad ends up with the same value)
  lp += __shfl_xor_sync(0xffffffffu, lp, 1);
  lp += __shfl_xor_sync(0xffffffffu, lp, 2);

  float g_lp = 0.f, surr = 0.f, vl = 0.f, dv_sum = 0.f, gmax = 0.f;
  if (valid) {
    float adv = __ldg(advantages + j);
    if (adv_stats) {  // per-minibatch normalisation (ppo.py:221-223), unbiased std
      const double n = (double)mb_orig;
      const double mean = adv_stats[0] / n;
      double var = (adv_stats[1] - adv_stats[0] * adv_stats[0] / n) / (n - 1.0);
      var = var > 0.0 ? var : 0.0;
      adv = (adv - (float)mean) / ((float)sqrt(var) + 1e-8f);
    }
    const float ratio = expf(lp - __ldg(old_logp + j));
    const float s1 = -adv * ratio;
    const float clamped = fminf(fmaxf(ratio, lo), hi);
    const float s2 = -adv * clamped;
    // aten maximum backward: grad to the larger input, grad/2 to both on ties
    const float g1 = s1 > s2 ? g : (s1 == s2 ? 0.5f * g : 0.f);
    const float g2 = s2 > s1 ? g : (s1 == s2 ? 0.5f * g : 0.f);
    const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;  // clamp backward mask (inclusive)
    g_lp = (g1 * (-adv) + g2 * (-adv) * in_range) * ratio;
    if (lead) {
      surr = fmaxf(s1, s2);
      const float v = value[i], R = __ldg(returns + j);
      float dv_out;
      if (use_clipped) {
        const float vo = __ldg(old_values + j);
        const float dv = v - vo;
        const float vclip = vo + fminf(fmaxf(dv, -clip), clip);
        const float e1 = v - R, e2 = vclip - R;
        const float l1 = e1 * e1, l2 = e2 * e2;
        vl = fmaxf(l1, l2);
        const float ga = l1 > l2 ? g : (l1 == l2 ? 0.5f * g : 0.f);
        const float gb = l2 > l1 ? g : (l1 == l2 ? 0.5f * g : 0.f);
        const float in_v = (dv >= -clip && dv <= clip) ? 1.f : 0.f;
        dv_out = c_v * (ga * 2.f * e1 + gb * 2.f * e2 * in_v);
      } else {
        const float e = R - v;
        vl = e * e;
        dv_out = c_v * g * (-2.f) * e;
      }
      dvalue[i * dvalue_stride] = dv_out;
      if (dvalue16 != nullptr) dvalue16[i * dvalue16_stride] = __float2half_rn(sat16(dv_out * scale16));
      dv_sum = dv_out;
      gmax = fabsf(dv_out);
    }
  }

  // dL/dmu for this thread's four actions and the per-action dL/dsigma column sums (lanes with equal (lane & 3))
  {
    float o[4], ds[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float s = c + q < A ? s_sigma[c + q] : 1.f;
      const float var = s * s;
      o[q] = g_lp * d[q] / var;
      ds[q] = mine ? g_lp * (d[q] * d[q] / (var * s) - 1.f / s) : 0.f;
    }
    if (mine) {
      gmax = fmaxf(fmaxf(gmax, fmaxf(fabsf(o[0]), fabsf(o[1]))), fmaxf(fabsf(o[2]), fabsf(o[3])));
      *(float4*)(dmu + i * dmu_stride + c) = make_float4(o[0], o[1], o[2], o[3]);
      if (dmu16 != nullptr) {  // scaled fp16 copy: the A operand of the fp16 output-layer backward
        const __half2 lo2 = __floats2half2_rn(sat16(o[0] * scale16), sat16(o[1] * scale16)), hi2 = __floats2half2_rn(sat16(o[2] * scale16), sat16(o[3] * scale16));
        uint2 w;
        w.x = *(const unsigned int*)&lo2;
        w.y = *(const unsigned int*)&hi2;
        *(uint2*)(dmu16 + i * dmu16_stride + c) = w;
      }
    }
    if (head.dz_a != nullptr) {   // uniform branch: fused data gradient of the two output layers
      const int rl = threadIdx.x >> 2;   // row of this quad inside the block
      {
        const __half2 lo2 = __floats2half2_rn(mine ? sat16(o[0] * scale16) : 0.f, mine ? sat16(o[1] * scale16) : 0.f);
        const __half2 hi2 = __floats2half2_rn(mine ? sat16(o[2] * scale16) : 0.f, mine ? sat16(o[3] * scale16) : 0.f);
        uint2 w;
        w.x = *(const unsigned int*)&lo2;
        w.y = *(const unsigned int*)&hi2;
        *(uint2*)(s_dm16 + rl * kHeadPitch + c) = w;
        if (qd == 0) s_dv[rl] = lead ? __half2float(__float2half_rn(sat16(dv_sum * scale16))) : 0.f;
      }
      __syncthreads();
      const int warp = threadIdx.x >> 5;
      const long long row0 = (long long)blockIdx.x * (kQuadThreads / 4);
      if (warp < 4) {
        // ---- actor: rows [16 warp, 16 warp + 16) x K3 columns on m16n8k16 MMAs
        const int gq = lane >> 2, tq = lane & 3, r0 = 16 * warp;
        const unsigned int a0 = *(const unsigned int*)(s_dm16 + (r0 + gq) * kHeadPitch + 2 * tq);
        const unsigned int a1 = *(const unsigned int*)(s_dm16 + (r0 + gq + 8) * kHeadPitch + 2 * tq);
        const unsigned int a2 = *(const unsigned int*)(s_dm16 + (r0 + gq) * kHeadPitch + 2 * tq + 8);
        const unsigned int a3 = *(const unsigned int*)(s_dm16 + (r0 + gq + 8) * kHeadPitch + 2 * tq + 8);
        const long long ra = row0 + r0 + gq, rb = ra + 8;
        for (int base = 0; base < head.K3; base += 32) {
          float acc[4][4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            // B fragment of MMA jj: n index gq <-> hidden unit base + (gq >> 1) * 8 + 2 jj + (gq & 1), so that this lane's
            // accumulator columns (n = 2 tq, 2 tq + 1) over jj = 0..3 are the 8 consecutive units base + 8 tq .. + 7
            const int col = base + (gq >> 1) * 8 + 2 * jj + (gq & 1);
            const unsigned int b0 = *(const unsigned int*)(s_wt + col * kHeadPitch + 2 * tq);
            const unsigned int b1 = *(const unsigned int*)(s_wt + col * kHeadPitch + 2 * tq + 8);
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                : "=f"(acc[jj][0]), "=f"(acc[jj][1]), "=f"(acc[jj][2]), "=f"(acc[jj][3])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
          }
#pragma unroll
          for (int half_ = 0; half_ < 2; ++half_) {
            const long long r = half_ ? rb : ra;
            if (r < rows) {
              const uint4 hv = *(const uint4*)(head.h_a + r * head.ldh + base + 8 * tq);
              const __half2* h2 = (const __half2*)&hv;
              uint4 ov;
              __half2* o2 = (__half2*)&ov;
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const float2 f = __half22float2(h2[jj]);
                o2[jj] = __floats2half2_rn(sat16(acc[jj][2 * half_] * act_grad_from_output(head.act, f.x)),
                                           sat16(acc[jj][2 * half_ + 1] * act_grad_from_output(head.act, f.y)));
              }
              *(uint4*)(head.dz_a + r * head.ldh + base + 8 * tq) = ov;
            }
          }
        }
      } else {
        // ---- critic: dZ3[r][k] = (S dL/dV[r]) * w[k] * act'(H3[r][k]); consecutive lanes walk one row in 16-byte chunks
        const int chunks = head.K3 >> 3, tasks = (kQuadThreads / 4) * chunks;
        for (int task = threadIdx.x - 128; task < tasks; task += 128) {
          const int rr = task / chunks, ch = task - rr * chunks;
          const long long r = row0 + rr;
          if (r >= rows) continue;
          const float dvr = s_dv[rr];
          const float4 v0 = *(const float4*)(s_wc + 8 * ch), v1 = *(const float4*)(s_wc + 8 * ch + 4);
          const float wv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          const uint4 hv = *(const uint4*)(head.h_c + r * head.ldh + 8 * ch);
          const __half2* h2 = (const __half2*)&hv;
          uint4 ov;
          __half2* o2 = (__half2*)&ov;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 f = __half22float2(h2[u]);
            o2[u] = __floats2half2_rn(sat16(dvr * wv[2 * u] * act_grad_from_output(head.act, f.x)),
                                      sat16(dvr * wv[2 * u + 1] * act_grad_from_output(head.act, f.y)));
          }
          *(uint4*)(head.dz_c + r * head.ldh + 8 * ch) = ov;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) ds[q] += __shfl_xor_sync(0xffffffffu, ds[q], off);
      if (db_mu != nullptr) {
        float sm = mine ? o[q] : 0.f;
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) sm += __shfl_xor_sync(0xffffffffu, sm, off);
        if (lane < 4 && c + q < A) atomicAdd(&s_dbmu[c + q], sm);
      }
      if (lane < 4 && c + q < A) atomicAdd(&s_dstd[c + q], ds[q]);
    }
  }

  loss_scale_track(scale_state, &s_gmax, gmax);
  double red[4] = {(double)kl, (double)surr, (double)vl, (double)dv_sum};
  block_sum<4>(red, s_red);   // (its barriers also order the shared-memory maximum)
  loss_scale_block_publish(scale_state, &s_gmax);
  double* const slot = ws->acc + (size_t)(blockIdx.x & (kLossSlots - 1)) * kLossAcc;
  if (threadIdx.x == 0) {
    atomicAdd(&slot[0], red[0]);
    atomicAdd(&slot[1], red[1]);
    atomicAdd(&slot[2], red[2]);
    atomicAdd(&slot[3], red[3]);
  }
  __syncthreads();
  if (threadIdx.x < A) {
    atomicAdd(&slot[4 + threadIdx.x], (double)s_dstd[threadIdx.x]);
    if (db_mu != nullptr) atomicAdd(&slot[4 + kMaxActions + threadIdx.x], (double)s_dbmu[threadIdx.x]);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last block: publish and clear (as in ppo_loss_kernel)
  loss_collect_slots(ws, s_tot);
  if (threadIdx.x < A) {
    const int a = threadIdx.x;
    const float s = s_sigma[a];
    // entropy bonus -c_e * mean_i sum_a log sigma_a  ->  d/dsigma_a = -c_e / sigma_a (x sigma for log-std)
    double gsum = s_tot[4 + a];
    if (std_type == RSLRL_STD_LOG) gsum = gsum * (double)s - (double)c_e;
    else gsum = gsum - (double)c_e / (double)s;
    dstd[a] += (float)gsum;
    if (db_mu != nullptr) db_mu[a] += (float)s_tot[4 + kMaxActions + a];
  }
  if (threadIdx.x == 0) {
    const double klm = s_tot[0] / (double)mb_orig;
    if (kl_out) *kl_out = (float)klm;
    if (db_value) *db_value += (float)s_tot[3];
    float ent = 0.f;
    for (int a = 0; a < A; ++a) ent += kEntropyConst + s_logs[a];
    if (loss_sums) {
      loss_sums[0] += s_tot[1] / (double)rows;
      loss_sums[1] += s_tot[2] / (double)rows;
      loss_sums[2] += (double)ent;
    }
    loss_scale_finalize(scale_state, scale16);
    __threadfence();
    ws->ticket = 0u;
  }
}

__global__ void __launch_bounds__(256)
sum_sumsq_kernel(const float* __restrict__ x, long long n, double* out, SumWorkspace* ws) {
  __shared__ double s_red[2 * 32];
  __shared__ bool s_last;
  double v[2] = {0.0, 0.0};
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const double a = (double)__ldg(x + i);
    v[0] += a;
    v[1] += a * a;
  }
  block_sum<2>(v, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(&ws->acc[0], v[0]);
    atomicAdd(&ws->acc[1], v[1]);
    __threadfence();
    s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    out[0] = *((volatile double*)&ws->acc[0]);
    out[1] = *((volatile double*)&ws->acc[1]);
    ws->acc[0] = ws->acc[1] = 0.0;
    __threadfence();
    ws->ticket = 0u;
  }
}

}  // namespace rslrl

extern "C" {

int rslrl_gaussian_act(const float* mu, const float* std_param, int std_type, const float* z, float* actions,
                       float* log_prob, float* sigma_out, int64_t n, int A, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_gaussian_act";
  RSLRL_CHECK_ARG(mu && std_param && z && actions && log_prob, what);
  RSLRL_CHECK_ARG(n > 0 && A > 0 && A <= kMaxActions, what);
  const int blocks = (int)((n + kLossThreads - 1) / kLossThreads);
  gaussian_act_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(mu, std_param, std_type, z, actions,
                                                                          log_prob, sigma_out, n, A);
  return finish_launch(what);
}

int64_t rslrl_loss_workspace_bytes(void) { return (int64_t)sizeof(rslrl::LossWorkspace); }

int rslrl_ppo_loss_fwd_bwd(const float* mu, const float* std_param, int std_type, const float* value,
                           const float* actions, const float* old_logp, const float* advantages,
                           const float* returns, const float* old_values, const float* old_mu,
                           const float* old_sigma, const double* adv_stats, float* dmu, int dmu_stride,
                           float* dvalue, int dvalue_stride, float* dstd, float* db_mu, float* db_value,
                           float* kl_out, double* loss_sums, void* workspace, int64_t rows,
                           int64_t mb_orig, int A, double clip_param, double value_loss_coef,
                           double entropy_coef, int use_clipped_value_loss, void* dmu16, int dmu16_stride,
                           void* dvalue16, int dvalue16_stride, double grad_scale16, float* scale_state,
                           const rslrl_loss_head_t* head_host, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_ppo_loss_fwd_bwd";
  RSLRL_CHECK_ARG(mu && std_param && value && actions && old_logp && advantages && returns && old_values, what);
  RSLRL_CHECK_ARG(old_mu && old_sigma && dmu && dvalue && dstd && workspace, what);
  RSLRL_CHECK_ARG(rows > 0 && mb_orig > 0 && rows % mb_orig == 0 && A > 0 && A <= kMaxActions, what);
  RSLRL_CHECK_ARG(dmu_stride >= A && dvalue_stride >= 1, what);
  const int blocks = (int)((rows + kLossThreads - 1) / kLossThreads);
  const bool vec = (A % 4 == 0) && (dmu_stride % 4 == 0) &&
                   ((((uintptr_t)mu | (uintptr_t)actions | (uintptr_t)old_mu | (uintptr_t)old_sigma |
                      (uintptr_t)dmu) & 15) == 0);
  static const bool quad_ok = [] { const char* e = getenv("RSLRL_LOSS_QUAD"); return !(e && e[0] == '0'); }();
  RSLRL_CHECK_ARG(dmu16 == nullptr || (dmu16_stride % 4 == 0 && ((uintptr_t)dmu16 & 7) == 0), what);
  // four threads per row while a thread per row could not fill the machine (measured at 393 216 rows: 58 us for the
  // row-per-thread kernel, 80 us for this one); beyond that the row-per-thread kernel
  static const long long quad_max_rows = [] {
    const char* e = getenv("RSLRL_LOSS_QUAD_MAX_ROWS");
    return e ? atoll(e) : (long long)kNumSMs * 1024;
  }();
  LossHead head{};
  if (head_host != nullptr && head_host->dz_actor != nullptr) {
    const rslrl_loss_head_t& h = *head_host;
    RSLRL_CHECK_ARG(h.w_actor && h.w_critic && h.h_actor && h.h_critic && h.dz_actor && h.dz_critic, what);
    RSLRL_CHECK_ARG(h.K3 > 0 && h.K3 <= kHeadMaxK3 && (h.K3 % 32) == 0 && h.ldh >= h.K3 && (h.ldh % 8) == 0, what);
    RSLRL_CHECK_ARG((((uintptr_t)h.h_actor | (uintptr_t)h.h_critic | (uintptr_t)h.dz_actor | (uintptr_t)h.dz_critic) & 15) == 0, what);
    if (!(vec && A <= 16 && quad_ok && rows <= (long long)kNumSMs * 1024)) {
      set_error("%s: the fused output-layer data gradient needs the four-threads-per-row kernel (A %% 4 == 0, A <= 16, "
                "rows <= %d)", what, kNumSMs * 1024);
      return 4;
    }
    head.w_a = h.w_actor; head.w_c = h.w_critic; head.h_a = (const __half*)h.h_actor; head.h_c = (const __half*)h.h_critic;
    head.dz_a = (__half*)h.dz_actor; head.dz_c = (__half*)h.dz_critic; head.K3 = h.K3; head.ldh = h.ldh; head.act = h.act;
  }
  if (vec && A <= 16 && quad_ok && rows <= quad_max_rows) {
    const int qblocks = (int)((rows + kQuadThreads / 4 - 1) / (kQuadThreads / 4));
    launch_chain_kernel(
        ppo_loss_quad_kernel, dim3(qblocks), dim3(kQuadThreads), 0, (cudaStream_t)stream, mu, std_param, std_type, value,
        actions, old_logp, advantages, returns, old_values, old_mu, old_sigma, adv_stats, dmu, dmu_stride, dvalue,
        dvalue_stride, dstd, db_mu, db_value, kl_out, loss_sums, (LossWorkspace*)workspace, (long long)rows,
        (long long)mb_orig, A, (float)clip_param, (float)value_loss_coef, (float)entropy_coef, use_clipped_value_loss,
        (__half*)dmu16, dmu16_stride, (__half*)dvalue16, dvalue16_stride, (float)grad_scale16, scale_state, head);
    return finish_launch(what);
  }
  if ((dmu16 != nullptr || dvalue16 != nullptr) && !vec) {
    set_error("%s: fp16 gradient outputs need A %% 4 == 0 and 16-byte aligned rows", what);
    return 4;
  }
  auto kern = vec ? ppo_loss_kernel<true> : ppo_loss_kernel<false>;
  launch_chain_kernel(
      kern, dim3(blocks), dim3(kLossThreads), 0, (cudaStream_t)stream, mu, std_param, std_type, value, actions, old_logp, advantages, returns, old_values, old_mu, old_sigma, adv_stats,
      dmu, dmu_stride, dvalue, dvalue_stride, dstd, db_mu, db_value, kl_out, loss_sums, (LossWorkspace*)workspace, (long long)rows,
      (long long)mb_orig, A, (float)clip_param,
      (float)value_loss_coef, (float)entropy_coef, use_clipped_value_loss, (__half*)dmu16, dmu16_stride, (__half*)dvalue16,
      dvalue16_stride, (float)grad_scale16, scale_state);
  return finish_launch(what);
}

int rslrl_sum_sumsq(const float* x, int64_t n, double* out, void* workspace, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_sum_sumsq";
  RSLRL_CHECK_ARG(x && out && workspace && n > 0, what);
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  sum_sumsq_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, out, (SumWorkspace*)workspace);
  return finish_launch(what);
}

}  // extern "C"
