// Rollout buffer kernels: single-launch transition store and the permuted minibatch gather.
//
// The rollout buffer is a structure of arrays in HBM: one dense [T,N,d] fp32 array per field
// (observation groups, actions, mu, sigma, values, log-probs, rewards, returns, advantages) and a
// [T,N] u8 array for dones -- byte-identical to the reference's tensors
// (rsl_rl/storage/rollout_storage.py:47-68).  Row `step` of every array is one contiguous span, so
// add_transitions is a set of straight copies: here one launch (grid.y = segment) with 16-byte
// vector accesses, fused with the reward bootstrap and the dones dtype conversion.
#include "common.cuh"

namespace rslrl {

constexpr int kCopyThreads = 256;

struct CopySeg {
  const float* src;
  float* dst;
  long long count;  // floats
};

struct StoreParams {
  int n_copy;
  CopySeg copy[RSLRL_MAX_SEGMENTS + 6];
  // sigma broadcast ([A] -> [N,A]); sigma == nullptr when sigma is copied as a plain segment
  const float* sigma;
  float* sigma_row;
  // rewards / dones
  const float* rewards;
  const float* values;
  const float* intrinsic;
  const void* time_outs;
  const void* dones;
  float* rewards_row;
  uint8_t* dones_row;
  int time_outs_dtype, dones_dtype;
  float gamma;
  int N, A;
};

__device__ __forceinline__ float load_flag(const void* p, int dtype, int i) {
  switch (dtype) {
    case RSLRL_U8: return (float)((const uint8_t*)p)[i];
    case RSLRL_I32: return (float)((const int32_t*)p)[i];
    case RSLRL_I64: return (float)((const long long*)p)[i];
    default: return ((const float*)p)[i];
  }
}

__global__ void __launch_bounds__(kCopyThreads) store_transition_kernel(const StoreParams p) {
  const int seg = blockIdx.y;
  const long long tid = (long long)blockIdx.x * kCopyThreads + threadIdx.x;
  const long long stride = (long long)gridDim.x * kCopyThreads;
  if (seg < p.n_copy) {
    const CopySeg s = p.copy[seg];
    const bool vec = (((uintptr_t)s.src | (uintptr_t)s.dst) & 15) == 0 && (s.count & 3) == 0;
    if (vec) {
      const float4* src = (const float4*)s.src;
      float4* dst = (float4*)s.dst;
      for (long long i = tid; i < (s.count >> 2); i += stride) dst[i] = __ldg(src + i);
    } else {
      for (long long i = tid; i < s.count; i += stride) s.dst[i] = __ldg(s.src + i);
    }
  } else if (seg == p.n_copy) {
    // rewards (+ intrinsic) + gamma * (values * time_outs)   [ppo.py:150-164], dones -> u8 [rollout_storage.py:86]
    for (long long i = tid; i < p.N; i += stride) {
      float r = __ldg(p.rewards + i);
      if (p.intrinsic) r = __fadd_rn(r, __ldg(p.intrinsic + i));
      if (p.time_outs) {
        const float to = load_flag(p.time_outs, p.time_outs_dtype, (int)i);
        r = __fadd_rn(r, __fmul_rn(p.gamma, __fmul_rn(__ldg(p.values + i), to)));
      }
      p.rewards_row[i] = r;
      const float d = load_flag(p.dones, p.dones_dtype, (int)i);
      p.dones_row[i] = (uint8_t)(int)d;  // Tensor.copy_ float/int -> uint8 truncation
    }
  } else if (p.sigma) {
    const long long total = (long long)p.N * p.A;
    for (long long i = tid; i < total; i += stride) p.sigma_row[i] = __ldg(p.sigma + (i % p.A));
  }
}

struct GatherField {
  const float* src;
  float* dst;
  int row_floats, dst_stride, dst_col;
};

struct GatherParams {
  GatherField f[RSLRL_MAX_SEGMENTS];
  const long long* idx;
  long long n_rows;
};

// Permuted row gather, all fields in one launch.  blockIdx.y < n_wide: one WARP per row of a wide field (>= 32 floats
// per row): the lanes stride the row, so every request is a coalesced 128/512-byte span of one source row, and each
// warp keeps kGatherRows rows in flight (their indices and first loads are issued before any store).  The last
// blockIdx.y: one THREAD per row for all narrow fields together (the index is read once per row).  No integer
// division anywhere: rows are enumerated by warp / thread id.
constexpr int kGatherRows = 4;

__device__ __forceinline__ void gather_wide(const GatherField f, const long long* __restrict__ idx, long long n_rows,
                                            long long warp_id, long long n_warps, int lane) {
  const bool vec = (f.row_floats & 3) == 0 && (f.dst_stride & 3) == 0 && (f.dst_col & 3) == 0 &&
                   (((uintptr_t)f.src | (uintptr_t)f.dst) & 15) == 0;
  if (vec) {
    // 16-byte chunks; a warp takes 8 rows at a time and walks their chunks as one flat range, so short rows (48 floats
    // = 12 chunks) still keep all 32 lanes busy; 4 loads in flight per lane
    constexpr int R = 8, U = 4;
    const int chunks = f.row_floats >> 2;
    for (long long j0 = warp_id * R; j0 < n_rows; j0 += n_warps * R) {
      const int nr = (int)((n_rows - j0 < R) ? n_rows - j0 : R);
      const long long s_lane = (lane < nr) ? __ldg(idx + j0 + lane) : 0;
      const int units = nr * chunks;
      for (int u0 = 0; u0 < units; u0 += 32 * U) {
        float4 v[U];
        int rr[U], cc[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int u = u0 + k * 32 + lane;
          rr[k] = u / chunks;
          cc[k] = u - rr[k] * chunks;
          const long long srow = __shfl_sync(0xffffffffu, s_lane, rr[k] & 7);
          if (u < units) v[k] = __ldg((const float4*)(f.src + (size_t)srow * f.row_floats) + cc[k]);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int u = u0 + k * 32 + lane;
          if (u < units) *((float4*)(f.dst + (size_t)(j0 + rr[k]) * f.dst_stride + f.dst_col) + cc[k]) = v[k];
        }
      }
    }
    return;
  }
  for (long long j0 = warp_id * kGatherRows; j0 < n_rows; j0 += n_warps * kGatherRows) {
    long long s[kGatherRows];
#pragma unroll
    for (int r = 0; r < kGatherRows; ++r) s[r] = (j0 + r < n_rows) ? __ldg(idx + j0 + r) : 0;
    for (int c0 = 0; c0 < f.row_floats; c0 += 32) {
      const int c = c0 + lane;
      float v[kGatherRows];
#pragma unroll
      for (int r = 0; r < kGatherRows; ++r)
        if (c < f.row_floats && j0 + r < n_rows) v[r] = __ldg(f.src + (size_t)s[r] * f.row_floats + c);
#pragma unroll
      for (int r = 0; r < kGatherRows; ++r)
        if (c < f.row_floats && j0 + r < n_rows) f.dst[(size_t)(j0 + r) * f.dst_stride + f.dst_col + c] = v[r];
    }
  }
}

__global__ void __launch_bounds__(kCopyThreads) permute_gather_kernel(const GatherParams p, const int n_wide,
                                                                      const int n_fields) {
  const long long tid = (long long)blockIdx.x * kCopyThreads + threadIdx.x;
  const long long stride = (long long)gridDim.x * kCopyThreads;
  if ((int)blockIdx.y < n_wide) {
    gather_wide(p.f[blockIdx.y], p.idx, p.n_rows, tid >> 5, stride >> 5, threadIdx.x & 31);
    return;
  }
  for (long long j = tid; j < p.n_rows; j += stride) {
    const long long s = __ldg(p.idx + j);
    for (int i = n_wide; i < n_fields; ++i) {
      const GatherField f = p.f[i];
      const float* src = f.src + (size_t)s * f.row_floats;
      float* dst = f.dst + (size_t)j * f.dst_stride + f.dst_col;
      if ((f.row_floats & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        for (int c = 0; c < f.row_floats; c += 4) *(float4*)(dst + c) = __ldg((const float4*)(src + c));
      } else {
        for (int c = 0; c < f.row_floats; ++c) dst[c] = __ldg(src + c);
      }
    }
  }
}

static int grid_for(long long units) {
  long long b = (units + kCopyThreads - 1) / kCopyThreads;
  const long long cap = (long long)kNumSMs * 8;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

}  // namespace rslrl

extern "C" {

int rslrl_store_transition(int n_groups, const float* const* obs_src_host, float* const* obs_dst_host,
                           const int* obs_dim_host, const float* actions, const float* mu, const float* sigma,
                           int sigma_stride, const float* values, const float* log_prob, const float* rewards,
                           const void* dones, int dones_dtype, const void* time_outs, int time_outs_dtype,
                           const float* intrinsic, double gamma, float* actions_row, float* mu_row,
                           float* sigma_row, float* values_row, float* log_prob_row, float* rewards_row,
                           uint8_t* dones_row, int N, int A, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_store_transition";
  RSLRL_CHECK_ARG(n_groups >= 0 && n_groups <= RSLRL_MAX_SEGMENTS, what);
  RSLRL_CHECK_ARG(N > 0 && A > 0 && rewards && dones && rewards_row && dones_row, what);
  RSLRL_CHECK_ARG(sigma_stride == 0 || sigma_stride == A, what);
  StoreParams p{};
  long long max_units = N;
  auto add = [&](const float* s, float* d, long long count) {
    if (!s || !d) return;
    p.copy[p.n_copy++] = CopySeg{s, d, count};
    if (count > max_units) max_units = count;
  };
  for (int g = 0; g < n_groups; ++g) add(obs_src_host[g], obs_dst_host[g], (long long)N * obs_dim_host[g]);
  add(actions, actions_row, (long long)N * A);
  add(mu, mu_row, (long long)N * A);
  add(values, values_row, N);
  add(log_prob, log_prob_row, N);
  if (sigma && sigma_row) {
    if (sigma_stride == A) {
      add(sigma, sigma_row, (long long)N * A);
    } else {
      p.sigma = sigma;
      p.sigma_row = sigma_row;
      if ((long long)N * A > max_units) max_units = (long long)N * A;
    }
  }
  p.rewards = rewards; p.values = values; p.intrinsic = intrinsic; p.time_outs = time_outs; p.dones = dones;
  p.rewards_row = rewards_row; p.dones_row = dones_row;
  p.time_outs_dtype = time_outs_dtype; p.dones_dtype = dones_dtype;
  p.gamma = (float)gamma; p.N = N; p.A = A;
  RSLRL_CHECK_ARG(!time_outs || values, what);
  dim3 grid(grid_for(max_units / 4 + 1), p.n_copy + 1 + (p.sigma ? 1 : 0));
  prefer_max_smem(store_transition_kernel);
  store_transition_kernel<<<grid, kCopyThreads, 0, (cudaStream_t)stream>>>(p);
  return finish_launch(what);
}

int rslrl_store_step(const rslrl_store_plan_t* plan, const float* const* obs_src_host, const float* rewards,
                     const void* dones, int dones_dtype, const void* time_outs, int time_outs_dtype,
                     const float* intrinsic, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_store_step";
  RSLRL_CHECK_ARG(plan && obs_src_host && rewards && dones, what);
  RSLRL_CHECK_ARG(plan->n_groups >= 0 && plan->n_groups <= RSLRL_MAX_SEGMENTS && plan->N > 0, what);
  RSLRL_CHECK_ARG(!time_outs || plan->values_row, what);
  StoreParams p{};
  long long max_units = plan->N;
  for (int g = 0; g < plan->n_groups; ++g) {
    RSLRL_CHECK_ARG(obs_src_host[g] && plan->obs_dst[g], what);
    const long long count = (long long)plan->N * plan->obs_dim[g];
    p.copy[p.n_copy++] = CopySeg{obs_src_host[g], plan->obs_dst[g], count};
    if (count > max_units) max_units = count;
  }
  p.rewards = rewards; p.values = plan->values_row; p.intrinsic = intrinsic; p.time_outs = time_outs; p.dones = dones;
  p.rewards_row = plan->rewards_row; p.dones_row = plan->dones_row;
  p.time_outs_dtype = time_outs_dtype; p.dones_dtype = dones_dtype;
  p.gamma = (float)plan->gamma; p.N = plan->N; p.A = plan->A;
  dim3 grid(grid_for(max_units / 4 + 1), p.n_copy + 1);
  prefer_max_smem(store_transition_kernel);
  store_transition_kernel<<<grid, kCopyThreads, 0, (cudaStream_t)stream>>>(p);
  return finish_launch(what);
}

int rslrl_permute_gather(int n_fields, const float* const* src_host, float* const* dst_host,
                         const int* row_floats_host, const int* dst_stride_host, const int* dst_col_host,
                         const int64_t* idx, int64_t n_rows, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_permute_gather";
  RSLRL_CHECK_ARG(n_fields > 0 && n_fields <= RSLRL_MAX_SEGMENTS && idx && n_rows >= 0, what);
  if (n_rows == 0) return 0;
  GatherParams p{};
  for (int i = 0; i < n_fields; ++i) {
    RSLRL_CHECK_ARG(src_host[i] && dst_host[i] && row_floats_host[i] > 0, what);
    RSLRL_CHECK_ARG(dst_stride_host[i] >= dst_col_host[i] + row_floats_host[i], what);
  }
  // wide fields first (one warp per row), narrow ones after (one thread per row, all of them together)
  int n_wide = 0, k = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < n_fields; ++i) {
      const bool wide = row_floats_host[i] >= 32;
      if (wide != (pass == 0)) continue;
      p.f[k++] = GatherField{src_host[i], dst_host[i], row_floats_host[i], dst_stride_host[i], dst_col_host[i]};
      if (wide) ++n_wide;
    }
  p.idx = (const long long*)idx;
  p.n_rows = n_rows;
  const bool has_narrow = n_wide < n_fields;
  // wide: n_rows / kGatherRows warps; narrow: n_rows threads
  const long long units = n_wide ? (n_rows + kGatherRows - 1) / kGatherRows * 32 : n_rows;
  dim3 grid(grid_for(units > n_rows ? units : n_rows), n_wide + (has_narrow ? 1 : 0));
  permute_gather_kernel<<<grid, kCopyThreads, 0, (cudaStream_t)stream>>>(p, n_wide, n_fields);
  return finish_launch(what);
}

}  // extern "C"
