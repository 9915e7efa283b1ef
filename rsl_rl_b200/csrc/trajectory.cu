// Trajectory bookkeeping for recurrent minibatches as a segmented scan over the done flags.
//
// Reference: split_and_pad_trajectories / unpad_trajectories (rsl_rl/utils/utils.py:78-141) and the
// hidden-state extraction of recurrent_mini_batch_generator (rsl_rl/storage/rollout_storage.py:
// 218-250).  The reference builds ~6000 tensor slices on the host per update (nonzero -> tolist ->
// split -> pad_sequence, 245 host syncs); here the (t, env) -> (trajectory, position) map is an
// integer scan computed once per update on the device and every consumer is a pure gather, so each
// output element is written exactly once with coalesced stores.  All results are bit-exact integer /
// byte work; float payloads are moved, never recomputed.
#include "common.cuh"

namespace rslrl {

constexpr int kTrajThreads = 256;

// 1) trajectories per env: 1 + number of dones in rows [0, T-1)
__global__ void traj_count_kernel(const uint8_t* __restrict__ dones, int T, int N, int32_t* __restrict__ count) {
  const int n = blockIdx.x * kTrajThreads + threadIdx.x;
  if (n >= N) return;
  int c = 1;
  for (int t = 0; t < T - 1; ++t) c += dones[(size_t)t * N + n] != 0;
  count[n] = c;
}

// 2) exclusive scan of count[N] -> offset[N+1]; one block walks the array in chunks with a carry
__global__ void __launch_bounds__(1024) traj_scan_kernel(const int32_t* __restrict__ count, int N,
                                                         int32_t* __restrict__ offset) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < N; base += 1024) {
    const int i = base + threadIdx.x;
    const int32_t v = i < N ? count[i] : 0;
    int32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int32_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int32_t carry = carry_s;
    const int32_t incl = x + (warp > 0 ? warp_tot[warp - 1] : 0) + carry;
    if (i < N) offset[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offset[N] = carry_s;
}

// 3) per-env walk: ids, positions and the per-trajectory tables
__global__ void traj_fill_kernel(const uint8_t* __restrict__ dones, int T, int N, const int32_t* __restrict__ offset,
                                 int32_t* __restrict__ traj_id, int32_t* __restrict__ traj_pos,
                                 int32_t* __restrict__ traj_len, int32_t* __restrict__ traj_env,
                                 int32_t* __restrict__ traj_start) {
  const int n = blockIdx.x * kTrajThreads + threadIdx.x;
  if (n >= N) return;
  int j = offset[n], start = 0;
  for (int t = 0; t < T; ++t) {
    const size_t o = (size_t)t * N + n;
    traj_id[o] = j;
    traj_pos[o] = t - start;
    if (dones[o] != 0 || t == T - 1) {
      traj_len[j] = t - start + 1;
      traj_env[j] = n;
      traj_start[j] = start;
      ++j;
      start = t + 1;
    }
  }
}

__global__ void traj_masks_kernel(const int32_t* __restrict__ traj_len, uint8_t* __restrict__ masks, int T,
                                  int first_traj, int n_traj_end, int capacity) {
  const long long i = (long long)blockIdx.x * kTrajThreads + threadIdx.x;
  if (i >= (long long)T * capacity) return;
  const int t = (int)(i / capacity), j = (int)(i - (long long)t * capacity);
  const int J = first_traj + j;
  masks[i] = (J < n_traj_end && traj_len[J] > t) ? 1 : 0;
}

// out[p, j, :] over D floats; one thread per float (float4 when D % 4 == 0)
template <int VEC>
__global__ void split_pad_kernel(const float* __restrict__ x, const int32_t* __restrict__ traj_len,
                                 const int32_t* __restrict__ traj_env, const int32_t* __restrict__ traj_start,
                                 float* __restrict__ padded, int T, int N, int D, int first_traj, int n_traj_end,
                                 int capacity) {
  const int dv = D / VEC;
  const long long total = (long long)T * capacity * dv;
  for (long long i = (long long)blockIdx.x * kTrajThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kTrajThreads) {
    const int c = (int)(i % dv);
    const long long pj = i / dv;
    const int j = (int)(pj % capacity), p = (int)(pj / capacity);
    const int J = first_traj + j;
    const bool live = J < n_traj_end && p < traj_len[J];
    if (VEC == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) v = __ldg((const float4*)(x + ((size_t)(traj_start[J] + p) * N + traj_env[J]) * D) + c);
      ((float4*)padded)[i] = v;
    } else {
      padded[i] = live ? __ldg(x + ((size_t)(traj_start[J] + p) * N + traj_env[J]) * D + c) : 0.f;
    }
  }
}

template <int VEC>
__global__ void unpad_kernel(const float* __restrict__ padded, const int32_t* __restrict__ traj_id,
                             const int32_t* __restrict__ traj_pos, float* __restrict__ out, int T, int N, int D,
                             int env0, int env1, int first_traj, int capacity) {
  const int dv = D / VEC, ne = env1 - env0;
  const long long total = (long long)T * ne * dv;
  for (long long i = (long long)blockIdx.x * kTrajThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kTrajThreads) {
    const int c = (int)(i % dv);
    const long long te = i / dv;
    const int e = (int)(te % ne), t = (int)(te / ne);
    const size_t o = (size_t)t * N + env0 + e;
    const size_t src = ((size_t)traj_pos[o] * capacity + (traj_id[o] - first_traj)) * D;
    if (VEC == 4) ((float4*)out)[i] = __ldg((const float4*)(padded + src) + c);
    else out[i] = __ldg(padded + src + c);
  }
}

__global__ void hidden_starts_kernel(const float* __restrict__ saved, const int32_t* __restrict__ traj_env,
                                     const int32_t* __restrict__ traj_start, float* __restrict__ out, int T, int L,
                                     int N, int H, int first_traj, int n_traj_end, int capacity) {
  const long long total = (long long)L * capacity * H;
  for (long long i = (long long)blockIdx.x * kTrajThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kTrajThreads) {
    const int h = (int)(i % H);
    const long long lj = i / H;
    const int j = (int)(lj % capacity), l = (int)(lj / capacity);
    const int J = first_traj + j;
    float v = 0.f;
    if (J < n_traj_end) v = __ldg(saved + (((size_t)traj_start[J] * L + l) * N + traj_env[J]) * H + h);
    out[i] = v;
  }
}

static int blocks_for(long long total) {
  long long b = (total + kTrajThreads - 1) / kTrajThreads;
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  return b < 1 ? 1 : (int)b;
}

}  // namespace rslrl

extern "C" {

int64_t rslrl_traj_workspace_bytes(int T, int N) {
  (void)T;
  return (int64_t)(N + 1) * sizeof(int32_t);
}

int rslrl_trajectory_index(const uint8_t* dones, int T, int N, int32_t* traj_id, int32_t* traj_pos,
                           int32_t* traj_len, int32_t* traj_env, int32_t* traj_start, int32_t* env_traj_offset,
                           void* workspace, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_trajectory_index";
  RSLRL_CHECK_ARG(dones && traj_id && traj_pos && traj_len && traj_env && traj_start && env_traj_offset && workspace,
                  what);
  RSLRL_CHECK_ARG(T > 0 && N > 0, what);
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* count = (int32_t*)workspace;
  const int blocks = (N + kTrajThreads - 1) / kTrajThreads;
  traj_count_kernel<<<blocks, kTrajThreads, 0, st>>>(dones, T, N, count);
  if (finish_launch(what)) return 1;
  traj_scan_kernel<<<1, 1024, 0, st>>>(count, N, env_traj_offset);
  if (finish_launch(what)) return 1;
  traj_fill_kernel<<<blocks, kTrajThreads, 0, st>>>(dones, T, N, env_traj_offset, traj_id, traj_pos, traj_len,
                                                    traj_env, traj_start);
  return finish_launch(what);
}

int rslrl_trajectory_masks(const int32_t* traj_len, uint8_t* masks, int T, int first_traj, int n_traj_end,
                           int capacity, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_trajectory_masks";
  RSLRL_CHECK_ARG(traj_len && masks && T > 0 && capacity >= 0, what);
  if (capacity == 0) return 0;
  const long long total = (long long)T * capacity;
  traj_masks_kernel<<<(int)((total + kTrajThreads - 1) / kTrajThreads), kTrajThreads, 0, (cudaStream_t)stream>>>(
      traj_len, masks, T, first_traj, n_traj_end, capacity);
  return finish_launch(what);
}

int rslrl_split_and_pad(const float* x, const int32_t* traj_len, const int32_t* traj_env,
                        const int32_t* traj_start, float* padded, int T, int N, int D, int first_traj,
                        int n_traj_end, int capacity, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_split_and_pad";
  RSLRL_CHECK_ARG(x && traj_len && traj_env && traj_start && padded && T > 0 && N > 0 && D > 0 && capacity >= 0, what);
  if (capacity == 0) return 0;
  const bool vec = D % 4 == 0 && (((uintptr_t)x | (uintptr_t)padded) & 15) == 0;
  const long long total = (long long)T * capacity * (vec ? D / 4 : D);
  if (vec)
    split_pad_kernel<4><<<blocks_for(total), kTrajThreads, 0, (cudaStream_t)stream>>>(
        x, traj_len, traj_env, traj_start, padded, T, N, D, first_traj, n_traj_end, capacity);
  else
    split_pad_kernel<1><<<blocks_for(total), kTrajThreads, 0, (cudaStream_t)stream>>>(
        x, traj_len, traj_env, traj_start, padded, T, N, D, first_traj, n_traj_end, capacity);
  return finish_launch(what);
}

int rslrl_unpad(const float* padded, const int32_t* traj_id, const int32_t* traj_pos, float* out, int T, int N,
                int D, int env0, int env1, int first_traj, int capacity, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_unpad";
  RSLRL_CHECK_ARG(padded && traj_id && traj_pos && out && T > 0 && N > 0 && D > 0, what);
  RSLRL_CHECK_ARG(env0 >= 0 && env1 > env0 && env1 <= N && capacity > 0, what);
  const bool vec = D % 4 == 0 && (((uintptr_t)out | (uintptr_t)padded) & 15) == 0;
  const long long total = (long long)T * (env1 - env0) * (vec ? D / 4 : D);
  if (vec)
    unpad_kernel<4><<<blocks_for(total), kTrajThreads, 0, (cudaStream_t)stream>>>(padded, traj_id, traj_pos, out, T, N,
                                                                                  D, env0, env1, first_traj, capacity);
  else
    unpad_kernel<1><<<blocks_for(total), kTrajThreads, 0, (cudaStream_t)stream>>>(padded, traj_id, traj_pos, out, T, N,
                                                                                  D, env0, env1, first_traj, capacity);
  return finish_launch(what);
}

int rslrl_hidden_at_starts(const float* saved, const int32_t* traj_env, const int32_t* traj_start, float* out,
                           int T, int L, int N, int H, int first_traj, int n_traj_end, int capacity,
                           rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_hidden_at_starts";
  RSLRL_CHECK_ARG(saved && traj_env && traj_start && out && T > 0 && L > 0 && N > 0 && H > 0 && capacity >= 0, what);
  if (capacity == 0) return 0;
  const long long total = (long long)L * capacity * H;
  hidden_starts_kernel<<<blocks_for(total), kTrajThreads, 0, (cudaStream_t)stream>>>(
      saved, traj_env, traj_start, out, T, L, N, H, first_traj, n_traj_end, capacity);
  return finish_launch(what);
}

}  // extern "C"
