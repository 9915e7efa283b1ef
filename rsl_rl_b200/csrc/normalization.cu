// Running-statistics observation normaliser: fused batch statistics + Chan update, and the normalise pass.
// Replaces rsl_rl/networks/normalization.py:44-66 (EmpiricalNormalization.forward / update: mean, var, 6 elementwise
// tensor ops and -- with `until` -- a device->host sync per env step).
//
//   rslrl_empirical_norm_update: ONE launch.  Phase 1: per-column sum and sum of squares of the [N,D] batch (lane <->
//   column, so a warp reads consecutive floats of a row; fp64 accumulation; per-block partials in a workspace).  Phase 2
//   (the last block to finish, ticket counter): partials summed in index order (deterministic), then the reference's
//   update in its fp32 operation order --
//       count += N;  rate = fp32(N) / fp32(count)
//       delta = mean_x - mean;  mean += rate * delta;  var += rate * (var_x - var + delta * (mean_x - mean_new))
//       std = sqrt(var)
//   with the `until` test done on the device (count is an int64 buffer in HBM: no host synchronisation).
//   HBM-bound: 4*N*D bytes read once.
//   rslrl_empirical_normalize: y = (x - mean) / (std + eps), fp32, the reference's operation order (bit-exact), written
//   with an arbitrary row pitch (e.g. straight into the zero-padded 16-byte-pitch input buffer of the fused forward).
#include "common.cuh"

namespace rslrl {

constexpr int kNormThreads = 256;
constexpr int kNormMaxChunks = 512;

struct NormUpdateWs {
  unsigned int ticket;
  unsigned int pad_;
  double partial[1];  // [chunks][D][2]
};

__global__ void __launch_bounds__(kNormThreads)
empirical_norm_update_kernel(const float* __restrict__ x, long long ldx, long long N, int D, int CW,
                             float* mean, float* var, float* stdv, long long* count, long long until,
                             NormUpdateWs* ws) {
  __shared__ double s_sum[kNormThreads], s_sq[kNormThreads];
  __shared__ bool s_last;
  const int col_in = threadIdx.x % CW, row_in = threadIdx.x / CW, rows_per_it = kNormThreads / CW;
  const int n_cg = (D + CW - 1) / CW;
  const int cg = blockIdx.x % n_cg, chunk = blockIdx.x / n_cg, chunks = gridDim.x / n_cg;
  const int col = cg * CW + col_in;
  const long long per = (N + chunks - 1) / chunks;
  const long long r0 = (long long)chunk * per, r1 = r0 + per < N ? r0 + per : N;
  double s = 0.0, q = 0.0;
  if (col < D) {
    for (long long r = r0 + row_in; r < r1; r += rows_per_it) {
      const double v = (double)__ldg(x + r * ldx + col);
      s += v;
      q += v * v;
    }
  }
  s_sum[threadIdx.x] = s;
  s_sq[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < CW && col < D) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < kNormThreads; i += CW) {   // fixed order
      a += s_sum[i];
      b += s_sq[i];
    }
    ws->partial[((size_t)chunk * D + col) * 2] = a;
    ws->partial[((size_t)chunk * D + col) * 2 + 1] = b;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const long long cnt_old = *count;
  const bool frozen = until >= 0 && cnt_old >= until;   // normalization.py:53-54
  if (!frozen) {
    const long long cnt_new = cnt_old + N;
    const float rate = (float)N / (float)cnt_new;        // int / LongTensor -> fp32 true division
    for (int c = threadIdx.x; c < D; c += kNormThreads) {
      double a = 0.0, b = 0.0;
      for (int k = 0; k < chunks; ++k) {
        a += *((volatile double*)&ws->partial[((size_t)k * D + c) * 2]);
        b += *((volatile double*)&ws->partial[((size_t)k * D + c) * 2 + 1]);
      }
      const double m = a / (double)N;
      double v = b / (double)N - m * m;                  // biased variance (unbiased=False)
      v = v > 0.0 ? v : 0.0;
      const float mean_x = (float)m, var_x = (float)v;
      const float mean_old = mean[c], var_old = var[c];
      const float delta = __fsub_rn(mean_x, mean_old);
      const float mean_new = __fadd_rn(mean_old, __fmul_rn(rate, delta));
      const float inner = __fadd_rn(__fsub_rn(var_x, var_old), __fmul_rn(delta, __fsub_rn(mean_x, mean_new)));
      const float var_new = __fadd_rn(var_old, __fmul_rn(rate, inner));
      mean[c] = mean_new;
      var[c] = var_new;
      stdv[c] = sqrtf(var_new);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!frozen) *count = cnt_old + N;
    __threadfence();
    ws->ticket = 0u;
  }
}

__global__ void __launch_bounds__(kNormThreads)
empirical_normalize_kernel(const float* __restrict__ x, long long ldx, long long N, int D,
                           const float* __restrict__ mean, const float* __restrict__ stdv, float eps,
                           float* __restrict__ y, long long ldy) {
  const long long total = N * D;
  for (long long i = (long long)blockIdx.x * kNormThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kNormThreads) {
    const long long r = i / D;
    const int c = (int)(i - r * D);
    y[r * ldy + c] = __fdiv_rn(__fsub_rn(__ldg(x + r * ldx + c), __ldg(mean + c)), __fadd_rn(__ldg(stdv + c), eps));
  }
}

static void norm_grid(long long N, int D, int& CW, int& n_cg, int& chunks) {
  CW = 32;
  if (D < 32) {
    CW = 1;
    while (CW < D) CW <<= 1;
  }
  n_cg = (D + CW - 1) / CW;
  const int rows_per_it = kNormThreads / CW;
  long long want = (2LL * kNumSMs + n_cg - 1) / n_cg;            // ~2 blocks per SM in total
  const long long max_chunks = (N + 4LL * rows_per_it - 1) / (4LL * rows_per_it);  // >= 4 rows per thread
  if (want > max_chunks) want = max_chunks;
  if (want > kNormMaxChunks) want = kNormMaxChunks;
  if (want < 1) want = 1;
  chunks = (int)want;
}

}  // namespace rslrl

extern "C" {

int64_t rslrl_empirical_norm_workspace_bytes(int D) {
  return (int64_t)sizeof(rslrl::NormUpdateWs) + (int64_t)rslrl::kNormMaxChunks * (D > 0 ? D : 1) * 2 * sizeof(double);
}

int rslrl_empirical_norm_update(const float* x, int64_t ldx, int64_t N, int D, float* mean, float* var, float* std,
                                int64_t* count, int64_t until, void* workspace, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_empirical_norm_update";
  RSLRL_CHECK_ARG(x && mean && var && std && count && workspace && N > 0 && D > 0 && ldx >= 1, what);
  int CW, n_cg, chunks;
  norm_grid(N, D, CW, n_cg, chunks);
  empirical_norm_update_kernel<<<n_cg * chunks, kNormThreads, 0, (cudaStream_t)stream>>>(
      x, ldx, N, D, CW, mean, var, std, (long long*)count, until, (NormUpdateWs*)workspace);
  return finish_launch(what);
}

int rslrl_empirical_normalize(const float* x, int64_t ldx, int64_t N, int D, const float* mean, const float* std,
                              double eps, float* y, int64_t ldy, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_empirical_normalize";
  RSLRL_CHECK_ARG(x && mean && std && y && N >= 0 && D > 0 && ldx >= 1 && ldy >= D, what);
  if (N == 0) return 0;
  long long blocks = (N * D + kNormThreads * 4 - 1) / (kNormThreads * 4);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (blocks < 1) blocks = 1;
  empirical_normalize_kernel<<<(int)blocks, kNormThreads, 0, (cudaStream_t)stream>>>(x, ldx, N, D, mean, std, (float)eps,
                                                                                     y, ldy);
  return finish_launch(what);
}

}  // extern "C"
