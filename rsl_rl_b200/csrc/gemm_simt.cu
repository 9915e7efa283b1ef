// FP32 CUDA-core GEMMs for nn.Linear forward / data-gradient / weight-gradient with fused epilogues.
//
// This is the RSLRL_PREC_FP32 arithmetic (the reference's own default on a GPU is cuBLAS SGEMM with
// TF32 off, so this is the mode that meets the 1e-5 parity class; tcgen05 has no fp32 MMA).  The
// tensor-core path lives in gemm_tc.cu.
//
// One templated kernel: C[M,N] (+)= sum_k A(m,k) B(n,k), 128x128x16 tiles, 256 threads, 8x8 register
// tile per thread, register-staged double buffering through shared memory.  Operands may be stored
// with the reduction index contiguous ("KC": X[M,K], W[N,K]) or strided ("MC": reduction index is the
// slow one), which covers all three products without materialising a transpose:
//   forward   Y  = act(X W^T + b)        A = X (KC),   B = W (KC)          epilogue: bias + activation
//   dgrad     dX = (dY W) * act'(H)      A = dY (KC),  B = W (MC)          epilogue: activation gradient
//   wgrad     dW += dY^T X, db += 1^T dY A = dY (MC),  B = X (MC), split-K epilogue: vector atomics
// Several problems (actor and critic) share one launch (blockIdx.y).
#include <cuda_fp16.h>

#include "common.cuh"

namespace rslrl {

constexpr int BM = 128, BN = 128, BK = 16, GEMM_THREADS = 256, PADW = 4;

enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_WGRAD = 2 };

struct GemmProblem {
  const float* A;
  const float* B;
  float* C;
  const float* bias;  // EPI_FWD
  const float* H;     // EPI_DGRAD
  float* colsum;      // EPI_WGRAD: db (row sums of A);  EPI_DGRAD: db_prev (column sums of C)
  int M, N, K, lda, ldb, ldc, ldh, act, splits, k_per_split;
};

struct GemmBatch {
  GemmProblem p[RSLRL_MAX_PROBLEMS];
};

// Loads a [ROWS x BK] operand tile into registers (2 float4 per thread) ...
template <bool KC>
__device__ __forceinline__ void load_tile(const float* __restrict__ base, int ld, int row0, int rows, int k0, int k1,
                                          bool vec_ok, float4 (&r)[2]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int f = t + i * GEMM_THREADS;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KC) {
      const int row = row0 + (f >> 2), k = k0 + ((f & 3) << 2);
      if (row < rows) {
        const float* src = base + (size_t)row * ld + k;
        if (vec_ok && k + 3 < k1) {
          v = __ldg((const float4*)src);
        } else {
          if (k < k1) v.x = __ldg(src);
          if (k + 1 < k1) v.y = __ldg(src + 1);
          if (k + 2 < k1) v.z = __ldg(src + 2);
          if (k + 3 < k1) v.w = __ldg(src + 3);
        }
      }
    } else {
      const int k = k0 + (f >> 5), row = row0 + ((f & 31) << 2);
      if (k < k1) {
        const float* src = base + (size_t)k * ld + row;
        if (vec_ok && row + 3 < rows) {
          v = __ldg((const float4*)src);
        } else {
          if (row < rows) v.x = __ldg(src);
          if (row + 1 < rows) v.y = __ldg(src + 1);
          if (row + 2 < rows) v.z = __ldg(src + 2);
          if (row + 3 < rows) v.w = __ldg(src + 3);
        }
      }
    }
    r[i] = v;
  }
}

// ... and stores it k-major into shared memory: S[k][row].
template <bool KC>
__device__ __forceinline__ void store_tile(float (*S)[BM + PADW], const float4 (&r)[2]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int f = t + i * GEMM_THREADS;
    if (KC) {
      const int row = f >> 2, k = (f & 3) << 2;
      S[k][row] = r[i].x;
      S[k + 1][row] = r[i].y;
      S[k + 2][row] = r[i].z;
      S[k + 3][row] = r[i].w;
    } else {
      const int k = f >> 5, row = (f & 31) << 2;
      *(float4*)&S[k][row] = r[i];
    }
  }
}

template <bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2) sgemm_kernel(const GemmBatch batch) {
  const GemmProblem& P = batch.p[blockIdx.y];
  const int tiles_m = (P.M + BM - 1) / BM, tiles_n = (P.N + BN - 1) / BN;
  const int tiles = tiles_m * tiles_n;
  if ((int)blockIdx.x >= tiles * P.splits) return;
  const int split = blockIdx.x / tiles, rem = blockIdx.x - split * tiles;
  const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
  const int m0 = tm * BM, n0 = tn * BN;
  const int k_begin = split * P.k_per_split;
  const int k_end = min(P.K, k_begin + P.k_per_split);

  __shared__ __align__(16) float As[2][BK][BM + PADW];
  __shared__ __align__(16) float Bs[2][BK][BN + PADW];

  const bool a_vec = (P.lda & 3) == 0 && ((uintptr_t)P.A & 15) == 0 && (A_KC ? (k_begin & 3) == 0 : true);
  const bool b_vec = (P.ldb & 3) == 0 && ((uintptr_t)P.B & 15) == 0 && (B_KC ? (k_begin & 3) == 0 : true);

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float rowsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool do_rowsum = EPI == EPI_WGRAD && P.colsum != nullptr && tn == 0 && tx == 0;

  float4 ra[2], rb[2];
  if (k_begin < k_end) {
    load_tile<A_KC>(P.A, P.lda, m0, P.M, k_begin, k_end, a_vec, ra);
    load_tile<B_KC>(P.B, P.ldb, n0, P.N, k_begin, k_end, b_vec, rb);
    store_tile<A_KC>(As[0], ra);
    store_tile<B_KC>(Bs[0], rb);
  }
  __syncthreads();
  int buf = 0;
  for (int kk = k_begin; kk < k_end; kk += BK) {
    const bool has_next = kk + BK < k_end;
    if (has_next) {
      load_tile<A_KC>(P.A, P.lda, m0, P.M, kk + BK, k_end, a_vec, ra);
      load_tile<B_KC>(P.B, P.ldb, n0, P.N, kk + BK, k_end, b_vec, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *(const float4*)&As[buf][k][ty * 4];
      const float4 a1 = *(const float4*)&As[buf][k][64 + ty * 4];
      const float4 b0 = *(const float4*)&Bs[buf][k][tx * 4];
      const float4 b1 = *(const float4*)&Bs[buf][k][64 + tx * 4];
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (EPI == EPI_WGRAD) {
        if (do_rowsum) {
#pragma unroll
          for (int i = 0; i < 8; ++i) rowsum[i] += a[i];
        }
      }
    }
    if (has_next) {
      store_tile<A_KC>(As[buf ^ 1], ra);
      store_tile<B_KC>(Bs[buf ^ 1], rb);
    }
    __syncthreads();
    buf ^= 1;
  }

  // ---- epilogue
  float dcol[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // EPI_DGRAD: per-thread column sums of its 8 rows
  const bool c_vec = (P.ldc & 3) == 0 && ((uintptr_t)P.C & 15) == 0;
  const bool h_vec = EPI == EPI_DGRAD && (P.ldh & 3) == 0 && ((uintptr_t)P.H & 15) == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= P.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      if (n >= P.N) continue;
      float o[4] = {acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      const int nv = min(4, P.N - n);
      float* dst = P.C + (size_t)m * P.ldc + n;
      if (EPI == EPI_FWD) {
        for (int q = 0; q < nv; ++q) {
          const float z = o[q] + (P.bias ? __ldg(P.bias + n + q) : 0.f);
          o[q] = act_forward(P.act, z);
        }
      } else if (EPI == EPI_DGRAD) {
        const float* hp = P.H + (size_t)m * P.ldh + n;
        if (P.act != RSLRL_ACT_IDENTITY) {
          if (h_vec && nv == 4) {
            const float4 h4 = __ldg((const float4*)hp);
            o[0] *= act_grad_from_output(P.act, h4.x);
            o[1] *= act_grad_from_output(P.act, h4.y);
            o[2] *= act_grad_from_output(P.act, h4.z);
            o[3] *= act_grad_from_output(P.act, h4.w);
          } else {
            for (int q = 0; q < nv; ++q) o[q] *= act_grad_from_output(P.act, __ldg(hp + q));
          }
        }
        for (int q = 0; q < nv; ++q) dcol[jh * 4 + q] += o[q];
      }
      if (EPI == EPI_WGRAD) {
        if (c_vec && nv == 4) {
          atomicAdd((float4*)dst, make_float4(o[0], o[1], o[2], o[3]));
        } else {
          for (int q = 0; q < nv; ++q) atomicAdd(dst + q, o[q]);
        }
      } else {
        if (c_vec && nv == 4) {
          *(float4*)dst = make_float4(o[0], o[1], o[2], o[3]);
        } else {
          for (int q = 0; q < nv; ++q) dst[q] = o[q];
        }
      }
    }
    if (EPI == EPI_WGRAD) {
      if (do_rowsum) atomicAdd(P.colsum + m, rowsum[i]);
    }
  }
  if (EPI == EPI_DGRAD) {
    if (P.colsum != nullptr) {
      // reduce the 16 thread-rows (ty) that share a column inside the warp pair first: lanes l and l^16
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = dcol[j] + __shfl_xor_sync(0xffffffffu, dcol[j], 16);
        const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        if ((threadIdx.x & 16) == 0 && n < P.N) atomicAdd(P.colsum + n, v);
      }
    }
  }
}

static int pick_splits(int total_tiles_all, int K) {
  // aim at >= 2 CTAs per SM in flight; keep at least 8 k-tiles per split
  int want = (2 * kNumSMs + total_tiles_all - 1) / total_tiles_all;
  int max_s = K / (BK * 8);
  if (max_s < 1) max_s = 1;
  if (want > max_s) want = max_s;
  return want < 1 ? 1 : want;
}

template <bool A_KC, bool B_KC, int EPI>
static int launch(const GemmBatch& batch, int n, cudaStream_t st, const char* what) {
  int max_ctas = 0;
  for (int i = 0; i < n; ++i) {
    const GemmProblem& P = batch.p[i];
    const int t = ((P.M + BM - 1) / BM) * ((P.N + BN - 1) / BN) * P.splits;
    if (t > max_ctas) max_ctas = t;
  }
  if (max_ctas == 0) return 0;
  sgemm_kernel<A_KC, B_KC, EPI><<<dim3(max_ctas, n), GEMM_THREADS, 0, st>>>(batch);
  return finish_launch(what);
}

int simt_linear_forward(int n, const rslrl_linear_fwd_t* pr, cudaStream_t st) {
  GemmBatch b{};
  for (int i = 0; i < n; ++i) {
    GemmProblem& P = b.p[i];
    P.A = pr[i].x; P.B = pr[i].w; P.C = pr[i].y; P.bias = pr[i].b;
    P.M = pr[i].M; P.N = pr[i].N; P.K = pr[i].K;
    P.lda = pr[i].ldx; P.ldb = pr[i].K; P.ldc = pr[i].ldy; P.act = pr[i].act;
    P.splits = 1; P.k_per_split = pr[i].K;
  }
  return launch<true, true, EPI_FWD>(b, n, st, "rslrl_linear_forward");
}

int simt_linear_backward_data(int n, const rslrl_linear_bwd_data_t* pr, cudaStream_t st) {
  GemmBatch b{};
  for (int i = 0; i < n; ++i) {
    GemmProblem& P = b.p[i];
    // dX[M,K] = sum_n dY[m,n] W[n,k]  ->  C[M, K] with reduction length N
    P.A = (const float*)pr[i].dy; P.B = (const float*)pr[i].w; P.C = (float*)pr[i].dx; P.H = (const float*)pr[i].h;
    P.colsum = pr[i].db_prev;
    P.M = pr[i].M; P.N = pr[i].K; P.K = pr[i].N;
    P.lda = pr[i].ldy; P.ldb = pr[i].K; P.ldc = pr[i].ldx; P.ldh = pr[i].ldh; P.act = pr[i].act;
    P.splits = 1; P.k_per_split = pr[i].N;
  }
  return launch<true, false, EPI_DGRAD>(b, n, st, "rslrl_linear_backward_data");
}

int simt_linear_backward_weight(int n, const rslrl_linear_bwd_weight_t* pr, cudaStream_t st) {
  GemmBatch b{};
  int tiles_all = 0;
  for (int i = 0; i < n; ++i) tiles_all += ((pr[i].N + BM - 1) / BM) * ((pr[i].K + BN - 1) / BN);
  for (int i = 0; i < n; ++i) {
    GemmProblem& P = b.p[i];
    // dW[N,K] += sum_m dY[m,n] X[m,k]  ->  C[N, K] with reduction length M
    P.A = (const float*)pr[i].dy; P.B = (const float*)pr[i].x; P.C = pr[i].dw; P.colsum = pr[i].db;
    P.M = pr[i].N; P.N = pr[i].K; P.K = pr[i].M;
    P.lda = pr[i].ldy; P.ldb = pr[i].ldx; P.ldc = pr[i].K;
    P.splits = pick_splits(tiles_all, pr[i].M);
    int kps = (pr[i].M + P.splits - 1) / P.splits;
    kps = ((kps + BK - 1) / BK) * BK;
    P.k_per_split = kps;
    P.splits = (pr[i].M + kps - 1) / kps;
  }
  return launch<false, false, EPI_WGRAD>(b, n, st, "rslrl_linear_backward_weight");
}

}  // namespace rslrl

// ======================================================================================================
// Narrow output layers (the 12-wide action head and the 1-wide value head).
// A 128-wide MMA tile is >90 % padding for them and a launch costs more than their arithmetic, so they get
// fp32 CUDA-core kernel shaped after its memory traffic for the BACKWARD pass (the forward stays on the GEMM path):
//   head_backward  dx = (dy W) * act'(h), dW += dy^T h, db += 1^T dy     warp per row, lane <-> 4 input columns
// It fuses what would be one dgrad and one wgrad launch and reads h / writes dx exactly once (2*M*K*4 bytes).
// Used by both arithmetic modes (it is exact fp32).
// ======================================================================================================
namespace rslrl {

struct HeadBwdProblem {
  const float* dy; const float* w; const void* h; void* dx; float* dw; float* db;
  int M, N, K, ldy, ldh, ldx, act;
  float dx_scale;
};
struct HeadBwdBatch { HeadBwdProblem p[RSLRL_MAX_PROBLEMS]; };

constexpr int HEADB_WARPS = 8;
constexpr int HEADB_MAX_N = 12;

// K <= 128 per pass (lane <-> 4 consecutive input columns); wider inputs loop over 128-column panels.
constexpr int HEADB_ROWS = 4;  // rows per warp iteration: all their loads are issued before any is consumed

// 4 consecutive elements of a row of h / dx, fp32 or fp16 storage
template <bool F16>
__device__ __forceinline__ float4 headb_load4(const void* base, size_t elem) {
  if constexpr (F16) {
    const uint2 raw = __ldg((const uint2*)((const __half*)base + elem));
    const float2 a = __half22float2(*(const __half2*)&raw.x), b = __half22float2(*(const __half2*)&raw.y);
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    return __ldg((const float4*)((const float*)base + elem));
  }
}
// raw (unconverted) form: what the software pipeline keeps in registers while the load is in flight
template <bool F16> struct HeadbRaw { using type = float4; };
template <> struct HeadbRaw<true> { using type = uint2; };
template <bool F16>
__device__ __forceinline__ typename HeadbRaw<F16>::type headb_load_raw(const void* base, size_t elem) {
  if constexpr (F16) return __ldg((const uint2*)((const __half*)base + elem));
  else return __ldg((const float4*)((const float*)base + elem));
}
__device__ __forceinline__ float4 headb_cvt(float4 r) { return r; }
__device__ __forceinline__ float4 headb_cvt(uint2 r) {
  const float2 a = __half22float2(*(const __half2*)&r.x), b = __half22float2(*(const __half2*)&r.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <bool F16>
__device__ __forceinline__ void headb_store4(void* base, size_t elem, float4 o) {
  if constexpr (F16) {
    uint2 raw;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(raw.x) : "f"(o.y), "f"(o.x));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(raw.y) : "f"(o.w), "f"(o.z));
    *(uint2*)((__half*)base + elem) = raw;
  } else {
    *(float4*)((float*)base + elem) = o;
  }
}

template <bool F16>
__global__ void __launch_bounds__(HEADB_WARPS * 32, 2) head_backward_kernel(const HeadBwdBatch batch) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_dw[HEADB_MAX_N][128];
  __shared__ __align__(16) float s_w[HEADB_MAX_N][128];
  __shared__ float s_db[HEADB_MAX_N];
  const HeadBwdProblem P = batch.p[blockIdx.y];
  const int N = P.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * HEADB_WARPS + warp, nwarps = gridDim.x * HEADB_WARPS;
  for (int k0 = 0; k0 < P.K; k0 += 128) {
    const int col = k0 + lane * 4;
    const bool col_ok = col < P.K;  // K % 4 == 0 is required by the dispatcher
    // the weight panel lives in shared memory (one conflict-free 16-byte load per output per iteration), which
    // leaves the registers to the rows in flight
    for (int i = threadIdx.x; i < HEADB_MAX_N * 128; i += HEADB_WARPS * 32) {
      const int a = i >> 7, c = k0 + (i & 127);
      (&s_w[0][0])[i] = (a < N && c < P.K) ? __ldg(P.w + (size_t)a * P.K + c) : 0.f;
      (&s_dw[0][0])[i] = 0.f;
    }
    if (threadIdx.x < HEADB_MAX_N) s_db[threadIdx.x] = 0.f;
    float dw[HEADB_MAX_N][4];
#pragma unroll
    for (int a = 0; a < HEADB_MAX_N; ++a) dw[a][0] = dw[a][1] = dw[a][2] = dw[a][3] = 0.f;
    float db_acc = 0.f;
    __syncthreads();
    // software pipeline: the loads of the next HEADB_ROWS rows are in flight while the current ones are consumed
    float dyn[HEADB_ROWS];
    typename HeadbRaw<F16>::type hvn[HEADB_ROWS];
    auto fetch = [&](int row) {
#pragma unroll
      for (int j = 0; j < HEADB_ROWS; ++j) {
        const int r = row + j * nwarps;
        const bool ok = r < P.M;
        dyn[j] = (ok && lane < N) ? __ldg(P.dy + (size_t)r * P.ldy + lane) : 0.f;
        if (ok && col_ok) hvn[j] = headb_load_raw<F16>(P.h, (size_t)r * P.ldh + col);
        else hvn[j] = typename HeadbRaw<F16>::type{};
      }
    };
    fetch(gwarp);
    for (int row = gwarp; row < P.M; row += HEADB_ROWS * nwarps) {
      float dyl[HEADB_ROWS];
      float4 hv[HEADB_ROWS], dx[HEADB_ROWS];
#pragma unroll
      for (int j = 0; j < HEADB_ROWS; ++j) {
        dyl[j] = dyn[j];
        hv[j] = headb_cvt(hvn[j]);
        dx[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (row + HEADB_ROWS * nwarps < P.M) fetch(row + HEADB_ROWS * nwarps);
#pragma unroll
      for (int j = 0; j < HEADB_ROWS; ++j) db_acc += dyl[j];
#pragma unroll
      for (int a = 0; a < HEADB_MAX_N; ++a) {
        if (a < N) {
          const float4 w4 = *(const float4*)&s_w[a][lane * 4];
#pragma unroll
          for (int j = 0; j < HEADB_ROWS; ++j) {
            const float d = __shfl_sync(0xffffffffu, dyl[j], a);
            dx[j].x = fmaf(d, w4.x, dx[j].x); dx[j].y = fmaf(d, w4.y, dx[j].y);
            dx[j].z = fmaf(d, w4.z, dx[j].z); dx[j].w = fmaf(d, w4.w, dx[j].w);
            dw[a][0] = fmaf(d, hv[j].x, dw[a][0]); dw[a][1] = fmaf(d, hv[j].y, dw[a][1]);
            dw[a][2] = fmaf(d, hv[j].z, dw[a][2]); dw[a][3] = fmaf(d, hv[j].w, dw[a][3]);
          }
        }
      }
      if (P.dx != nullptr && col_ok) {
#pragma unroll
        for (int j = 0; j < HEADB_ROWS; ++j) {
          const int r = row + j * nwarps;
          if (r < P.M) {
            float4 o = dx[j];
            if (P.act == RSLRL_ACT_ELU) {  // ELU'(z) from the output h: 1 for h > 0, h + 1 otherwise
              o.x *= fminf(hv[j].x, 0.f) + 1.f; o.y *= fminf(hv[j].y, 0.f) + 1.f;
              o.z *= fminf(hv[j].z, 0.f) + 1.f; o.w *= fminf(hv[j].w, 0.f) + 1.f;
            } else if (P.act != RSLRL_ACT_IDENTITY) {
              o.x *= act_grad_from_output(P.act, hv[j].x); o.y *= act_grad_from_output(P.act, hv[j].y);
              o.z *= act_grad_from_output(P.act, hv[j].z); o.w *= act_grad_from_output(P.act, hv[j].w);
            }
            o.x *= P.dx_scale; o.y *= P.dx_scale; o.z *= P.dx_scale; o.w *= P.dx_scale;
            headb_store4<F16>(P.dx, (size_t)r * P.ldx + col, o);
          }
        }
      }
    }
    // block-level reduction of the weight / bias gradient partials, then one atomic per element per block
#pragma unroll
    for (int a = 0; a < HEADB_MAX_N; ++a) {
      if (a < N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&s_dw[a][lane * 4 + j], dw[a][j]);
      }
    }
    if (k0 == 0 && lane < N) atomicAdd(&s_db[lane], db_acc);
    __syncthreads();
    for (int i = threadIdx.x; i < N * 128; i += HEADB_WARPS * 32) {
      const int a = i >> 7, c = k0 + (i & 127);
      if (c < P.K) atomicAdd(P.dw + (size_t)a * P.K + c, s_dw[a][i & 127]);
    }
    if (k0 == 0 && P.db != nullptr && threadIdx.x < N) atomicAdd(P.db + threadIdx.x, s_db[threadIdx.x]);
    __syncthreads();
  }
}

// returns -1 when the problems do not fit the head kernel (caller uses the general GEMM path)
int head_linear_backward(int n, const rslrl_linear_bwd_head_t* pr, cudaStream_t st) {
  HeadBwdBatch b{};
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_bwd_head_t& q = pr[i];
    if (q.N > HEADB_MAX_N || (q.K & 3) || (q.ldh & 3) || ((uintptr_t)q.h & 7) || ((uintptr_t)q.w & 15)) return -1;
    if (!q.f16 && ((uintptr_t)q.h & 15)) return -1;
    if (q.dx && ((q.ldx & 3) || ((uintptr_t)q.dx & (q.f16 ? 7 : 15)))) return -1;
    if ((q.f16 != 0) != (pr[0].f16 != 0)) return -1;
    b.p[i] = HeadBwdProblem{q.dy, q.w, q.h, q.dx, q.dw, q.db, q.M, q.N, q.K, q.ldy, q.ldh, q.ldx, q.act,
                            q.dx_scale == 0.f ? 1.f : q.dx_scale};
  }
  // one block per SM per problem: the weight-gradient partials end in global atomics (N x K per block), so fewer,
  // longer-running blocks keep the per-address contention at 148
  if (pr[0].f16) launch_chain_kernel(head_backward_kernel<true>, dim3(kNumSMs, n), dim3(HEADB_WARPS * 32), 0, st, b);
  else launch_chain_kernel(head_backward_kernel<false>, dim3(kNumSMs, n), dim3(HEADB_WARPS * 32), 0, st, b);
  return finish_launch("rslrl_linear_backward_head");
}

}  // namespace rslrl
