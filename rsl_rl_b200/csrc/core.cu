// Library plumbing: error channel, launch counter, version.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include <cuda_fp16.h>

#include "common.cuh"

namespace rslrl {

// fp32 -> fp16 copy of a strided 2-D array, 4 elements per thread when the row geometry allows it
__global__ void __launch_bounds__(256) convert_f32_f16_kernel(const float* __restrict__ src, long long ld_src,
                                                             __half* __restrict__ dst, long long ld_dst, long long rows,
                                                             int cols, int cols_pad, int vec) {
  pdl_trigger();
  pdl_wait();
  const int per_row = vec ? cols_pad / 4 : cols_pad;
  const long long total = rows * per_row;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / per_row;
    const int c = (int)(i - r * per_row);
    if (vec) {
      const int c4 = c * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 < cols) v = __ldg((const float4*)(src + r * ld_src + c4));  // cols % 4 == 0 in vector mode
      __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
      uint2 o;
      o.x = *(unsigned int*)&lo;
      o.y = *(unsigned int*)&hi;
      *(uint2*)(dst + r * ld_dst + c4) = o;
    } else {
      dst[r * ld_dst + c] = __float2half_rn(c < cols ? __ldg(src + r * ld_src + c) : 0.f);
    }
  }
}

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

bool carveout_enabled() {
  static const bool on = [] { const char* e = getenv("RSLRL_SMEM_CARVEOUT"); return e != nullptr && e[0] == '1'; }();
  return on;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  // default OFF: measured on B200 (profiles/experiments/notes/r1_pdl.md) the programmatic edges made the graph-replayed
  // iteration slower (6.08 vs 5.63 ms); the hooks stay for further experiments (RSLRL_PDL=1)
  static const bool on = [] { const char* e = getenv("RSLRL_PDL"); return e && e[0] == '1'; }();
  return on;
}

}  // namespace rslrl

extern "C" {

int rslrl_abi_version(void) { return RSLRL_ABI_VERSION; }
const char* rslrl_last_error(void) { return rslrl::g_err; }
unsigned long long rslrl_launch_count(void) { return rslrl::g_launches.load(); }
void rslrl_reset_launch_count(void) { rslrl::g_launches.store(0); }
void rslrl_add_launch_count(unsigned long long n) { rslrl::count_launch(n); }

int rslrl_fill_zero(void* p, int64_t bytes, rslrl_stream_t stream) {
  RSLRL_CHECK_ARG(p != nullptr && bytes >= 0, "rslrl_fill_zero");
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    rslrl::set_error("rslrl_fill_zero: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int rslrl_convert_f32_f16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int cols,
                          int cols_pad, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_convert_f32_f16";
  RSLRL_CHECK_ARG(src && dst && rows >= 0 && cols > 0 && cols_pad >= cols, what);
  RSLRL_CHECK_ARG(rows <= 1 || (ld_src >= cols && ld_dst >= cols_pad), what);
  if (rows == 0) return 0;
  const int vec = (cols % 4 == 0 && cols_pad % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 &&
                   ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0) ? 1 : 0;
  const long long total = rows * (long long)(vec ? cols_pad / 4 : cols_pad);
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  cudaError_t e = launch_chain_kernel(convert_f32_f16_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, src,
                                      (long long)ld_src, (__half*)dst, (long long)ld_dst, (long long)rows, cols, cols_pad, vec);
  (void)e;
  return finish_launch(what);
}

}  // extern "C"
