// Library plumbing: error channel, launch counter, version.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace rslrl {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

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

}  // extern "C"
