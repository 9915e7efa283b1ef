/* TEST INFRASTRUCTURE -- plain-C restatement of RolloutStorage.compute_returns
 * (reference rsl_rl/storage/rollout_storage.py:127-149), used as a second, independent checker of
 * oracle/ppo_oracle.py::compute_returns and as the CPU baseline for the GAE bandwidth figure.
 * Not linked into the product library.  Build: `make -C oracle` (also done by __graft_entry__.build()).
 *
 * rewards/values/returns/advantages: [T*N] fp32 (row t = step t), dones: [T*N] u8, last_values: [N].
 * The fp32 operation order is the reference's: delta = (r + ((nnt*gamma)*v_next)) - v;
 * adv = delta + (((nnt*gamma)*lam)*adv); ret = adv + v; advantages = ret - v; then the global
 * (x - mean) / (std_unbiased + 1e-8) with fp64 moments. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#if defined(__GNUC__)
#pragma GCC optimize("no-fast-math")
#pragma STDC FP_CONTRACT OFF
#endif

int gae_oracle(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
               float* returns, float* advantages, int T, int N, float gamma, float lam, int normalize) {
  if (T <= 0 || N <= 0) return 1;
  float* adv = (float*)calloc((size_t)N, sizeof(float));
  if (!adv) return 2;
  for (int t = T - 1; t >= 0; --t) {
    const float* next_v = (t == T - 1) ? last_values : values + (size_t)(t + 1) * N;
    const size_t o = (size_t)t * N;
    for (int n = 0; n < N; ++n) {
      volatile float nnt = 1.0f - (float)dones[o + n];
      volatile float ng = nnt * gamma;
      volatile float boot = ng * next_v[n];
      volatile float sum = rewards[o + n] + boot;
      volatile float delta = sum - values[o + n];
      volatile float ngl = ng * lam;
      volatile float carry = ngl * adv[n];
      volatile float a = delta + carry;
      adv[n] = a;
      volatile float ret = a + values[o + n];
      returns[o + n] = ret;
      volatile float x = ret - values[o + n];
      advantages[o + n] = x;
    }
  }
  free(adv);
  if (normalize) {
    const size_t B = (size_t)T * N;
    double s = 0.0, ss = 0.0;
    for (size_t i = 0; i < B; ++i) { s += advantages[i]; ss += (double)advantages[i] * advantages[i]; }
    const double mean = s / (double)B;
    double var = (ss - s * s / (double)B) / ((double)B - 1.0);
    if (var < 0) var = 0;
    const float m = (float)mean, d = (float)sqrt(var) + 1e-8f;
    for (size_t i = 0; i < B; ++i) advantages[i] = (advantages[i] - m) / d;
  }
  return 0;
}
