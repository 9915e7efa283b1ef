// Fused forward of a 3-hidden-layer MLP (x -> N1 -> N2 -> N3 -> N4) on sm_100a: one persistent kernel, one CTA
// per SM, each work item = (net, 128-row tile).  All four layers run on tcgen05 (kind::tf32) with the
// accumulators in TMEM; the activations h1, h2, h3 never leave the SM between layers:
//
//   L1   acc1[c&1] (128 TMEM columns, double buffered) = x_tile W1[c*128:(c+1)*128,:]^T   per 128-column chunk c
//        epilogue: + bias, activation, round to tf32 (rna) -> "granules" (128 rows x 32 columns, 16 KB, K-major
//        128-byte swizzle) in shared memory = the A operand of the next layer, read in place by the MMA
//   L2   acc2 (N2 <= 256 columns) += granule x W2[:, k-block]^T  while the epilogue produces the next chunk
//   L3   acc3 (aliases acc1[0])   likewise from the h2 granules;  L4  acc4 (aliases acc1[1], 16 columns) from h3
//   update mode: a store warp copies every granule to h1/h2/h3 in HBM with TMA (they are the saved activations of
//   the backward pass); rollout mode stores nothing but the N4 outputs.
//
// Warps: 0 TMA producer (weight ring: x and weight k-blocks in MMA consumption order), 1 MMA issuer + TMEM owner,
// 2 granule store warp, 3.. epilogue warps (E = 16: one (row-quarter, 32-column block) unit each per 128-column
// chunk).  Shared memory: weight ring (32 KB stages), granule ring (16 KB each), biases of both nets, barriers.
//
// Numerics are those of the layer-by-layer tcgen05 path (gemm_tc.cu): operands rounded fp32 -> tf32 to nearest
// (TMA does it for x and the weights, the epilogue for the activations), fp32 accumulation in the same k order.
// Replaces rsl_rl/networks/mlp.py:106-114 (MLP.forward) for actor and critic in one launch.
#include <cstdlib>

#include "tc_ptx.cuh"

namespace rslrl {

bool tc_make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_cols,
                 int box_rows, bool operand);
bool tc_make_map16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
                   int box_rows);
int tc_num_sms();

constexpr int MF_EPI_WARPS = 16;
constexpr int MF_FIRST_EPI_WARP = 3;
constexpr int MF_THREADS = (MF_FIRST_EPI_WARP + MF_EPI_WARPS) * 32;
#ifndef RSLRL_MF_W_STAGES
#define RSLRL_MF_W_STAGES 3
#endif
#ifndef RSLRL_MF_GRANULES
#define RSLRL_MF_GRANULES 7
#endif
constexpr int MF_W_STAGES = RSLRL_MF_W_STAGES;
constexpr int MF_W_STAGE_BYTES = 32768;
constexpr int MF_GRANULES = RSLRL_MF_GRANULES;
constexpr int MF_GRANULE_BYTES = 16384;
constexpr int MF_MAX_BIAS = 512 + 256 + 128 + 16;

struct MfNet {
  CUtensorMap map_x, map_w1, map_w2, map_w3, map_w4;
  CUtensorMap map_h1, map_h2, map_h3;  // stores (valid when store_h)
  const float* b[4];
  float* y;
  int M, K0, N1, N2, N3, N4, ldy, act, store_h;
  // fused Gaussian head (rollout): actions = z * sigma + mu, log-prob, sigma rows; head_z == nullptr: off
  const float* head_z;
  const float* head_std;
  float* head_actions;
  float* head_logp;
  float* head_sigma;
  float* head_actions_host;
  int head_std_type;
};

struct alignas(64) MfBatch {
  MfNet net[2];
};

// Optional in-kernel timeline (build with -DRSLRL_TC_TRACE; read with profiles/experiments/mlp3_trace.py): per CTA,
// slot 0 entry, 1 prologue done, then 9 slots per tile j (first three tiles): L1(0) issued, E1(0) done, last L1
// chunk's epilogue done, acc2 published, E2 done, acc3 published, E3 done, acc4 published, E4 done.
#ifdef RSLRL_TC_TRACE
__device__ unsigned long long g_mf_trace[148 * 32];
__device__ __forceinline__ void mf_trace(int slot) {
  if (slot < 32) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_mf_trace[blockIdx.x * 32 + slot] = t;
  }
}
#define MF_TRACE(cond, slot) do { if (cond) mf_trace(slot); } while (0)
#else
#define MF_TRACE(cond, slot) do { } while (0)
#endif

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// round-to-nearest (ties away) to tf32 with two integer ops: add half an ulp of the 10-bit mantissa, clear the 13
// low bits (same result as cvt.rna.tf32.f32 for finite inputs; activations are finite)
__device__ __forceinline__ float tf32_round_bits(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// ELU(z) = max(z, exp(min(z, 0)) - 1): exp(z) - 1 >= z everywhere and is 0 for z >= 0 (5 instructions, one MUFU)
__device__ __forceinline__ float elu_max_form(float z) {
  return fmaxf(z, ex2_approx(fminf(z, 0.f) * 1.4426950408889634f) - 1.f);
}

// ---- 2-CTA cluster helpers: the two CTAs of a cluster work on adjacent row tiles of the SAME net, so they consume the
// same weight k-blocks in the same order; each CTA fetches half of every weight k-block and TMA-multicasts it into
// both CTAs' rings (half the L2->SM weight traffic per SM, which is what bounds the layer-1/2 phase of a tile).
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// MMA completion -> arrive on the barrier at the same offset in every CTA of the mask
__device__ __forceinline__ void tcgen05_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// F16: the activations leave the epilogue as fp16 (granules of 128 rows x 64 columns, still 128-byte rows), layers 2-4
// run as kind::f16 MMAs against the fp16 weight shadows and the saved activations are stored as fp16.  Layer 1 is
// tf32 in both modes (x and W1 are fp32 in HBM).
// MC: launched as clusters of two CTAs with multicast weight loads (see the helpers above).
template <bool F16, bool MC>
__global__ void __launch_bounds__(MF_THREADS, 1) mlp3_forward_kernel(const __grid_constant__ MfBatch batch,
                                                                     const int n_nets) {
  const int cta_rank = MC ? (int)(blockIdx.x & 1) : 0;   // cluster dims (2,1,1): rank = blockIdx.x % 2
  constexpr int GK = F16 ? 64 : 32;    // activation columns per granule = k-block of layers 2-4
  constexpr int KB2 = 128 / GK;        // granules per 128-column chunk of layer 1
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* const wring = smem_raw + pad;
  uint8_t* const gring = wring + MF_W_STAGES * MF_W_STAGE_BYTES;
  float* const bias_s = (float*)(gring + MF_GRANULES * MF_GRANULE_BYTES);  // [2][MF_MAX_BIAS]
  float* const head_s = bias_s + 2 * MF_MAX_BIAS;                          // [2][3][16]: sigma, log sigma, 2 sigma^2
  uint64_t* const w_full = (uint64_t*)(head_s + 2 * 48);
  uint64_t* const w_empty = w_full + MF_W_STAGES;
  uint64_t* const g_full = w_empty + MF_W_STAGES;
  uint64_t* const g_empty = g_full + MF_GRANULES;
  uint64_t* const a_full = g_empty + MF_GRANULES;  // [3]: TMEM buffer 0, buffer 1, acc2
  uint64_t* const a_empty = a_full + 3;            // [3]
  uint32_t* const tmem_ptr_smem = (uint32_t*)(a_empty + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  MF_TRACE(threadIdx.x == 0, 0);

  pdl_trigger();
  if (warp == 0 && elect_one()) {
    for (int i = 0; i < n_nets; ++i) {
      const MfNet& P = batch.net[i];
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&P.map_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&P.map_w1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&P.map_w2) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&P.map_w3) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&P.map_w4) : "memory");
    }
    for (int s = 0; s < MF_W_STAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], MC ? 2 : 1);   // MMA commits of every CTA that receives the stage
    }
    for (int g = 0; g < MF_GRANULES; ++g) {
      mbar_init(&g_full[g], F16 ? 8 : 4);   // the row-quarter warps (x 2 column halves in fp16) that write a granule
      mbar_init(&g_empty[g], 2);  // MMA commit + store warp
    }
    for (int a = 0; a < 3; ++a) {
      mbar_init(&a_full[a], 1);
      mbar_init(&a_empty[a], MF_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (MC) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast into this CTA
  pdl_wait();  // everything above touched no global memory; the parameters below were written by the previous kernels
  // biases of every layer of every net -> shared memory (read by the epilogues as broadcast 16-byte loads)
  for (int i = threadIdx.x; i < n_nets * MF_MAX_BIAS; i += MF_THREADS) {
    const int ni = i / MF_MAX_BIAS, o = i - ni * MF_MAX_BIAS;
    const MfNet& P = batch.net[ni];
    float v = 0.f;
    if (o < 512) { if (o < P.N1 && P.b[0]) v = __ldg(P.b[0] + o); }
    else if (o < 768) { if (o - 512 < P.N2 && P.b[1]) v = __ldg(P.b[1] + o - 512); }
    else if (o < 896) { if (o - 768 < P.N3 && P.b[2]) v = __ldg(P.b[2] + o - 768); }
    else { if (o - 896 < P.N4 && P.b[3]) v = __ldg(P.b[3] + o - 896); }
    bias_s[i] = v;
  }
  if (threadIdx.x < n_nets * 16) {   // Gaussian head constants (same expressions as gaussian_act_kernel)
    const int ni = threadIdx.x >> 4, a = threadIdx.x & 15;
    const MfNet& P = batch.net[ni];
    if (P.head_z != nullptr && a < P.N4) {
      const float p = __ldg(P.head_std + a);
      const float sg = P.head_std_type == RSLRL_STD_LOG ? expf(p) : p;
      head_s[ni * 48 + a] = sg;
      head_s[ni * 48 + 16 + a] = logf(sg);
      head_s[ni * 48 + 32 + a] = 2.f * (sg * sg);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  MF_TRACE(threadIdx.x == 0, 1);

  // work items: (net, 128-row tile), net fastest.  MC: a cluster takes the pair (net, tiles 2k and 2k+1) - a phantom
  // tile past the end of an odd tile count loads zeros and stores nothing (TMA clips, the output rows are guarded)
  const int tiles_m = (batch.net[0].M + TC_BLOCK_M - 1) / TC_BLOCK_M;
  const int it0 = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int it_step = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int total = MC ? n_nets * ((tiles_m + 1) / 2) : tiles_m * n_nets;
  auto item_m0 = [&](int t) { return (MC ? (t / n_nets) * 2 + cta_rank : t / n_nets) * TC_BLOCK_M; };

  if (warp == 0) {
    // ===================================================================== TMA producer (weight ring)
    if (elect_one()) {
      int it = 0;
      auto stage = [&](int& s) -> uint8_t* {
        s = it % MF_W_STAGES;
        mbar_wait(&w_empty[s], ((it / MF_W_STAGES) & 1) ^ 1);
        ++it;
        return wring + (size_t)s * MF_W_STAGE_BYTES;
      };
      for (int t = it0; t < total; t += it_step) {
        const MfNet& P = batch.net[t % n_nets];
        const int m0 = item_m0(t);
        const int kb0 = (P.K0 + TC_BLOCK_K - 1) / TC_BLOCK_K, C1 = P.N1 / 128;
        // a weight k-block of `rows` rows arrives as two half-row boxes: fetched and multicast by one CTA each (MC), or
        // both fetched locally
        auto wload = [&](uint8_t* dst, const CUtensorMap* map, int s, int col, int row0, int rows) {
          const int half = rows / 2;
          if (MC) {
            tma_load_2d_mc(dst + (size_t)cta_rank * half * 128, map, &w_full[s], col, row0 + cta_rank * half, (uint16_t)3);
          } else {
            tma_load_2d(dst, map, &w_full[s], col, row0);
            tma_load_2d(dst + (size_t)half * 128, map, &w_full[s], col, row0 + half);
          }
        };
        auto l1 = [&](int c) {  // x k-block + the chunk's W1 k-block per stage
          for (int kb = 0; kb < kb0; ++kb) {
            int s;
            uint8_t* dst = stage(s);
            mbar_expect_tx(&w_full[s], 2u * 16384u);
            tma_load_2d(dst, &P.map_x, &w_full[s], kb * TC_BLOCK_K, m0);
            wload(dst + 16384, &P.map_w1, s, kb * TC_BLOCK_K, c * 128, 128);
          }
        };
        auto l2p = [&](int c) {
          for (int kb = 0; kb < KB2; ++kb) {
            int s;
            uint8_t* dst = stage(s);
            mbar_expect_tx(&w_full[s], (uint32_t)P.N2 * 128u);
            wload(dst, &P.map_w2, s, (c * KB2 + kb) * GK, 0, P.N2);
          }
        };
        l1(0);
        for (int c = 0; c < C1; ++c) {
          if (c + 1 < C1) l1(c + 1);
          l2p(c);
        }
        // the narrow layers pack several k-blocks into one 32 KB stage (bytes in flight, not stage count, is what
        // the weight stream needs)
        const int nkb3 = P.N2 / GK, pack3 = min(MF_W_STAGE_BYTES / (P.N3 * 128), nkb3);
        for (int kb = 0; kb < nkb3; kb += pack3) {
          int s;
          uint8_t* dst = stage(s);
          mbar_expect_tx(&w_full[s], (uint32_t)(pack3 * P.N3) * 128u);
          for (int i = 0; i < pack3; ++i)
            wload(dst + (size_t)i * P.N3 * 128, &P.map_w3, s, (kb + i) * GK, 0, P.N3);
        }
        {
          const int nkb4 = P.N3 / GK;
          int s;
          uint8_t* dst = stage(s);
          mbar_expect_tx(&w_full[s], (uint32_t)nkb4 * 16u * 128u);
          for (int i = 0; i < nkb4; ++i) tma_load_2d(dst + i * 2048, &P.map_w4, &w_full[s], i * GK, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int it = 0;          // weight ring position
    int gn = 0;          // granule counter
    uint32_t use[3] = {0, 0, 0};  // uses of TMEM buffer 0 / buffer 1 / acc2 so far
    // one k-block (32 tf32 columns) of MMAs: A at `sa`, B at `sb` (both K-major, 128-byte swizzle)
    // (a 128-byte k-block is 4 MMAs in both element types: 4 x 8 tf32 or 4 x 16 fp16)
    auto kblock = [&](uint32_t d_tmem, uint32_t sa, uint32_t sb, uint32_t idesc, bool first, bool f16) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
          const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024, kLayoutSw128);
          const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024, kLayoutSw128);
          if (f16) umma_f16(d_tmem, da, db, idesc, (first && k == 0) ? 0u : 1u);
          else umma_tf32(d_tmem, da, db, idesc, (first && k == 0) ? 0u : 1u);
        }
      }
      __syncwarp();
    };
    auto wait_w = [&](int& s) -> uint32_t {
      s = it % MF_W_STAGES;
      mbar_wait(&w_full[s], (it / MF_W_STAGES) & 1);
      ++it;
      return smem_u32(wring + (size_t)s * MF_W_STAGE_BYTES);
    };
    auto wait_g = [&](int& g) -> uint32_t {
      g = gn % MF_GRANULES;
      mbar_wait(&g_full[g], (gn / MF_GRANULES) & 1);
      ++gn;
      return smem_u32(gring + (size_t)g * MF_GRANULE_BYTES);
    };
    auto acquire_acc = [&](int a) {
      mbar_wait(&a_empty[a], (use[a] & 1) ^ 1);
      tcgen05_fence_after();
    };
    auto publish_acc = [&](int a) {
      if (elect_one()) tcgen05_commit(&a_full[a]);
      __syncwarp();
      ++use[a];
    };
    int tj = -1;
    for (int t = it0; t < total; t += it_step) {
      ++tj;
      const MfNet& P = batch.net[t % n_nets];
      const int kb0 = (P.K0 + TC_BLOCK_K - 1) / TC_BLOCK_K, C1 = P.N1 / 128;
      const uint32_t idesc1 = make_idesc(128, false, false);
      const uint32_t idesc2 = F16 ? make_idesc_f16(P.N2, false, false) : make_idesc(P.N2, false, false);
      const uint32_t idesc3 = F16 ? make_idesc_f16(P.N3, false, false) : make_idesc(P.N3, false, false);
      const uint32_t idesc4 = F16 ? make_idesc_f16(16, false, false) : make_idesc(16, false, false);
      auto l1 = [&](int c) {
        const int a = c & 1;
        acquire_acc(a);
        for (int kb = 0; kb < kb0; ++kb) {
          int s;
          const uint32_t base = wait_w(s);
          tcgen05_fence_after();
          kblock(tmem_base + (uint32_t)a * 128u, base, base + 16384u, idesc1, kb == 0, false);
          if (elect_one()) { if (MC) tcgen05_commit_mc(&w_empty[s], (uint16_t)3); else tcgen05_commit(&w_empty[s]); }
          __syncwarp();
        }
        publish_acc(a);
      };
      // consume `n` granules against weight k-blocks into the accumulator at column `col`
      // `pack` k-blocks of `kb_bytes` each share one weight stage
      auto from_granules = [&](int n, uint32_t col, uint32_t idesc, bool first_overall, int pack, uint32_t kb_bytes) {
        for (int kb = 0; kb < n; kb += pack) {
          int s;
          const uint32_t wb = wait_w(s);
          for (int i = 0; i < pack; ++i) {
            int g;
            const uint32_t ga = wait_g(g);
            tcgen05_fence_after();
            kblock(tmem_base + col, ga, wb + (uint32_t)i * kb_bytes, idesc, first_overall && kb + i == 0, F16);
            if (elect_one()) tcgen05_commit(&g_empty[g]);
            __syncwarp();
          }
          if (elect_one()) { if (MC) tcgen05_commit_mc(&w_empty[s], (uint16_t)3); else tcgen05_commit(&w_empty[s]); }
          __syncwarp();
        }
      };
      l1(0);
      MF_TRACE(lane == 0, 2 + 9 * tj);
      acquire_acc(2);
      for (int c = 0; c < C1; ++c) {
        if (c + 1 < C1) l1(c + 1);
        from_granules(KB2, 256u, idesc2, c == 0, 1, 0u);
      }
      publish_acc(2);
      MF_TRACE(lane == 0, 2 + 9 * tj + 3);
      acquire_acc(0);
      from_granules(P.N2 / GK, 0u, idesc3, true, min(MF_W_STAGE_BYTES / (P.N3 * 128), P.N2 / GK), (uint32_t)P.N3 * 128u);
      publish_acc(0);
      MF_TRACE(lane == 0, 2 + 9 * tj + 5);
      acquire_acc(1);
      from_granules(P.N3 / GK, 128u, idesc4, true, P.N3 / GK, 2048u);
      publish_acc(1);
      MF_TRACE(lane == 0, 2 + 9 * tj + 7);
    }
  } else if (warp == 2) {
    // ===================================================================== granule store warp
    int gn = 0;
    for (int t = it0; t < total; t += it_step) {
      const MfNet& P = batch.net[t % n_nets];
      const int m0 = item_m0(t);
      const int n1 = P.N1 / GK, n2 = P.N2 / GK, n3 = P.N3 / GK;
      for (int i = 0; i < n1 + n2 + n3; ++i, ++gn) {
        const int g = gn % MF_GRANULES;
        mbar_wait(&g_full[g], (gn / MF_GRANULES) & 1);
        if (lane == 0) {
          if (P.store_h) {
            const CUtensorMap* map = i < n1 ? &P.map_h1 : (i < n1 + n2 ? &P.map_h2 : &P.map_h3);
            const int col = (i < n1 ? i : (i < n1 + n2 ? i - n1 : i - n1 - n2)) * GK;
            tma_store_2d(map, gring + (size_t)g * MF_GRANULE_BYTES, col, m0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          mbar_arrive(&g_empty[g]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // ===================================================================== epilogue warps
    const int ew = warp - MF_FIRST_EPI_WARP;
    const int q = warp & 3;    // TMEM sub-partition = row quarter of the tile
    const int j = ew >> 2;     // column-block slot 0..3
    const uint32_t sw = (uint32_t)(lane & 7);
    int gbase = 0;
    uint32_t use[3] = {0, 0, 0};
    int tj = -1;
    for (int t = it0; t < total; t += it_step) {
      ++tj;
      const int ni = t % n_nets;
      const MfNet& P = batch.net[ni];
      const int m0 = item_m0(t);
      const int C1 = P.N1 / 128;
      const float* const bias_net = bias_s + ni * MF_MAX_BIAS;
      const int act = P.act;

      // one unit: TMEM columns [col, col+32) of this warp's 32 rows -> bias + activation -> granule `gidx`
      // (fp16: two units share a granule; `hsel` picks the 64-byte half of the 128-byte rows)
      auto unit = [&](uint32_t col, const float* bias, int gidx, int hsel, int acc, bool last_read) {
        uint32_t v[32];
        TMEM_LD_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + col, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (last_read) {  // hand the accumulator back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_empty[acc]);
        }
        const int g = gidx % MF_GRANULES;
        mbar_wait(&g_empty[g], ((gidx / MF_GRANULES) & 1) ^ 1);
        uint8_t* const my_row = gring + (size_t)g * MF_GRANULE_BYTES + q * 4096 + lane * 128;
        if constexpr (F16) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float z[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 b4 = *(const float4*)(bias + 16 * hh + 4 * u);
              z[4 * u + 0] = __uint_as_float(v[16 * hh + 4 * u + 0]) + b4.x;
              z[4 * u + 1] = __uint_as_float(v[16 * hh + 4 * u + 1]) + b4.y;
              z[4 * u + 2] = __uint_as_float(v[16 * hh + 4 * u + 2]) + b4.z;
              z[4 * u + 3] = __uint_as_float(v[16 * hh + 4 * u + 3]) + b4.w;
            }
            if (act == RSLRL_ACT_ELU) {
#pragma unroll
              for (int i = 0; i < 16; ++i) z[i] = elu_max_form(z[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) z[i] = act_forward(act, z[i]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)  // 16-byte chunk = 8 fp16 columns
              *(uint4*)(my_row + (((uint32_t)(4 * hsel + 2 * hh + u) ^ sw) << 4)) =
                  make_uint4(pack_f16x2(z[8 * u], z[8 * u + 1]), pack_f16x2(z[8 * u + 2], z[8 * u + 3]),
                             pack_f16x2(z[8 * u + 4], z[8 * u + 5]), pack_f16x2(z[8 * u + 6], z[8 * u + 7]));
          }
        } else if (act == RSLRL_ACT_ELU) {
          // branch-free fast path, two halves of 16 columns: 16 independent chains per lane (ILP hides the MUFU and
          // shared-memory latencies; only four epilogue warps share a scheduler)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float z[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 b4 = *(const float4*)(bias + 16 * hh + 4 * u);
              z[4 * u + 0] = __uint_as_float(v[16 * hh + 4 * u + 0]) + b4.x;
              z[4 * u + 1] = __uint_as_float(v[16 * hh + 4 * u + 1]) + b4.y;
              z[4 * u + 2] = __uint_as_float(v[16 * hh + 4 * u + 2]) + b4.z;
              z[4 * u + 3] = __uint_as_float(v[16 * hh + 4 * u + 3]) + b4.w;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = tf32_round_bits(elu_max_form(z[i]));
#pragma unroll
            for (int u = 0; u < 4; ++u)
              *(float4*)(my_row + (((4 * hh + u) ^ sw) << 4)) = make_float4(z[4 * u], z[4 * u + 1], z[4 * u + 2], z[4 * u + 3]);
          }
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 b4 = *(const float4*)(bias + 4 * u);
            float4 o;
            o.x = tf32_rna(act_forward(act, __uint_as_float(v[4 * u + 0]) + b4.x));
            o.y = tf32_rna(act_forward(act, __uint_as_float(v[4 * u + 1]) + b4.y));
            o.z = tf32_rna(act_forward(act, __uint_as_float(v[4 * u + 2]) + b4.z));
            o.w = tf32_rna(act_forward(act, __uint_as_float(v[4 * u + 3]) + b4.w));
            *(float4*)(my_row + ((u ^ sw) << 4)) = o;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&g_full[g]);
      };
      auto wait_acc = [&](int a) {
        mbar_wait(&a_full[a], use[a] & 1);
        tcgen05_fence_after();
        ++use[a];
      };
      auto release_only = [&](int a) {  // a warp without a unit in this accumulator still owes its arrival
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_empty[a]);
      };

      // L1: chunk c lives in TMEM buffer c & 1, columns [ (c&1)*128 + 32 j, +32 )
      for (int c = 0; c < C1; ++c) {
        const int a = c & 1;
        wait_acc(a);
        unit((uint32_t)a * 128u + 32u * j, bias_net + c * 128 + 32 * j, gbase + (F16 ? j >> 1 : j), j & 1, a, true);
        gbase += KB2;
        MF_TRACE(ew == 0 && lane == 0 && c == 0, 2 + 9 * tj + 1);
        MF_TRACE(ew == 0 && lane == 0 && c == C1 - 1, 2 + 9 * tj + 2);
      }
      // L2: acc2 at columns 256.., N2 / 32 column blocks, this warp takes j, j + 4
      {
        const int ncb = P.N2 / 32;
        wait_acc(2);
        if (j >= ncb) release_only(2);
        for (int cb = j; cb < ncb; cb += 4)
          unit(256u + 32u * cb, bias_net + 512 + 32 * cb, gbase + (F16 ? cb >> 1 : cb), cb & 1, 2, cb + 4 >= ncb);
        gbase += F16 ? ncb / 2 : ncb;
        MF_TRACE(ew == 0 && lane == 0, 2 + 9 * tj + 4);
      }
      // L3: acc3 in buffer 0
      {
        const int ncb = P.N3 / 32;
        wait_acc(0);
        if (j >= ncb) release_only(0);
        else unit(32u * j, bias_net + 768 + 32 * j, gbase + (F16 ? j >> 1 : j), j & 1, 0, true);
        gbase += F16 ? ncb / 2 : ncb;
        MF_TRACE(ew == 0 && lane == 0, 2 + 9 * tj + 6);
      }
      // L4: acc4 in buffer 1 (16 columns): the four j == 0 warps write the N4 outputs
      {
        wait_acc(1);
        if (j == 0) {
          uint32_t v[16];
          TMEM_LD_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + 128u, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          release_only(1);
          const int row = m0 + q * 32 + lane;
          if (row < P.M) {
            float* const yr = P.y + (size_t)row * P.ldy;
            const float* const b4 = bias_net + 896;
            if (P.head_z == nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < P.N4) yr[i] = __uint_as_float(v[i]) + b4[i];
            } else {
              // Gaussian policy head on this lane's row (arithmetic of gaussian_act_kernel, bit for bit)
              const float* const hs = head_s + ni * 48;
              const float* const zr = P.head_z + (size_t)row * P.N4;
              float* const ar = P.head_actions + (size_t)row * P.N4;
              float* const ah = P.head_actions_host ? P.head_actions_host + (size_t)row * P.N4 : nullptr;
              float lp = 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (i < P.N4) {
                  const float mean = __uint_as_float(v[i]) + b4[i], sg = hs[i];
                  const float a = __fadd_rn(__fmul_rn(__ldg(zr + i), sg), mean);
                  yr[i] = mean;
                  ar[i] = a;
                  if (ah) ah[i] = a;
                  const float d = a - mean;
                  lp += -(d * d) / hs[32 + i] - hs[16 + i] - 0.91893853320467274178f;
                  if (P.head_sigma) P.head_sigma[(size_t)row * P.N4 + i] = sg;
                }
              }
              P.head_logp[row] = lp;
            }
          }
        } else {
          release_only(1);
        }
        MF_TRACE(ew == 0 && lane == 0, 2 + 9 * tj + 8);
      }
    }
    tcgen05_fence_before();
  }

  __syncthreads();
  if (MC) cluster_sync_all();  // the peer may still multicast commits into this CTA's barriers until it is done too
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static_assert(sizeof(rslrl_mlp3_fwd_t) == 216, "rslrl_mlp3_fwd_t layout is mirrored by _native.Mlp3Fwd (ctypes)");

static bool ok16(const void* p, long long ld) { return ((uintptr_t)p & 15) == 0 && (ld % 4) == 0; }

static int mlp3_launch(const MfBatch& b, int n, bool f16, int M, cudaStream_t st);

// Validates the nets and encodes the tensor maps.  0 ok, 1 error, -1 not eligible (caller runs the layers one by one)
static int mlp3_build(int n, const rslrl_mlp3_fwd_t* nets, MfBatch& b, bool& f16) {
  const char* what = "rslrl_mlp3_forward";
  if (n < 1 || n > 2) return -1;
  f16 = nets[0].precision == RSLRL_PREC_FP16;
  for (int i = 0; i < n; ++i) {
    const rslrl_mlp3_fwd_t& q = nets[i];
    if (q.M != nets[0].M) return -1;
    if (q.N1 % 128 || q.N1 < 128 || q.N1 > 512 || q.N2 % 128 || q.N2 < 128 || q.N2 > 256 || q.N3 % 32 || q.N3 < 32 ||
        q.N3 > 128 || q.N4 < 1 || q.N4 > 16 || q.K0 < 1)
      return -1;
    if (!ok16(q.x, q.ldx) || !ok16(q.w[0], q.ldw0)) return -1;
    if (!f16 && (!ok16(q.w[1], q.N1) || !ok16(q.w[2], q.N2) || !ok16(q.w[3], q.N3))) return -1;
    const bool store = q.h[0] != nullptr;
    if (store && (!q.h[1] || !q.h[2])) return -1;
    if ((q.precision == RSLRL_PREC_FP16) != f16) return -1;
    MfNet& P = b.net[i];
    bool ok = tc_make_map(&P.map_x, q.x, q.M, q.K0, q.ldx, 32, 128, true) &&
              tc_make_map(&P.map_w1, (const float*)q.w[0], q.N1, q.K0, q.ldw0, 32, 64, true);
    if (f16) {
      // fp16 rows: 16-byte alignment = pitch % 8; N3 must fill whole 64-column granules
      if (q.N3 % 64 || q.N1 % 8 || q.N2 % 8 || ((uintptr_t)q.w[1] & 15) || ((uintptr_t)q.w[2] & 15) || ((uintptr_t)q.w[3] & 15))
        return -1;
      if (store)
        for (int l = 0; l < 3; ++l)
          if (((uintptr_t)q.h[l] & 15) || q.ldh[l] % 8) return -1;
      ok = ok && tc_make_map16(&P.map_w2, q.w[1], q.N2, q.N1, q.N1, 64, q.N2 / 2) &&
           tc_make_map16(&P.map_w3, q.w[2], q.N3, q.N2, q.N2, 64, q.N3 / 2) &&
           tc_make_map16(&P.map_w4, q.w[3], q.N4, q.N3, q.N3, 64, 16);
      if (ok && store)
        ok = tc_make_map16(&P.map_h1, q.h[0], q.M, q.N1, q.ldh[0], 64, 128) &&
             tc_make_map16(&P.map_h2, q.h[1], q.M, q.N2, q.ldh[1], 64, 128) &&
             tc_make_map16(&P.map_h3, q.h[2], q.M, q.N3, q.ldh[2], 64, 128);
    } else {
      if (store && (!ok16(q.h[0], q.ldh[0]) || !ok16(q.h[1], q.ldh[1]) || !ok16(q.h[2], q.ldh[2]))) return -1;
      ok = ok && tc_make_map(&P.map_w2, (const float*)q.w[1], q.N2, q.N1, q.N1, 32, q.N2 / 2, true) &&
           tc_make_map(&P.map_w3, (const float*)q.w[2], q.N3, q.N2, q.N2, 32, q.N3 / 2, true) &&
           tc_make_map(&P.map_w4, (const float*)q.w[3], q.N4, q.N3, q.N3, 32, 16, true);
      if (ok && store)
        ok = tc_make_map(&P.map_h1, (const float*)q.h[0], q.M, q.N1, q.ldh[0], 32, 128, false) &&
             tc_make_map(&P.map_h2, (const float*)q.h[1], q.M, q.N2, q.ldh[1], 32, 128, false) &&
             tc_make_map(&P.map_h3, (const float*)q.h[2], q.M, q.N3, q.ldh[2], 32, 128, false);
    }
    if (!ok) {
      set_error("%s: cuTensorMapEncodeTiled failed", what);
      return 1;
    }
    for (int l = 0; l < 4; ++l) P.b[l] = q.b[l];
    P.y = q.y; P.M = q.M; P.K0 = q.K0; P.N1 = q.N1; P.N2 = q.N2; P.N3 = q.N3; P.N4 = q.N4; P.ldy = q.ldy;
    P.act = q.act; P.store_h = store ? 1 : 0;
    P.head_z = q.head_z; P.head_std = q.head_std; P.head_actions = q.head_actions; P.head_logp = q.head_logp;
    P.head_sigma = q.head_sigma; P.head_std_type = q.head_std_type; P.head_actions_host = q.head_actions_host;
    if (q.head_z != nullptr && (!q.head_std || !q.head_actions || !q.head_logp)) {
      set_error("%s: head_z needs head_std, head_actions and head_logp", what);
      return 1;
    }
  }
  return 0;
}

int mlp3_forward_fused(int n, const rslrl_mlp3_fwd_t* nets, cudaStream_t st) {
  MfBatch b{};
  bool f16 = false;
  const int rc = mlp3_build(n, nets, b, f16);
  if (rc != 0) return rc;
  return mlp3_launch(b, n, f16, nets[0].M, st);
}

struct Mlp3Plan {
  MfBatch b;
  int n;
  bool f16;
  rslrl_mlp3_fwd_t nets[2];
};

static int mlp3_launch(const MfBatch& b, int n, bool f16, int M, cudaStream_t st) {
  const char* what = "rslrl_mlp3_forward";
  const size_t smem = MF_W_STAGES * MF_W_STAGE_BYTES + MF_GRANULES * MF_GRANULE_BYTES + 2 * MF_MAX_BIAS * 4 + 2 * 48 * 4 +
                      (2 * MF_W_STAGES + 2 * MF_GRANULES + 6) * 8 + 16 + 1024;
  // clusters of two CTAs with multicast weight loads: opt-in (RSLRL_MF_MULTICAST=1).  Measured on B200
  // (profiles/experiments/notes/r1_mlp3_multicast.md): 53.3 us vs 51.2 us without - the per-SM ingest path, not the L2
  // read bandwidth, bounds the weight stream, and multicast delivers the same bytes into every SM.
  static const bool mc_env = [] { const char* e = getenv("RSLRL_MF_MULTICAST"); return e && e[0] == '1'; }();
  const int tiles_m = (M + TC_BLOCK_M - 1) / TC_BLOCK_M;
  const bool mc = mc_env && (tc_num_sms() % 2 == 0);
  using Kern = void (*)(const MfBatch, const int);
  const Kern kern = f16 ? (mc ? (Kern)mlp3_forward_kernel<true, true> : (Kern)mlp3_forward_kernel<true, false>)
                        : (mc ? (Kern)mlp3_forward_kernel<false, true> : (Kern)mlp3_forward_kernel<false, false>);
  static bool attr_set[4] = {false, false, false, false};
  const int ki = (f16 ? 2 : 0) + (mc ? 1 : 0);
  if (!attr_set[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return 1;
    }
    attr_set[ki] = true;
  }
  if (mc) {
    const int pairs = n * ((tiles_m + 1) / 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tc_num_sms() / 2 * 2);
    cfg.blockDim = dim3(MF_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent kernel: never more clusters than can be co-resident (GPCs with an odd SM count strand one SM each)
    static int max_clusters[4] = {0, 0, 0, 0};
    if (max_clusters[ki] == 0) {
      int mcl = 0;
      if (cudaOccupancyMaxActiveClusters(&mcl, kern, &cfg) != cudaSuccess || mcl < 1) mcl = tc_num_sms() / 2;
      (void)cudaGetLastError();
      max_clusters[ki] = mcl;
    }
    cfg.gridDim = dim3(2 * (pairs < max_clusters[ki] ? pairs : max_clusters[ki]));
    cudaLaunchKernelEx(&cfg, kern, b, n);
  } else {
    const int total = tiles_m * n;
    const int grid = total < tc_num_sms() ? total : tc_num_sms();
    launch_chain_kernel(kern, dim3(grid), dim3(MF_THREADS), smem, st, b, n);
  }
  return finish_launch(what);
}

}  // namespace rslrl

extern "C" int rslrl_mlp3_forward(int n, const rslrl_mlp3_fwd_t* nets_host, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_mlp3_forward";
  RSLRL_CHECK_ARG(n > 0 && n <= RSLRL_MAX_PROBLEMS && nets_host, what);
  for (int i = 0; i < n; ++i) {
    const rslrl_mlp3_fwd_t& p = nets_host[i];
    RSLRL_CHECK_ARG(p.x && p.y && p.w[0] && p.w[1] && p.w[2] && p.w[3] && p.M > 0 && p.ldx >= p.K0 && p.ldw0 >= p.K0 &&
                        p.ldy >= p.N4,
                    what);
  }
  const int rc = mlp3_forward_fused(n, nets_host, (cudaStream_t)stream);
  if (rc < 0) {
    set_error("%s: shape not eligible (needs N1 in {128..512} %% 128, N2 in {128,256}, N3 <= 128 %% 32 (%% 64 in fp16), "
              "N4 <= 16, 16-byte aligned rows, equal M and precision)", what);
    return 4;
  }
  return rc;
}

extern "C" int rslrl_mlp3_plan_create(int n, const rslrl_mlp3_fwd_t* nets_host, void** plan_out) {
  using namespace rslrl;
  const char* what = "rslrl_mlp3_plan_create";
  RSLRL_CHECK_ARG(n > 0 && n <= 2 && nets_host && plan_out, what);
  for (int i = 0; i < n; ++i) {
    const rslrl_mlp3_fwd_t& p = nets_host[i];
    RSLRL_CHECK_ARG(p.x && p.y && p.w[0] && p.w[1] && p.w[2] && p.w[3] && p.M > 0 && p.ldx >= p.K0 && p.ldw0 >= p.K0 &&
                        p.ldy >= p.N4,
                    what);
  }
  Mlp3Plan* plan = new Mlp3Plan();
  plan->n = n;
  for (int i = 0; i < n; ++i) plan->nets[i] = nets_host[i];
  const int rc = mlp3_build(n, nets_host, plan->b, plan->f16);
  if (rc != 0) {
    delete plan;
    if (rc < 0) {
      set_error("%s: shape not eligible for the fused forward", what);
      return 4;
    }
    return rc;
  }
  *plan_out = plan;
  return 0;
}

extern "C" int rslrl_mlp3_plan_run(void* plan_v, const float* const* x_host, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_mlp3_plan_run";
  RSLRL_CHECK_ARG(plan_v && x_host, what);
  Mlp3Plan* plan = (Mlp3Plan*)plan_v;
  for (int i = 0; i < plan->n; ++i) {
    rslrl_mlp3_fwd_t& q = plan->nets[i];
    RSLRL_CHECK_ARG(x_host[i] != nullptr && ((uintptr_t)x_host[i] & 15) == 0, what);
    if (x_host[i] != q.x) {   // only the input descriptor depends on the observation pointer
      q.x = x_host[i];
      if (!tc_make_map(&plan->b.net[i].map_x, q.x, q.M, q.K0, q.ldx, 32, 128, true)) {
        set_error("%s: cuTensorMapEncodeTiled failed", what);
        return 1;
      }
    }
  }
  return mlp3_launch(plan->b, plan->n, plan->f16, plan->nets[0].M, (cudaStream_t)stream);
}

extern "C" void rslrl_mlp3_plan_destroy(void* plan) { delete (rslrl::Mlp3Plan*)plan; }

#ifdef RSLRL_TC_TRACE
extern "C" int rslrl_debug_read_trace_mlp3(unsigned long long* host) {
  return (int)cudaMemcpyFromSymbol(host, rslrl::g_mf_trace, sizeof(unsigned long long) * 148 * 32);
}
#endif
