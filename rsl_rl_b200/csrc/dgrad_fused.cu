// Fused data-gradient chain of the two wide hidden layers of a 3-hidden-layer MLP (fp16-storage path) on sm_100a:
//   dZ2 = (dZ3 W3) * act'(H2)          [M,N3] x [N3,N2]
//   dZ1 = (dZ2 W2) * act'(H1)          [M,N2] x [N2,N1]
// in ONE persistent tcgen05 launch per minibatch for actor and critic (work item = (net, 128-row tile)), instead of two
// launches of the layer-by-layer kernel: the dZ2 tile never comes back from HBM - the epilogue of the first product
// multiplies the TMA-loaded H2 chunk in place by the accumulator, which makes that shared-memory buffer at the same time
// (a) the staging tile of the dZ2 bulk store (the weight gradient of layer 2 needs dZ2) and (b) the K-major A operand
// ("granule", 128 rows x 64 fp16 columns, 128-byte swizzled rows) of the second product.
//
//   warp 0      TMA producer: operand ring of three 32 KB stages: per tile one stage with the whole dZ3 tile (<= 2 k-blocks of
//               16 KB), the W3 k-blocks, then the W2 k-blocks of every 256-column pass over N1 (96 KB of weights in
//               flight: the ring depth, not the MMA, bounds this kernel - shared memory is full)
//   warp 1      MMA issuer / TMEM owner: two 256-column accumulators used alternately by the "jobs" of the stream
//               (first product, then one job per 256-column pass of the second product, tile after tile), so that the
//               MMAs of job k+1 run while the epilogue drains job k
//   warp 2      H loader: H2 chunks straight into the granule buffers, H1 chunks through a ring of four 16 KB buffers
//   warps 3-18  epilogue (16): TMEM -> registers -> x act'(H) in place -> fence -> bulk store; the 64-column chunk c of a
//               job belongs to the four row-quarter warps of group c & 3
// Measured on the way (C2 minibatch, both nets, 24 576 rows; the two layer-by-layer launches it replaces: 18 + 32 us):
// 8 epilogue warps and one accumulator per product 46 us; 16 epilogue warps + accumulators alternating by job 43 us; three
// 32 KB ring stages instead of two 48 KB ones 41 us; second product in passes of 128 columns alternating the epilogue
// group pairs and the halves of the H1 ring 39.8 us (this version); H1 fetched by the epilogue lanes straight from global
// memory instead of the TMA ring 56 us (uncoalesced per-lane rows).  Shared memory is full (96 KB operand ring + 64 KB
// granules + 64 KB H1 ring): the ring depth, not the MMA, bounds the kernel.
// Replaces the autograd of rsl_rl/networks/mlp.py:106-114 for the two middle layers (same arithmetic as two calls of
// tc_gemm_persistent_kernel<dgrad, fp16>: fp16 operands, fp32 accumulation, fp16 results with saturation).
#include <cuda.h>
#include <cuda_fp16.h>

#include "tc_ptx.cuh"

namespace rslrl {

bool tc_make_map16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
                   int box_rows);
int tc_num_sms();

constexpr int DC_EPI_WARPS = 16;
constexpr int DC_FIRST_EPI_WARP = 3;
constexpr int DC_THREADS = (DC_FIRST_EPI_WARP + DC_EPI_WARPS) * 32;
constexpr int DC_OP_STAGES = 3;
constexpr int DC_OP_STAGE_BYTES = 32768;   // one weight k-block (64 reduction rows x 256 columns), or the dZ3 tile
constexpr int DC_GRANULES = 4;             // N2 / 64 <= 4
constexpr int DC_H_SLOTS = 4;
constexpr int DC_CHUNK_BYTES = 16384;      // 128 rows x 64 fp16

struct DcNet {
  CUtensorMap map_dz3, map_w3, map_w2, map_h2, map_h1, map_dz2, map_dz1;
  int M, N1, N2, N3, act;
};

struct alignas(64) DcBatch {
  DcNet net[2];
};

// Optional in-kernel timeline (build with -DRSLRL_DC_TRACE; profiles/experiments/dgrad_chain_trace.py): per CTA 64 slots:
// 0 entry, 1 prologue done, then per tile j < 3 (base 2 + 20 j): +0 S1 operands landed, +1 S1 issued, +2+p pass p issued
// (MMA warp), +8 E1 done, +9+p epilogue of pass p done (epilogue warp 0 / warp 8 for odd passes).
#ifdef RSLRL_DC_TRACE
__device__ unsigned long long g_dc_trace[148 * 64];
__device__ __forceinline__ void dc_trace(int slot) {
  if (slot < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_dc_trace[blockIdx.x * 64 + slot] = t;
  }
}
#define DC_TRACE(cond, slot) do { if (cond) dc_trace(slot); } while (0)
#else
#define DC_TRACE(cond, slot) do { } while (0)
#endif

__device__ __forceinline__ void dc_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(DC_THREADS, 1) dgrad_chain16_kernel(const __grid_constant__ DcBatch batch, const int n_nets) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* const opring = smem_raw + pad;
  uint8_t* const gran = opring + DC_OP_STAGES * DC_OP_STAGE_BYTES;
  uint8_t* const hring = gran + DC_GRANULES * DC_CHUNK_BYTES;
  uint64_t* const op_full = (uint64_t*)(hring + DC_H_SLOTS * DC_CHUNK_BYTES);
  uint64_t* const op_empty = op_full + DC_OP_STAGES;
  uint64_t* const g_ready = op_empty + DC_OP_STAGES;   // granule c holds dZ2 (A operand of the second product)
  uint64_t* const g_free = g_ready + DC_GRANULES;      // granule buffer c may be overwritten (next tile's H2 chunk)
  uint64_t* const h2_full = g_free + DC_GRANULES;      // H2 chunk c landed in granule buffer c
  uint64_t* const hr_full = h2_full + DC_GRANULES;     // H1 ring
  uint64_t* const hr_empty = hr_full + DC_H_SLOTS;
  uint64_t* const acc_full = hr_empty + DC_H_SLOTS;    // [2]: the two alternating accumulators
  uint64_t* const acc_empty = acc_full + 2;            // [2]
  uint32_t* const tmem_ptr_smem = (uint32_t*)(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DC_TRACE(threadIdx.x == 0, 0);
  if (warp == 0 && elect_one()) {
    for (int i = 0; i < n_nets; ++i) {
      const DcNet& P = batch.net[i];
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&P.map_dz3) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&P.map_w3) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&P.map_w2) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&P.map_h2) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&P.map_h1) : "memory");
    }
    for (int s = 0; s < DC_OP_STAGES; ++s) {
      mbar_init(&op_full[s], 1);
      mbar_init(&op_empty[s], 1);
    }
    for (int g = 0; g < DC_GRANULES; ++g) {
      mbar_init(&g_ready[g], 4);   // the four row-quarter warps that produce the granule
      mbar_init(&g_free[g], 5);    // MMA commit after its last use + the four warps' stores have read it
      mbar_init(&h2_full[g], 1);
    }
    for (int s = 0; s < DC_H_SLOTS; ++s) {
      mbar_init(&hr_full[s], 1);
      mbar_init(&hr_empty[s], 4);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], DC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  DC_TRACE(threadIdx.x == 0, 1);

  const int tiles_m = (batch.net[0].M + TC_BLOCK_M - 1) / TC_BLOCK_M;
  const int total = tiles_m * n_nets;
  // shapes are the same for both nets (checked on the host)
  const int N1 = batch.net[0].N1, N2 = batch.net[0].N2, N3 = batch.net[0].N3;
  // The second product runs in passes of 128 output columns = 2 chunks: consecutive passes alternate between the two
  // accumulators AND between the epilogue group pairs (0,1) / (2,3), and their H1 chunks between the two halves of the
  // four-slot ring - so the H1 chunks of pass p+1 are in flight while pass p is being multiplied and stored.
  constexpr int NB2 = 128;                  // columns of the second product per pass
  const int NH = N1 / NB2;                  // passes over the granules
  const int nkb1 = N3 / 64, nkb2 = N2 / 64; // k-blocks of the two products (nkb2 = 2 or 4)
  const int nch1 = N2 / 64;
  constexpr int nch2 = NB2 / 64;

  if (warp == 0) {
    // ===================================================================== TMA producer (operand ring)
    if (elect_one()) {
      int it = 0;
      auto stage = [&](int& s) -> uint8_t* {
        s = it % DC_OP_STAGES;
        mbar_wait(&op_empty[s], ((it / DC_OP_STAGES) & 1) ^ 1);
        ++it;
        return opring + (size_t)s * DC_OP_STAGE_BYTES;
      };
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const DcNet& P = batch.net[t % n_nets];
        const int m0 = (t / n_nets) * TC_BLOCK_M;
        {   // the dZ3 tile: nkb1 <= 2 k-blocks of 128 rows x 64 columns in one stage
          int s;
          uint8_t* dst = stage(s);
          mbar_expect_tx(&op_full[s], (uint32_t)nkb1 * 16384u);
          for (int kb = 0; kb < nkb1; ++kb) tma_load_2d(dst + kb * 16384, &P.map_dz3, &op_full[s], kb * 64, m0);
        }
        for (int kb = 0; kb < nkb1; ++kb) {   // W3 k-block (64 reduction rows x N2 columns)
          int s;
          uint8_t* dst = stage(s);
          mbar_expect_tx(&op_full[s], (uint32_t)N2 * 128u);
          for (int j = 0; j < N2 / 64; ++j) tma_load_2d(dst + j * 8192, &P.map_w3, &op_full[s], j * 64, kb * 64);
        }
        for (int h = 0; h < NH; ++h)
          for (int kb = 0; kb < nkb2; kb += 2) {   // two W2 k-blocks (64 reduction rows x 128 columns, 16 KB each) per stage
            int s;
            uint8_t* dst = stage(s);
            mbar_expect_tx(&op_full[s], 32768u);
            for (int i = 0; i < 2; ++i)
              for (int jj = 0; jj < 2; ++jj)
                tma_load_2d(dst + i * 16384 + jj * 8192, &P.map_w2, &op_full[s], h * NB2 + jj * 64, (kb + i) * 64);
          }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int it = 0, j = 0;
    uint32_t job = 0;   // accumulation jobs so far: job n uses accumulator n & 1 (TMEM columns 256 (n & 1))
    const uint32_t idesc1 = make_idesc_f16(N2, false, true);
    const uint32_t idesc2 = make_idesc_f16(NB2, false, true);   // 128 columns per pass
    auto kblock = [&](uint32_t d_tmem, uint32_t sa, uint32_t sb, uint32_t idesc, bool first) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 4 x K = 16 per 64-element k-block
          const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024, kLayoutSw128);          // K-major A
          const uint64_t db = make_smem_desc(sb + k * 2048, 8192, 1024, kLayoutSw128);      // MN-major B, 64 x 64 boxes
          umma_f16(d_tmem, da, db, idesc, (first && k == 0) ? 0u : 1u);
        }
      }
      __syncwarp();
    };
    auto acquire = [&]() -> uint32_t {   // accumulator of the next job, drained by the epilogue of job - 2
      const uint32_t a = job & 1u;
      mbar_wait(&acc_empty[a], ((job >> 1) & 1u) ^ 1u);
      tcgen05_fence_after();
      return tmem_base + a * 256u;
    };
    auto publish = [&]() {
      if (elect_one()) tcgen05_commit(&acc_full[job & 1u]);
      __syncwarp();
      ++job;
    };
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
      // ---- first product: A = the dZ3 tile (one stage, kept until the last k-block has been issued), B = W3 k-blocks
      {
        const uint32_t d = acquire();
        const int sa_stage = it % DC_OP_STAGES;
        mbar_wait(&op_full[sa_stage], (it / DC_OP_STAGES) & 1);
        ++it;
        DC_TRACE(lane == 0 && j < 3, 2 + 20 * j);
        const uint32_t a_base = smem_u32(opring + (size_t)sa_stage * DC_OP_STAGE_BYTES);
        for (int kb = 0; kb < nkb1; ++kb, ++it) {
          const int s = it % DC_OP_STAGES;
          mbar_wait(&op_full[s], (it / DC_OP_STAGES) & 1);
          tcgen05_fence_after();
          kblock(d, a_base + (uint32_t)kb * 16384u, smem_u32(opring + (size_t)s * DC_OP_STAGE_BYTES), idesc1, kb == 0);
          if (elect_one()) {
            tcgen05_commit(&op_empty[s]);
            if (kb == nkb1 - 1) tcgen05_commit(&op_empty[sa_stage]);
          }
          __syncwarp();
        }
        publish();
        DC_TRACE(lane == 0 && j < 3, 2 + 20 * j + 1);
      }
      // ---- second product, NH passes of 128 columns, A = the granules the first epilogue leaves in shared memory
      for (int h = 0; h < NH; ++h) {
        const uint32_t d = acquire();
        for (int kb = 0; kb < nkb2; kb += 2, ++it) {
          const int s = it % DC_OP_STAGES;
          mbar_wait(&op_full[s], (it / DC_OP_STAGES) & 1);
          const uint32_t base = smem_u32(opring + (size_t)s * DC_OP_STAGE_BYTES);
          for (int i = 0; i < 2; ++i) {
            if (h == 0) mbar_wait(&g_ready[kb + i], j & 1);
            tcgen05_fence_after();
            kblock(d, smem_u32(gran + (size_t)(kb + i) * DC_CHUNK_BYTES), base + (uint32_t)i * 16384u, idesc2, kb + i == 0);
          }
          if (elect_one()) {
            tcgen05_commit(&op_empty[s]);
            if (h == NH - 1) {   // last readers of these two granules in this tile
              tcgen05_commit(&g_free[kb]);
              tcgen05_commit(&g_free[kb + 1]);
            }
          }
          __syncwarp();
        }
        publish();
        DC_TRACE(lane == 0 && j < 3 && h < 6, 2 + 20 * j + 2 + h);
      }
    }
  } else if (warp == 2) {
    // ===================================================================== H loader
    if (elect_one()) {
      int j = 0, g = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
        const DcNet& P = batch.net[t % n_nets];
        const int m0 = (t / n_nets) * TC_BLOCK_M;
        for (int c = 0; c < nch1; ++c) {   // H2 chunk c -> granule buffer c (multiplied in place by the first epilogue)
          mbar_wait(&g_free[c], (j & 1) ^ 1);
          mbar_expect_tx(&h2_full[c], (uint32_t)DC_CHUNK_BYTES);
          tma_load_2d(gran + (size_t)c * DC_CHUNK_BYTES, &P.map_h2, &h2_full[c], c * 64, m0);
        }
        for (int h = 0; h < NH; ++h)
          for (int c = 0; c < nch2; ++c, ++g) {   // H1 chunks of this pass through the ring
            const int s = g % DC_H_SLOTS;
            mbar_wait(&hr_empty[s], ((uint32_t)(g / DC_H_SLOTS) & 1u) ^ 1u);
            mbar_expect_tx(&hr_full[s], (uint32_t)DC_CHUNK_BYTES);
            tma_load_2d(hring + (size_t)s * DC_CHUNK_BYTES, &P.map_h1, &hr_full[s], h * NB2 + c * 64, m0);
          }
      }
    }
  } else {
    // ===================================================================== epilogue warps
    const int ew = warp - DC_FIRST_EPI_WARP;
    const int q = warp & 3;      // TMEM sub-partition = row quarter of the tile
    const int grp = ew >> 2;     // chunk group 0..3: this warp takes the chunks c with (c & 3) == grp
    const uint32_t sw = (uint32_t)(lane & 7);
    int j = 0, g = 0;
    uint32_t job = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
      const DcNet& P = batch.net[t % n_nets];
      const int m0 = (t / n_nets) * TC_BLOCK_M;
      const int row0 = m0 + q * 32;
      const bool rows_live = row0 < P.M;
      const int act = P.act;

      // one job: wait for its accumulator; this warp's chunk (if any): TMEM columns [64 c, +64) of its 32 rows times
      // act'(H) in place in the buffer that holds the H chunk, bulk store; the accumulator is handed back right after the
      // TMEM read.  granule: first product (H2 in granule buffer c, result = A operand of the second product).
      auto run_job = [&](bool granule, int nch, int first_grp, const CUtensorMap* map_out, int col_base) {
        const uint32_t a = job & 1u;
        mbar_wait(&acc_full[a], (job >> 1) & 1u);
        tcgen05_fence_after();
        ++job;
        const int c = grp - first_grp;   // the job's chunks 0 .. nch-1 belong to the groups first_grp .. first_grp + nch-1
        if (c < 0 || c >= nch) {   // no chunk for this warp in this job: it still owes its arrival
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) dc_arrive(&acc_empty[a]);
          return;
        }
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + a * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 64);
        TMEM_LD_32x32(taddr, v0);
        TMEM_LD_32x32(taddr + 32u, v1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) dc_arrive(&acc_empty[a]);
        const int slot = (g + c) % DC_H_SLOTS;
        uint8_t* const buf = granule ? gran + (size_t)c * DC_CHUNK_BYTES : hring + (size_t)slot * DC_CHUNK_BYTES;
        uint8_t* const blk = buf + q * 4096;
        uint8_t* const my_row = blk + lane * 128;
        if (granule) mbar_wait(&h2_full[c], (uint32_t)(j & 1));
        else mbar_wait(&hr_full[slot], (uint32_t)((g + c) / DC_H_SLOTS) & 1u);
        auto scale_chunk = [&](auto grad) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            uint4* const hp = (uint4*)(my_row + ((u ^ sw) << 4));
            const uint4 h = *hp;
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
            uint32_t ow[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 hf = unpack_f16x2(hw[e]);
              const float xa = __uint_as_float(u < 4 ? v0[(8 * u + 2 * e) & 31] : v1[(8 * u + 2 * e) & 31]);
              const float xb = __uint_as_float(u < 4 ? v0[(8 * u + 2 * e + 1) & 31] : v1[(8 * u + 2 * e + 1) & 31]);
              ow[e] = pack_f16x2(xa * grad(hf.x), xb * grad(hf.y));
            }
            *hp = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          }
        };
        if (act == RSLRL_ACT_ELU) scale_chunk([](float hv) { return fminf(hv, 0.f) + 1.f; });
        else scale_chunk([act](float hv) { return act_grad_from_output(act, hv); });
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (granule) dc_arrive(&g_ready[c]);   // the granule may be consumed by the second product
          if (rows_live) tma_store_2d(map_out, blk, col_base + c * 64, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          // (one chunk per warp and job: nothing else to do until the next accumulator is ready, so the buffer is given
          //  back as soon as the store has read it)
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          dc_arrive(granule ? &g_free[c] : &hr_empty[slot]);
        }
        __syncwarp();
      };

      // ---- first epilogue: accumulator x act'(H2) -> dZ2 granules (+ bulk store of dZ2)
      run_job(true, nch1, 0, &P.map_dz2, 0);
      DC_TRACE(ew == 0 && lane == 0 && j < 3, 2 + 20 * j + 8);
      // ---- second epilogue per pass: accumulator x act'(H1) -> dZ1 (bulk store); passes alternate between the group pairs
      for (int h = 0; h < NH; ++h) {
        run_job(false, nch2, 2 * (h & 1), &P.map_dz1, h * NB2);
        g += nch2;
        DC_TRACE(ew == 8 * (h & 1) && lane == 0 && j < 3 && h < 6, 2 + 20 * j + 9 + h);
      }
    }
    tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static bool dc_ok16(const void* p, long long ld) { return ((uintptr_t)p & 15) == 0 && (ld % 8) == 0; }

}  // namespace rslrl

extern "C" int rslrl_dgrad_chain16(int n, const rslrl_dgrad_chain_t* nets_host, rslrl_stream_t stream) {
  using namespace rslrl;
  const char* what = "rslrl_dgrad_chain16";
  RSLRL_CHECK_ARG(n >= 1 && n <= 2 && nets_host, what);
  DcBatch b{};
  for (int i = 0; i < n; ++i) {
    const rslrl_dgrad_chain_t& q = nets_host[i];
    RSLRL_CHECK_ARG(q.dz3 && q.w3 && q.w2 && q.h2 && q.h1 && q.dz2 && q.dz1 && q.M > 0, what);
    const bool shapes = q.N3 % 64 == 0 && q.N3 >= 64 && q.N3 <= 128 && q.N2 % 128 == 0 && q.N2 <= 256 && q.N1 % 128 == 0 &&
                        q.N1 <= 512 && q.M == nets_host[0].M && q.N1 == nets_host[0].N1 &&
                        q.N2 == nets_host[0].N2 && q.N3 == nets_host[0].N3;
    const bool aligned = dc_ok16(q.dz3, q.ld_dz3) && dc_ok16(q.w3, q.N2) && dc_ok16(q.w2, q.N1) && dc_ok16(q.h2, q.ldh2) &&
                         dc_ok16(q.h1, q.ldh1) && dc_ok16(q.dz2, q.ld_dz2) && dc_ok16(q.dz1, q.ld_dz1);
    if (!shapes || !aligned) {
      set_error("%s: shape not eligible (needs N3 in {64, 128}, N2 in {128, 256}, N1 in {128, 256, 512}, equal shapes "
                "for all nets, 16-byte aligned fp16 rows)", what);
      return 4;
    }
    DcNet& P = b.net[i];
    const bool ok = tc_make_map16(&P.map_dz3, q.dz3, q.M, q.N3, q.ld_dz3, 64, TC_BLOCK_M) &&
                    tc_make_map16(&P.map_w3, q.w3, q.N3, q.N2, q.N2, 64, 64) &&
                    tc_make_map16(&P.map_w2, q.w2, q.N2, q.N1, q.N1, 64, 64) &&
                    tc_make_map16(&P.map_h2, q.h2, q.M, q.N2, q.ldh2, 64, TC_BLOCK_M) &&
                    tc_make_map16(&P.map_h1, q.h1, q.M, q.N1, q.ldh1, 64, TC_BLOCK_M) &&
                    tc_make_map16(&P.map_dz2, q.dz2, q.M, q.N2, q.ld_dz2, 64, 32) &&
                    tc_make_map16(&P.map_dz1, q.dz1, q.M, q.N1, q.ld_dz1, 64, 32);
    if (!ok) {
      set_error("%s: cuTensorMapEncodeTiled failed", what);
      return 1;
    }
    P.M = q.M; P.N1 = q.N1; P.N2 = q.N2; P.N3 = q.N3; P.act = q.act;
  }
  const size_t smem = (size_t)DC_OP_STAGES * DC_OP_STAGE_BYTES + (size_t)(DC_GRANULES + DC_H_SLOTS) * DC_CHUNK_BYTES +
                      (2 * DC_OP_STAGES + 3 * DC_GRANULES + 2 * DC_H_SLOTS + 4) * 8 + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(dgrad_chain16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute failed", what);
      (void)cudaGetLastError();
      return 1;
    }
    attr_set = true;
  }
  const int tiles = (nets_host[0].M + TC_BLOCK_M - 1) / TC_BLOCK_M * n;
  const int grid = tiles < tc_num_sms() ? tiles : tc_num_sms();
  launch_chain_kernel(dgrad_chain16_kernel, dim3(grid), dim3(DC_THREADS), smem, (cudaStream_t)stream, b, n);
  return finish_launch(what);
}

#ifdef RSLRL_DC_TRACE
extern "C" int rslrl_debug_read_trace_dgrad_chain(unsigned long long* host) {
  return (int)cudaMemcpyFromSymbol(host, rslrl::g_dc_trace, sizeof(unsigned long long) * 148 * 64);
}
#endif
