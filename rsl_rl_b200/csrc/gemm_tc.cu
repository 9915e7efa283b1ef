// TF32 tensor-core GEMMs for nn.Linear forward / data-gradient / weight-gradient on sm_100a:
// tcgen05.mma (kind::tf32) with the accumulator in TMEM, operands staged in shared memory by TMA
// (cp.async.bulk.tensor, 128-byte swizzle), mbarrier producer/consumer pipeline, warp-specialised CTA.
//
//   warp 0      TMA producer  (one elected lane issues the bulk tensor copies of a k-block)
//   warp 1      MMA issuer    (allocates TMEM, one elected lane issues 4 x tcgen05.mma per k-block,
//                              tcgen05.commit releases the smem stage / publishes the accumulator)
//   warps 2..5  epilogue      (tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue ->
//                              smem transpose -> 128-byte coalesced global stores / vector atomics)
//
// One kernel covers the three products through the operand "major" bits of the instruction descriptor
// and matching shared-memory descriptors, so no transposed copies of weights or activations exist:
//   forward  Y  = act(X W^T + b)     A = X  [M,K]  K-major     B = W [N,K]  K-major
//   dgrad    dX = (dY W) * act'(H)   A = dY [M,N]  K-major     B = W [N,K]  MN-major (N is the reduction)
//   wgrad    dW += dY^T X            A = dY [M,N]  MN-major    B = X [M,K]  MN-major (M is the reduction, split-K)
// K-major tiles use SWIZZLE_128B (rows of 32 tf32), MN-major tf32 tiles need the 32-byte-atom flavour
// (UMMA layout SWIZZLE_128B_BASE32B <-> TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
// Out-of-range rows/columns (M, N or K tails, N = 12 or 1 heads) are zero-filled by TMA and masked in the
// epilogue, so every shape of the MLP goes through the same tiles.  Operands whose row pitch is not a
// multiple of 16 bytes cannot be described to TMA; those problems are reported back to the caller
// (linear.cu runs them on the fp32 CUDA-core kernel, which is strictly more accurate).
//
// Inputs are fp32 in HBM; the tensor core consumes them as tf32.  The bias-gradient of a layer is the
// column sum of the dY it receives, so it is accumulated by the kernel that PRODUCES that dY (the dgrad
// epilogue here, the loss kernel for the output layers) instead of re-reading dY.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rslrl {

enum { TC_FWD = 0, TC_DGRAD = 1, TC_WGRAD = 2 };

struct TcProblem {
  CUtensorMap map_a, map_b;
  CUtensorMap map_c;  // output, box {32 cols, 32 rows} (store, or reduce-add for split-K wgrad)
  CUtensorMap map_h;  // dgrad: saved activation H, same geometry as C
  float* C;
  const float* bias;  // fwd
  float* db;          // wgrad: += sum over the batch of dY (bias gradient); may be null
  int M, N, K;        // C is [M,N], reduction length K
  int ldc, act, tma_out;
  float out_scale;    // wgrad: dW, db += out_scale * (...)  (undoes the power-of-two gradient scaling of the fp16 path)
  const float* out_scale_dev;  // optional device-resident factor multiplied into out_scale (dynamic gradient scale)
  int block_n, tiles_m, tiles_n, splits, kb_per_split, num_kb;
};

constexpr int TC_MAX_PROBLEMS = 8;  // 8 x 640 B of kernel parameters (CUDA >= 12.1 allows up to 32 KB)

struct alignas(64) TcBatch {
  TcProblem p[TC_MAX_PROBLEMS];
};

// ------------------------------------------------------------------------------- persistent kernel
// One CTA per SM loops over the tiles of all grouped problems (static round-robin schedule): the prologue
// (barrier init, TMEM allocation, descriptor prefetch) is paid once, the TMA producer runs ahead into the
// next tile while the current one is still in its epilogue, and the accumulator is double buffered in TMEM
// (2 x 256 columns) so that the epilogue of tile i overlaps the MMAs of tile i+1.  Tiles are 128 x block_n
// with block_n up to 256.
//   warp 0       TMA producer (operands)     warp 1   MMA issuer / TMEM owner      warps 2..9  epilogue (8)
//   warps 10..13 wgrad: bias-gradient reducers over the dY tiles; dgrad: warp 10 streams the saved
//                activations H through a ring of 8 chunk buffers (128 rows x 32 columns each) ahead of the
//                epilogue, which multiplies in place and bulk-stores from the same buffer.
constexpr int TCP_EPI_WARPS = 8;
constexpr int TCP_THREADS = (2 + TCP_EPI_WARPS + 4) * 32;
constexpr int TCP_MAX_STAGES = 6;
constexpr int TCP_RING_BUDGET = 150 * 1024;
constexpr int TCP_STAGING = TCP_EPI_WARPS * 2 * 4096;  // fwd / wgrad: two 32x32 fp32 slots per epilogue warp
constexpr int TCP_H_CHUNKS = 8;                        // dgrad: ring of H chunk buffers, 16 KB each
constexpr int TCP_H_BYTES = TCP_H_CHUNKS * 16384;
constexpr int TCP_RING_BUDGET_DGRAD = 96 * 1024;
// (measured: a single staging slot per epilogue warp + a 4-stage 192 KB ring is slower than 2 slots + 3 stages -- the
// serialised epilogue, 8 us instead of 3 us per 128x256 tile, costs more than the deeper ring gains)
// (3xTF32: the mainloop is three times longer per tile, so one staging slot is enough and its 32 KB buy a third ring stage)
__host__ __device__ constexpr int tcp_staging_slots(int mode, bool f16, bool x3 = false) { return x3 ? 1 : 2; }
__host__ __device__ constexpr int tcp_staging_bytes(int mode, bool f16, bool x3 = false) {
  return mode == 1 ? (x3 ? 4 * 16384 : TCP_H_BYTES) : TCP_EPI_WARPS * tcp_staging_slots(mode, f16, x3) * 4096;
}
__host__ __device__ constexpr int tcp_ring_budget(int mode, bool f16, bool x3 = false) {
  return mode == 1 ? (x3 ? TCP_RING_BUDGET : TCP_RING_BUDGET_DGRAD) : (x3 ? 192 * 1024 : TCP_RING_BUDGET);
}
constexpr int TCP_MAX_BLOCK_N = 256;
// X3 = "3xTF32": fp32-class products on the tf32 tensor cores.  The operands arrive as raw fp32 (FLOAT32 tensor maps:
// the MMA truncates them to tf32, hi = x & 0xffffe000); four extra converter warps write lo = x - hi (exact in fp32) of
// every staged k-block into a second copy of the stage, and every k-step issues three MMAs into the same accumulator:
// A_lo B_hi + A_hi B_lo + A_hi B_hi (the dropped lo x lo term and the truncation of lo are ~2^-21 relative).
constexpr int TCP_X3_WARPS = 4;
constexpr int TCP_X3_BLOCK_N = 128;                    // stage = 2 x (16 KB + 16 KB): two stages in the ring
__host__ __device__ constexpr int tcp_threads(bool x3) { return TCP_THREADS + (x3 ? TCP_X3_WARPS * 32 : 0); }
__host__ __device__ constexpr int tcp_h_chunks(bool x3) { return x3 ? 4 : TCP_H_CHUNKS; }

struct TcpScalars {  // per-problem scalars mirrored into shared memory once per CTA
  float* C;
  const float* bias;
  float* db;
  float out_scale;
  int M, N, ldc, act, tma_out, block_n, tiles_m, tiles_n, splits, kb_per_split, num_kb, tile_count;
};

// Optional in-kernel timeline (build with -DRSLRL_TC_TRACE, read with profiles/experiments/gemm_trace.py):
// per CTA 32 globaltimer slots: 0 entry, 1 prologue done, 2 first TMA, 3 first full barrier, 4+2j mainloop of the
// CTA's j-th tile issued, 5+2j its epilogue issued (epilogue warp 0), 31 exit.
#ifdef RSLRL_TC_TRACE
__device__ unsigned long long g_tc_trace[148 * 32];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(slot) do { if ((slot) < 31 || (slot) == 31) g_tc_trace[blockIdx.x * 32 + ((slot) < 31 ? (slot) : 31)] = gtime(); } while (0)
#define TRACE_TILE(slot) do { if ((slot) < 31) g_tc_trace[blockIdx.x * 32 + (slot)] = gtime(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#define TRACE_TILE(slot) do { } while (0)
#endif

// F16 = operands (and the dgrad's H input / dX output) are fp16 in HBM: kind::f16 MMAs, k-blocks of 64 elements (the
// same 128-byte rows), MN-major operands in the plain 128-byte swizzle (64-element atoms, 8 KB TMA boxes).
template <int MODE, bool F16, bool X3 = false>
__global__ void __launch_bounds__(tcp_threads(X3), 1) tc_gemm_persistent_kernel(const __grid_constant__ TcBatch batch,
                                                                                const int n_problems) {
  static_assert(!(F16 && X3), "3xTF32 is an fp32-operand mode");
  constexpr int HCH = tcp_h_chunks(X3);        // dgrad: H chunk buffers in the ring
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr bool A_MN = MODE == TC_WGRAD;
  constexpr bool B_MN = MODE != TC_FWD;
  constexpr int BK = F16 ? 64 : 32;            // elements per k-block (one 128-byte row)
  constexpr int MNB = F16 ? 64 : 32;           // MN-major operands: elements per TMA box along MN
  constexpr uint32_t MN_BOX = F16 ? 8192u : 4096u;    // bytes per MN-major box (MNB x BK elements)
  constexpr uint32_t MN_KADV = F16 ? 2048u : 1024u;   // MN-major: bytes per MMA k-step (16 / 8 rows of 128 bytes)
  constexpr uint32_t MN_SBO = F16 ? 1024u : 512u;
  constexpr uint32_t MN_LAYOUT = F16 ? kLayoutSw128 : kLayoutSw128Base32;
  constexpr int CH = (MODE == TC_DGRAD && F16) ? 64 : 32;  // epilogue chunk width in output columns

  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* const ring = smem_raw + pad;
  // the ring is sized for the widest problem of the batch; every problem uses the same stage pitch
  int max_bn = 16;
  for (int i = 0; i < n_problems; ++i) max_bn = max(max_bn, batch.p[i].block_n);
  const uint32_t a_bytes = TC_BLOCK_M * TC_BLOCK_K * 4;
  const uint32_t raw_bytes = a_bytes + (uint32_t)max_bn * TC_BLOCK_K * 4;   // operands as TMA delivers them
  const uint32_t stage_bytes = X3 ? 2u * raw_bytes : raw_bytes;             // X3: + the lo = x - tf32(x) copies behind them
  int stages = tcp_ring_budget(MODE, F16, X3) / stage_bytes;
  stages = stages > TCP_MAX_STAGES ? TCP_MAX_STAGES : stages;
  uint8_t* const staging = ring + (size_t)stages * stage_bytes;  // 1024-byte aligned (stage_bytes is a multiple of 2048)
  uint64_t* const full_bar = (uint64_t*)(staging + tcp_staging_bytes(MODE, F16, X3));
  uint64_t* const empty_bar = full_bar + TCP_MAX_STAGES;
  uint64_t* const conv_bar = empty_bar + TCP_MAX_STAGES;    // X3: lo copies of the stage written
  uint64_t* const tmem_full = conv_bar + TCP_MAX_STAGES;    // [2]
  uint64_t* const tmem_empty = tmem_full + 2;               // [2]
  uint64_t* const h_full = tmem_empty + 2;                  // [TCP_H_CHUNKS]  (dgrad)
  uint64_t* const h_empty = h_full + TCP_H_CHUNKS;          // [TCP_H_CHUNKS]
  uint32_t* const tmem_ptr_smem = (uint32_t*)(h_empty + TCP_H_CHUNKS);
  TcpScalars* const sp = (TcpScalars*)(tmem_ptr_smem + 4);  // [TC_MAX_PROBLEMS]
  float* const bias_stage = (float*)(sp + TC_MAX_PROBLEMS);               // [TCP_EPI_WARPS][32], 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TRACE(0);
  pdl_trigger();

  if (threadIdx.x < n_problems) {
    const TcProblem& P = batch.p[threadIdx.x];
    TcpScalars q;
    q.C = P.C; q.bias = P.bias; q.db = P.db; q.out_scale = P.out_scale; q.M = P.M; q.N = P.N; q.ldc = P.ldc; q.act = P.act;
    q.tma_out = P.tma_out; q.block_n = P.block_n; q.tiles_m = P.tiles_m; q.tiles_n = P.tiles_n; q.splits = P.splits;
    q.kb_per_split = P.kb_per_split; q.num_kb = P.num_kb; q.tile_count = P.tiles_m * P.tiles_n * P.splits;
    sp[threadIdx.x] = q;
  }
  if (warp == 0 && elect_one()) {
    for (int i = 0; i < n_problems; ++i) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&batch.p[i].map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&batch.p[i].map_b) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&batch.p[i].map_c) : "memory");
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MODE == TC_WGRAD ? 5 : 1);  // MMA commit (+ the 4 reducer warps in wgrad)
      mbar_init(&conv_bar[s], TCP_X3_WARPS);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], TCP_EPI_WARPS);
    }
    for (int w = 0; w < HCH; ++w) {
      mbar_init(&h_full[w], 1);
      mbar_init(&h_empty[w], 4);  // the four row-group warps that consume a chunk
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();  // the prologue above touched no global memory; the operands below come from the previous kernels
  if (MODE == TC_WGRAD && threadIdx.x < n_problems && batch.p[threadIdx.x].out_scale_dev != nullptr)
    sp[threadIdx.x].out_scale *= __ldcg(batch.p[threadIdx.x].out_scale_dev);  // 1 / (dynamic scale of this minibatch)
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) TRACE(1);

  int total_tiles = 0;
  for (int i = 0; i < n_problems; ++i) total_tiles += sp[i].tile_count;

  // tile id -> (problem, split, tile_m, tile_n); tile_n fastest so that concurrently running CTAs share A rows
  auto decode = [&](int t, int& pi, int& split, int& tm, int& tn) {
    pi = 0;
    while (t >= sp[pi].tile_count) { t -= sp[pi].tile_count; ++pi; }
    const int tiles = sp[pi].tiles_m * sp[pi].tiles_n;
    split = t / tiles;
    const int rem = t - split * tiles;
    tm = rem / sp[pi].tiles_n;
    tn = rem - tm * sp[pi].tiles_n;
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      int it = 0;  // running k-block counter over all tiles of this CTA
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int pi, split, tm, tn;
        decode(t, pi, split, tm, tn);
        const CUtensorMap* map_a = &batch.p[pi].map_a;
        const CUtensorMap* map_b = &batch.p[pi].map_b;
        const int block_n = sp[pi].block_n;
        const int m0 = tm * TC_BLOCK_M, n0 = tn * block_n;
        const int kb0 = split * sp[pi].kb_per_split;
        const int kb1 = min(sp[pi].num_kb, kb0 + sp[pi].kb_per_split);
        const uint32_t tx_bytes = a_bytes + (uint32_t)block_n * TC_BLOCK_K * 4;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % stages;
          mbar_wait(&empty_bar[s], ((it / stages) & 1) ^ 1);
          uint8_t* sa = ring + (size_t)s * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          if (it == 0) TRACE(2);
          mbar_expect_tx(&full_bar[s], tx_bytes);
          const int k0 = kb * BK;
          if (!A_MN) {
            tma_load_2d(sa, map_a, &full_bar[s], k0, m0);
          } else {
            for (int j = 0; j < TC_BLOCK_M / MNB; ++j) tma_load_2d(sa + j * MN_BOX, map_a, &full_bar[s], m0 + j * MNB, k0);
          }
          if (!B_MN) {
            tma_load_2d(sb, map_b, &full_bar[s], k0, n0);
          } else {
            for (int j = 0; j < block_n / MNB; ++j) tma_load_2d(sb + j * MN_BOX, map_b, &full_bar[s], n0 + j * MNB, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int it = 0, j = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
      int pi, split, tm, tn;
      decode(t, pi, split, tm, tn);
      const int block_n = sp[pi].block_n;
      const int kb0 = split * sp[pi].kb_per_split;
      const int kb1 = min(sp[pi].num_kb, kb0 + sp[pi].kb_per_split);
      const uint32_t idesc = F16 ? make_idesc_f16(block_n, A_MN, B_MN) : make_idesc(block_n, A_MN, B_MN);
      const int as = j & 1;
      const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
      mbar_wait(&tmem_empty[as], ((j >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
      tcgen05_fence_after();
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % stages;
        mbar_wait(X3 ? &conv_bar[s] : &full_bar[s], (it / stages) & 1);
        if (it == 0 && lane == 0) TRACE(3);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
            const uint64_t da = A_MN ? make_smem_desc(sa + k * MN_KADV, MN_BOX, MN_SBO, MN_LAYOUT)
                                     : make_smem_desc(sa + k * 32, 16, 1024, kLayoutSw128);
            const uint64_t db = B_MN ? make_smem_desc(sb + k * MN_KADV, MN_BOX, MN_SBO, MN_LAYOUT)
                                     : make_smem_desc(sb + k * 32, 16, 1024, kLayoutSw128);
            if constexpr (X3) {
              // small terms first: A_lo B_hi + A_hi B_lo, then the leading product
              const uint64_t da_lo = A_MN ? make_smem_desc(sa + raw_bytes + k * MN_KADV, MN_BOX, MN_SBO, MN_LAYOUT)
                                          : make_smem_desc(sa + raw_bytes + k * 32, 16, 1024, kLayoutSw128);
              const uint64_t db_lo = B_MN ? make_smem_desc(sb + raw_bytes + k * MN_KADV, MN_BOX, MN_SBO, MN_LAYOUT)
                                          : make_smem_desc(sb + raw_bytes + k * 32, 16, 1024, kLayoutSw128);
              // The correction terms go to their OWN accumulator (columns 128.. of this TMEM buffer; tiles are at most 128
              // wide in this mode): the tensor core truncates when it adds into an fp32 accumulator, a coherent bias that
              // grows with the number of accumulations -- the main chain now sees one addition per k-step instead of three,
              // and the epilogue adds the two accumulators with one rounding.
              const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
              umma_tf32(d_tmem + 128u, da_lo, db, idesc, acc);
              umma_tf32(d_tmem + 128u, da, db_lo, idesc, 1u);
              umma_tf32(d_tmem, da, db, idesc, acc);
            } else if (F16) {
              umma_f16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            } else {
              umma_tf32(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
        }
        __syncwarp();
        if (elect_one()) {
          tcgen05_commit(&empty_bar[s]);
          if (kb == kb1 - 1) tcgen05_commit(&tmem_full[as]);
        }
        __syncwarp();
      }
      if (lane == 0) TRACE_TILE(4 + 2 * j);
    }
  } else if (warp < 2 + TCP_EPI_WARPS) {
    // ===================================================================== epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;      // TMEM sub-partition
    const int half = ew >> 2;    // chunk parity handled by this warp
    constexpr int SLOTS = tcp_staging_slots(MODE, F16, X3);
    uint8_t* const slot_base = staging + (size_t)ew * SLOTS * 4096;   // fwd / wgrad staging slots of this warp
    const uint32_t sw = (uint32_t)(lane & 7);
    int j = 0;
    uint32_t n_issued = 0;       // chunks processed by this warp so far (selects the staging slot)
    int gbase = 0;               // dgrad: running chunk index of this CTA (H ring position)
    int prev_h_slot = -1;        // dgrad: H buffer whose store is still being read
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
      int pi, split, tm, tn;
      decode(t, pi, split, tm, tn);
      const TcpScalars P = sp[pi];
      const CUtensorMap* map_c = &batch.p[pi].map_c;
      const int block_n = P.block_n;
      const int m0 = tm * TC_BLOCK_M, n0 = tn * block_n;
      const int nchunks = (block_n + CH - 1) / CH;
      const int row0 = m0 + q * 32;
      const int as = j & 1;
      const bool rows_live = row0 < P.M;

      mbar_wait(&tmem_full[as], (j >> 1) & 1);
      tcgen05_fence_after();

      for (int c = half; c < nchunks; c += 2) {
        if constexpr (MODE == TC_DGRAD && F16) {
          // 64 output columns per chunk: the fp16 H chunk (128 rows x 128 bytes, streamed in by warp 10) is multiplied
          // in place by the accumulator and becomes the staging tile of the fp16 dX bulk store
          const int c0 = c * 64;
          uint32_t v0[32], v1[32];
          const uint32_t taddr = tmem_base + (uint32_t)as * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
          TMEM_LD_32x32(taddr, v0);
          TMEM_LD_32x32(taddr + 32u, v1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c + 2 >= nchunks) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[as])) : "memory");
          }
          const int g = gbase + c;
          const int hs = g % HCH;
          uint8_t* const blk = staging + (size_t)hs * 16384 + q * 4096;
          uint8_t* const my_row = blk + lane * 128;
          mbar_wait(&h_full[hs], (uint32_t)(g / HCH) & 1u);
          // branch-free inner loop per activation (the act code is uniform: hoisted out of the 64-element loop)
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
          if (P.act == RSLRL_ACT_ELU) scale_chunk([](float hv) { return fminf(hv, 0.f) + 1.f; });
          else scale_chunk([act = P.act](float hv) { return act_grad_from_output(act, hv); });
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (rows_live) tma_store_2d(map_c, blk, n0 + c0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (prev_h_slot >= 0) {
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&h_empty[prev_h_slot])) : "memory");
            }
          }
          prev_h_slot = hs;
          continue;
        }
        const int c0 = c * 32;
        const int ncols = min(32, block_n - c0);
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (uint32_t)as * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        if (ncols == 32) TMEM_LD_32x32(taddr, v);
        else TMEM_LD_32x16(taddr, v);
        uint32_t w[X3 ? 32 : 1];   // 3xTF32: the accumulator of the correction terms A_lo B_hi + A_hi B_lo (columns 128..)
        if constexpr (X3) {
          if (ncols == 32) TMEM_LD_32x32(taddr + 128u, w);
          else TMEM_LD_32x16(taddr + 128u, w);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c + 2 >= nchunks) {  // last TMEM read of this tile by this warp: hand the accumulator back
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[as])) : "memory");
        }
        float x[32];
        if (ncols == 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = (i < 16) ? __uint_as_float(v[i]) : 0.f;
        }
        if constexpr (X3) {   // main + correction accumulator: one rounding
          if (ncols == 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] += __uint_as_float(w[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] += __uint_as_float(w[i]);
          }
        }

        if (MODE == TC_DGRAD) {
          // H chunk (128 rows x 32 cols, 128-byte swizzle) streamed in by warp 10; this warp owns rows [32q, 32q+32)
          const int g = gbase + c;
          const int hs = g % HCH;
          uint8_t* const blk = staging + (size_t)hs * 16384 + q * 4096;
          uint8_t* const my_row = blk + lane * 128;
          mbar_wait(&h_full[hs], (uint32_t)(g / HCH) & 1u);
          auto scale_chunk = [&](auto grad) {  // act code is uniform: one branch per chunk, none per element
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float4* const hp = (float4*)(my_row + ((u ^ sw) << 4));
              const float4 h = *hp;
              float4 o;
              o.x = x[4 * u + 0] * grad(h.x);
              o.y = x[4 * u + 1] * grad(h.y);
              o.z = x[4 * u + 2] * grad(h.z);
              o.w = x[4 * u + 3] * grad(h.w);
              *hp = o;  // in place: the buffer becomes the staging tile of the bulk store
            }
          };
          // ELU'(z) from the output: 1 for h > 0, h + 1 otherwise = min(h,0) + 1
          if (P.act == RSLRL_ACT_ELU) scale_chunk([](float hv) { return fminf(hv, 0.f) + 1.f; });
          else scale_chunk([act = P.act](float hv) { return act_grad_from_output(act, hv); });
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (rows_live) tma_store_2d(map_c, blk, n0 + c0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (prev_h_slot >= 0) {  // the previous chunk's store has read its buffer: give it back to the H producer
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&h_empty[prev_h_slot])) : "memory");
            }
          }
          prev_h_slot = hs;
          continue;
        }

        const uint32_t slot = SLOTS == 2 ? (n_issued & 1) : 0;
        uint8_t* const slot_ptr = slot_base + slot * 4096;
        uint8_t* const my_row = slot_ptr + lane * 128;
        if (MODE == TC_FWD) {
          // bias of this 32-column chunk: one coalesced load, staged in 128 B of warp-private smem, read back
          // as 8 broadcast 16-byte loads (instead of 32 shuffles)
          float* const bs = bias_stage + ew * 32;
          bs[lane] = (P.bias != nullptr && n0 + c0 + lane < P.N && lane < ncols) ? __ldg(P.bias + n0 + c0 + lane) : 0.f;
          __syncwarp();
          float b[32];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 b4 = *(const float4*)(bs + 4 * u);
            b[4 * u] = b4.x; b[4 * u + 1] = b4.y; b[4 * u + 2] = b4.z; b[4 * u + 3] = b4.w;
          }
          __syncwarp();
          if (P.act == RSLRL_ACT_ELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = elu_fast(x[i] + b[i]);
          } else if (P.act == RSLRL_ACT_IDENTITY) {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = x[i] + b[i];
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = act_forward(P.act, x[i] + b[i]);
          }
        }

        if (MODE == TC_WGRAD && P.out_scale != 1.f) {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] *= P.out_scale;
        }
        if (P.tma_out) {
          if (lane == 0) {  // the slot's previous store (SLOTS chunks ago) must have been read
            if (SLOTS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *(float4*)(my_row + ((u ^ sw) << 4)) = make_float4(x[4 * u], x[4 * u + 1], x[4 * u + 2], x[4 * u + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (rows_live) {
              if (MODE == TC_WGRAD) tma_reduce_add_2d(map_c, slot_ptr, n0 + c0, row0);
              else tma_store_2d(map_c, slot_ptr, n0 + c0, row0);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else {
          const int row = row0 + lane;
          if (row < P.M) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = n0 + c0 + i;
              if (i < ncols && col < P.N) {
                if (MODE == TC_WGRAD) atomicAdd(P.C + (size_t)row * P.ldc + col, x[i]);
                else P.C[(size_t)row * P.ldc + col] = x[i];
              }
            }
          }
        }
        ++n_issued;
      }
      if (half >= nchunks) {  // this warp has no chunk in this tile but still owes its arrival
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[as])) : "memory");
      }
      gbase += nchunks;
      if (ew == 0 && lane == 0) TRACE_TILE(5 + 2 * j);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    tcgen05_fence_before();
  } else if (X3 && warp >= TCP_THREADS / 32) {
    // ===================================================================== 3xTF32 converter warps
    // lo = x - (x & 0xffffe000) for every 16-byte chunk of the staged A and B tiles (element-wise: the swizzled
    // layouts of the raw tiles and of their lo copies are identical)
    const int ct = (warp - TCP_THREADS / 32) * 32 + lane;   // 0 .. 127
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int pi, split, tm, tn;
      decode(t, pi, split, tm, tn);
      const int kb0 = split * sp[pi].kb_per_split;
      const int kb1 = min(sp[pi].num_kb, kb0 + sp[pi].kb_per_split);
      const int chunks = (int)(a_bytes >> 4) + sp[pi].block_n * 8;   // 16-byte chunks of A (16 KB) + B (block_n x 128 B)
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % stages;
        mbar_wait(&full_bar[s], (it / stages) & 1);
        const uint4* const src = (const uint4*)(ring + (size_t)s * stage_bytes);
        uint4* const dst = (uint4*)(ring + (size_t)s * stage_bytes + raw_bytes);
#pragma unroll 4
        for (int c = ct; c < chunks; c += TCP_X3_WARPS * 32) {
          const uint4 v = src[c];
          uint4 o;
          // lo = x - hi is exact; it is rounded to NEAREST tf32 here (the MMA would truncate it: a coherent bias)
          o.x = (__float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
          o.y = (__float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
          o.z = (__float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
          o.w = (__float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
          dst[c] = o;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv_bar[s])) : "memory");
      }
    }
  } else if (MODE == TC_DGRAD) {
    // ===================================================================== H producer (dgrad), warp 10
    if (warp == 2 + TCP_EPI_WARPS && elect_one()) {
      int g = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int pi, split, tm, tn;
        decode(t, pi, split, tm, tn);
        const CUtensorMap* map_h = &batch.p[pi].map_h;
        const int block_n = sp[pi].block_n;
        const int m0 = tm * TC_BLOCK_M, n0 = tn * block_n;
        const int nchunks = (block_n + CH - 1) / CH;
        for (int c = 0; c < nchunks; ++c, ++g) {
          const int hs = g % HCH;
          mbar_wait(&h_empty[hs], ((uint32_t)(g / HCH) & 1u) ^ 1u);
          mbar_expect_tx(&h_full[hs], 16384u);
          tma_load_2d(staging + (size_t)hs * 16384, map_h, &h_full[hs], n0 + c * CH, m0);  // box {CH cols, 128 rows}
        }
      }
    }
  } else if (MODE == TC_WGRAD) {
    // ===================================================================== bias-gradient reducers (wgrad)
    const int f = (warp - 2 - TCP_EPI_WARPS) * 32 + lane;  // feature 0..127 of the A (dY) tile
    const uint32_t foff = (uint32_t)(f >> 5) * 4096u + (uint32_t)(f & 7) * 4u;
    const uint32_t fchunk = (uint32_t)((f & 31) >> 3);
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int pi, split, tm, tn;
      decode(t, pi, split, tm, tn);
      const int kb0 = split * sp[pi].kb_per_split;
      const int kb1 = min(sp[pi].num_kb, kb0 + sp[pi].kb_per_split);
      float* const db_ptr = sp[pi].db;
      const bool active = db_ptr != nullptr && tn == 0;
      if constexpr (F16) {
        // thread <-> (16-byte chunk c = 8 features, row group g): 8 x LDS.128 per k-block instead of 64 two-byte loads;
        // rows g, g + 8, ... share (row & 7) = g, so the swizzled chunk offset is a per-thread constant
        const int c = f & 15, g = f >> 4;
        const uint32_t base = (uint32_t)(c >> 3) * 8192u + (uint32_t)g * 128u + ((uint32_t)((c & 7) ^ g) << 4);
        float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % stages;
          mbar_wait(&full_bar[s], (it / stages) & 1);
          if (active) {
            const uint8_t* sa = ring + (size_t)s * stage_bytes + base;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 v = *(const uint4*)(sa + i * 1024);
              const float2 p0 = unpack_f16x2(v.x), p1 = unpack_f16x2(v.y), p2 = unpack_f16x2(v.z), p3 = unpack_f16x2(v.w);
              acc8[0] += p0.x; acc8[1] += p0.y; acc8[2] += p1.x; acc8[3] += p1.y;
              acc8[4] += p2.x; acc8[5] += p2.y; acc8[6] += p3.x; acc8[7] += p3.y;
            }
          }
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
        }
        if (active) {  // the two row groups of a warp share a chunk: fold them, then one atomic per feature per warp
#pragma unroll
          for (int e = 0; e < 8; ++e) acc8[e] += __shfl_xor_sync(0xffffffffu, acc8[e], 16);
          if (lane < 16) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int feat = tm * TC_BLOCK_M + c * 8 + e;
              if (feat < sp[pi].M) atomicAdd(db_ptr + feat, acc8[e] * sp[pi].out_scale);
            }
          }
        }
        continue;
      }
      float acc = 0.f;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % stages;
        mbar_wait(&full_bar[s], (it / stages) & 1);
        if (active) {
          const uint8_t* sa = ring + (size_t)s * stage_bytes;
#pragma unroll
          for (int r = 0; r < TC_BLOCK_K; ++r)
            acc += *(const float*)(sa + foff + (uint32_t)r * 128u + ((fchunk ^ (uint32_t)(r & 3)) << 5));
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
      }
      const int feat = tm * TC_BLOCK_M + f;
      if (active && feat < sp[pi].M) atomicAdd(db_ptr + feat, acc * sp[pi].out_scale);
    }
  }

  __syncthreads();
  if (threadIdx.x == 0) TRACE(31);
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int colsum_accumulate(const float* x, int M, int N, int ld, float* out, cudaStream_t st);

// -------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch ld (elements); box {box_cols, box_rows}.
// operand = true : MMA operand, encoded as TFLOAT32 so that TMA rounds fp32 -> tf32 to nearest on the way in
// operand = false: plain fp32 tile (epilogue output / saved activation), no conversion
static bool make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_cols,
                     int box_rows, bool atom32, bool operand = true) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  static const int dtype_sel = [] { const char* e = getenv("RSLRL_TMA_DTYPE"); return e ? atoi(e) : 1; }();  // default: TFLOAT32 = TMA rounds fp32 -> tf32 to nearest (measured: profiles/experiments)
  const CUtensorMapDataType dt = !operand          ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype_sel == 1 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                                 : dtype_sel == 2 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32_FTZ
                                                  : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(map, dt, 2, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 2-D fp16 tensor [rows, cols] with row pitch ld (elements); box {box_cols <= 64, box_rows}, 128-byte swizzle.
static bool make_map16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
                       int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
bool tc_make_map16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
                   int box_rows) {
  return make_map16(map, base, rows, cols, ld, box_cols, box_rows);
}
static bool tma_ok16(const void* p, long long ld) { return ((uintptr_t)p & 15) == 0 && (ld % 8) == 0; }

// exported to the other tcgen05 translation units (mlp_fused.cu)
bool tc_make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_cols,
                 int box_rows, bool operand) {
  return make_map(map, base, rows, cols, ld, box_cols, box_rows, false, operand);
}

static bool tma_ok(const void* p, long long ld) { return ((uintptr_t)p & 15) == 0 && (ld % 4) == 0; }

static int max_block_n() { return TCP_MAX_BLOCK_N; }

static int pick_block_n(int N) {
  if (N >= max_block_n()) return max_block_n();
  int b = (N + 31) / 32 * 32;  // MN-major operands come in 32-element atoms; keep it simple for K-major too
  if (N <= 16) b = 16;
  return b;
}

int tc_num_sms();
static int num_sms() { return tc_num_sms(); }
int tc_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = kNumSMs;
  }
  return n;
}

template <int MODE, bool F16 = false, bool X3 = false>
static int launch_tc_persistent(TcBatch& b, int n, cudaStream_t st, const char* what) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_persistent_kernel<MODE, F16, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return 1;
    }
    attr_set = true;
  }
  int total = 0, max_bn = 16;
  for (int i = 0; i < n; ++i) {
    total += b.p[i].tiles_m * b.p[i].tiles_n * b.p[i].splits;
    if (b.p[i].block_n > max_bn) max_bn = b.p[i].block_n;
  }
  if (total == 0) return 0;
  const size_t stage = (size_t)(TC_BLOCK_M + max_bn) * TC_BLOCK_K * 4 * (X3 ? 2 : 1);
  size_t stages = tcp_ring_budget(MODE, F16, X3) / stage;
  if (stages > TCP_MAX_STAGES) stages = TCP_MAX_STAGES;
  if (stages < 2) {
    set_error("%s: tile too wide for the operand ring", what);
    return 1;
  }
  const size_t smem = stages * stage + tcp_staging_bytes(MODE, F16, X3) +
                      (3 * TCP_MAX_STAGES + 4 + 2 * TCP_H_CHUNKS) * 8 + 16 + TC_MAX_PROBLEMS * sizeof(TcpScalars) +
                      TCP_EPI_WARPS * 128 + 1024;
  const int grid = total < num_sms() ? total : num_sms();
  launch_chain_kernel(tc_gemm_persistent_kernel<MODE, F16, X3>, dim3(grid), dim3(tcp_threads(X3)), smem, st, b, n);
  return finish_launch(what);
}

template <int MODE, bool F16 = false, bool X3 = false>
static int launch_tc(TcBatch& b, int n, cudaStream_t st, const char* what) {
  for (int i = 0; i < n; ++i)
    if (b.p[i].out_scale == 0.f) b.p[i].out_scale = 1.f;
  return launch_tc_persistent<MODE, F16, X3>(b, n, st, what);
}

// Each tc_* entry returns 0 on success, 1 on error, and -1 if a problem cannot be described to TMA
// (caller falls back to the fp32 kernel for the whole group).
int tc_linear_forward(int n, const rslrl_linear_fwd_t* pr, cudaStream_t st, bool x3) {
  const char* what = x3 ? "rslrl_linear_forward(tf32x3)" : "rslrl_linear_forward(tf32)";
  if (n > 2) return -1;
  TcBatch b{};
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_fwd_t& q = pr[i];
    if (!tma_ok(q.x, q.ldx) || !tma_ok(q.w, q.K)) return -1;
    TcProblem& P = b.p[i];
    P.block_n = pick_block_n(q.N);
    if (x3 && P.block_n > TCP_X3_BLOCK_N) P.block_n = TCP_X3_BLOCK_N;
    if (!make_map(&P.map_a, q.x, q.M, q.K, q.ldx, TC_BLOCK_K, TC_BLOCK_M, false, !x3) ||
        !make_map(&P.map_b, q.w, q.N, q.K, q.K, TC_BLOCK_K, P.block_n, false, !x3)) {
      set_error("%s: cuTensorMapEncodeTiled failed", what);
      return 1;
    }
    P.C = q.y; P.bias = q.b; P.M = q.M; P.N = q.N; P.K = q.K; P.ldc = q.ldy; P.act = q.act;
    P.tma_out = tma_ok(q.y, q.ldy) && make_map(&P.map_c, q.y, q.M, q.N, q.ldy, 32, 32, false, false);
    P.tiles_m = (q.M + TC_BLOCK_M - 1) / TC_BLOCK_M;
    P.tiles_n = (q.N + P.block_n - 1) / P.block_n;
    P.num_kb = (q.K + TC_BLOCK_K - 1) / TC_BLOCK_K;
    P.splits = 1; P.kb_per_split = P.num_kb;
  }
  return x3 ? launch_tc<TC_FWD, false, true>(b, n, st, what) : launch_tc<TC_FWD>(b, n, st, what);
}

int tc_linear_backward_data(int n, const rslrl_linear_bwd_data_t* pr, cudaStream_t st, bool x3) {
  const char* what = x3 ? "rslrl_linear_backward_data(tf32x3)" : "rslrl_linear_backward_data(tf32)";
  if (n > 2) return -1;
  TcBatch b{};
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_bwd_data_t& q = pr[i];
    // dX[M,K] = dY[M,N] W[N,K]: reduction N; B = W is MN-major (K contiguous) -> K must be whole 32-atoms
    if (!tma_ok(q.dy, q.ldy) || !tma_ok(q.w, q.K) || (q.K % 32) != 0) return -1;
    TcProblem& P = b.p[i];
    const int bn_cap = x3 ? TCP_X3_BLOCK_N : max_block_n();
    P.block_n = q.K >= bn_cap ? bn_cap : q.K;
    if (q.K % P.block_n != 0) return -1;
    if (!make_map(&P.map_a, (const float*)q.dy, q.M, q.N, q.ldy, TC_BLOCK_K, TC_BLOCK_M, false, !x3) ||
        !make_map(&P.map_b, (const float*)q.w, q.N, q.K, q.K, 32, TC_BLOCK_K, true, !x3)) {
      set_error("%s: cuTensorMapEncodeTiled failed", what);
      return 1;
    }
    P.C = (float*)q.dx; P.M = q.M; P.N = q.K; P.K = q.N; P.ldc = q.ldx; P.act = q.act;
    if (!tma_ok(q.dx, q.ldx)) return -1;
    // (act == IDENTITY with an h of the output's shape: the factor is 1 whatever h holds -- used for the data gradient
    //  that leaves an MLP towards a recurrent memory; without any h tensor the exact fp32 kernel runs)
    if (q.h == nullptr || !tma_ok(q.h, q.ldh)) return -1;
    if (!make_map(&P.map_c, (const float*)q.dx, q.M, q.K, q.ldx, 32, 32, false, false) ||
        !make_map(&P.map_h, (const float*)q.h, q.M, q.K, q.ldh, 32, TC_BLOCK_M, false, false)) {
      set_error("%s: cuTensorMapEncodeTiled failed", what);
      return 1;
    }
    P.tma_out = 1;
    P.tiles_m = (q.M + TC_BLOCK_M - 1) / TC_BLOCK_M;
    P.tiles_n = q.K / P.block_n;
    P.num_kb = (q.N + TC_BLOCK_K - 1) / TC_BLOCK_K;
    P.splits = 1; P.kb_per_split = P.num_kb;
  }
  const int rc = x3 ? launch_tc<TC_DGRAD, false, true>(b, n, st, what) : launch_tc<TC_DGRAD>(b, n, st, what);
  if (rc) return rc;
  for (int i = 0; i < n; ++i)  // optional fused-API output: bias gradient of the previous layer
    if (pr[i].db_prev && colsum_accumulate((const float*)pr[i].dx, pr[i].M, pr[i].K, pr[i].ldx, pr[i].db_prev, st)) return 1;
  return 0;
}

int tc_linear_backward_weight(int n, const rslrl_linear_bwd_weight_t* pr, cudaStream_t st, bool x3) {
  const char* what = x3 ? "rslrl_linear_backward_weight(tf32x3)" : "rslrl_linear_backward_weight(tf32)";
  if (n > 2) return -1;
  TcBatch b{};
  int tiles_all = 0;
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_bwd_weight_t& q = pr[i];
    // dW[N,K] += dY[M,N]^T X[M,K]: reduction M; both operands MN-major
    if (!tma_ok(q.dy, q.ldy) || !tma_ok(q.x, q.ldx)) return -1;
    TcProblem& P = b.p[i];
    P.block_n = pick_block_n(q.K);
    if (P.block_n < 32) P.block_n = 32;
    if (x3 && P.block_n > TCP_X3_BLOCK_N) P.block_n = TCP_X3_BLOCK_N;
    P.tiles_m = (q.N + TC_BLOCK_M - 1) / TC_BLOCK_M;
    P.tiles_n = (q.K + P.block_n - 1) / P.block_n;
    tiles_all += P.tiles_m * P.tiles_n;
  }
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_bwd_weight_t& q = pr[i];
    TcProblem& P = b.p[i];
    if (!make_map(&P.map_a, (const float*)q.dy, q.M, q.N, q.ldy, 32, TC_BLOCK_K, true, !x3) ||
        !make_map(&P.map_b, (const float*)q.x, q.M, q.K, q.ldx, 32, TC_BLOCK_K, true, !x3)) {
      set_error("%s: cuTensorMapEncodeTiled failed", what);
      return 1;
    }
    P.C = q.dw; P.db = q.db; P.M = q.N; P.N = q.K; P.K = q.M; P.ldc = q.K; P.out_scale = q.scale; P.out_scale_dev = q.scale_dev;
    P.tma_out = tma_ok(q.dw, q.K) && make_map(&P.map_c, q.dw, q.N, q.K, q.K, 32, 32, false, false);
    P.num_kb = (q.M + TC_BLOCK_K - 1) / TC_BLOCK_K;
    int want = (2 * kNumSMs + tiles_all - 1) / tiles_all;   // ~2 CTAs worth of work per SM
    int max_s = P.num_kb / 8;                               // at least 8 k-blocks (256 rows) per split
    if (max_s < 1) max_s = 1;
    if (want > max_s) want = max_s;
    if (want < 1) want = 1;
    P.kb_per_split = (P.num_kb + want - 1) / want;
    P.splits = (P.num_kb + P.kb_per_split - 1) / P.kb_per_split;
  }
  return x3 ? launch_tc<TC_WGRAD, false, true>(b, n, st, what) : launch_tc<TC_WGRAD>(b, n, st, what);
}

// ---- fp16-storage variants (RSLRL_PREC_FP16): same kernel, F16 = true.  -1 = shape not eligible (no fallback: the
// caller reports it, the operands are fp16 and no other kernel reads them).
int tc16_linear_backward_data(int n, const rslrl_linear_bwd_data_t* pr, cudaStream_t st) {
  const char* what = "rslrl_linear_backward_data(fp16)";
  if (n > 2) return -1;
  TcBatch b{};
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_bwd_data_t& q = pr[i];
    // dX[M,K] = dY[M,N] W[N,K]: A = dY K-major (k-blocks of 64 of N), B = W16 MN-major in 64x64 boxes
    if (!tma_ok16(q.dy, q.ldy) || !tma_ok16(q.w, q.K) || !tma_ok16(q.dx, q.ldx) || !tma_ok16(q.h, q.ldh)) return -1;
    if ((q.K % 64) != 0 || q.act == RSLRL_ACT_IDENTITY || q.h == nullptr || q.db_prev != nullptr) return -1;
    TcProblem& P = b.p[i];
    P.block_n = q.K >= max_block_n() ? max_block_n() : q.K;
    if (q.K % P.block_n != 0) return -1;
    if (!make_map16(&P.map_a, q.dy, q.M, q.N, q.ldy, 64, TC_BLOCK_M) ||
        !make_map16(&P.map_b, q.w, q.N, q.K, q.K, 64, 64) ||
        !make_map16(&P.map_c, q.dx, q.M, q.K, q.ldx, 64, 32) ||
        !make_map16(&P.map_h, q.h, q.M, q.K, q.ldh, 64, TC_BLOCK_M)) {
      set_error("%s: cuTensorMapEncodeTiled failed", what);
      return 1;
    }
    P.C = nullptr; P.M = q.M; P.N = q.K; P.K = q.N; P.ldc = q.ldx; P.act = q.act; P.tma_out = 1;
    P.tiles_m = (q.M + TC_BLOCK_M - 1) / TC_BLOCK_M;
    P.tiles_n = q.K / P.block_n;
    P.num_kb = (q.N + 63) / 64;
    P.splits = 1; P.kb_per_split = P.num_kb;
  }
  return launch_tc<TC_DGRAD, true>(b, n, st, what);
}

// Relative cost of one k-block of a wgrad tile: proportional to the bytes it streams (128 + block_n rows of 128 B)
// when the ring is full-width, but never below the latency floor of a ring stage (measured with the in-kernel
// timeline: 0.82 us for a 48 KB k-block, 0.72 us for a 24 KB one -> floor at 0.88 of the full-width cost).
static double wgrad16_kb_cost(int block_n) {
  const double bytes = TC_BLOCK_M + block_n, floor_ = 0.88 * (TC_BLOCK_M + TCP_MAX_BLOCK_N);
  return bytes > floor_ ? bytes : floor_;
}

int tc16_linear_backward_weight(int n, const rslrl_linear_bwd_weight_t* pr, cudaStream_t st) {
  const char* what = "rslrl_linear_backward_weight(fp16)";
  if (n > TC_MAX_PROBLEMS) return -1;
  TcBatch b{};
  int tiles_all = 0;
  double cost_all = 0.0;  // smem bytes streamed by all tiles of all problems (the mainloops are L2->SM bound)
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_bwd_weight_t& q = pr[i];
    // dW[N,K] += scale * dY[M,N]^T X[M,K]: reduction M in k-blocks of 64 rows; both operands MN-major 64x64 boxes
    if (!tma_ok16(q.dy, q.ldy) || !tma_ok16(q.x, q.ldx)) return -1;
    TcProblem& P = b.p[i];
    P.block_n = q.K >= max_block_n() ? max_block_n() : (q.K + 63) / 64 * 64;
    P.tiles_m = (q.N + TC_BLOCK_M - 1) / TC_BLOCK_M;
    P.tiles_n = (q.K + P.block_n - 1) / P.block_n;
    tiles_all += P.tiles_m * P.tiles_n;
    cost_all += (double)P.tiles_m * P.tiles_n * ((q.M + 63) / 64) * wgrad16_kb_cost(P.block_n);
  }
  for (int i = 0; i < n; ++i) {
    const rslrl_linear_bwd_weight_t& q = pr[i];
    TcProblem& P = b.p[i];
    // (a weight matrix whose rows are not 16-byte aligned, e.g. 235 inputs, takes the vector-atomic epilogue)
    const bool c_tma = tma_ok(q.dw, q.K);
    if (!make_map16(&P.map_a, q.dy, q.M, q.N, q.ldy, 64, 64) || !make_map16(&P.map_b, q.x, q.M, q.K, q.ldx, 64, 64) ||
        (c_tma && !make_map(&P.map_c, q.dw, q.N, q.K, q.K, 32, 32, false, false))) {
      set_error("%s: cuTensorMapEncodeTiled failed", what);
      return 1;
    }
    P.C = q.dw; P.db = q.db; P.M = q.N; P.N = q.K; P.K = q.M; P.ldc = q.K; P.out_scale = q.scale; P.out_scale_dev = q.scale_dev; P.tma_out = c_tma ? 1 : 0;
    P.num_kb = (q.M + 63) / 64;
  }
  // k-splits: every work item (tile x k-range) of every grouped problem streams about the same number of bytes, and
  // there are at most items_per_sm * #SMs of them (one more would add a whole extra round to the static round-robin
  // schedule): grow the per-item byte budget until the item count fits
  static const int items_per_sm = [] { const char* e = getenv("RSLRL_WGRAD_ITEMS_PER_SM"); return e ? atoi(e) : 2; }();
  const int budget = items_per_sm * num_sms();
  double unit = cost_all / budget;
  for (int iter = 0; iter < 64; ++iter, unit *= 1.03) {
    int total = 0;
    for (int i = 0; i < n; ++i) {
      TcProblem& P = b.p[i];
      int kps = (int)(unit / wgrad16_kb_cost(P.block_n) + 0.999);
      if (kps < 4) kps = 4;                                 // at least 4 k-blocks (256 rows) per split
      if (kps > P.num_kb) kps = P.num_kb;
      P.kb_per_split = kps;
      P.splits = (P.num_kb + kps - 1) / kps;
      total += P.tiles_m * P.tiles_n * P.splits;
    }
    if (total <= budget) break;
  }
  return launch_tc<TC_WGRAD, true>(b, n, st, what);
}

// out[n] += sum_m x[m, n]  (only used when a caller asks the tensor-core wgrad for db)
__global__ void colsum_kernel(const float* __restrict__ x, int M, int N, int ld, float* __restrict__ out) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float acc = 0.f;
  if (n < N)
    for (int r = r0 + (threadIdx.x >> 5); r < r1; r += 8) acc += __ldg(x + (size_t)r * ld + n);
  __shared__ float s[8][33];
  s[threadIdx.x >> 5][threadIdx.x & 31] = acc;
  __syncthreads();
  if (threadIdx.x < 32 && n < N) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s[w][threadIdx.x];
    atomicAdd(out + n, t);
  }
}

int colsum_accumulate(const float* x, int M, int N, int ld, float* out, cudaStream_t st) {
  int ysplit = (M + 1023) / 1024;
  if (ysplit > 256) ysplit = 256;
  colsum_kernel<<<dim3((N + 31) / 32, ysplit), 256, 0, st>>>(x, M, N, ld, out);
  return finish_launch("rslrl_linear_backward_weight(tf32, db)");
}

}  // namespace rslrl

#ifdef RSLRL_TC_TRACE
extern "C" int rslrl_debug_read_trace(unsigned long long* host) {
  return (int)cudaMemcpyFromSymbol(host, rslrl::g_tc_trace, sizeof(unsigned long long) * 148 * 32);
}
#endif
