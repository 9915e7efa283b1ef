/*
 * rslrl_b200.h -- C ABI of the B200-native PPO learning-iteration kernels.
 *
 * The reference (kaixi287/rsl_rl v3.1.0) is pure Python on PyTorch and has no FFI of its own; the
 * drop-in boundary is its Python class API.  This header is the boundary one
 * level below: what the Python mirror in rsl_rl_b200/ binds with ctypes.  Every entry point cites
 * the reference code whose arithmetic it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the parameter name ends in `_host`;
 *   - all tensors are dense row-major fp32 unless stated; sizes are element counts;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work (no sync, no alloc);
 *   - return value 0 = ok, non-zero = error (message via rslrl_last_error());
 *   - no function falls back to the host: without a CUDA device every launch returns an error.
 */
#ifndef RSLRL_B200_H_
#define RSLRL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSLRL_ABI_VERSION 3
#define RSLRL_MAX_SEGMENTS 16 /* max obs groups / gather fields per launch */
#define RSLRL_MAX_PROBLEMS 8  /* max GEMM problems grouped into one launch (fp16 wgrad: 4 layers x actor, critic) */

typedef void* rslrl_stream_t;

/* ---- library plumbing ------------------------------------------------------------------------ */
int rslrl_abi_version(void);
const char* rslrl_last_error(void);
/* number of kernel launches enqueued through this library since load / last reset */
unsigned long long rslrl_launch_count(void);
void rslrl_reset_launch_count(void);
/* adds `n` to the launch counter: used by the host when it replays a captured CUDA graph of n of
 * this library's kernels */
void rslrl_add_launch_count(unsigned long long n);

/* dtype codes for integer/bool inputs coming from the environment (VecEnv.step, env/vec_env.py:56-98) */
enum { RSLRL_U8 = 0, RSLRL_I32 = 1, RSLRL_I64 = 2, RSLRL_F32 = 3 };
/* activation codes (rsl_rl/utils/utils.py:18-49 resolve_nn_activation) */
enum {
  RSLRL_ACT_IDENTITY = 0, RSLRL_ACT_ELU = 1, RSLRL_ACT_RELU = 2, RSLRL_ACT_TANH = 3, RSLRL_ACT_SIGMOID = 4,
  RSLRL_ACT_LRELU = 5, RSLRL_ACT_SELU = 6, RSLRL_ACT_SOFTPLUS = 7
};
/* GEMM arithmetic: fp32 CUDA-core FMA (1e-5 parity class) or TF32 tcgen05 tensor cores (1e-3 class) */
enum { RSLRL_PREC_FP32 = 0, RSLRL_PREC_TF32 = 1, RSLRL_PREC_FP16 = 2, RSLRL_PREC_TF32X3 = 3 };
/* RSLRL_PREC_TF32X3 ("3xTF32"): fp32 operands split on chip into hi = tf32(x) and lo = x - hi, three tensor-core products
 * (lo*hi + hi*lo + hi*hi) into one fp32 accumulator: the 1e-5 error class of the reference's fp32 arithmetic
 * (torch.backends.cuda.matmul.allow_tf32 = False) at tensor-core speed.  Accepted by the rslrl_linear_* entry points. */
/* RSLRL_PREC_FP16 ("16-bit activation storage"): the saved activations h, the layer gradients dy / dx, the
 * minibatch observations x and a shadow copy of the weights are IEEE fp16 in HBM (10-bit mantissa = the precision
 * tf32 gives the same operands) and the MMAs are tcgen05 kind::f16 with fp32 accumulation.  Master weights, weight /
 * bias gradients, biases and the network outputs stay fp32.  Layer gradients carry a power-of-two scale (set where
 * the chain starts, rslrl_linear_bwd_head_t.dx_scale, undone by rslrl_linear_bwd_weight_t.scale) so that they stay
 * in fp16's normal range.  In this mode the pointers documented "fp32, or fp16 with RSLRL_PREC_FP16" below are
 * `const __half*`; row strides are always in elements. */
/* action-noise parameterisation (rsl_rl/modules/actor_critic.py:88-93,133-138) */
enum { RSLRL_STD_SCALAR = 0, RSLRL_STD_LOG = 1 };

/* ---- rollout side ---------------------------------------------------------------------------- */

/* Gaussian head of PPO.act (rsl_rl/algorithms/ppo.py:133-137; modules/actor_critic.py:133-146,170-171;
 * torch.distributions.Normal.sample/log_prob):
 *   sigma[a]   = std[a]            (RSLRL_STD_SCALAR)  |  exp(log_std[a])  (RSLRL_STD_LOG)
 *   actions    = z * sigma + mu    (two roundings, as normal_().mul_().add_())
 *   log_prob   = sum_a( -(a-mu)^2/(2 sigma^2) - log sigma - log sqrt(2 pi) )
 * mu,z,actions,sigma_out: [n,A]; log_prob: [n].  sigma_out may be NULL. */
int rslrl_gaussian_act(const float* mu, const float* std_param, int std_type, const float* z, float* actions,
                       float* log_prob, float* sigma_out, int64_t n, int A, rslrl_stream_t stream);

/* RolloutStorage.add_transitions (rsl_rl/storage/rollout_storage.py:77-103) fused with the reward
 * handling of PPO.process_env_step (rsl_rl/algorithms/ppo.py:150-164):
 *   rewards_row = rewards (+ intrinsic) + gamma * (values * time_outs)
 * One launch writes row `step` of every [T,N,.] buffer; the *_row pointers already address that row.
 * obs_src_host/obs_dst_host/obs_dim_host: host arrays with one entry per observation group.
 * sigma_stride: 0 -> sigma is [A] and is broadcast, A -> sigma is [N,A].
 * dones / time_outs: dtype codes above; time_outs and intrinsic may be NULL.
 * A (source, row) pair of actions / mu / sigma / values / log_prob that is NULL is skipped (values may only be NULL when
 * time_outs is): the distillation rollout buffer (rollout_storage.py:57-59,88-90: observations, actions,
 * privileged_actions, rewards, dones) is stored with the teacher's actions in the `mu` segment and the PPO-only
 * fields absent. */
int rslrl_store_transition(int n_groups, const float* const* obs_src_host, float* const* obs_dst_host,
                           const int* obs_dim_host, const float* actions, const float* mu, const float* sigma,
                           int sigma_stride, const float* values, const float* log_prob, const float* rewards,
                           const void* dones, int dones_dtype, const void* time_outs, int time_outs_dtype,
                           const float* intrinsic, double gamma, float* actions_row, float* mu_row,
                           float* sigma_row, float* values_row, float* log_prob_row, float* rewards_row,
                           uint8_t* dones_row, int N, int A, rslrl_stream_t stream);

/* The same launch for the rollout fast path (PPO.act has already written actions / mean / sigma / value / log-prob of the
 * step in place): everything that is fixed for a step index lives in a plan the host builds once; per env step only the
 * source pointers are passed.  Observations, rewards (+ intrinsic) with the time-out bootstrap and dones move. */
typedef struct {
  float* obs_dst[RSLRL_MAX_SEGMENTS]; /* row `step` of every observation group's [T,N,d] buffer */
  int obs_dim[RSLRL_MAX_SEGMENTS];
  const float* values_row;            /* row `step` of values (read for the bootstrap) */
  float* rewards_row;
  uint8_t* dones_row;
  double gamma;
  int n_groups, N, A;
} rslrl_store_plan_t;
int rslrl_store_step(const rslrl_store_plan_t* plan, const float* const* obs_src_host, const float* rewards,
                     const void* dones, int dones_dtype, const void* time_outs, int time_outs_dtype,
                     const float* intrinsic, rslrl_stream_t stream);

/* RolloutStorage.compute_returns (rsl_rl/storage/rollout_storage.py:127-149): reverse GAE scan over T
 * with done masking, returns, advantages = returns - values, and (normalize != 0) the global
 * rewards/values/returns/advantages: [T,N] fp32, dones: [T,N] u8, last_values: [N].
 * (adv - mean) / (std_unbiased + 1e-8) normalisation, in ONE kernel.  N % 16 == 0 and T <= 32: inputs staged through
 * shared memory with bulk copies; up to 8 192 envs the grid is one thread-block cluster (statistics over distributed
 * shared memory), above that a cooperative launch with a grid barrier; other shapes: register / grid-stride kernels with
 * the same barrier.  Returns and un-normalised advantages are bit-identical to the reference's fp32 op order.
 * workspace: rslrl_gae_workspace_bytes(N) bytes of device scratch, zero-initialised ONCE by the caller
 * (the kernel re-arms it); one workspace must not be shared by launches that may run concurrently. */
int64_t rslrl_gae_workspace_bytes(int N);
int rslrl_gae_norm(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                   float* returns, float* advantages, void* workspace, int T, int N, double gamma, double lam,
                   int normalize, rslrl_stream_t stream);

/* ---- minibatch construction ------------------------------------------------------------------ */

/* RolloutStorage.mini_batch_generator's `x.flatten(0,1)[batch_idx]` for several fields at once
 * (rsl_rl/storage/rollout_storage.py:168-197).  For field f and output row j:
 *   dst[f][j*dst_stride[f] + dst_col[f] + c] = src[f][idx[j]*row_floats[f] + c],  c < row_floats[f]
 * (dst_stride/dst_col let several observation groups be concatenated into one matrix, which also
 * absorbs the torch.cat of modules/actor_critic.py:158-168).  idx: int64 [n_rows]. */
int rslrl_permute_gather(int n_fields, const float* const* src_host, float* const* dst_host,
                         const int* row_floats_host, const int* dst_stride_host, const int* dst_col_host,
                         const int64_t* idx, int64_t n_rows, rslrl_stream_t stream);

/* ---- dense layers (rsl_rl/networks/mlp.py:106-114 forward; autograd backward of nn.Linear + activation) */

typedef struct {
  const float* x; /* [M,K], row stride ldx */
  const float* w; /* [N,K] row-major (nn.Linear.weight), row stride K */
  const float* b; /* [N] or NULL */
  float* y;       /* [M,N], row stride ldy */
  int M, N, K, ldx, ldy;
  int act;        /* RSLRL_ACT_* applied to x W^T + b */
} rslrl_linear_fwd_t;

typedef struct {
  const void* dy; /* [M,N], row stride ldy : gradient w.r.t. this layer's pre-activation output (fp32, or fp16 with
                     RSLRL_PREC_FP16 -- as are w, h and dx) */
  const void* w;  /* [N,K] */
  const void* h;  /* [M,K], row stride ldh : this layer's INPUT (= previous layer's activation output) */
  void* dx;       /* [M,K], row stride ldx : gradient w.r.t. previous layer's pre-activation */
  int M, N, K, ldy, ldh, ldx;
  int act;         /* activation of the PREVIOUS layer: dx = (dy W) * act'(.) evaluated from h */
  float* db_prev;  /* [K] or NULL: += column sums of dx = bias gradient of the previous layer (fused into the
                      epilogue that produces dx, so dx is not re-read for it) */
} rslrl_linear_bwd_data_t;

typedef struct {
  const void* dy; /* [M,N], row stride ldy (fp32, or fp16 with RSLRL_PREC_FP16 -- as is x) */
  const void* x;  /* [M,K], row stride ldx */
  float* dw;      /* [N,K] accumulated:  dw += scale * dy^T x */
  float* db;      /* [N]   accumulated:  db += scale * colsum(dy); may be NULL */
  int M, N, K, ldy, ldx;
  float scale;    /* 0 is read as 1; the fp16 path passes 1 / dx_scale of the chain's head */
  const float* scale_dev; /* NULL, or a device float multiplied into `scale` when the kernel starts: element [1] of the
                             rslrl_ppo_loss_fwd_bwd scale_state (1 / the dynamic gradient scale of this minibatch) */
} rslrl_linear_bwd_weight_t;

typedef struct {
  const float* dy; /* [M,N], row stride ldy */
  const float* w;  /* [N,K] */
  const void* h;   /* [M,K], row stride ldh : the layer's input (previous activation output); fp16 if f16 != 0 */
  void* dx;        /* [M,K], row stride ldx, or NULL : dx_scale * (dy W) * act'(h); fp16 if f16 != 0 */
  float* dw;       /* [N,K] accumulated (never scaled) */
  float* db;       /* [N] accumulated, or NULL */
  int M, N, K, ldy, ldh, ldx;
  int act;         /* activation of the previous layer */
  int f16;         /* h and dx are fp16 */
  float dx_scale;  /* 0 is read as 1 */
} rslrl_linear_bwd_head_t;

/* Whole 3-hidden-layer MLP forward in ONE launch (rsl_rl/networks/mlp.py:106-114 for actor and critic):
 *   y = W4 act(W3 act(W2 act(W1 x + b1) + b2) + b3) + b4
 * on tcgen05 (tf32) with the activations kept in shared / tensor memory between the layers.  h[0..2] (all three or
 * none) receive the hidden activations for the backward pass; NULL (rollout, inference) stores nothing but y. */
typedef struct {
  const float* x;    /* [M,K0], row stride ldx */
  const void* w[4];  /* nn.Linear weights: [N1,K0] (row stride ldw0), [N2,N1], [N3,N2], [N4,N3] row-major; w[0] is
                        always fp32, w[1..3] are the fp16 shadows when precision == RSLRL_PREC_FP16 */
  const float* b[4]; /* biases (entries may be NULL) */
  void* h[3];        /* [M,N1], [M,N2], [M,N3] with row strides ldh[i], or all NULL (fp16 with RSLRL_PREC_FP16) */
  float* y;          /* [M,N4], row stride ldy */
  int M, K0, N1, N2, N3, N4;
  int ldx, ldw0, ldh[3], ldy;
  int act;           /* RSLRL_ACT_* of the three hidden layers */
  int precision;     /* RSLRL_PREC_TF32 (0 is read as TF32) or RSLRL_PREC_FP16: layer 1 always runs in tf32 on the
                        fp32 x / w[0]; layers 2-4 consume the on-chip activations as tf32 or fp16 */
  /* Optional Gaussian policy head fused into the output epilogue (PPO.act, rsl_rl/algorithms/ppo.py:129-140 ->
   * modules/actor_critic.py:130-146): with head_z != NULL the N4 outputs y = mu are also turned into
   *   head_actions[M,N4] = head_z * sigma + mu,  head_logp[M] = sum log N(actions; mu, sigma),  head_sigma[M,N4] = sigma
   * (sigma from head_std[N4] as in rslrl_gaussian_act; same arithmetic, bit for bit).  head_sigma may be NULL.
   * head_actions_host (may be NULL): a second destination of the actions, e.g. PINNED HOST memory mapped into the device
   * address space: a host-resident simulator then reads the actions after one stream synchronisation, without a
   * device-to-host copy (the `actions.to(self.env.device)` of on_policy_runner.py:104). */
  const float* head_z;
  const float* head_std;
  int head_std_type;
  float* head_actions;
  float* head_logp;
  float* head_sigma;
  float* head_actions_host;
} rslrl_mlp3_fwd_t;

/* Each call runs `n` independent problems (e.g. actor and critic) in one launch; *_host arrays live
 * in host memory and are copied into the kernel parameters. */
int rslrl_linear_forward(int n, const rslrl_linear_fwd_t* probs_host, int precision, rslrl_stream_t stream);
int rslrl_linear_backward_data(int n, const rslrl_linear_bwd_data_t* probs_host, int precision,
                               rslrl_stream_t stream);
int rslrl_linear_backward_weight(int n, const rslrl_linear_bwd_weight_t* probs_host, int precision,
                                 rslrl_stream_t stream);
/* Fused data-gradient chain of the two wide hidden layers of a 3-hidden-layer MLP on the fp16-storage path
 * (autograd of rsl_rl/networks/mlp.py:106-114 for layers 3 and 2, both nets in one launch):
 *   dz2 = (dz3 w3) * act'(h2),   dz1 = (dz2 w2) * act'(h1)
 * The dz2 tile stays in shared memory between the two products (it is still stored: the weight gradient of layer 2
 * needs it).  All tensors fp16 with 16-byte aligned rows; w3 [N3,N2] and w2 [N2,N1] are the fp16 weight shadows.
 * Returns 4 when the shape is not eligible (N3 in {64,128}, N2 in {128,256}, N1 in {128,256,512}, equal shapes for all nets):
 * the caller then issues two rslrl_linear_backward_data launches. */
typedef struct {
  const void* dz3; /* [M,N3], row stride ld_dz3 */
  const void* w3;
  const void* w2;
  const void* h2;  /* [M,N2], row stride ldh2: saved activation of hidden layer 2 */
  const void* h1;  /* [M,N1], row stride ldh1 */
  void* dz2;       /* [M,N2] out, row stride ld_dz2 */
  void* dz1;       /* [M,N1] out, row stride ld_dz1 */
  int M, N1, N2, N3, ld_dz3, ldh2, ldh1, ld_dz2, ld_dz1, act;
} rslrl_dgrad_chain_t;
int rslrl_dgrad_chain16(int n, const rslrl_dgrad_chain_t* nets_host, rslrl_stream_t stream);

/* Backward of a NARROW layer (N <= 12: the action and value heads), data- and weight-gradient fused in one fp32
 * launch that reads h and writes dx exactly once.  Returns a non-zero code (message via rslrl_last_error) if a
 * problem does not qualify (N > 12, K % 4 != 0, unaligned rows); callers then use the two general calls. */
int rslrl_linear_backward_head(int n, const rslrl_linear_bwd_head_t* probs_host, rslrl_stream_t stream);
/* Returns 4 when a net's shape is not eligible (N1 in 128..512 step 128, N2 in {128,256}, N3 <= 128 step 32,
 * N4 <= 16, 16-byte aligned rows, same M for all nets): the caller then runs the layers one by one. */
int rslrl_mlp3_forward(int n, const rslrl_mlp3_fwd_t* nets_host, rslrl_stream_t stream);
/* Launch plan of rslrl_mlp3_forward for the rollout (PPO.act is called once per env step with new observation
 * pointers but otherwise identical arguments): create() validates the nets and encodes every TMA descriptor once;
 * run() only re-encodes the descriptor of an input whose pointer changed (x_host[i] replaces nets[i].x) and launches.
 * Weight / bias / output pointers are captured by value: re-create the plan when they move. */
int rslrl_mlp3_plan_create(int n, const rslrl_mlp3_fwd_t* nets_host, void** plan_out);
int rslrl_mlp3_plan_run(void* plan, const float* const* x_host, rslrl_stream_t stream);
void rslrl_mlp3_plan_destroy(void* plan);

/* ---- LSTM cell (rsl_rl/networks/memory.py:27-37 -> nn.LSTM) ------------------------------------------------ */

/* One time step of an LSTM layer after its two projections (PyTorch gate order i, f, g, o along the 4H axis):
 *   xg [n,4H] = x_t W_ih^T + b_ih (hoisted over all steps), hg [n,4H] = h_{t-1} W_hh^T, b_hh [4H] or NULL, c_prev [n,H]
 *   -> gates_act [n,4H] (activated gates, saved for the backward pass; NULL for inference), c_out [n,H], h_out [n,H].
 * H % 4 == 0, every pointer 16-byte aligned, rows dense. */
int rslrl_lstm_cell_forward(const float* xg, const float* hg, const float* b_hh, const float* c_prev, float* gates_act,
                            float* c_out, float* h_out, int64_t n, int H, rslrl_stream_t stream);
/* Backward of the same step: dh_out (gradient of the layer output at t, may be NULL), dh_rec (gradient arriving through
 * h_t from step t+1, may be NULL), dc_next (may be NULL) -> dgates [n,4H] (gradient of the pre-activation gates: the
 * operand of the dW_ih / dW_hh / db products and of dh_{t-1} = dgates W_hh) and dc_prev [n,H]. */
int rslrl_lstm_cell_backward(const float* dh_out, const float* dh_rec, const float* dc_next, const float* gates_act,
                             const float* c_prev, const float* c_cur, float* dgates, float* dc_prev, int64_t n, int H,
                             rslrl_stream_t stream);

/* Whole-layer drivers: the time loop runs in C, one grouped recurrent-projection launch + one cell launch per step for
 * 1 or 2 LSTM layers of equal shape (the actor's and the critic's memory advance together).
 * Forward: xg [T*n,4H] hoisted input projection, h_all / c_all [(T+1),n,H] with slot 0 = initial state (filled by the
 * caller) and slot t+1 = state after step t, gates [T,n,4H] or NULL (inference), hg [n,4H] scratch.
 * Backward through time: dout [T,n,H] / dhT / dcT (each may be NULL), w_hh_t [H,4H] = W_hh transposed (dh_{t-1} =
 * dgates_t W_hh runs as a forward product) -> dgates [T,n,4H]; dh_scratch / dc_scratch [2][n,H], slot 0 of each = d h_0 /
 * d c_0 on return (d h_0 only when need_dh0).  The weight gradients are two products over all T*n rows
 * (rslrl_linear_backward_weight with dgates against the layer input and against h_all[0..T-1]). */
typedef struct {
  const float* xg; const float* w_hh; const float* b_hh;
  float* h_all; float* c_all; float* gates; float* hg;
} rslrl_lstm_fwd_t;
typedef struct {
  const float* dout; const float* dhT; const float* dcT; const float* gates; const float* c_all; const float* w_hh_t;
  float* dgates; float* dh_scratch; float* dc_scratch;
} rslrl_lstm_bwd_t;
int rslrl_lstm_layers_forward(int n_nets, const rslrl_lstm_fwd_t* nets_host, int T, int64_t n, int H, int precision,
                              rslrl_stream_t stream);
int rslrl_lstm_layers_backward(int n_nets, const rslrl_lstm_bwd_t* nets_host, int need_dh0, int T, int64_t n, int H,
                               int precision, rslrl_stream_t stream);

/* ---- running-statistics normaliser (rsl_rl/networks/normalization.py:44-66) ---------------------------- */

/* EmpiricalNormalization.update(x) in one launch: batch mean / biased variance of x [N,D] (row stride ldx) folded into the
 * running buffers mean/var/std [D] and count (int64, device) with the reference's Chan update in its fp32 operation order;
 * `until` < 0 disables the freeze test, otherwise nothing changes once count >= until (tested on the device).
 * workspace: rslrl_empirical_norm_workspace_bytes(D) bytes, zero-initialised once. */
int64_t rslrl_empirical_norm_workspace_bytes(int D);
int rslrl_empirical_norm_update(const float* x, int64_t ldx, int64_t N, int D, float* mean, float* var, float* std,
                                int64_t* count, int64_t until, void* workspace, rslrl_stream_t stream);
/* EmpiricalNormalization.forward: y = (x - mean) / (std + eps), bit-exact in fp32; y row stride ldy >= D. */
int rslrl_empirical_normalize(const float* x, int64_t ldx, int64_t N, int D, const float* mean, const float* std,
                              double eps, float* y, int64_t ldy, rslrl_stream_t stream);

/* ---- PPO objective (rsl_rl/algorithms/ppo.py:255-315), forward + hand-derived backward ---------- */

/* rows = num_aug * mb_orig samples (symmetry augmentation repeats the per-sample scalars,
 * ppo.py:240-244: they are indexed i % mb_orig here instead of being materialised).
 *   mu [rows,A], value [rows], actions [rows,A]
 *   old_logp/advantages/returns/old_values [mb_orig], old_mu/old_sigma [mb_orig,A]
 *   std_param [A] (std or log_std according to std_type)
 *   adv_stats: NULL, or double[2] = {sum, sumsq} of `advantages` for the per-minibatch
 *              normalisation of ppo.py:221-223
 * outputs
 *   dmu [rows,A] (row stride dmu_stride >= A), dvalue [rows] (element stride dvalue_stride >= 1):
 *              dL/dmu, dL/dvalue of loss = surr + c_v*vloss - c_e*entropy; the strides let the caller keep
 *              16-byte row pitches for TMA (e.g. dvalue_stride = 4)
 *   dstd [A]   += dL/dstd_param   (accumulated)
 *   db_mu [A], db_value [1]: NULL or += column sums of dmu / dvalue (= bias gradients of the two output layers)
 *   kl_out [1]  = mean KL(old||new) over the first mb_orig rows as fp32 (ppo.py:262-269); may alias
 *                 the tail slot of the flat gradient buffer so that it rides the gradient all-reduce
 *   loss_sums  double[3] += {surrogate_mean, value_loss_mean, entropy_mean} of this minibatch
 *   workspace  rslrl_loss_workspace_bytes() bytes, zero-initialised once by the caller
 *   dmu16 / dvalue16 (may be NULL; need A % 4 == 0, A <= 16, 16-byte aligned rows): fp16 copies
 *              grad_scale16 * dmu, grad_scale16 * dvalue (row strides in elements, multiples of 8) = the scaled output
 *              gradients the RSLRL_PREC_FP16 backward starts from
 *   scale_state (may be NULL) device float[8], zero-initialised except [2] = the first scale: DYNAMIC power-of-two scale
 *              replacing grad_scale16 -- [0] scale S used by the most recent launch, [1] 1/S (pass its address as
 *              rslrl_linear_bwd_weight_t.scale_dev of the same minibatch), [2] S of the next launch (derived from the
 *              largest |gradient| this launch saw, so that S*max sits 16x below fp16's limit), [3] scratch,
 *              [4] uint32 count of launches whose scaled gradients saturated fp16 (the caller reports it) */
/* Optional fused DATA gradient of the two output layers (RSLRL_PREC_FP16 chain, A % 4 == 0, A <= 16, rows <= 148 * 1024):
 * with dz_actor != NULL the loss kernel also writes, as fp16 scaled like dmu16,
 *   dz_actor [rows,K3] = (dL/dmu W_actor) * act'(h_actor),  dz_critic [rows,K3] = (dL/dV W_critic) * act'(h_critic)
 * i.e. the gradient at the output of the last hidden layer (what rslrl_linear_backward_data on the K = A / K = 1 output
 * layers produced).  w_* are the fp32 output-layer weights [A,K3] / [1,K3], h_* the saved fp16 activations (row stride
 * ldh, as dz_*), K3 % 32 == 0, K3 <= 128, `act` the hidden activation code. */
typedef struct {
  const float* w_actor; const float* w_critic;
  const void* h_actor; const void* h_critic;
  void* dz_actor; void* dz_critic;
  int K3, ldh, act;
} rslrl_loss_head_t;
int64_t rslrl_loss_workspace_bytes(void);
int rslrl_ppo_loss_fwd_bwd(const float* mu, const float* std_param, int std_type, const float* value,
                           const float* actions, const float* old_logp, const float* advantages,
                           const float* returns, const float* old_values, const float* old_mu,
                           const float* old_sigma, const double* adv_stats, float* dmu, int dmu_stride,
                           float* dvalue, int dvalue_stride, float* dstd, float* db_mu, float* db_value,
                           float* kl_out, double* loss_sums, void* workspace, int64_t rows,
                           int64_t mb_orig, int A, double clip_param, double value_loss_coef,
                           double entropy_coef, int use_clipped_value_loss, void* dmu16, int dmu16_stride,
                           void* dvalue16, int dvalue16_stride, double grad_scale16, float* scale_state,
                           const rslrl_loss_head_t* head_host, rslrl_stream_t stream);

/* sum and sum of squares of x[0..n) into out[2] (double, overwritten): statistics for
 * ppo.py:221-223 per-minibatch advantage normalisation. */
int rslrl_sum_sumsq(const float* x, int64_t n, double* out, void* workspace, rslrl_stream_t stream);

/* ---- Random Network Distillation (rsl_rl/modules/rnd.py:113-135, get_intrinsic_reward) ---------------- */

/* One small MLP: num_layers Linear layers, w[l] is [dims[l+1], dims[l]] row-major, b[l] [dims[l+1]] or NULL, act[l] the
 * RSLRL_ACT_* code applied after layer l (0 after the last one).  Widths <= 128, at most 4 layers. */
typedef struct {
  const float* w[4];
  const float* b[4];
  int act[4];
  int dims[5];
  int num_layers;
} rslrl_rnd_net_t;

/* out[i] = scale * || target(x_i) - predictor(x_i) ||_2 for n rows of x (row stride ldx, floats): both embedding
 * networks, the gap, the norm and the weight in ONE launch (weights resident in shared memory, fp32 FMA). */
int rslrl_rnd_reward(const rslrl_rnd_net_t* target, const rslrl_rnd_net_t* predictor, const float* x, int64_t ldx,
                     int64_t n, float* out, double scale, rslrl_stream_t stream);

/* ---- optimiser (rsl_rl/algorithms/ppo.py:260-294 KL-adaptive LR, :380 clip_grad_norm_, :381 Adam) -- */

/* Device-resident optimiser scalars, one per parameter group (policy / RND predictor). */
typedef struct {
  double lr;          /* learning rate used by the NEXT adam kernel */
  double clip_scale;  /* clip_coef / world_size, applied to the (summed) gradient */
  double total_norm;  /* L2 norm of the averaged gradient before clipping */
  double step_size;   /* lr / (1 - beta1^t) */
  double bc2_sqrt;    /* sqrt(1 - beta2^t) */
  double kl_mean;     /* last averaged KL (diagnostic) */
  long long step;     /* Adam step count t */
  long long pad_;
} rslrl_opt_state_t;

/* Pass 1 over the flat gradient buffer g[0..P):  total_norm = ||g / world||_2,
 * clip_scale = min(1, max_norm / (total_norm + 1e-6)) / world   (torch clip_grad_norm_ semantics;
 * max_norm <= 0 disables clipping), then -- if adaptive != 0 -- the KL rule of ppo.py:280-284 on
 * kl = *kl_sum / world (lr /= 1.5 if kl > 2 d, lr *= 1.5 if 0 < kl < d/2, clamped to [1e-5,1e-2];
 * round_lr_fp32 != 0 reproduces the fp32 broadcast of ppo.py:288-290 in multi-GPU mode), then
 * step += 1 and the bias corrections.  All scalars stay on the device in *state. */
int64_t rslrl_optim_workspace_bytes(int64_t P);
int rslrl_grad_norm_lr(const float* g, int64_t P, const float* kl_sum, rslrl_opt_state_t* state, void* workspace,
                       int world_size, double max_grad_norm, int adaptive, double desired_kl, int round_lr_fp32,
                       double beta1, double beta2, rslrl_stream_t stream);

/* Pass 2: torch.optim.Adam (weight_decay=0, amsgrad=False) on the flat buffers with g' = g*clip_scale:
 *   m += (g'-m)(1-beta1);  v = beta2 v + (1-beta2) g'^2;  p -= step_size * m / (sqrt(v)/bc2_sqrt + eps)
 * zero_grad != 0 clears g afterwards (the backward kernels accumulate). */
int rslrl_adam_step(float* p, float* g, float* m, float* v, int64_t P, const rslrl_opt_state_t* state,
                    double beta1, double beta2, double eps, int zero_grad, rslrl_stream_t stream);

/* ---- data-parallel optimiser step in ONE launch (rsl_rl/algorithms/ppo.py:441-469 reduce_parameters, :272-290 KL
 * all-reduce + lr broadcast, :380 clip_grad_norm_, :381 Adam.step).
 * Every rank keeps its gradient buffer [P floats | 4-float tail whose first float is the minibatch KL] in memory that
 * its peers can address (NVLink peer / symmetric memory); peer_grads_host[k] is rank k's buffer as seen from this
 * rank (entry `rank` = the local buffer), peer_flags_host[k] rank k's zero-initialised flag buffer of
 * rslrl_allreduce_adam_flag_bytes() bytes (unused when world == 1).  The kernel sums the gradients of all ranks
 * (fixed rank order: bit-identical on every rank), derives ||mean g||, the clip scale and the KL-adaptive learning
 * rate exactly like rslrl_grad_norm_lr, applies Adam like rslrl_adam_step and clears the local gradients once the
 * peers have read them.  Must be launched by every rank the same number of times (flags are epoch stamped); the
 * grid (<= 148 blocks) has to be co-resident.  workspace: rslrl_allreduce_adam_workspace_bytes(), zero-initialised.
 * p16_shadow (may be NULL): fp16 copy of p, same element offsets, rewritten by the Adam phase (RSLRL_PREC_FP16).
 * peer_results_host (may be NULL): per rank a peer-addressable buffer of P floats.  When given, the reduction runs in two
 * steps inside every slice -- rank r sums sub-slice r over all ranks and writes the result into every rank's buffer,
 * then all ranks read the reduced slice locally -- which moves 2 (W-1)/W x P floats per rank instead of (W-1) x P
 * (worth it from 4 ranks up) and makes the sums bit-identical on all ranks by construction.
 * peer_stage_host (may be NULL): per rank a peer-addressable staging buffer of 2 * world * (P + 4) floats.  When given the
 * exchange is PUSHED: every rank writes its own gradients (its share of every slice) into the owners' staging buffers
 * and flags them, so a step costs one NVLink hop instead of the pull protocol's two ("is your backward complete?" +
 * the read round trip), and nobody reads a peer's gradient buffer any more. */
int64_t rslrl_allreduce_adam_workspace_bytes(void);
int64_t rslrl_allreduce_adam_flag_bytes(void);
int rslrl_allreduce_clip_adam(int world, int rank, const float* const* peer_grads_host,
                              unsigned int* const* peer_flags_host, float* const* peer_results_host,
                              float* const* peer_stage_host, float* p, float* m, float* v, int64_t P,
                              rslrl_opt_state_t* state, void* workspace, double max_grad_norm, int adaptive,
                              double desired_kl, int round_lr_fp32, int use_kl, double beta1, double beta2, double eps,
                              int zero_grad, void* p16_shadow, rslrl_stream_t stream);

/* ---- recurrent minibatches (rsl_rl/utils/utils.py:78-141; storage/rollout_storage.py:206-260) ---- */

/* Segmented scan over dones[T,N] (u8).  Trajectories are enumerated env-major (all pieces of env 0,
 * then env 1, ...); a piece ends at a done step and at t = T-1 (utils.py:97-98).  Outputs (int32):
 *   traj_id [T,N], traj_pos [T,N]   trajectory index and position inside it for every (t, env)
 *   traj_len / traj_env / traj_start [T*N capacity]  per trajectory: length, env, first time step
 *   env_traj_offset [N+1]            exclusive scan of trajectories per env; [N] = n_traj
 * workspace: rslrl_traj_workspace_bytes(T,N) bytes. */
int64_t rslrl_traj_workspace_bytes(int T, int N);
int rslrl_trajectory_index(const uint8_t* dones, int T, int N, int32_t* traj_id, int32_t* traj_pos,
                           int32_t* traj_len, int32_t* traj_env, int32_t* traj_start, int32_t* env_traj_offset,
                           void* workspace, rslrl_stream_t stream);
/* masks[t*capacity + j] = (first_traj + j < n_traj_end && traj_len[first_traj + j] > t), u8, j < capacity
 * (utils.py:129 trajectory_masks) */
int rslrl_trajectory_masks(const int32_t* traj_len, uint8_t* masks, int T, int first_traj, int n_traj_end,
                           int capacity, rslrl_stream_t stream);
/* split_and_pad_trajectories (utils.py:106-127) for trajectories [first_traj, first_traj+capacity):
 *   padded[p, j, :] = x[traj_start[J]+p, traj_env[J], :] if J = first_traj+j < n_traj_end and p < traj_len[J]
 *                     else 0;   x: [T,N,D], padded: [T,capacity,D] (every element written once). */
int rslrl_split_and_pad(const float* x, const int32_t* traj_len, const int32_t* traj_env,
                        const int32_t* traj_start, float* padded, int T, int N, int D, int first_traj,
                        int n_traj_end, int capacity, rslrl_stream_t stream);
/* unpad_trajectories (utils.py:134-141): out[t, env-env0, :] = padded[traj_pos[t,env], traj_id[t,env]-first_traj, :]
 * for env in [env0, env1); out: [T, env1-env0, D]. */
int rslrl_unpad(const float* padded, const int32_t* traj_id, const int32_t* traj_pos, float* out, int T, int N,
                int D, int env0, int env1, int first_traj, int capacity, rslrl_stream_t stream);
/* hidden states at trajectory starts (rollout_storage.py:238-250): saved [T,L,N,H] ->
 *   out[l, j, :] = saved[traj_start[J], l, traj_env[J], :],  J = first_traj + j < n_traj_end; out: [L,capacity,H] */
int rslrl_hidden_at_starts(const float* saved, const int32_t* traj_env, const int32_t* traj_start, float* out,
                           int T, int L, int N, int H, int first_traj, int n_traj_end, int capacity,
                           rslrl_stream_t stream);

/* ---- small utilities ------------------------------------------------------------------------- */
int rslrl_fill_zero(void* p, int64_t bytes, rslrl_stream_t stream);
/* dst[r, c] = (fp16, round to nearest even, saturating) src[r, c] for r < rows, c < cols; columns [cols, cols_pad) of
 * dst are written as zero.  Row strides in elements.  Used for the fp16 shadow of the flat parameter buffer
 * (rows = 1) and of the permuted minibatch observations (RSLRL_PREC_FP16). */
int rslrl_convert_f32_f16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int cols,
                          int cols_pad, rslrl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RSLRL_B200_H_ */
