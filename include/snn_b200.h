/*
 * snn_b200.h — C ABI of the B200 (sm_100a) spiking-actor hot path.
 *
 * Drop-in boundary for the population-coded spiking actor of
 * thisisnotahuman/FullySNNforLeggedRobots.  Every entry point replaces one
 * PyTorch-level call of the reference; citations are relative to
 *   AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/modules/actor_critic.py   ("AMP:<line>")
 *   AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/storage/rollout_storage.py ("STO:<line>")
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - every pointer is a DEVICE pointer unless the name ends in _host
 *   - every buffer is allocated and owned by the caller; kernels never allocate
 *   - `stream` is a cudaStream_t passed as void*
 *   - return 0 on success, a negative SNN_E* code otherwise; snn_last_error() gives text
 *   - there is no CPU fallback: SNN_EARCH is returned when the device is not sm_100
 */
#ifndef SNN_B200_H
#define SNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNN_MAX_LAYERS 8      /* hidden layers + output population layer */

enum {
  SNN_OK = 0,
  SNN_EINVAL = -1,   /* bad descriptor / null pointer / unsupported size */
  SNN_EARCH = -2,    /* device is not compute capability 10.x */
  SNN_ECUDA = -3,    /* a CUDA runtime call or launch failed */
  SNN_ENOMEM = -4    /* caller-provided workspace / trace too small */
};

/* Static description of one PopSpikeActor (AMP:274-298).  POD, copied by value. */
typedef struct SnnDesc {
  int32_t obs_dim;                 /* O: actor input width (AMP:277)                       */
  int32_t act_dim;                 /* A: number of actions                                 */
  int32_t en_pop;                  /* encoder population per obs dim (reference: 10)        */
  int32_t de_pop;                  /* decoder population per action  (reference: 10)        */
  int32_t num_hidden;              /* number of hidden LIF layers (AMP:192-214)            */
  int32_t hidden[SNN_MAX_LAYERS];  /* hidden layer widths                                  */
  int32_t spike_ts;                /* T: spike timesteps (reference: 5), 1..32             */
  int32_t dec_tanh;                /* 1: tanh on decoder output (CASSIE copy :147)         */
  float enc_eps;                   /* added to std^2 in the encoder (HUMANOID copy :107)   */
  int32_t fwd_planes;              /* bf16 planes per fp32 weight in forward: 3 = lossless
                                      ("fp32-parity mode"), 2 or 1 = faster, approximate   */
} SnnDesc;

/* Parameter pointers, fp32, in the reference's own layouts (state-dict tensors as they are). */
typedef struct SnnParams {
  const float* enc_mean;                 /* actor.encoder.mean   [1,O,P]  (AMP:85-89)       */
  const float* enc_std;                  /* actor.encoder.std    [1,O,P]  (AMP:90-92)       */
  const float* weight[SNN_MAX_LAYERS];   /* hidden_layers.i.weight / out_pop_layer.weight [N,K] */
  const float* bias[SNN_MAX_LAYERS];     /* same, .bias [N]                                 */
  const float* dec_w;                    /* actor.decoder.decoder.weight [A,1,P] (AMP:137)  */
  const float* dec_b;                    /* actor.decoder.decoder.bias   [A]                */
} SnnParams;

/* Gradient outputs (fp32, same shapes as SnnParams).  All are OVERWRITTEN, not accumulated. */
typedef struct SnnGrads {
  float* enc_mean;
  float* enc_std;
  float* weight[SNN_MAX_LAYERS];
  float* bias[SNN_MAX_LAYERS];
  float* dec_w;
  float* dec_b;
  float* obs;                            /* d loss / d obs [B,O]; may be NULL (RMA latent path only) */
} SnnGrads;

const char* snn_last_error(void);
int snn_abi_version(void);
/* bit 0: built with -DSNN_ABLATE (tools build with work-skipping SNN_DBG switches and barrier-wait cycle counters).
 * The release library returns 0; bench.py and the tests refuse any other value. */
int snn_build_flags(void);

/* Sizes of the caller-owned buffers, in bytes.  B = batch rows of the call. */
size_t snn_packed_weights_bytes(const SnnDesc* d);           /* bf16 plane images of all layers */
size_t snn_trace_bytes(const SnnDesc* d, int64_t B);         /* spike bits + 16-bit voltage-code traces */
size_t snn_workspace_bytes(const SnnDesc* d, int64_t B, int backward);

/* Split every fp32 weight matrix into bf16 hi/mid/lo planes laid out as swizzled tensor-core
 * operand tiles (for W and W^T).  Must be re-run after every optimizer step.
 * Replaces nothing in the reference: it is the price of feeding nn.Linear weights (AMP:208-214)
 * to tcgen05 MMAs without losing fp32 mantissa bits. */
int snn_pack_weights(const SnnDesc* d, const SnnParams* p, void* packed, void* stream);

/* PopSpikeActor.forward under inference_mode (AMP:300-320; called from ActorCritic.act /
 * act_inference, AMP:410-426): obs [B,O] -> mu [B,A].  No traces are kept. */
int snn_fwd_infer(const SnnDesc* d, const SnnParams* p, const void* packed, const float* obs,
                  int64_t B, float* mu, void* workspace, size_t workspace_bytes, void* stream);

/* The rollout step's forward as ONE kernel launch (PPO.act, AMP/.../algorithms/ppo.py:90-102 ->
 * ActorCritic.act, AMP:410-415 -> PopSpikeActor.forward, AMP:300-320): encoder, every LIF layer and the decoder run
 * inside one CTA per batch tile; spike words live in shared memory and never touch HBM, so there is no workspace.
 * mu is bit-identical to snn_fwd_infer's.  With noise != NULL (standard-normal draws [B,A] made by the caller) the
 * kernel also emits what Normal(mu, exp(log_std)).sample() / log_prob().sum(-1) / stddev give the reference
 * (AMP:398-421): actions = mu + sigma * noise, logp [B], sigma [B,A]  (act_dim <= 64); with noise == NULL the four
 * pointers after it are ignored.  snn_fwd_act_supported: 1 when the fused kernel takes this geometry (T * 16 <= 256
 * accumulator columns per neuron tile, layers <= 512 wide, operand buffers within shared memory), else 0 — callers
 * then use snn_fwd_infer.  snn_fwd_act_preferred: supported AND the batch is at most two waves of tiles (rollout-sized
 * batches); larger batches are faster on the layered kernels of snn_fwd_infer. */
int snn_fwd_act_supported(const SnnDesc* d, int64_t B);
int snn_fwd_act_preferred(const SnnDesc* d, int64_t B);
int snn_fwd_act(const SnnDesc* d, const SnnParams* p, const void* packed, const float* obs, int64_t B, float* mu,
                const float* noise, const float* log_std, float* actions, float* logp, float* sigma, void* stream);

/* Same forward with the traces BPTT needs (what autograd would save: AMP:154-167, 216-232).  mu may be NULL: the decoder
 * pass is skipped (the fused PPO step decodes inside snn_bwd_ppo_top). */
int snn_fwd_train(const SnnDesc* d, const SnnParams* p, const void* packed, const float* obs,
                  int64_t B, float* mu, void* trace, size_t trace_bytes,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Surrogate-gradient BPTT: what loss.backward() does to the reference graph.
 * g_mu [B,A] is d loss / d mu; mu is the forward output (needed for the tanh variant). */
int snn_bwd(const SnnDesc* d, const SnnParams* p, const void* packed, const float* obs,
            int64_t B, const float* mu, const float* g_mu, const void* trace,
            const SnnGrads* g, void* workspace, size_t workspace_bytes, void* stream);

/* ---- The fused PPO step's split of snn_bwd (r2).  Between the actor's last LIF layer and its backward chain the
 * reference runs the decoder, the whole PPO head and their autograd (AMP/.../algorithms/ppo.py:132-171,
 * modules/actor_critic.py:140-150, 398-421).  Nothing of the POLICY terms needs the critic, so the actor's chain does them
 * itself, right before the output layer's BPTT scan:
 *   snn_bwd_ppo_top  decoder (mu from the recorded output spikes; snn_fwd_train may be called with mu == NULL) + Normal
 *                    log-prob / ratio / clipped surrogate / KL / d loss / d mu per sample in one kernel, then the output
 *                    layer's scan.  Partial sums go to head->part (snn_ppo_head_part_bytes); mu_out / g_mu_out receive
 *                    mu and d loss / d mu (the scan reads them).
 *   snn_bwd_rest     every stage of snn_bwd below the output layer, dW, the deferred reductions.  Same workspace.
 *   snn_ppo_value_loss  the clipped value loss and g_value on the critic's side (ppo.py:160-168); part >= ceil(B/256)*4 bytes.
 *   snn_ppo_finalize    joins both partial sets: stats[0..6] and g_log_std as snn_ppo_head; kl_out (nullable) = KL mean.
 * Loss means are over B samples (the caller's data-parallel scale goes through snn_clip_adam's grad_scale). */
typedef struct SnnPpoActorHead {
  const float* log_std;     /* [A] */
  const float* actions;     /* [B,A] */
  const float* old_logp;    /* [B] */
  const float* adv;         /* [B] */
  const float* old_mu;      /* [B,A] */
  const float* old_sigma;   /* [B,A] */
  float clip;
  float* mu_out;            /* [B,A] */
  float* g_mu_out;          /* [B,A] */
  float* part;              /* snn_ppo_head_part_bytes(d, B) bytes */
  size_t part_bytes;
} SnnPpoActorHead;
size_t snn_ppo_head_part_bytes(const SnnDesc* d, int64_t B);
int snn_bwd_ppo_top(const SnnDesc* d, const SnnParams* p, const void* packed, const float* obs, int64_t B,
                    const SnnPpoActorHead* head, const void* trace, void* workspace, size_t workspace_bytes, void* stream);
int snn_bwd_rest(const SnnDesc* d, const SnnParams* p, const void* packed, const float* obs, int64_t B, const void* trace,
                 const SnnGrads* g, void* workspace, size_t workspace_bytes, void* stream);
int snn_ppo_value_loss(const float* value, const float* returns, const float* target_v, int64_t B, float clip, float value_coef,
                       int clipped_value, float* g_value, void* part, size_t part_bytes, void* stream);
int snn_ppo_finalize(const SnnDesc* d, int64_t B, const void* head_part, const void* value_part, const float* log_std,
                     float ent_coef, float* g_log_std, float* stats, float* kl_out, void* stream);

/* RolloutStorage.compute_returns (STO:124-138): reverse GAE scan over S steps and N envs plus
 * the global advantage normalisation.  rewards/values/returns/advantages are [S,N] fp32,
 * dones [S,N] uint8, last_values [N].  scratch: >= 16 KiB. */
int snn_gae(const float* rewards, const uint8_t* dones, const float* values,
            const float* last_values, int64_t S, int64_t N, float gamma, float lam,
            float* returns, float* advantages, int normalize, void* scratch, void* stream);

/* Minibatch permutation (STO:148-184: `tensor[batch_idx]` for 9 tensors, per epoch and minibatch): gathers rows
 * idx[0..n) of `count` (<= 12) row-major fields in ONE launch: dst[k][i] = src[k][idx[i]], rows of row_bytes[k]
 * bytes (multiples of 4).  idx is int64 on the device; src/dst/row_bytes are host arrays of device pointers / sizes. */
int snn_gather_rows(int count, const void* const* src, void* const* dst, const int32_t* row_bytes,
                    const int64_t* idx, int64_t n, void* stream);

/* ---- dense ELU MLP with one output: the critic (AMP:359-368, evaluate :428-430) -------------------------------
 * Linear -> ELU -> ... -> Linear(., 1).  Hidden layers run on tcgen05 with hi/lo bf16 operand planes (three
 * products, ~2^-16 relative); the 1-wide output layer and its backward are SIMT.  weight[i] / bias[i] for
 * i < num_hidden are the hidden layers [N,K] / [N]; index num_hidden is the output layer [1,N] / [1]. */
typedef struct SnnMlpDesc {
  int32_t in_dim;
  int32_t num_hidden;
  int32_t hidden[SNN_MAX_LAYERS];
  int32_t out_dim;                    /* 0 or 1: the critic (1-wide SIMT head); 2..128: linear output layer [out_dim, N] on
                                         the tensor cores (RMA env_factor_encoder / adaptation_module) */
} SnnMlpDesc;
typedef struct SnnMlpParams {
  const float* weight[SNN_MAX_LAYERS + 1];
  const float* bias[SNN_MAX_LAYERS + 1];
} SnnMlpParams;
typedef struct SnnMlpGrads {
  float* weight[SNN_MAX_LAYERS + 1];
  float* bias[SNN_MAX_LAYERS + 1];
} SnnMlpGrads;
size_t snn_mlp_packed_bytes(const SnnMlpDesc* d);
size_t snn_mlp_trace_bytes(const SnnMlpDesc* d, int64_t B);
size_t snn_mlp_workspace_bytes(const SnnMlpDesc* d, int64_t B);
int snn_mlp_pack(const SnnMlpDesc* d, const SnnMlpParams* p, void* packed, void* stream);
/* value [B] = critic(x [B,in_dim]); trace keeps the operand planes of x and of every hidden activation */
int snn_mlp_fwd(const SnnMlpDesc* d, const SnnMlpParams* p, const void* packed, const float* x, int64_t B,
                float* value, void* trace, size_t trace_bytes, void* stream);
/* gradients of sum_b g_value[b] * value[b] w.r.t. every weight / bias (overwritten) */
int snn_mlp_bwd(const SnnMlpDesc* d, const SnnMlpParams* p, const void* packed, int64_t B, const float* g_value,
                const void* trace, const SnnMlpGrads* g, void* workspace, size_t workspace_bytes, void* stream);
/* General form (RMA fork, RMA/.../rsl_rl/modules/actor_critic.py:394-418, 549-553, 633-634: the actor and the critic
 * read cat(obs, latent), latent = env_factor_encoder(privileged_obs), so both need a gradient into their input and
 * the encoder receives the SUM of the two):
 *   fwd_ex:  y [B, out_dim] (row stride y_stride floats: lands inside a wider cat buffer) = mlp(x [B, in_dim], row
 *            stride x_stride);
 *   bwd_ex:  gradients of sum_b <g_y[b] + g_y2[b], y[b]> (g_y2 nullable; strided rows) w.r.t. every weight / bias, and,
 *            when d_x is not null, w.r.t. the input rows (d_x [B, in_dim], row stride dx_stride, overwritten). */
int snn_mlp_fwd_ex(const SnnMlpDesc* d, const SnnMlpParams* p, const void* packed, const float* x, int64_t x_stride,
                   int64_t B, float* y, int64_t y_stride, void* trace, size_t trace_bytes, void* stream);
int snn_mlp_bwd_ex(const SnnMlpDesc* d, const SnnMlpParams* p, const void* packed, int64_t B, const float* g_y,
                   int64_t g_stride, const float* g_y2, int64_t g2_stride, const void* trace, const SnnMlpGrads* g,
                   float* d_x, int64_t dx_stride, void* workspace, size_t workspace_bytes, void* stream);

/* Fused PPO head (AMP/.../algorithms/ppo.py:132-171 with modules/actor_critic.py:398-421): Normal log-prob,
 * entropy, KL to the rollout policy, clipped surrogate and clipped value loss, forward AND closed-form
 * gradients in one pass.  Inputs [B,A] / [B] fp32.  Outputs: g_mu [B,A], g_value [B], g_log_std [A] =
 * d loss / d (mu, value, log_std) for loss = mean(surrogate) + value_coef*mean(value loss) - ent_coef*mean(entropy);
 * stats[0..3] = surrogate mean, value-loss mean, KL mean, entropy; stats[4..6] accumulate surrogate, value loss,
 * step count across calls (caller zeroes them).  scratch >= ceil(B/256)*(4+A)*4 bytes. */
int snn_ppo_head(const float* mu, const float* log_std, const float* value, const float* actions,
                 const float* old_logp, const float* adv, const float* returns, const float* target_v,
                 const float* old_mu, const float* old_sigma, int64_t B, int A, float clip, float value_coef,
                 float ent_coef, int clipped_value, float* g_mu, float* g_value, float* g_log_std, float* stats,
                 void* scratch, size_t scratch_bytes, void* stream);

/* Rollout step tail (PPO.act, ppo.py:90-102 -> actor_critic.py:398-421): sigma = exp(log_std) broadcast to [B,A],
 * actions = mu + sigma * noise (noise ~ N(0,1) drawn by the caller), logp[b] = sum_a Normal(mu, sigma).log_prob(action)
 * in torch.distributions.Normal's arithmetic order.  A <= 64. */
int snn_sample_logprob(const float* mu, const float* log_std, const float* noise, int64_t B, int A, float* actions,
                       float* logp, float* sigma, void* stream);

/* Adaptive-KL learning-rate rule (ppo.py:145-148) on a DEVICE learning rate: no host round trip.
 * kl_mean = kl[0] * kl_scale. */
int snn_lr_adapt(const float* kl, float kl_scale, float desired_kl, float* lr, void* stream);

/* clip_grad_norm_(max_norm) + Adam.step (ppo.py:176-177) over one flat fp32 buffer of n parameters.
 * grads are first multiplied by grad_scale (1/world after a SUM all-reduce).  lr and step live on the device;
 * step is incremented by this call.  norm_out (nullable) receives the pre-clip gradient norm.  scratch >= 2 KiB.
 * The betas are doubles: torch.optim.Adam forms (1 - beta) in double before rounding to fp32 (1 - 0.999 in fp32
 * arithmetic is 1.3e-5 off), and the second-moment buffer must match it. */
int snn_clip_adam(float* params, const float* grads, float* m, float* v, int64_t n, const float* lr, int* step,
                  double beta1, double beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                  float* norm_out, void* scratch, void* stream);
/* The same with the adaptive-KL learning-rate rule of snn_lr_adapt applied first, in the same launch (kl nullable:
 * no schedule).  lr is updated in place.  One kernel: lr rule, gradient norm, grid barrier, clip + Adam
 * (AMP/.../algorithms/ppo.py:145-148, 176-177). */
int snn_clip_adam_kl(float* params, const float* grads, float* m, float* v, int64_t n, float* lr, int* step,
                     double beta1, double beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                     const float* kl, float kl_scale, float desired_kl, float* norm_out, void* scratch, void* stream);

/* Data parallel (one process per GPU, one node): the gradient all-reduce INSIDE the optimizer kernel, over NVLink peer
 * memory instead of an NCCL call between the backward pass and snn_clip_adam.  Every rank owns a staging
 * buffer of 2 * stage_stride floats and a flag array of 8 ints (snn_p2p_alloc: cudaMalloc + zero + CUDA IPC handle; the
 * handles are exchanged by the caller, e.g. torch.distributed.all_gather_object, and opened with snn_p2p_open).
 * snn_clip_adam_kl_p2p: grads[0..n_reduce) <- sum over the ranks (rank order: bit-identical on every rank), then the
 * adaptive-KL rule on the summed grads[kl_index] (kl_index < 0: none), gradient norm, clip, Adam over params[0..n), all in
 * one launch.  stage / flags: host arrays of `world` device pointers (entry `rank` = the local buffers); epoch: a device
 * int, zero at start, owned by the kernel.  Every rank must make the same sequence of calls.  world 2..8. */
int snn_p2p_alloc(size_t bytes, void** ptr, void* handle64);
int snn_p2p_open(const void* handle64, void** ptr);
int snn_p2p_close(void* ptr);
int snn_p2p_free(void* ptr);
int snn_clip_adam_kl_p2p(float* params, float* grads, float* m, float* v, int64_t n, int64_t n_reduce, float* lr, int* step,
                         double beta1, double beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                         int64_t kl_index, float kl_scale, float desired_kl, float* norm_out, void* scratch,
                         void* const* stage, void* const* flags, int rank, int world, int64_t stage_stride, int* epoch,
                         void* stream);

/* Adaptation-module regression of the RMA fork (RMA/.../rsl_rl/algorithms/ppo.py:200-213): loss = F.mse_loss(pred,
 * target) over n = B*D contiguous elements, g = d loss / d pred = 2 (pred - target) / n; loss_accum[0] += loss.
 * scratch >= 1 KiB. */
int snn_mse_grad(const float* pred, const float* target, int64_t n, float* g, float* loss_accum, void* scratch,
                 void* stream);

/* Optional per-launch device timing used by bench.py's roofline: while enabled, every kernel launch is
 * bracketed by CUDA events on its stream.  Slot = kind + 4 * layer, kind: 0 = forward scan-GEMM, 1 = backward
 * dX scan-GEMM, 2 = dW GEMM, 3 = all other (SIMT) kernels (layer 0).  snn_profile_read synchronises on the
 * recorded events and returns, per slot (32 of them), accumulated milliseconds, ALGORITHMIC flops (2*B*K*N*T,
 * unpadded, one product per MAC) and launch counts since the last snn_profile_enable.  snn_launch_count counts
 * every kernel this library launched. */
int snn_profile_enable(int on);
int snn_profile_read(double* ms32, double* flops32, int64_t* launches32);
int64_t snn_launch_count(void);

/* Debug / test hooks: unpack internal buffers into plain tensors (tests only). */
/* -DSNN_ABLATE tools build only (SNN_EINVAL otherwise): device buffer of 16 * 148 * 16 uint64 receiving per-role
 * barrier-wait cycle counts (NULL = off) */
int snn_debug_set_cycle_buffer(void* buf);
int snn_debug_trace_layout(const SnnDesc* d, int64_t B, int64_t* out, int n_out);

/* Device-time of the kernels of the last fwd/bwd call is measured by the caller with CUDA
 * events on `stream`; the library itself never synchronises. */

#ifdef __cplusplus
}
#endif
#endif /* SNN_B200_H */
