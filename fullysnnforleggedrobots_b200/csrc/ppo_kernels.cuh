// ppo_kernels.cuh — the PPO head and the optimizer as fused kernels.
//
//   ppo_head_kernel / ppo_head_finalize_kernel
//       everything between the network outputs and their gradients in PPO.update
//       (AMP/.../algorithms/ppo.py:132-171 + modules/actor_critic.py:398-421): Normal log-prob, entropy,
//       KL to the rollout policy, clipped surrogate, clipped value loss — forward AND the closed-form
//       gradient w.r.t. mu, value and log_std, in one pass over the minibatch.  ~45 PyTorch kernels
//       and their autograd nodes in the reference.
//   lr_adapt_kernel
//       the adaptive-KL learning-rate schedule (ppo.py:139-151) without the host round trip.
//   adam_fused_kernel
//       adaptive lr + nn.utils.clip_grad_norm_ + torch.optim.Adam.step (ppo.py:145-148, 176-177) over ONE flat fp32
//       buffer in one launch (grid barrier between the norm and the update).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace snn {

struct PpoHeadParams {
  const float* mu;          // [B,A] actor output
  const float* log_std;     // [A]
  const float* value;       // [B]
  const float* actions;     // [B,A]
  const float* old_logp;    // [B]
  const float* adv;         // [B]
  const float* returns;     // [B]
  const float* target_v;    // [B] values stored in the rollout
  const float* old_mu;      // [B,A]
  const float* old_sigma;   // [B,A]
  int B, A;
  float clip, value_coef, ent_coef;
  int clipped_value;
  float inv_count;          // 1 / (B * world): loss means are over the GLOBAL minibatch
  float* g_mu;              // [B,A]
  float* g_value;           // [B]
  float* part;              // [gridDim.x][4 + A] partial sums: surrogate, value loss, kl, (unused), g_log_std[a]
};

constexpr int kHeadMaxA = 64;
constexpr float kHalfLog2Pi = 0.91893853320467274178f;

__device__ __forceinline__ float block_sum_256(float x, float* sh) {
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[w] = x;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 8) t = sh[threadIdx.x];
  if (threadIdx.x < 32) {
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in thread 0
}

__global__ void __launch_bounds__(256) ppo_head_kernel(const PpoHeadParams p) {
  __shared__ float s_sigma[kHeadMaxA], s_inv_var[kHeadMaxA], s_log_sigma[kHeadMaxA];
  if (threadIdx.x < p.A) {
    const float ls = p.log_std[threadIdx.x];
    const float s = expf(ls);
    s_sigma[threadIdx.x] = s;
    s_inv_var[threadIdx.x] = 1.f / (s * s);
    s_log_sigma[threadIdx.x] = logf(s);
  }
  __syncthreads();
  const int b = blockIdx.x * 256 + threadIdx.x;
  const bool live = b < p.B;
  float sur = 0.f, vl = 0.f, kl = 0.f, g_logp = 0.f;
  if (live) {
    const float* mu = p.mu + static_cast<size_t>(b) * p.A;
    const float* act = p.actions + static_cast<size_t>(b) * p.A;
    const float* omu = p.old_mu + static_cast<size_t>(b) * p.A;
    const float* osg = p.old_sigma + static_cast<size_t>(b) * p.A;
    float logp = 0.f;
    for (int a = 0; a < p.A; ++a) {
      const float d = act[a] - mu[a];
      logp += -(d * d) * 0.5f * s_inv_var[a] - s_log_sigma[a] - kHalfLog2Pi;
      const float os = osg[a], dm = omu[a] - mu[a];
      kl += logf(s_sigma[a] / os + 1.e-5f) + (os * os + dm * dm) * 0.5f * s_inv_var[a] - 0.5f;
    }
    const float ad = p.adv[b];
    const float ratio = expf(logp - p.old_logp[b]);
    const float s1 = -ad * ratio;
    const float s2 = -ad * fminf(fmaxf(ratio, 1.f - p.clip), 1.f + p.clip);
    sur = fmaxf(s1, s2);
    g_logp = (s1 >= s2 ? -ad : 0.f) * ratio * p.inv_count;     // d mean(sur) / d logp
    // value loss
    const float v = p.value[b], R = p.returns[b];
    const float e1 = v - R;
    float gv;
    if (p.clipped_value) {
      const float tv = p.target_v[b];
      const float dv = v - tv;
      const float vc = tv + fminf(fmaxf(dv, -p.clip), p.clip);
      const float e2 = vc - R;
      const float l1 = e1 * e1, l2 = e2 * e2;
      vl = fmaxf(l1, l2);
      const float inside = (dv >= -p.clip && dv <= p.clip) ? 1.f : 0.f;
      const float g1 = 2.f * e1, g2 = 2.f * e2 * inside;
      gv = l1 > l2 ? g1 : (l1 < l2 ? g2 : 0.5f * (g1 + g2));
    } else {
      vl = e1 * e1;
      gv = 2.f * e1;
    }
    p.g_value[b] = gv * p.value_coef * p.inv_count;
    float* gm = p.g_mu + static_cast<size_t>(b) * p.A;
    for (int a = 0; a < p.A; ++a) gm[a] = g_logp * (act[a] - mu[a]) * s_inv_var[a];
  }
  // all 3 + A block sums at once: warp shuffles per quantity, one shared-memory exchange, ONE barrier (fifteen
  // back-to-back two-barrier block reductions were most of this kernel's 11 us); fixed order -> reproducible
  __shared__ float s_w[8][4 + kHeadMaxA];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto warp_sum = [](float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  };
  {
    const float a0 = warp_sum(sur), a1 = warp_sum(vl), a2 = warp_sum(kl);
    if (lane == 0) { s_w[warp][0] = a0; s_w[warp][1] = a1; s_w[warp][2] = a2; s_w[warp][3] = 0.f; }
  }
  for (int a = 0; a < p.A; ++a) {
    float x = 0.f;
    if (live) {
      const float d = p.actions[static_cast<size_t>(b) * p.A + a] - p.mu[static_cast<size_t>(b) * p.A + a];
      x = g_logp * (d * d * s_inv_var[a] - 1.f);      // d logp / d log_std = (a-mu)^2/sigma^2 - 1
    }
    x = warp_sum(x);
    if (lane == 0) s_w[warp][4 + a] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4 + p.A) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_w[w][threadIdx.x];
    p.part[static_cast<size_t>(blockIdx.x) * (4 + p.A) + threadIdx.x] = t;
  }
}

// stats: [0] surrogate mean (local), [1] value-loss mean (local), [2] kl mean (local), [3] entropy mean,
//        [4] accumulated surrogate, [5] accumulated value loss, [6] number of accumulated steps
// g_log_std[a] = sum_b(...) - ent_coef / world   (entropy bonus: d(-c * mean_b sum_a log sigma_a)/d log_std = -c)
__global__ void __launch_bounds__(1024) ppo_head_finalize_kernel(const float* __restrict__ part, int nparts, int A, int B,
                                                                 const float* __restrict__ log_std, float ent_coef, float inv_world,
                                                                 float* __restrict__ g_log_std, float* __restrict__ stats) {
  // one warp per column of the partials (3 loss sums, entropy, A log-std gradients): lanes stride over the rows,
  // fixed-order shuffle reduction in double
  const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x >> 5; i < 4 + A; i += nwarps) {
    if (i == 3) {
      if (lane == 0) {
        float e = 0.f;
        for (int a = 0; a < A; ++a) e += 0.5f + kHalfLog2Pi + log_std[a];
        stats[3] = e;
      }
      continue;
    }
    double acc = 0.0;
    for (int s = lane; s < nparts; s += 32) acc += part[static_cast<size_t>(s) * (4 + A) + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane != 0) continue;
    if (i < 3) {
      const float mean = static_cast<float>(acc / B);
      stats[i] = mean;
      if (i == 0) stats[4] += mean;
      if (i == 1) stats[5] += mean;
      if (i == 2) stats[6] += 1.f;
    } else {
      g_log_std[i - 4] = static_cast<float>(acc) - ent_coef * inv_world;
    }
  }
}

// ---- the same head split along the two chains of the fused PPO step (r2): the policy terms are computed inside
// actor_head_kernel (simt_kernels.cuh) on the actor's stream, the value loss here on the critic's stream, and
// ppo_finalize_kernel joins their partial sums off the actor's critical path.
// value loss (ppo.py:160-168) and its gradient: g_value[b] = d (value_coef * mean(vl)) / d value[b]; part[blk] = sum of vl
__global__ void __launch_bounds__(256) value_loss_kernel(const float* __restrict__ value, const float* __restrict__ returns,
                                                        const float* __restrict__ target_v, int B, float clip, float value_coef,
                                                        int clipped_value, float inv_count, float* __restrict__ g_value,
                                                        float* __restrict__ part) {
  __shared__ float sh[8];
  const int b = blockIdx.x * 256 + threadIdx.x;
  float vl = 0.f;
  if (b < B) {
    const float v = value[b], R = returns[b];
    const float e1 = v - R;
    float gv;
    if (clipped_value) {
      const float tv = target_v[b];
      const float dv = v - tv;
      const float vc = tv + fminf(fmaxf(dv, -clip), clip);
      const float e2 = vc - R;
      const float l1 = e1 * e1, l2 = e2 * e2;
      vl = fmaxf(l1, l2);
      const float inside = (dv >= -clip && dv <= clip) ? 1.f : 0.f;
      const float g1 = 2.f * e1, g2 = 2.f * e2 * inside;
      gv = l1 > l2 ? g1 : (l1 < l2 ? g2 : 0.5f * (g1 + g2));
    } else {
      vl = e1 * e1;
      gv = 2.f * e1;
    }
    g_value[b] = gv * value_coef * inv_count;
  }
  const float t = block_sum_256(vl, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// head_part [nrows][4 + A] (actor_head_kernel: column 0 surrogate, 2 kl, 4.. log-std gradient sums), value_part [nrows_v]
// -> stats / g_log_std as ppo_head_finalize_kernel; kl_out (nullable) receives the KL mean as well (the tail of the
// gradient bucket: it rides in the data-parallel all-reduce and feeds the adaptive learning rate).
// One block of 1024 threads: thread = rows r, r + 1024, ... of head_part, all columns of a row in registers (the rows are
// contiguous: coalesced), then a fixed-order block reduction per column in double (thousands of rows: one per 8 samples).
__global__ void __launch_bounds__(1024) ppo_finalize_kernel(const float* __restrict__ head_part, int nrows,
                                                            const float* __restrict__ value_part, int nrows_v, int A, int B,
                                                            const float* __restrict__ log_std, float ent_coef, float inv_world,
                                                            float* __restrict__ g_log_std, float* __restrict__ stats,
                                                            float* __restrict__ kl_out) {
  __shared__ double s_w[32][4 + kHeadMaxA];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nc = 4 + A;
  // column i of this thread's rows (column 1 = the value loss comes from value_part)
  for (int c0 = 0; c0 < nc; c0 += 8) {
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0;
    for (int r = threadIdx.x; r < nrows; r += 1024) {
      const float* row = head_part + static_cast<size_t>(r) * nc + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < nc) acc[j] += row[j];
    }
    if (c0 == 0) {
      acc[1] = 0.0;
      for (int r = threadIdx.x; r < nrows_v; r += 1024) acc[1] += value_part[r];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], o);
      if (lane == 0 && c0 + j < nc) s_w[warp][c0 + j] = acc[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < nc) {
    const int i = threadIdx.x;
    double acc = 0.0;
    for (int w = 0; w < 32; ++w) acc += s_w[w][i];
    if (i == 3) {
      float e = 0.f;
      for (int a = 0; a < A; ++a) e += 0.5f + kHalfLog2Pi + log_std[a];
      stats[3] = e;
    } else if (i < 3) {
      const float mean = static_cast<float>(acc / B);
      stats[i] = mean;
      if (i == 0) stats[4] += mean;
      if (i == 1) stats[5] += mean;
      if (i == 2) { stats[6] += 1.f; if (kl_out != nullptr) kl_out[0] = mean; }
    } else {
      g_log_std[i - 4] = static_cast<float>(acc) - ent_coef * inv_world;
    }
  }
}

// ppo.py:145-148.  kl_mean = kl[0] * kl_scale (kl_scale = 1/world when kl[0] is a sum of rank means).
__global__ void lr_adapt_kernel(const float* __restrict__ kl, float kl_scale, float desired_kl, float* __restrict__ lr) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float k = kl[0] * kl_scale;
    float l = lr[0];
    if (k > desired_kl * 2.0f) l = fmaxf(1e-5f, l / 1.5f);
    else if (k < desired_kl / 2.0f && k > 0.0f) l = fminf(1e-2f, l * 1.5f);
    lr[0] = l;
  }
}

// adaptive lr (optional) + clip_grad_norm_(max_norm) + Adam (torch semantics: lerp first moment, eps outside the
// bias-corrected sqrt) in ONE launch — lr_adapt / sumsq / clip_adam were three dependent launches on the critical path of
// every minibatch step.  The grid is at most one block per SM, so all blocks are resident and a grid-wide barrier
// (counter + spin) can separate the two phases:
//   phase A  partial[blockIdx.x] = sum of (g * grad_scale)^2 over the block's slice (fixed order); block 0 also applies the
//            adaptive-KL schedule (ppo.py:145-148) to the DEVICE learning rate, advances the step counter and derives
//            Adam's bias corrections once (double-precision pow / sqrt): coefs[0] = lr / (1 - beta1^t), [1] = sqrt(1 - beta2^t)
//   barrier
//   phase B  every block re-derives the global norm from the partials (parallel, fixed order) and updates its slice.
// counters[0] = barrier arrivals, counters[1] = blocks done; the last block zeroes both (re-armed for the next launch).
constexpr int kAdamCoefs = 384;      // offset (floats) of the two coefficients in the scratch buffer; >= the partials
__global__ void __launch_bounds__(256) adam_fused_kernel(float* __restrict__ prm, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        float* __restrict__ partial, int* __restrict__ step,
                                                        float* __restrict__ lr, const float* __restrict__ kl, float kl_scale,
                                                        float desired_kl, double beta1d, double beta2d, float eps,
                                                        float weight_decay, float max_norm, float grad_scale,
                                                        float* __restrict__ norm_out, int* __restrict__ counters) {
  __shared__ float sh[8];
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  __shared__ double s_tot[8];
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;                 // contiguous slice per block
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * per, i1 = i0 + per < n ? i0 + per : n;
  if (blockIdx.x == 0 && threadIdx.x == 32) {      // a lane of warp 1: warp 0's lane 0 finishes the block sum below
    float l = lr[0];
    if (kl != nullptr) {
      const float k = kl[0] * kl_scale;
      if (k > desired_kl * 2.0f) l = fmaxf(1e-5f, l / 1.5f);
      else if (k < desired_kl / 2.0f && k > 0.0f) l = fminf(1e-2f, l * 1.5f);
      lr[0] = l;
    }
    const int s1 = step[0] + 1;
    const double t = static_cast<double>(s1);
    const double bc1 = 1.0 - pow(beta1d, t);
    const double bc2 = 1.0 - pow(beta2d, t);
    partial[kAdamCoefs] = static_cast<float>(static_cast<double>(l) / bc1);
    partial[kAdamCoefs + 1] = static_cast<float>(sqrt(bc2));
    step[0] = s1;
  }
  // four independent elements per trip in both phases: the loops are pure load latency (a one-element grid-stride loop
  // ran eight dependent round trips per thread: 10 us for 1.2 MB)
  float acc = 0.f;
  for (int64_t b = i0 + threadIdx.x; b < i1; b += 1024) {
    float x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = b + u * 256 < i1 ? g[b + u * 256] * grad_scale : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += x[u] * x[u];
  }
  const float t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
  // ---- grid barrier
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(counters, 1);
    while (atomicAdd(counters, 0) < static_cast<int>(gridDim.x)) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
  // torch.optim.Adam: exp_avg.lerp_(g, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2) with the
  // python-double (1 - beta) rounded to fp32 once
  const float beta2 = static_cast<float>(beta2d);
  const float omb1 = static_cast<float>(1.0 - beta1d), omb2 = static_cast<float>(1.0 - beta2d);
  {
    double tt = 0.0;
    for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += 256) tt += __ldcg(partial + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tt += __shfl_down_sync(0xffffffffu, tt, o);
    if ((threadIdx.x & 31) == 0) s_tot[threadIdx.x >> 5] = tt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += s_tot[i];
    const float norm = static_cast<float>(sqrt(tot));
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(1.f, max_norm / (norm + 1e-6f));
    s_coef = coef;
    s_step_size = __ldcg(partial + kAdamCoefs);
    s_bc2_sqrt = __ldcg(partial + kAdamCoefs + 1);
    if (blockIdx.x == 0 && norm_out != nullptr) norm_out[0] = norm;
  }
  __syncthreads();
  const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  for (int64_t b = i0 + threadIdx.x; b < i1; b += 1024) {
    float gi[4], pi[4], mi[4], vi[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = b + u * 256;
      const bool on = i < i1;
      gi[u] = on ? g[i] : 0.f; pi[u] = on ? prm[i] : 0.f; mi[u] = on ? m[i] : 0.f; vi[u] = on ? v[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = b + u * 256;
      if (i >= i1) continue;
      float gg = gi[u] * grad_scale * coef;
      if (weight_decay != 0.f) gg = gg + weight_decay * pi[u];
      const float mm = mi[u] + (gg - mi[u]) * omb1;
      const float vv = vi[u] * beta2 + gg * gg * omb2;
      m[i] = mm;
      v[i] = vv;
      const float denom = sqrtf(vv) / bc2_sqrt + eps;
      prm[i] = pi[u] - step_size * (mm / denom);
    }
  }
  if (threadIdx.x == 0 && atomicAdd(counters + 1, 1) == static_cast<int>(gridDim.x) - 1) {
    counters[0] = 0;
    counters[1] = 0;
  }
}

// ------------------------------------------------------------------ data parallel: gradient all-reduce + optimizer in ONE kernel
// Data-parallel PPO sums the flat gradient bucket (and the KL statistic in its tail) over the ranks between the backward
// pass and the optimizer (one NCCL all-reduce per minibatch step until r2: ~30 us of a 600 us step, nothing to
// overlap it with).  Here the collective is part of the optimizer kernel, over NVLink peer memory:
//   phase 0  every rank copies its bucket into its own STAGING buffer (peer-mapped, double-buffered by step parity), grid
//            barrier, then announces the step number in every peer's flag array (st.release.sys) and waits for all peers'
//            announcements (ld.acquire.sys, bounded at ~30 s: a lost peer traps instead of hanging the GPU)
//   phase A  the sum over the ranks, every element in rank order 0..W-1 and read back identically by every rank, so the
//            ranks' parameters stay bit-identical: two ranks read both buckets (one shot); more ranks reduce-scatter
//            (every rank sums its segment in place in its staging buffer), exchange flags once more and all-gather.  The
//            result goes back into the local bucket (the all-reduce's result) while the squared norm is accumulated;
//            block 0 applies the adaptive learning-rate rule to the summed KL tail and derives Adam's step coefficients
//   grid barrier, phase B = clip + Adam as in adam_fused_kernel.
// A rank re-writes a staging buffer two steps later, after the flag exchange of the step in between, which every peer
// enters only once it has finished reading: no barrier at the end.  1.2 MB per rank at C1.
constexpr int kP2PMaxWorld = 8;
struct P2PPeers {
  float* stage[kP2PMaxWorld];      // rank q's staging buffer [2][stage_stride] (own entry: local memory)
  int* flags[kP2PMaxWorld];        // rank q's flag array [kP2PMaxWorld]: slot r = last step announced by rank r
  int rank, world;
  long long stage_stride;          // floats per parity half
};
__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(int* counter, int target) {      // thread 0 of every block; all blocks resident
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1);
    while (atomicAdd(counter, 0) < target) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
}
__global__ void __launch_bounds__(256) p2p_adam_kernel(float* __restrict__ prm, float* __restrict__ g, float* __restrict__ m,
                                                      float* __restrict__ v, int64_t n, int64_t n_reduce,
                                                      float* __restrict__ partial, int* __restrict__ step, float* __restrict__ lr,
                                                      int64_t kl_index, float kl_scale, float desired_kl, double beta1d,
                                                      double beta2d, float eps, float weight_decay, float max_norm,
                                                      float grad_scale, float* __restrict__ norm_out, int* __restrict__ counters,
                                                      int* __restrict__ epoch, const P2PPeers peers, int one_shot) {
  __shared__ float sh[8];
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  __shared__ double s_tot[8];
  const int W = peers.world;
  const int ep = epoch[0] + 1;                                           // this step's number (written back at the end)
  const long long half = static_cast<long long>(ep & 1) * peers.stage_stride;
  const int64_t per = (n_reduce + gridDim.x - 1) / gridDim.x;            // contiguous slice per block
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * per, i1 = i0 + per < n_reduce ? i0 + per : n_reduce;
  // announce `what` to every peer (block 0) and wait until every peer has announced at least as much (all blocks)
  auto exchange = [&](int what) {
    if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(peers.flags[threadIdx.x] + peers.rank, what);
    if (threadIdx.x < W) {
      const int* f = peers.flags[peers.rank] + threadIdx.x;
      const long long t0 = clock64();
      while (ld_acquire_sys(f) - what < 0) {
        if (clock64() - t0 > 60000000000LL) __trap();                     // a peer never arrived (~30 s): fail, do not hang
        __nanosleep(32);
      }
    }
    __syncthreads();
  };
  // ---- phase 0: publish the local bucket
  {
    float* mine = peers.stage[peers.rank] + half;
    for (int64_t b = i0 + threadIdx.x; b < i1; b += 1024) {
      float x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = b + u * 256 < i1 ? g[b + u * 256] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (b + u * 256 < i1) mine[b + u * 256] = x[u];
    }
    __threadfence_system();
  }
  grid_barrier(counters, static_cast<int>(gridDim.x));
  exchange(2 * ep);
  // ---- phase A: sum over the ranks in rank order.  Two ranks: every rank reads both buckets (one shot).  More: every
  // rank sums ITS segment of the bucket from all ranks, in place in its staging buffer, a second exchange, then every
  // rank gathers the segments from their owners (reduce-scatter + all-gather: 2 x the bucket over NVLink per GPU instead
  // of W x).  Either way every element is one sum in rank order, read back identically by every rank.
  const bool two_shot = W > 2 && !one_shot;      // one_shot: the two-rank scheme at any world size (validation switch)
  const int64_t seg = (n_reduce + W - 1) / W;
  if (two_shot) {
    const int64_t s0 = seg * peers.rank, s1 = s0 + seg < n_reduce ? s0 + seg : n_reduce;
    const int64_t sper = (seg + gridDim.x - 1) / gridDim.x;
    const int64_t j0 = s0 + static_cast<int64_t>(blockIdx.x) * sper, j1 = j0 + sper < s1 ? j0 + sper : s1;
    float* mine = peers.stage[peers.rank] + half;
    for (int64_t b = j0 + threadIdx.x; b < j1; b += 1024) {
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      for (int q = 0; q < W; ++q) {
        const float* src = peers.stage[q] + half;
        float y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) y[u] = b + u * 256 < j1 ? ld_relaxed_sys(src + b + u * 256) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] += y[u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (b + u * 256 < j1) mine[b + u * 256] = x[u];
    }
    __threadfence_system();
    grid_barrier(counters + 3, static_cast<int>(gridDim.x));
    exchange(2 * ep + 1);
  }
  if (blockIdx.x == 0 && threadIdx.x == 32) {      // a lane of warp 1: warp 0's lane 0 finishes the block sum below
    float l = lr[0];
    if (kl_index >= 0) {
      float ks = 0.f;
      if (two_shot) {
        int owner = static_cast<int>(kl_index / seg);
        ks = ld_relaxed_sys(peers.stage[owner < W ? owner : W - 1] + half + kl_index);
      } else {
        for (int q = 0; q < W; ++q) ks += ld_relaxed_sys(peers.stage[q] + half + kl_index);
      }
      const float k = ks * kl_scale;
      if (k > desired_kl * 2.0f) l = fmaxf(1e-5f, l / 1.5f);
      else if (k < desired_kl / 2.0f && k > 0.0f) l = fminf(1e-2f, l * 1.5f);
      lr[0] = l;
    }
    const int s1 = step[0] + 1;
    const double t = static_cast<double>(s1);
    const double bc1 = 1.0 - pow(beta1d, t);
    const double bc2 = 1.0 - pow(beta2d, t);
    partial[kAdamCoefs] = static_cast<float>(static_cast<double>(l) / bc1);
    partial[kAdamCoefs + 1] = static_cast<float>(sqrt(bc2));
    step[0] = s1;
  }
  float acc = 0.f;
  for (int64_t b = i0 + threadIdx.x; b < i1; b += 1024) {
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (two_shot) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = b + u * 256;
        if (i < i1) {
          const int owner = static_cast<int>(i / seg);
          x[u] = ld_relaxed_sys(peers.stage[owner] + half + i);
        }
      }
    } else {
      for (int q = 0; q < W; ++q) {
        const float* src = peers.stage[q] + half;
        float y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) y[u] = b + u * 256 < i1 ? ld_relaxed_sys(src + b + u * 256) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] += y[u];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = b + u * 256;
      if (i < i1) {
        g[i] = x[u];
        if (i < n) { const float z = x[u] * grad_scale; acc += z * z; }
      }
    }
  }
  const float t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
  grid_barrier(counters + 1, static_cast<int>(gridDim.x));
  // ---- phase B: clip + Adam (adam_fused_kernel)
  const float beta2 = static_cast<float>(beta2d);
  const float omb1 = static_cast<float>(1.0 - beta1d), omb2 = static_cast<float>(1.0 - beta2d);
  {
    double tt = 0.0;
    for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += 256) tt += __ldcg(partial + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tt += __shfl_down_sync(0xffffffffu, tt, o);
    if ((threadIdx.x & 31) == 0) s_tot[threadIdx.x >> 5] = tt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += s_tot[i];
    const float norm = static_cast<float>(sqrt(tot));
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(1.f, max_norm / (norm + 1e-6f));
    s_coef = coef;
    s_step_size = __ldcg(partial + kAdamCoefs);
    s_bc2_sqrt = __ldcg(partial + kAdamCoefs + 1);
    if (blockIdx.x == 0 && norm_out != nullptr) norm_out[0] = norm;
  }
  __syncthreads();
  const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const int64_t j1 = i1 < n ? i1 : n;
  for (int64_t b = i0 + threadIdx.x; b < j1; b += 1024) {
    float gi[4], pi[4], mi[4], vi[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = b + u * 256;
      const bool on = i < j1;
      gi[u] = on ? g[i] : 0.f; pi[u] = on ? prm[i] : 0.f; mi[u] = on ? m[i] : 0.f; vi[u] = on ? v[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = b + u * 256;
      if (i >= j1) continue;
      float gg = gi[u] * grad_scale * coef;
      if (weight_decay != 0.f) gg = gg + weight_decay * pi[u];
      const float mm = mi[u] + (gg - mi[u]) * omb1;
      const float vv = vi[u] * beta2 + gg * gg * omb2;
      m[i] = mm;
      v[i] = vv;
      const float denom = sqrtf(vv) / bc2_sqrt + eps;
      prm[i] = pi[u] - step_size * (mm / denom);
    }
  }
  if (threadIdx.x == 0 && atomicAdd(counters + 2, 1) == static_cast<int>(gridDim.x) - 1) {
    counters[0] = 0; counters[1] = 0; counters[2] = 0; counters[3] = 0;
    epoch[0] = ep;
  }
}

// ------------------------------------------------------------------ adaptation-module regression (RMA fork)
// F.mse_loss(pred, target) and its gradient (RMA/.../rsl_rl/algorithms/ppo.py:200-213): loss = mean over all n = B*D
// elements of (pred - target)^2; g = 2 (pred - target) / n.  Block partial sums in a fixed order (bit-reproducible).
__global__ void __launch_bounds__(256) mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                      int64_t n, float scale, float* __restrict__ g,
                                                      float* __restrict__ partial) {
  __shared__ float sh[8];
  float acc = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * 256) {
    const float d = pred[i] - target[i];
    acc += d * d;
    g[i] = d * scale;
  }
  const float t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void mse_finalize_kernel(const float* __restrict__ partial, int nparts, double inv_n, float* __restrict__ loss_accum) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nparts; ++i) t += partial[i];
    loss_accum[0] += static_cast<float>(t * inv_n);
  }
}

// ------------------------------------------------------------------ rollout step: sample + log-prob in one pass
// ActorCritic.act / get_actions_log_prob / action_std (actor_critic.py:398-421) on the actor's (mu, log_std):
//   sigma = exp(log_std) ; action = mu + sigma * noise ; log_prob = sum_a Normal(mu, sigma).log_prob(action)
// with torch.distributions.Normal's arithmetic order: -((x - mu)^2) / (2 sigma^2) - log(sigma) - log(sqrt(2 pi)).
// One thread per sample (A <= 64); noise is drawn by the caller (torch's Philox stream, CUDA-graph safe).
__global__ void __launch_bounds__(128) sample_logprob_kernel(const float* __restrict__ mu, const float* __restrict__ log_std,
                                                            const float* __restrict__ noise, int B, int A,
                                                            float* __restrict__ actions, float* __restrict__ logp,
                                                            float* __restrict__ sigma_out) {
  __shared__ float s_sig[64], s_two_var[64], s_logsig[64];
  if (threadIdx.x < A) {
    const float sg = expf(log_std[threadIdx.x]);
    s_sig[threadIdx.x] = sg;
    s_two_var[threadIdx.x] = 2.f * (sg * sg);
    s_logsig[threadIdx.x] = logf(sg);
  }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float lp = 0.f;
  for (int a = 0; a < A; ++a) {
    const size_t i = static_cast<size_t>(b) * A + a;
    const float m = mu[i], sg = s_sig[a];
    const float x = __fadd_rn(m, __fmul_rn(sg, noise[i]));
    const float d = __fsub_rn(x, m);
    lp += __fsub_rn(__fsub_rn(__fdiv_rn(-(d * d), s_two_var[a]), s_logsig[a]), 0.91893853320467267f);
    actions[i] = x;
    sigma_out[i] = sg;
  }
  logp[b] = lp;
}

}  // namespace snn
