// simt_kernels.cuh — the HBM-bound elementwise / reduction kernels around the tensor-core GEMMs:
// weight packing, population encoder, decoder, top-layer BPTT scan, encoder / decoder gradients,
// GAE.  All are plain coalesced SIMT kernels; none of them is GEMM-shaped.
#pragma once
#include "plan.cuh"
#include "umma.cuh"

namespace snn {

// ------------------------------------------------------------------ weight packing
// W [N][K] fp32  ->  w_img: 3 bf16 planes (hi, mid, lo; hi+mid+lo == W exactly) as swz64 image with
// rows n (forward B operand) ; wt_img: 2 planes (hi, lo) of W^T as swz64 image with rows k
// (backward dX B operand) ; bias padded to NP.
__device__ __forceinline__ void split3(float w, float& h, float& m, float& l) {
  h = bf16_round(w);
  const float r1 = w - h;
  m = bf16_round(r1);
  l = bf16_round(r1 - m);
}

// all layers of a network in ONE launch: blockIdx.y selects the layer
struct PackJob { const float* W; const float* bias; int N, K, NP, KP; uint8_t* w_img; size_t w_plane; uint8_t* wt_img; size_t wt_plane; float* bias_pad; };
// enc_row >= 0: blockIdx.y == enc_row builds the encoder table of the fused rollout kernel (act_fused.cuh) instead:
// per encoder neuron k  {mean, var = std^2 + eps, -0.5 * rcp(var), obs index k / P as int bits (-1: padding neuron)} —
// the same per-neuron constants encode_kernel derives in shared memory
struct PackJobs {
  PackJob j[SNN_MAX_LAYERS + 1];
  int enc_row;
  const float* enc_mean; const float* enc_std; float4* enc_tab; int K0, K0P, P; float eps;
};
__global__ void pack_w_kernel(const PackJobs jobs) {
  if (static_cast<int>(blockIdx.y) == jobs.enc_row) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < jobs.K0P) {
      const bool live = k < jobs.K0;
      const float sk = live ? jobs.enc_std[k] : 1.f;
      const float var = __fadd_rn(__fmul_rn(sk, sk), jobs.eps);
      jobs.enc_tab[k] = make_float4(live ? jobs.enc_mean[k] : 0.f, var, -0.5f * __frcp_rn(var),
                                    __int_as_float(live ? k / jobs.P : -1));
    }
    return;
  }
  const PackJob& jb = jobs.j[blockIdx.y];
  const float* __restrict__ W = jb.W;
  const float* __restrict__ bias = jb.bias;
  const int N = jb.N, K = jb.K, NP = jb.NP, KP = jb.KP;
  uint8_t* __restrict__ w_img = jb.w_img;
  uint8_t* __restrict__ wt_img = jb.wt_img;
  const size_t w_plane = jb.w_plane, wt_plane = jb.wt_plane;
  float* __restrict__ bias_pad = jb.bias_pad;
  if (static_cast<int>(blockIdx.x * blockDim.x) >= NP * (KP / 8)) return;      // grid.x is sized for the largest layer
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int kp8 = KP / 8;
  if (idx < NP * kp8) {   // forward image: one 16-byte piece = 8 consecutive k of row n
    const int n = idx / kp8, k0 = (idx - n * kp8) * 8;
    float h[8], m[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      const float w = (n < N && k < K) ? W[static_cast<size_t>(n) * K + k] : 0.f;
      split3(w, h[j], m[j], l[j]);
    }
    const size_t off = swz64_offset(n, k0, NP);
    *reinterpret_cast<uint4*>(w_img + off) =
        make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
    *reinterpret_cast<uint4*>(w_img + w_plane + off) =
        make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
    *reinterpret_cast<uint4*>(w_img + 2 * w_plane + off) =
        make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
  }
  const int np8 = NP / 8;
  if (idx < KP * np8) {   // transposed image: one piece = 8 consecutive n of row k
    const int k = idx / np8, n0 = (idx - k * np8) * 8;
    float h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + j;
      const float w = (n < N && k < K) ? W[static_cast<size_t>(n) * K + k] : 0.f;
      h[j] = bf16_round(w);
      l[j] = w - h[j];
    }
    const size_t off = swz64_offset(k, n0, KP);
    *reinterpret_cast<uint4*>(wt_img + off) =
        make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
    *reinterpret_cast<uint4*>(wt_img + wt_plane + off) =
        make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
  }
  if (idx < NP) bias_pad[idx] = idx < N ? bias[idx] : 0.f;
}

// ------------------------------------------------------------------ population encoder
// actor_critic.py:94-122: a = exp(-1/2 (x-mean)^2 / std^2) ; T steps of u += a; s = u > .999; u -= .999 s.
// One warp = 32 consecutive encoder neurons of one sample; __ballot packs its spike word.
__device__ __forceinline__ float pop_activation(float x, float mean, float std, float eps) {
  const float d = __fsub_rn(x, mean);
  const float num = __fmul_rn(-0.5f, __fmul_rn(d, d));
  const float var = __fadd_rn(__fmul_rn(std, std), eps);
  const float q = __fdiv_rn(num, var);
  return static_cast<float>(exp(static_cast<double>(q)));   // correctly rounded fp32 exp
}

// fp32 activation for gradient use (no spike decision depends on its last ulp)
__device__ __forceinline__ float pop_activation_fast(float x, float mean, float std, float eps) {
  const float d = x - mean;
  return expf(-0.5f * d * d / (std * std + eps));
}

template <int TT>
__device__ __forceinline__ uint32_t regular_spikes(float a, int T, float& margin) {
  float u = 0.f;
  uint32_t bits = 0;
  margin = 1.f;
  if (TT > 0) {
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      u = __fadd_rn(u, a);
      margin = fminf(margin, fabsf(u - 0.999f));
      const bool s = u > 0.999f;
      bits |= s ? (1u << t) : 0u;
      u = s ? __fsub_rn(u, 0.999f) : u;
    }
  } else {
    for (int t = 0; t < T; ++t) {
      u = __fadd_rn(u, a);
      margin = fminf(margin, fabsf(u - 0.999f));
      if (u > 0.999f) { bits |= 1u << t; u = __fsub_rn(u, 0.999f); }
    }
  }
  return bits;
}

// batch-tile geometry of the actor's time-stepped tensors (plan.cuh): sample-step (t, b) -> tile * CT + t * Bt + b % Bt
struct TileGeom { int Bt, CT, ntile; long long Rt; };

// TT = 5: fully unrolled for the reference's spike_ts; TT = 0: generic T.
// Thread = (sample, spike word): it evaluates the word's 32 encoder neurons (independent chains -> ILP, no
// __ballot), and a warp = 32 consecutive samples stores each timestep's words as one coalesced run.
// grid (ceil(Bcap / 256), K0P / 32), block 256.
template <int TT>
__global__ void __launch_bounds__(256) encode_kernel(const float* __restrict__ obs, const float* __restrict__ mean,
                                                    const float* __restrict__ std, int B, TileGeom tg, int O, int P,
                                                    int K0, int T, float eps, uint32_t* __restrict__ bits) {
  __shared__ float s_mk[32], s_var[32], s_nh[32];
  __shared__ int s_o[32];
  const int kw = blockIdx.y;
  if (threadIdx.x < 32) {
    const int k = kw * 32 + threadIdx.x;
    const bool k_live = k < K0;
    const float sk = k_live ? std[k] : 1.f;
    const float var = __fadd_rn(__fmul_rn(sk, sk), eps);
    s_mk[threadIdx.x] = k_live ? mean[k] : 0.f;
    s_var[threadIdx.x] = var;
    s_nh[threadIdx.x] = -0.5f * __frcp_rn(var);
    s_o[threadIdx.x] = k_live ? k / P : -1;                 // -1: padding neuron, never spikes
  }
  __syncthreads();
  const int b = blockIdx.x * 256 + threadIdx.x;
  if (b >= tg.ntile * tg.Bt) return;
  const bool b_live = b < B;
  const int nt = TT > 0 ? TT : T;
  // Pass 1 (cheap, all 32 neurons): which receptive fields can fire at all?  A neuron whose activation a satisfies
  // a * T < 0.999 never reaches the threshold; with Gaussian fields a tenth of the population range apart that is
  // ~7 of 8 neurons.  The test uses a 20 % guard band (a > 0.8 * 0.999 / T), far wider than the 1e-6 error of the
  // fast exponent, so the skipped neurons are exactly the silent ones (their membrane stays >= 0.19 below threshold).
  const float q_min = __logf(0.8f * 0.999f / static_cast<float>(nt));
  uint32_t active = 0;
  {
    int o_prev = -2;
    float x = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int o = s_o[i];
      if (b_live && o >= 0) {
        if (o != o_prev) { x = obs[static_cast<size_t>(b) * O + o]; o_prev = o; }
        const float d = __fsub_rn(x, s_mk[i]);
        active |= (d * d * s_nh[i] > q_min) ? (1u << i) : 0u;
      }
    }
  }
  // Pass 2 (only the candidates; a warp iterates max-over-lanes candidate count times, ~8 instead of 32): the spike
  // generator, accumulated straight into the T output words of this (sample, word)
  uint32_t w[TT > 0 ? TT : 32];
#pragma unroll
  for (int t = 0; t < (TT > 0 ? TT : 32); ++t) w[t] = 0;
  while (active) {
    const int i = __ffs(active) - 1;
    active &= active - 1;
    const float x = obs[static_cast<size_t>(b) * O + s_o[i]];
    const float d = __fsub_rn(x, s_mk[i]);
    // fast path: a within ~1e-6 relative of the reference's exp(((-0.5 d^2) / var)); the spike train can only
    // differ from the exact one if some membrane value comes within ~5e-6 of the threshold
    float margin;
    uint32_t sp = regular_spikes<TT>(__expf(d * d * s_nh[i]), T, margin);
    if (margin < 4e-5f) {
      // reference arithmetic order (actor_critic.py:105): ((-0.5 * d^2) / std^2), then a correctly rounded exp
      const float q = __fdiv_rn(__fmul_rn(-0.5f, __fmul_rn(d, d)), s_var[i]);
      sp = regular_spikes<0>(static_cast<float>(exp(static_cast<double>(q))), T, margin);
    }
    if (TT > 0) {
#pragma unroll
      for (int t = 0; t < TT; ++t) w[t] |= ((sp >> t) & 1u) << i;
    } else {
      for (int t = 0; t < T; ++t) w[t] |= ((sp >> t) & 1u) << i;
    }
  }
  const int tile = b / tg.Bt, bl = b - tile * tg.Bt;
  uint32_t* dst = bits + static_cast<size_t>(kw) * tg.Rt + static_cast<size_t>(tile) * tg.CT + bl;
  if (TT > 0) {
#pragma unroll
    for (int t = 0; t < TT; ++t) dst[static_cast<size_t>(t) * tg.Bt] = w[t];
  } else {
    for (int t = 0; t < T; ++t) dst[static_cast<size_t>(t) * tg.Bt] = w[t];
  }
}

// ------------------------------------------------------------------ decoder
// actor_critic.py:140-150: rate = spike count / T ; mu[b,a] = sum_p w[a,p] rate[b, a P + p] + bias[a] (; tanh)
__global__ void __launch_bounds__(256) decode_kernel(const uint32_t* __restrict__ bits, TileGeom tg, int B, int A,
                                                    int P, int T, const float* __restrict__ dec_w,
                                                    const float* __restrict__ dec_b, int dec_tanh, float* __restrict__ mu) {
  // one thread per (action, sample), samples fastest: the <= 2 spike words an action's population spans are
  // loaded coalesced across samples
  // rate = count / T takes T + 1 values: one IEEE division per table entry instead of one per (neuron, sample)
  // (the kernel is instruction-bound; same quotients, bit for bit)
  __shared__ float s_rate[kMaxT + 1];
  if (threadIdx.x <= T) s_rate[threadIdx.x] = __fdiv_rn(static_cast<float>(threadIdx.x), static_cast<float>(T));
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * A) return;
  const int a0 = idx / B, b = idx - a0 * B;
  const long long R = tg.Rt;
  const int tile = b / tg.Bt;
  const size_t r0 = static_cast<size_t>(tile) * tg.CT + (b - tile * tg.Bt);     // sample-step (0, b)
  for (int a = a0; a == a0; ++a) {
    const int n0 = a * P;
    const int w0 = n0 >> 5, w1 = (n0 + P - 1) >> 5, sh = n0 & 31;
    float acc = 0.f;
    if (P <= 16) {
      int cnt[16];
#pragma unroll
      for (int pp = 0; pp < 16; ++pp) cnt[pp] = 0;
      for (int t = 0; t < T; ++t) {
        const size_t r = r0 + static_cast<size_t>(t) * tg.Bt;
        const uint32_t lo = __ldg(bits + static_cast<size_t>(w0) * R + r);
        const uint32_t hi = w1 != w0 ? __ldg(bits + static_cast<size_t>(w1) * R + r) : 0u;
        const uint32_t x = __funnelshift_r(lo, hi, sh);      // bits n0 .. n0+31 of this sample at step t
#pragma unroll
        for (int pp = 0; pp < 16; ++pp) cnt[pp] += (x >> pp) & 1u;
      }
#pragma unroll
      for (int pp = 0; pp < 16; ++pp)
        if (pp < P) acc = __fmaf_rn(dec_w[n0 + pp], s_rate[cnt[pp]], acc);
    } else {
      for (int pp = 0; pp < P; ++pp) {
        const int n = n0 + pp;
        int cnt = 0;
        for (int t = 0; t < T; ++t)
          cnt += (__ldg(bits + static_cast<size_t>(n >> 5) * R + r0 + static_cast<size_t>(t) * tg.Bt) >> (n & 31)) & 1u;
        acc = __fmaf_rn(dec_w[n], s_rate[cnt], acc);
      }
    }
    acc = __fadd_rn(acc, dec_b[a]);
    mu[static_cast<size_t>(b) * A + a] = dec_tanh ? tanhf(acc) : acc;
  }
}

// ------------------------------------------------------------------ BPTT of the output population layer
// g_up[t] = G[b,a] w_dec[a,p] / T for every t (out_act = sum_t s_t / T feeds the grouped conv), then the recurrence of
// SURVEY.md 8a.  Orientation: thread = (PAIR of neurons, one sample), warp = the 64 neurons of one chunk of one sample,
// block = kTopRowsSPB samples (the dX epilogue's thread = neuron orientation is dictated by the TMEM lanes; here
// nothing dictates it).  Every request is a full row segment: a warp loads the 64 voltages of a sample-step as two
// 128-byte lines and stores each g_c plane row as ONE 128-byte line (4-byte bf16x2 stores).  The state per thread is
// two short recurrences (~40 registers -> full occupancy; the kernel is latency-bound).
// Writes the g_c hi/lo planes (swz64 image, rows = sample-steps) and per-block partials of the bias gradient and of the
// decoder gradients:  dec_part[blk][0][n] = sum_b G[b,a(n)] rate[b,n]  (d w_dec),  dec_part[blk][1][n] = sum_b G[b,a(n)]
// (d bias, read at n = a * P); the block's samples are added in shared memory in a fixed order.
// grid (Bcap / kTopRowsSPB, NP / 64), block 32 * kTopRowsSPB.
constexpr int kTopRowsSPB = 8;       // Bt % 16 == 0: the samples of a block lie in one tile (16: row reduction 12.8 -> 7.1 us but this kernel 36 -> 40 us: a wash)
template <int TT>
__global__ void __launch_bounds__(32 * kTopRowsSPB) top_scan_rows_kernel(const float* __restrict__ g_mu, const float* __restrict__ mu,
                                                           int dec_tanh, const float* __restrict__ dec_w,
                                                           const uint16_t* __restrict__ volt, int B, TileGeom tg, int A,
                                                           int P, int NP, int T, uint8_t* __restrict__ g_img,
                                                           size_t g_plane, float* __restrict__ db_part /* [gridDim.x][NP] */,
                                                           float* __restrict__ dec_part /* [gridDim.x][2][NP] */,
                                                           uint32_t p_magic /* ceil(2^20 / P), or 0: divide */) {
  constexpr int SPB = kTopRowsSPB;
  __shared__ float s_red[SPB][6][32];
  __shared__ float s_rate[kMaxT + 1];        // count / T for every possible spike count (the kernel is issue-bound)
  if (threadIdx.x <= T) s_rate[threadIdx.x] = __fdiv_rn(static_cast<float>(threadIdx.x), static_cast<float>(T));
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.y;
  const int n0 = chunk * 64 + 2 * lane;                        // this thread's neurons n0, n0 + 1
  // sample blocks are walked last to first: the forward pass wrote this layer's voltages in ascending order, so its
  // tail is what the L2 still holds (and the dX kernel that follows starts on the g_c rows written last here)
  const int b = static_cast<int>(gridDim.x - 1 - blockIdx.x) * SPB + w;      // its sample (< Bcap)
  const int tile = b / tg.Bt, bl = b - tile * tg.Bt;
  const int nt = TT > 0 ? TT : T;
  // the kernel is bound by instruction issue: no IEEE division per neuron (G w / T as a reciprocal product, within 1 ulp)
  // and n / P by a multiply-shift (exact for n < 2^20 / P: api.cu passes the magic only then)
  const float inv_T = 1.f / static_cast<float>(nt);
  const bool live = b < B;
  const bool l0 = live && n0 < A * P, l1 = live && n0 + 1 < A * P;
  const int a0 = l0 ? (p_magic ? static_cast<int>((static_cast<uint32_t>(n0) * p_magic) >> 20) : n0 / P) : 0;
  const int a1 = l1 ? (p_magic ? static_cast<int>((static_cast<uint32_t>(n0 + 1) * p_magic) >> 20) : (n0 + 1) / P) : 0;
  float G0 = 0.f, G1 = 0.f, gup0 = 0.f, gup1 = 0.f;
  if (l0) {
    G0 = g_mu[static_cast<size_t>(b) * A + a0];
    if (dec_tanh) { const float m = mu[static_cast<size_t>(b) * A + a0]; G0 = G0 * (1.f - m * m); }
    gup0 = G0 * dec_w[n0] * inv_T;
  }
  if (l1) {
    G1 = g_mu[static_cast<size_t>(b) * A + a1];
    if (dec_tanh) { const float m = mu[static_cast<size_t>(b) * A + a1]; G1 = G1 * (1.f - m * m); }
    gup1 = G1 * dec_w[n0 + 1] * inv_T;
  }
  const size_t r0 = static_cast<size_t>(tile) * tg.CT + bl;    // sample-step (0, b); step t is r0 + t * Bt
  // voltage codes (umma.cuh: volt_code2) [NP/32][Rt/8][32 neurons][8 sample-steps]; Bt % 16 == 0, so every timestep's
  // row has r0's phase r0 % 8.  The 8 warps of a block are 8 consecutive samples: together they read whole 16-byte pieces.
  const uint16_t* vsrc = volt + ((static_cast<size_t>(n0 >> 5) * tg.Rt + (r0 & ~static_cast<size_t>(7))) * 32) + (n0 & 31) * 8 + (r0 & 7);
  const size_t vstep = static_cast<size_t>(tg.Bt) * 32;        // codes between timesteps
  const uint32_t piece = static_cast<uint32_t>(lane >> 2);
  // Bt % 16 == 0: every timestep's row r0 + t * Bt has the swizzle phase of r0, so the in-row offset is a constant
  uint8_t* const grow = g_img + (static_cast<size_t>(chunk) * tg.Rt + r0) * 128 + (lane & 3) * 4 +
                        ((piece ^ static_cast<uint32_t>(r0 & 7)) << 4);
  const size_t gstep = static_cast<size_t>(tg.Bt) * 128;
  // the thread's two neurons as one packed fp32 pair (FFMA2 / FADD2 / FMUL2: each lane rounded like the scalar operation)
  f32x2 gvn = 0ull, gcn = 0ull, db = 0ull;
  const f32x2 gup = f2_pack(gup0, gup1);
  const f32x2 k_nc75 = f2_splat(-0.75f * kVoltInv), k_o75 = f2_splat(0.375f * kVoltInv), k_half = f2_splat(0.5f),
              k_mid = f2_splat(-32767.5f), k_magic = f2_splat(-8388608.f);
  int cnt0 = 0, cnt1 = 0;
  auto step = [&](int t, uint2 v) {
    // voltage codes (umma.cuh) as floats t in [0, 65535]; u = t - 32767.5: spike <=> u > 0, window <=> |u| < 32767
    const f32x2 t2 = f2_add(f2_pack(__uint_as_float(0x4B000000u | (live ? v.x : 0u)), __uint_as_float(0x4B000000u | (live ? v.y : 0u))), k_magic);
    const f32x2 u2 = f2_add(t2, k_mid);
    const bool sp0 = f2_lo(u2) > 0.f, sp1 = f2_hi(u2) > 0.f;
    cnt0 += sp0 ? 1 : 0;
    cnt1 += sp1 ? 1 : 0;
    const f32x2 win2 = f2_pack(fabsf(f2_lo(u2)) < 32767.f ? 1.f : 0.f, fabsf(f2_hi(u2)) < 32767.f ? 1.f : 0.f);
    const f32x2 keep2 = f2_pack(sp0 ? 0.f : 0.75f, sp1 ? 0.f : 0.75f);
    const f32x2 gs2 = f2_fma(gvn, f2_fma(t2, k_nc75, k_o75), gup);          // g_up - g_v(t+1) * 0.75 v
    const f32x2 gv2 = f2_fma(gvn, keep2, f2_mul(gs2, win2));
    const f32x2 gc2 = f2_fma(k_half, gcn, gv2);
    gvn = gv2; gcn = gc2;
    db = f2_add(db, gc2);
    const float gc0 = f2_lo(gc2), gc1 = f2_hi(gc2);
    const uint32_t hp = pack_bf16x2(gc0, gc1);
    const uint32_t lp = pack_bf16x2(gc0 - __uint_as_float(hp << 16), gc1 - __uint_as_float(hp & 0xFFFF0000u));
    uint8_t* q = grow + static_cast<size_t>(t) * gstep;
    *reinterpret_cast<uint32_t*>(q) = hp;
    *reinterpret_cast<uint32_t*>(q + g_plane) = lp;
  };
  if (TT > 0) {
    uint2 v[TT > 0 ? TT : 1];
#pragma unroll
    for (int t = 0; t < TT; ++t)                                                           // all loads up front
      v[t] = make_uint2(__ldg(vsrc + static_cast<size_t>(t) * vstep), __ldg(vsrc + static_cast<size_t>(t) * vstep + 8));
#pragma unroll
    for (int t = TT - 1; t >= 0; --t) step(t, v[t]);
  } else {
    uint2 vn = make_uint2(__ldg(vsrc + static_cast<size_t>(T - 1) * vstep), __ldg(vsrc + static_cast<size_t>(T - 1) * vstep + 8));
    for (int t = T - 1; t >= 0; --t) {
      const uint2 v = vn;
      if (t > 0) vn = make_uint2(__ldg(vsrc + static_cast<size_t>(t - 1) * vstep), __ldg(vsrc + static_cast<size_t>(t - 1) * vstep + 8));
      step(t, v);
    }
  }
  // the padding sample-steps [TB, CT) of the tile must read as zero in the dW GEMM: the tile's first block clears them
  if (bl - w == 0) {      // first block of its tile; its warps share the rows
    const int TB = nt * tg.Bt;
    for (int cc = TB + w; cc < tg.CT; cc += SPB) {
      const size_t r = static_cast<size_t>(tile) * tg.CT + cc;
      uint8_t* q = g_img + (static_cast<size_t>(chunk) * tg.Rt + r) * 128 + (lane & 3) * 4 + ((piece ^ static_cast<uint32_t>(r & 7)) << 4);
      *reinterpret_cast<uint32_t*>(q) = 0u;
      *reinterpret_cast<uint32_t*>(q + g_plane) = 0u;
    }
  }
  s_red[w][0][lane] = f2_lo(db); s_red[w][1][lane] = f2_hi(db);
  s_red[w][4][lane] = G0; s_red[w][5][lane] = G1;
  __syncthreads();                            // also publishes s_rate
  s_red[w][2][lane] = G0 * s_rate[cnt0]; s_red[w][3][lane] = G1 * s_rate[cnt1];
  __syncthreads();
  if (threadIdx.x < 192) {
    const int k = threadIdx.x >> 5;                             // which of the six sums; lane = neuron pair
    float acc = 0.f;
#pragma unroll
    for (int g = 0; g < SPB; ++g) acc += s_red[g][k][lane];
    const int n = n0 + (k & 1);
    if (k < 2) db_part[static_cast<size_t>(blockIdx.x) * NP + n] = acc;
    else dec_part[(static_cast<size_t>(blockIdx.x) * 2 + (k >= 4 ? 1 : 0)) * NP + n] = acc;
  }
}

// ------------------------------------------------------------------ decoder + PPO actor head, one launch
// PPO.update between the actor's last LIF layer and its backward chain (AMP/.../algorithms/ppo.py:132-171 for the policy
// terms, modules/actor_critic.py:140-150 for the decoder): until r2 that was decode_kernel -> ppo_head_kernel ->
// ppo_head_finalize_kernel -> (copy) -> top_scan_rows_kernel, with a join with the critic's stream in the middle, on the
// critical path of every minibatch step.  Nothing in the actor's chain needs the critic: the surrogate loss, the KL
// statistic and d loss / d mu depend on (mu, log_std) and the rollout fields only.  This kernel does the decoder and
// those policy terms; the value loss goes through value_loss_kernel on the critic's stream and ppo_finalize_kernel joins
// the partial sums on a third branch.  top_scan_rows_kernel follows directly.
// (A version that also ran the output layer's scan in the same blocks was bit-identical but no faster than the separate
// kernels: 60 us, bound by the barriers between its phases at 5 blocks per SM; profiles/r2t, r2v.)
// Block = 32 samples; thread = one (sample, action) item for the decoder's bit counting (dense lanes), one sample for the
// log-prob / ratio / surrogate.  part[blk][4 + A] = block sums: surrogate, (unused), kl, (unused), g_log_std[a].
struct ActorHeadParams {
  const uint32_t* bits;     // output layer's spike words [NP/32][Rt]
  const float* dec_w; const float* dec_b; int dec_tanh;
  const float* log_std; const float* actions; const float* old_logp; const float* adv; const float* old_mu; const float* old_sigma;
  float clip, inv_count;
  float* mu;                // [B, A] out
  float* g_mu;              // [B, A] out: d loss / d mu
  float* part;              // [gridDim.x][4 + A]
  int B, A, P, T;
  TileGeom tg;
};
constexpr int kHeadSPB = 16;               // samples per block: 16 x 12 actions = 192 items, one pass of the 256 threads
constexpr int kHeadMaxItems = kHeadSPB * 64;
__global__ void __launch_bounds__(256) actor_head_kernel(const ActorHeadParams p) {
  constexpr int SPB = kHeadSPB;
  __shared__ float s_rate[kMaxT + 1];
  __shared__ float s_sig[64], s_ivar[64], s_lsig[64];
  __shared__ float s_lp[kHeadMaxItems], s_kl[kHeadMaxItems], s_d[kHeadMaxItems];      // per item: log-prob / kl terms (then the log-std term), action - mu
  __shared__ float s_glp[SPB], s_sur[SPB], s_klb[SPB];
  const int A = p.A, P = p.P, nt = p.T;
  const TileGeom tg = p.tg;
  const int b0 = blockIdx.x * SPB;
  const int nitems = SPB * A;
  if (threadIdx.x <= nt) s_rate[threadIdx.x] = __fdiv_rn(static_cast<float>(threadIdx.x), static_cast<float>(nt));
  if (threadIdx.x >= 64 && threadIdx.x < 64 + A) {
    const int a = threadIdx.x - 64;
    const float sg = expf(__ldg(p.log_std + a));
    s_sig[a] = sg;
    s_ivar[a] = 1.f / (sg * sg);
    s_lsig[a] = logf(sg);
  }
  // requested now, used after the decoder: the kernel is a chain of dependent round trips (one block = one wave)
  float adv_b = 0.f, olp_b = 0.f;
  if (threadIdx.x < SPB && b0 + threadIdx.x < p.B) {
    adv_b = __ldg(p.adv + b0 + threadIdx.x);
    olp_b = __ldg(p.old_logp + b0 + threadIdx.x);
  }
  __syncthreads();
  // ---- decoder + per-action Normal terms, thread = item (sample-major: the [B, A] fields are read as one contiguous run)
  for (int item = threadIdx.x; item < nitems; item += 256) {
    const int sm = item / A, a = item - sm * A;
    const int b = b0 + sm;
    float lp = 0.f, kl = 0.f, dd = 0.f;
    if (b < p.B) {
      const size_t gi0 = static_cast<size_t>(b) * A + a;
      const float act_i = __ldg(p.actions + gi0), osg_i = __ldg(p.old_sigma + gi0), omu_i = __ldg(p.old_mu + gi0);
      const int tile = b / tg.Bt;
      const size_t r0 = static_cast<size_t>(tile) * tg.CT + (b - tile * tg.Bt);      // sample-step (0, b); step t is r0 + t * Bt
      const int n0 = a * P;
      float acc = 0.f;
      if (P <= 16) {
        const int w0 = n0 >> 5, w1 = (n0 + P - 1) >> 5, sh = n0 & 31;
        int cnt[16];
#pragma unroll
        for (int pp = 0; pp < 16; ++pp) cnt[pp] = 0;
        for (int t = 0; t < nt; ++t) {
          const size_t r = r0 + static_cast<size_t>(t) * tg.Bt;
          const uint32_t lo = __ldg(p.bits + static_cast<size_t>(w0) * tg.Rt + r);
          const uint32_t hi = w1 != w0 ? __ldg(p.bits + static_cast<size_t>(w1) * tg.Rt + r) : 0u;
          const uint32_t x = __funnelshift_r(lo, hi, sh);      // bits n0 .. n0+31 of this sample at step t
#pragma unroll
          for (int pp = 0; pp < 16; ++pp) cnt[pp] += (x >> pp) & 1u;
        }
#pragma unroll
        for (int pp = 0; pp < 16; ++pp)
          if (pp < P) acc = __fmaf_rn(__ldg(p.dec_w + n0 + pp), s_rate[cnt[pp]], acc);
      } else {
        for (int pp = 0; pp < P; ++pp) {
          const int n = n0 + pp;
          int cnt = 0;
          for (int t = 0; t < nt; ++t)
            cnt += (__ldg(p.bits + static_cast<size_t>(n >> 5) * tg.Rt + r0 + static_cast<size_t>(t) * tg.Bt) >> (n & 31)) & 1u;
          acc = __fmaf_rn(__ldg(p.dec_w + n), s_rate[cnt], acc);
        }
      }
      acc = __fadd_rn(acc, __ldg(p.dec_b + a));                 // decode_kernel's arithmetic, bit for bit
      const float m = p.dec_tanh ? tanhf(acc) : acc;
      const size_t gi = static_cast<size_t>(b) * A + a;
      p.mu[gi] = m;
      // Normal(mu, sigma) terms of this action (ppo_head_kernel's formulas)
      const float sg = s_sig[a], iv = s_ivar[a];
      const float d = act_i - m;
      lp = -(d * d) * 0.5f * iv - s_lsig[a] - 0.91893853320467274178f;
      const float os = osg_i, dm = omu_i - m;
      kl = logf(sg / os + 1.e-5f) + (os * os + dm * dm) * 0.5f * iv - 0.5f;
      dd = d;
    }
    s_lp[item] = lp;
    s_kl[item] = kl;
    s_d[item] = dd;
  }
  __syncthreads();
  // ---- per sample: log-prob, ratio, clipped surrogate (ppo.py:153-158), thread = sample
  if (threadIdx.x < SPB) {
    const int b = b0 + threadIdx.x;
    float logp = 0.f, kl = 0.f, sur = 0.f, g_logp = 0.f;
    if (b < p.B) {
      for (int a = 0; a < A; ++a) { logp += s_lp[threadIdx.x * A + a]; kl += s_kl[threadIdx.x * A + a]; }
      const float ad = adv_b;
      const float ratio = expf(logp - olp_b);
      const float s1 = -ad * ratio;
      const float s2 = -ad * fminf(fmaxf(ratio, 1.f - p.clip), 1.f + p.clip);
      sur = fmaxf(s1, s2);
      g_logp = (s1 >= s2 ? -ad : 0.f) * ratio * p.inv_count;     // d mean(sur) / d logp
    }
    s_glp[threadIdx.x] = g_logp; s_sur[threadIdx.x] = sur; s_klb[threadIdx.x] = kl;
  }
  __syncthreads();
  // ---- d loss / d mu and the log-std gradient terms, thread = item
  for (int item = threadIdx.x; item < nitems; item += 256) {
    const int sm = item / A, a = item - sm * A;
    const int b = b0 + sm;
    float gls = 0.f;
    if (b < p.B) {
      const size_t gi = static_cast<size_t>(b) * A + a;
      const float d = s_d[item], iv = s_ivar[a], g_logp = s_glp[sm];
      p.g_mu[gi] = g_logp * d * iv;
      gls = g_logp * (d * d * iv - 1.f);                          // d logp / d log_std = (a-mu)^2/sigma^2 - 1
    }
    s_lp[item] = gls;
  }
  __syncthreads();
  if (threadIdx.x < 4 + A) {      // block sums in a fixed order
    const int i = threadIdx.x;
    float t = 0.f;
    if (i == 0) { for (int g = 0; g < SPB; ++g) t += s_sur[g]; }
    else if (i == 2) { for (int g = 0; g < SPB; ++g) t += s_klb[g]; }
    else if (i >= 4) { for (int g = 0; g < SPB; ++g) t += s_lp[g * A + (i - 4)]; }
    p.part[static_cast<size_t>(blockIdx.x) * (4 + A) + i] = t;
  }
}

// ------------------------------------------------------------------ minibatch permutation: all fields in one pass
// RolloutStorage.mini_batch_generator gathers 9 tensors by the same random row indices, 20 times per update
// (rollout_storage.py:174-182); here every field is gathered ONCE per update, and all of them by this one kernel:
// warp = one output row, it copies that row of every field (4-byte words; rows are 4..192 bytes).
struct GatherJob { const uint32_t* src; uint32_t* dst; int words; };
struct GatherJobs { GatherJob j[12]; int count; };
__global__ void __launch_bounds__(256) gather_rows_kernel(const GatherJobs jobs, const long long* __restrict__ idx, long long n) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const long long src_row = idx[row];
  for (int k = 0; k < jobs.count; ++k) {
    const GatherJob jb = jobs.j[k];
    const uint32_t* s = jb.src + src_row * jb.words;
    uint32_t* d = jb.dst + row * jb.words;
    for (int c = lane; c < jb.words; c += 32) d[c] = __ldg(s + c);
  }
}

// ------------------------------------------------------------------ encoder input gradient
// (the encoder PARAMETER gradients are produced by the DX_ENC epilogue, nm_gemm.cuh)
// d obs[b,o] = sum_p g_a a (-(x-mean))/var   (RMA: the actor input is cat(obs, latent) and needs grad)
__global__ void enc_dobs_kernel(const float* __restrict__ obs, const float* __restrict__ mean,
                                const float* __restrict__ std, const float* __restrict__ g_a, int B, int O, int P,
                                int K0P, float eps, float* __restrict__ d_obs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * O) return;
  const int b = idx / O, o = idx - b * O;
  const float x = obs[idx];
  float acc = 0.f;
  for (int pp = 0; pp < P; ++pp) {
    const int k = o * P + pp;
    const float mk = mean[k], sk = std[k];
    const float var = __fadd_rn(__fmul_rn(sk, sk), eps);
    const float a = pop_activation_fast(x, mk, sk, eps);
    acc += g_a[(static_cast<size_t>(b >> 4) * K0P + k) * 16 + (b & 15)] * a * (-(x - mk)) / var;
  }
  d_obs[idx] = acc;
}

// ALL deferred reductions of a backward pass in ONE launch (blockIdx.y selects the job): the row reductions of the
// per-tile / per-block partial rows (bias, encoder and decoder gradients) and the split-K reductions of the dW tiles.
// A job with thousands of partial rows (the top-layer scan writes one row per 8 samples: 3072 at B = 24576) is reduced
// in two levels inside the same launch: blockIdx.z = segment of seg_rows rows -> one row of a scratch matrix; the LAST
// segment block of the job to finish (a counter per job) sums the scratch rows in a fixed order -> dst.  Until r2 these
// were three dependent launches (split-K, segments, finals) on the critical path of every minibatch step.
struct RowJob {
  const float* src; float* dst; float* tmp;      // tmp: [nseg][width] scratch (nseg > 1)
  int nparts, n, col_step, nseg, seg_rows, width;
  unsigned long long stride;
};
struct PartJob { const float* partial; float* dW; int splits, n_tiles, N, K; };
struct ReduceJobs { RowJob r[12]; PartJob p[8]; int nr, np; int* counters; };
__global__ void __launch_bounds__(1024) multi_reduce_kernel(const ReduceJobs jobs) {   // block (32, 32)
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (static_cast<int>(blockIdx.y) >= jobs.nr) {
    // split-K partial tiles [tile][split][128][256] -> dW[n][k]: thread = one element, warp = 32 consecutive k
    const PartJob jb = jobs.p[blockIdx.y - jobs.nr];
    const int k = blockIdx.x * 32 + threadIdx.x, n = blockIdx.z * 32 + threadIdx.y;
    if (k >= jb.K || n >= jb.N) return;
    const int tile = (n >> 7) * jb.n_tiles + (k >> 8);
    const float* src = jb.partial + (static_cast<size_t>(tile) * jb.splits * 128 + (n & 127)) * 256 + (k & 255);
    // eight independent loads in flight per thread, summed in split order (fixed: bitwise reproducible)
    float acc = 0.f;
    int sp = 0;
    for (; sp + 8 <= jb.splits; sp += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[static_cast<size_t>(sp + u) * 128 * 256];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; sp < jb.splits; ++sp) acc += src[static_cast<size_t>(sp) * 128 * 256];
    jb.dW[static_cast<size_t>(n) * jb.K + k] = acc;
    return;
  }
  const RowJob jb = jobs.r[blockIdx.y];
  const bool two_level = jb.nseg > 1;
  const int ncols = two_level ? jb.width : jb.n;           // the segment pass keeps every column of the strided window
  if (static_cast<int>(blockIdx.x) * 32 >= ncols || static_cast<int>(blockIdx.z) >= jb.nseg) return;
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int r0 = two_level ? blockIdx.z * jb.seg_rows : 0;
  const int nrows = two_level ? min(jb.seg_rows, jb.nparts - r0) : jb.nparts;
  float acc = 0.f;
  if (i < ncols) {
    // eight independent loads in flight per thread (the pass is pure load latency); fixed summation order
    const float* src = jb.src + static_cast<size_t>(r0) * jb.stride + static_cast<size_t>(i) * (two_level ? 1 : jb.col_step);
    int s = threadIdx.y;
    for (; s + 7 * 32 < nrows; s += 8 * 32) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[static_cast<size_t>(s + u * 32) * jb.stride];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; s < nrows; s += 32) acc += src[static_cast<size_t>(s) * jb.stride];
  }
  __shared__ float sh[32][33];
  __shared__ int s_last;
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < ncols) {
    float t = 0.f;
    for (int y = 0; y < 32; ++y) t += sh[y][threadIdx.x];
    if (two_level) jb.tmp[static_cast<size_t>(blockIdx.z) * jb.width + i] = t;
    else jb.dst[i] = t;
  }
  if (!two_level) return;
  // last segment block of this job: the final pass over the nseg scratch rows (fixed order: bitwise reproducible)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int total = ((jb.width + 31) / 32) * jb.nseg;
    s_last = atomicAdd(jobs.counters + blockIdx.y, 1) == total - 1 ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = tid; c < jb.n; c += 1024) {
    const float* src = jb.tmp + static_cast<size_t>(c) * jb.col_step;
    float t = 0.f;
    for (int z = 0; z < jb.nseg; ++z) t += __ldcg(src + static_cast<size_t>(z) * jb.width);
    jb.dst[c] = t;
  }
  if (tid == 0) jobs.counters[blockIdx.y] = 0;      // re-armed for the next launch that draws this counter slot
}

// ------------------------------------------------------------------ GAE (rollout_storage.py:124-138)
// Thread per environment, reverse scan over S steps; [S][N] layout makes every step a coalesced row.
__global__ void gae_scan_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones,
                                const float* __restrict__ values, const float* __restrict__ last_values, int S,
                                int N, float gamma, float lam, float* __restrict__ returns,
                                float* __restrict__ adv, double* __restrict__ part_sum) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double local = 0.0;
  if (n < N) {
    float a = 0.f;
    float next_v = last_values[n];
    for (int t = S - 1; t >= 0; --t) {
      const size_t i = static_cast<size_t>(t) * N + n;
      const float v = values[i];
      const float nt = 1.0f - static_cast<float>(dones[i]);
      const float delta = __fsub_rn(__fadd_rn(rewards[i], __fmul_rn(__fmul_rn(nt, gamma), next_v)), v);
      a = __fadd_rn(delta, __fmul_rn(__fmul_rn(__fmul_rn(nt, gamma), lam), a));
      const float ret = __fadd_rn(a, v);
      returns[i] = ret;
      const float ad = __fsub_rn(ret, v);
      adv[i] = ad;
      local += static_cast<double>(ad);
      next_v = v;
    }
  }
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) s += sh[w];
    part_sum[blockIdx.x] = s;
  }
}

__global__ void gae_var_kernel(const float* __restrict__ adv, int64_t count, const double* __restrict__ part_sum,
                               int nparts_sum, double* __restrict__ part_sq) {
  double total = 0.0;
  for (int i = 0; i < nparts_sum; ++i) total += part_sum[i];
  const double mean = total / static_cast<double>(count);
  double local = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double d = static_cast<double>(adv[i]) - mean;
    local += d * d;
  }
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) s += sh[w];
    part_sq[blockIdx.x] = s;
  }
}

__global__ void gae_norm_kernel(float* __restrict__ adv, int64_t count, const double* __restrict__ part_sum,
                                int nparts_sum, const double* __restrict__ part_sq, int nparts_sq) {
  double total = 0.0, sq = 0.0;
  for (int i = 0; i < nparts_sum; ++i) total += part_sum[i];
  for (int i = 0; i < nparts_sq; ++i) sq += part_sq[i];
  const float mean = static_cast<float>(total / static_cast<double>(count));
  const float stdv = static_cast<float>(sqrt(sq / static_cast<double>(count > 1 ? count - 1 : 1)));   // unbiased
  const float denom = stdv + 1e-8f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    adv[i] = (adv[i] - mean) / denom;
}

}  // namespace snn
