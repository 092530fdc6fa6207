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

__global__ void pack_w_kernel(const float* __restrict__ W, const float* __restrict__ bias, int N, int K, int NP,
                              int KP, uint8_t* __restrict__ w_img, size_t w_plane, uint8_t* __restrict__ wt_img,
                              size_t wt_plane, float* __restrict__ bias_pad) {
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

// TT = 5: fully unrolled for the reference's spike_ts; TT = 0: generic T
template <int TT>
__global__ void __launch_bounds__(256) encode_kernel(const float* __restrict__ obs, const float* __restrict__ mean,
                                                    const float* __restrict__ std, int B, int Bpad, int O, int P,
                                                    int K0, int T, float eps, int64_t R, uint32_t* __restrict__ bits) {
  // block (32 neurons, 8 warps) covers 128 rows: each warp walks 16 rows (few, fat blocks: the per-row work is tiny)
  const int k = blockIdx.x * 32 + threadIdx.x;
  const bool k_live = k < K0;
  const int o = static_cast<int>((static_cast<float>(k) + 0.5f) * __frcp_rn(static_cast<float>(P)));   // k / P
  const float mk = k_live ? mean[k] : 0.f, sk = k_live ? std[k] : 1.f;
  const float var = __fadd_rn(__fmul_rn(sk, sk), eps);
  const float nh_inv_var = -0.5f * __frcp_rn(var);
  const int b0 = blockIdx.y * 128 + threadIdx.y;
  const float* xp = obs + static_cast<size_t>(b0) * O + o;
  uint32_t* wp = bits + static_cast<size_t>(blockIdx.x) * R + b0;
  const int nt = TT > 0 ? TT : T;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int b = b0 + i * 8;
    if (b >= Bpad) break;
    uint32_t sp = 0;
    if ((b < B) && k_live) {
      const float d = __fsub_rn(xp[static_cast<size_t>(i) * 8 * O], mk);
      // fast path: a within ~1e-6 relative of the reference's exp(((-0.5 d^2) / var)); the spike train can only
      // differ from the exact one if some membrane value comes within ~5e-6 of the threshold
      float margin;
      sp = regular_spikes<TT>(__expf(d * d * nh_inv_var), T, margin);
      if (margin < 4e-5f) {
        // reference arithmetic order (actor_critic.py:105): ((-0.5 * d^2) / std^2), then a correctly rounded exp
        const float q = __fdiv_rn(__fmul_rn(-0.5f, __fmul_rn(d, d)), var);
        sp = regular_spikes<0>(static_cast<float>(exp(static_cast<double>(q))), T, margin);
      }
    }
    uint32_t* w = wp + i * 8;
    for (int t = 0; t < nt; ++t) {
      const uint32_t word = __ballot_sync(0xffffffffu, (sp >> t) & 1u);
      if (threadIdx.x == 0) w[static_cast<size_t>(t) * Bpad] = word;
    }
  }
}

// ------------------------------------------------------------------ decoder
// actor_critic.py:140-150: rate = spike count / T ; mu[b,a] = sum_p w[a,p] rate[b, a P + p] + bias[a] (; tanh)
__global__ void __launch_bounds__(256) decode_kernel(const uint32_t* __restrict__ bits, int64_t R, int B, int Bpad, int A,
                                                    int P, int T, const float* __restrict__ dec_w,
                                                    const float* __restrict__ dec_b, int dec_tanh, float* __restrict__ mu) {
  // one thread per (action, sample), samples fastest: the <= 2 spike words an action's population spans are
  // loaded coalesced across samples
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * A) return;
  const int a0 = idx / B, b = idx - a0 * B;
  const float invT = static_cast<float>(T);
  for (int a = a0; a == a0; ++a) {
    const int n0 = a * P;
    const int w0 = n0 >> 5, w1 = (n0 + P - 1) >> 5, sh = n0 & 31;
    float acc = 0.f;
    if (P <= 16) {
      int cnt[16];
#pragma unroll
      for (int pp = 0; pp < 16; ++pp) cnt[pp] = 0;
      for (int t = 0; t < T; ++t) {
        const size_t r = static_cast<size_t>(t) * Bpad + b;
        const uint32_t lo = __ldg(bits + static_cast<size_t>(w0) * R + r);
        const uint32_t hi = w1 != w0 ? __ldg(bits + static_cast<size_t>(w1) * R + r) : 0u;
        const uint32_t x = __funnelshift_r(lo, hi, sh);      // bits n0 .. n0+31 of this sample at step t
#pragma unroll
        for (int pp = 0; pp < 16; ++pp) cnt[pp] += (x >> pp) & 1u;
      }
#pragma unroll
      for (int pp = 0; pp < 16; ++pp)
        if (pp < P) acc = __fmaf_rn(dec_w[n0 + pp], __fdiv_rn(static_cast<float>(cnt[pp]), invT), acc);
    } else {
      for (int pp = 0; pp < P; ++pp) {
        const int n = n0 + pp;
        int cnt = 0;
        for (int t = 0; t < T; ++t)
          cnt += (__ldg(bits + static_cast<size_t>(n >> 5) * R + static_cast<size_t>(t) * Bpad + b) >> (n & 31)) & 1u;
        acc = __fmaf_rn(dec_w[n], __fdiv_rn(static_cast<float>(cnt), invT), acc);
      }
    }
    acc = __fadd_rn(acc, dec_b[a]);
    mu[static_cast<size_t>(b) * A + a] = dec_tanh ? tanhf(acc) : acc;
  }
}

// ------------------------------------------------------------------ BPTT of the output population layer
// g_up[t] = G[b,a] w_dec[a,p] / T for every t (out_act = sum_t s_t / T feeds the grouped conv), then the
// recurrence of SURVEY.md §8a.  Writes g_c hi/lo planes as swz64 image and per-block bias partials.
__device__ __forceinline__ float warp_col_sums16(float (&x)[16], int lane) {   // see scan_gemm.cuh
#pragma unroll
  for (int h = 16, n = 16; h >= 2; h >>= 1, n >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? x[i] : x[i + n / 2];
      const float keep = up ? x[i + n / 2] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return x[0] + __shfl_xor_sync(0xffffffffu, x[0], 1);
}

// one BPTT step of 16 neurons of one row: updates the recurrence state and writes the hi/lo planes of g_c
__device__ __forceinline__ void top_scan_step(const float (&gup)[16], const float (&v)[16], float (&gvn)[16], float (&gcn)[16],
                                              float (&dbacc)[16], uint8_t* g_img, size_t g_plane, int64_t R, size_t r, int n0) {
  float gc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float gs = gup[j] - gvn[j] * (0.75f * v[j]);
    const float win = fabsf(__fsub_rn(v[j], 0.5f)) < 0.5f ? 1.f : 0.f;
    const float keep = v[j] > 0.5f ? 0.f : 0.75f;
    const float gv = gs * win + gvn[j] * keep;
    gc[j] = gv + 0.5f * gcn[j];
    gvn[j] = gv; gcn[j] = gc[j];
    dbacc[j] += gc[j];
  }
  float hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { hi[j] = bf16_round(gc[j]); lo[j] = gc[j] - hi[j]; }
  const size_t base = static_cast<size_t>(n0 >> 6) * R * 128 + (r >> 3) * 1024 + (r & 7) * 128;
  const int p0 = (n0 & 63) >> 3;
  const int sw = static_cast<int>(r & 7);
  *reinterpret_cast<uint4*>(g_img + base + ((p0 ^ sw) << 4)) =
      make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]), pack_bf16x2(hi[6], hi[7]));
  *reinterpret_cast<uint4*>(g_img + base + (((p0 + 1) ^ sw) << 4)) =
      make_uint4(pack_bf16x2(hi[8], hi[9]), pack_bf16x2(hi[10], hi[11]), pack_bf16x2(hi[12], hi[13]), pack_bf16x2(hi[14], hi[15]));
  *reinterpret_cast<uint4*>(g_img + g_plane + base + ((p0 ^ sw) << 4)) =
      make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]), pack_bf16x2(lo[6], lo[7]));
  *reinterpret_cast<uint4*>(g_img + g_plane + base + (((p0 + 1) ^ sw) << 4)) =
      make_uint4(pack_bf16x2(lo[8], lo[9]), pack_bf16x2(lo[10], lo[11]), pack_bf16x2(lo[12], lo[13]), pack_bf16x2(lo[14], lo[15]));
}

__global__ void __launch_bounds__(256) top_scan_kernel(const float* __restrict__ g_mu, const float* __restrict__ mu,
                                                      int dec_tanh, const float* __restrict__ dec_w,
                                                      const float* __restrict__ volt, int B, int Bpad, int A, int P,
                                                      int NP, int T, int64_t R, uint8_t* __restrict__ g_img,
                                                      size_t g_plane, float* __restrict__ db_part,
                                                      float* __restrict__ dec_part /* [gridDim.x][2][NP] */) {
  // block = 8 warps x 32 rows; lane = row, warp w walks the 16-column groups w, w+8, ...  grid = Bpad / 32.
  // Also produces the decoder gradients: dec_part[blk][0][n] = sum_b G[b,a(n)] rate[b,n]  (d w_dec),
  // dec_part[blk][1][n] = sum_b G[b,a(n)] (d bias, read at n = a * P).
  // A warp reads 32 rows x 64 B = one contiguous 2 KiB run of the voltage trace per (t, group).
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane;
  const bool row_live = b < B;
  const float Tf = static_cast<float>(T);
  for (int g16 = wid; g16 < NP / 16; g16 += 8) {
    const int n0 = g16 * 16;
    float gup[16], Gc[16], cnt[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = n0 + j;
      gup[j] = 0.f; Gc[j] = 0.f; cnt[j] = 0.f;
      if (row_live && n < A * P) {
        const int a = n / P;
        float G = g_mu[static_cast<size_t>(b) * A + a];
        if (dec_tanh) { const float m = mu[static_cast<size_t>(b) * A + a]; G = G * (1.f - m * m); }
        Gc[j] = G;
        gup[j] = __fdiv_rn(G * dec_w[n], Tf);
      }
    }
    float gvn[16], gcn[16], dbacc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { gvn[j] = 0.f; gcn[j] = 0.f; dbacc[j] = 0.f; }
    const float* vbase = volt + (static_cast<size_t>(g16) * Bpad + b) * 16;
    const size_t vstep = static_cast<size_t>(NP >> 4) * Bpad * 16;
    {
      // the voltage loads do not depend on the recurrence: keep the next (earlier) timestep's load in flight
      float4 vn[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        vn[j] = row_live ? __ldg(reinterpret_cast<const float4*>(vbase + static_cast<size_t>(T - 1) * vstep) + j)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int t = T - 1; t >= 0; --t) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[4 * j] = vn[j].x; v[4 * j + 1] = vn[j].y; v[4 * j + 2] = vn[j].z; v[4 * j + 3] = vn[j].w; }
#pragma unroll
        for (int j = 0; j < 16; ++j) cnt[j] += v[j] > 0.5f ? 1.f : 0.f;       // output spike count -> rate
        if (t > 0 && row_live) {
#pragma unroll
          for (int j = 0; j < 4; ++j) vn[j] = __ldg(reinterpret_cast<const float4*>(vbase + static_cast<size_t>(t - 1) * vstep) + j);
        }
        top_scan_step(gup, v, gvn, gcn, dbacc, g_img, g_plane, R, static_cast<size_t>(t) * Bpad + b, n0);
      }
    }
    const float tot = warp_col_sums16(dbacc, lane);
    if ((lane & 1) == 0) db_part[static_cast<size_t>(blockIdx.x) * NP + n0 + (lane >> 1)] = tot;
    float dwd[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) dwd[j] = Gc[j] * __fdiv_rn(cnt[j], Tf);
    const float tw = warp_col_sums16(dwd, lane);
    const float tb = warp_col_sums16(Gc, lane);
    if ((lane & 1) == 0) {
      float* dst = dec_part + (static_cast<size_t>(blockIdx.x) * 2) * NP + n0 + (lane >> 1);
      dst[0] = tw;
      dst[NP] = tb;
    }
  }
}

// (decoder gradients are produced by top_scan_kernel)

// ------------------------------------------------------------------ encoder gradients
// From g_a = d loss / d pop_act [Bpad][K0P]:
//   d mean = sum_b g_a a (x-mean)/var ; d std = sum_b g_a a (x-mean)^2 std / var^2 ; (var = std^2 + eps)
// Block = 64 rows, threads stride over k; partials [gridDim.x][2][K0].
__global__ void __launch_bounds__(256) enc_bwd_kernel(const float* __restrict__ obs, const float* __restrict__ mean,
                                                     const float* __restrict__ std, const float* __restrict__ g_a,
                                                     int B, int O, int P, int K0, int K0P, float eps,
                                                     float* __restrict__ part /* [gridDim.y][2][K0] */) {
  // block (32 encoder neurons, 8 row lanes); grid = (ceil(K0/32), ceil(B/64)); each thread walks 8 rows
  const int k = blockIdx.x * 32 + threadIdx.x;
  const bool live = k < K0;
  float dm = 0.f, ds = 0.f;
  if (live) {
    const float mk = mean[k], sk = std[k];
    const float inv_var = 1.f / (sk * sk + eps);
    const int o = k / P;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = blockIdx.y * 64 + j * 8 + threadIdx.y;
      if (b < B) {
        const float x = obs[static_cast<size_t>(b) * O + o];
        const float d = x - mk;
        const float a = expf(-0.5f * d * d * inv_var);
        const float ga = g_a[static_cast<size_t>(b) * K0P + k] * a;
        dm += ga * d * inv_var;
        ds += ga * d * d * inv_var * inv_var * sk;
      }
    }
  }
  __shared__ float sh[2][8][33];
  sh[0][threadIdx.y][threadIdx.x] = dm;
  sh[1][threadIdx.y][threadIdx.x] = ds;
  __syncthreads();
  if (threadIdx.y < 2 && live) {
    float t = 0.f;
    for (int y = 0; y < 8; ++y) t += sh[threadIdx.y][y][threadIdx.x];
    part[(static_cast<size_t>(blockIdx.y) * 2 + threadIdx.y) * K0 + k] = t;
  }
}

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
    acc += g_a[static_cast<size_t>(b) * K0P + k] * a * (-(x - mk)) / var;
  }
  d_obs[idx] = acc;
}

// out[i] = sum_s part[s * stride + i]; block (32 columns, 8 part lanes), fixed summation order
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ part, int nparts, size_t stride,
                                                         int n, float* __restrict__ out) {
  const int i = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (i < n)
    for (int s = threadIdx.y; s < nparts; s += 8) acc += part[static_cast<size_t>(s) * stride + i];
  __shared__ float sh[8][33];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    float t = 0.f;
    for (int y = 0; y < 8; ++y) t += sh[y][threadIdx.x];
    out[i] = t;
  }
}

// Several independent row reductions / split-K reductions in ONE launch (blockIdx.y resp. .z selects the job):
// the backward pass defers all of its small reductions to the end instead of launching ten tiny kernels.
struct RowJob { const float* src; float* dst; int nparts; int n; unsigned long long stride; int col_step; };
struct RowJobs { RowJob j[12]; int count; };
__global__ void __launch_bounds__(1024) multi_reduce_rows_kernel(const RowJobs jobs) {   // block (32, 32)
  const RowJob jb = jobs.j[blockIdx.y];
  const int i = blockIdx.x * 32 + threadIdx.x;
  if (blockIdx.x * 32 >= jb.n) return;
  float acc = 0.f;
  if (i < jb.n)
    for (int s = threadIdx.y; s < jb.nparts; s += 32) acc += jb.src[static_cast<size_t>(s) * jb.stride + static_cast<size_t>(i) * jb.col_step];
  __shared__ float sh[32][33];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < jb.n) {
    float t = 0.f;
    for (int y = 0; y < 32; ++y) t += sh[y][threadIdx.x];
    jb.dst[i] = t;
  }
}

struct PartJob { const float* partial; float* dW; int splits, n_tiles, N, K; };
struct PartJobs { PartJob j[8]; int count; };
__global__ void multi_reduce_partials_kernel(const PartJobs jobs) {
  const PartJob jb = jobs.j[blockIdx.z];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (k >= jb.K || n >= jb.N) return;
  const int tile = (n >> 7) * jb.n_tiles + (k >> 8);
  const float* src = jb.partial + (static_cast<size_t>(tile) * jb.splits * 128 + (n & 127)) * 256 + (k & 255);
  float acc = 0.f;
  for (int s = 0; s < jb.splits; ++s) acc += src[static_cast<size_t>(s) * 128 * 256];
  jb.dW[static_cast<size_t>(n) * jb.K + k] = acc;
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
