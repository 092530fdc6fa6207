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

__device__ __forceinline__ uint32_t regular_spikes(float a, int T, float& margin) {
  float u = 0.f;
  uint32_t bits = 0;
  margin = 1.f;
  for (int t = 0; t < T; ++t) {
    u = __fadd_rn(u, a);
    margin = fminf(margin, fabsf(u - 0.999f));
    if (u > 0.999f) { bits |= 1u << t; u = __fsub_rn(u, 0.999f); }
  }
  return bits;
}

__global__ void encode_kernel(const float* __restrict__ obs, const float* __restrict__ mean,
                              const float* __restrict__ std, int B, int Bpad, int O, int P, int K0, int T,
                              float eps, int64_t R, uint32_t* __restrict__ bits) {
  const int k = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.y * blockDim.y + threadIdx.y;
  if (b >= Bpad) return;
  uint32_t sp = 0;
  if ((b < B) && (k < K0)) {
    const float x = obs[static_cast<size_t>(b) * O + k / P], mk = mean[k], sk = std[k];
    // reference arithmetic order (actor_critic.py:105): ((-0.5 * d^2) / std^2), then exp
    const float d = __fsub_rn(x, mk);
    const float q = __fdiv_rn(__fmul_rn(-0.5f, __fmul_rn(d, d)), __fadd_rn(__fmul_rn(sk, sk), eps));
    float margin;
    sp = regular_spikes(expf(q), T, margin);
    // expf is good to ~2 ulp; only a membrane within ~1e-5 of the threshold could see the difference:
    // redo those few with the correctly rounded exp
    if (margin < 1e-5f) sp = regular_spikes(static_cast<float>(exp(static_cast<double>(q))), T, margin);
  }
  for (int t = 0; t < T; ++t) {
    const uint32_t word = __ballot_sync(0xffffffffu, (sp >> t) & 1u);
    if (threadIdx.x == 0) bits[static_cast<size_t>(blockIdx.x) * R + static_cast<size_t>(t) * Bpad + b] = word;
  }
}

// ------------------------------------------------------------------ decoder
// actor_critic.py:140-150: rate = spike count / T ; mu[b,a] = sum_p w[a,p] rate[b, a P + p] + bias[a] (; tanh)
__global__ void decode_kernel(const uint32_t* __restrict__ bits, int64_t R, int B, int Bpad, int A, int P, int T,
                              const float* __restrict__ dec_w, const float* __restrict__ dec_b, int dec_tanh,
                              float* __restrict__ mu) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * A) return;
  const int b = idx / A, a = idx - b * A;
  float acc = 0.f;
  for (int pp = 0; pp < P; ++pp) {
    const int n = a * P + pp;
    int cnt = 0;
    for (int t = 0; t < T; ++t)
      cnt += (bits[static_cast<size_t>(n >> 5) * R + static_cast<size_t>(t) * Bpad + b] >> (n & 31)) & 1u;
    const float rate = __fdiv_rn(static_cast<float>(cnt), static_cast<float>(T));
    acc = __fmaf_rn(dec_w[n], rate, acc);
  }
  acc = __fadd_rn(acc, dec_b[a]);
  mu[idx] = dec_tanh ? tanhf(acc) : acc;
}

// ------------------------------------------------------------------ BPTT of the output population layer
// g_up[t] = G[b,a] w_dec[a,p] / T for every t (out_act = sum_t s_t / T feeds the grouped conv), then the
// recurrence of SURVEY.md §8a.  Writes g_c hi/lo planes as swz64 image and per-block bias partials.
__global__ void __launch_bounds__(256) top_scan_kernel(const float* __restrict__ g_mu, const float* __restrict__ mu,
                                                      int dec_tanh, const float* __restrict__ dec_w,
                                                      const float* __restrict__ volt, int B, int Bpad, int A, int P,
                                                      int NP, int T, int64_t R, uint8_t* __restrict__ g_img,
                                                      size_t g_plane, float* __restrict__ db_part) {
  // blockDim = (NP/8 column groups [<= 32], 256/(NP/8) rows); block covers 32 rows; grid = (Bpad/32)
  const int cg = threadIdx.x, ry = threadIdx.y, rows_per_it = blockDim.y;
  const int n0 = cg * 8;
  float w[8];
  int aidx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = n0 + j;
    const bool live = n < A * P;
    w[j] = live ? dec_w[n] : 0.f;
    aidx[j] = live ? n / P : -1;
  }
  float dbacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dbacc[j] = 0.f;
  const float invT = static_cast<float>(T);
  for (int rb = ry; rb < 32; rb += rows_per_it) {
    const int b = blockIdx.x * 32 + rb;
    const bool row_live = b < B;
    float gup[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gup[j] = 0.f;
      if (row_live && aidx[j] >= 0) {
        float G = g_mu[static_cast<size_t>(b) * A + aidx[j]];
        if (dec_tanh) { const float m = mu[static_cast<size_t>(b) * A + aidx[j]]; G = G * (1.f - m * m); }
        gup[j] = __fdiv_rn(G * w[j], invT);
      }
    }
    float gvn[8], gcn[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { gvn[j] = 0.f; gcn[j] = 0.f; }
    for (int t = T - 1; t >= 0; --t) {
      const size_t r = static_cast<size_t>(t) * Bpad + b;
      float v[8];
      if (row_live) {
        const float4* src = reinterpret_cast<const float4*>(
            volt + ((static_cast<size_t>(t) * (NP >> 4) + (n0 >> 4)) * Bpad + b) * 16 + (n0 & 15));
        const float4 f0 = __ldg(src), f1 = __ldg(src + 1);
        v[0] = f0.x; v[1] = f0.y; v[2] = f0.z; v[3] = f0.w; v[4] = f1.x; v[5] = f1.y; v[6] = f1.z; v[7] = f1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gs = gup[j] - gvn[j] * (0.75f * v[j]);
        const float win = fabsf(__fsub_rn(v[j], 0.5f)) < 0.5f ? 1.f : 0.f;
        const float keep = v[j] > 0.5f ? 0.f : 0.75f;
        const float gv = gs * win + gvn[j] * keep;
        const float gc = gv + 0.5f * gcn[j];
        gvn[j] = gv; gcn[j] = gc;
        dbacc[j] += gc;
        hi[j] = bf16_round(gc);
        lo[j] = gc - hi[j];
      }
      const size_t off = swz64_offset(static_cast<int64_t>(r), n0, R);
      *reinterpret_cast<uint4*>(g_img + off) =
          make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]), pack_bf16x2(hi[6], hi[7]));
      *reinterpret_cast<uint4*>(g_img + g_plane + off) =
          make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]), pack_bf16x2(lo[6], lo[7]));
    }
  }
  // reduce the bias gradient over the block's rows
  __shared__ float sh[256 * 8];
  float* mine = sh + (ry * blockDim.x + cg) * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) mine[j] = dbacc[j];
  __syncthreads();
  if (ry == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = 0.f;
      for (int y = 0; y < rows_per_it; ++y) acc += sh[(y * blockDim.x + cg) * 8 + j];
      db_part[static_cast<size_t>(blockIdx.x) * NP + n0 + j] = acc;
    }
  }
}

// ------------------------------------------------------------------ decoder gradients
// d w_dec[a,p] = sum_b G[b,a] rate[b, aP+p] ; d bias[a] = sum_b G[b,a].  Block = 128 rows; partials out.
__global__ void __launch_bounds__(256) dec_bwd_kernel(const float* __restrict__ g_mu, const float* __restrict__ mu,
                                                     int dec_tanh, const uint32_t* __restrict__ bits, int64_t R, int B,
                                                     int Bpad, int A, int P, int T,
                                                     float* __restrict__ part /* [gridDim.y][A*P + A] */) {
  // block (32 columns, 8 row lanes); grid = (ceil((AP + A)/32), ceil(B/64)); each thread walks 8 rows
  const int AP = A * P;
  const int i = blockIdx.x * 32 + threadIdx.x;
  const bool col_live = i < AP + A;
  const bool is_w = i < AP;
  const int a = col_live ? (is_w ? i / P : i - AP) : 0;
  float acc = 0.f;
  if (col_live) {
    for (int j = 0; j < 8; ++j) {
      const int b = blockIdx.y * 64 + j * 8 + threadIdx.y;
      if (b >= B) break;
      float G = g_mu[static_cast<size_t>(b) * A + a];
      if (dec_tanh) { const float m = mu[static_cast<size_t>(b) * A + a]; G = G * (1.f - m * m); }
      if (is_w) {
        int cnt = 0;
        for (int t = 0; t < T; ++t)
          cnt += (bits[static_cast<size_t>(i >> 5) * R + static_cast<size_t>(t) * Bpad + b] >> (i & 31)) & 1u;
        acc += G * __fdiv_rn(static_cast<float>(cnt), static_cast<float>(T));
      } else {
        acc += G;
      }
    }
  }
  __shared__ float sh[8][33];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col_live) {
    float t = 0.f;
    for (int y = 0; y < 8; ++y) t += sh[y][threadIdx.x];
    part[static_cast<size_t>(blockIdx.y) * (AP + A) + i] = t;
  }
}

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
struct RowJob { const float* src; float* dst; int nparts; int n; unsigned long long stride; };
struct RowJobs { RowJob j[12]; int count; };
__global__ void __launch_bounds__(256) multi_reduce_rows_kernel(const RowJobs jobs) {
  const RowJob jb = jobs.j[blockIdx.y];
  const int i = blockIdx.x * 32 + threadIdx.x;
  if (blockIdx.x * 32 >= jb.n) return;
  float acc = 0.f;
  if (i < jb.n)
    for (int s = threadIdx.y; s < jb.nparts; s += 8) acc += jb.src[static_cast<size_t>(s) * jb.stride + i];
  __shared__ float sh[8][33];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < jb.n) {
    float t = 0.f;
    for (int y = 0; y < 8; ++y) t += sh[y][threadIdx.x];
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
