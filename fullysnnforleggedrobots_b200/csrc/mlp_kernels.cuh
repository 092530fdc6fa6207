// mlp_kernels.cuh — SIMT ends of the dense ELU critic (AMP/.../modules/actor_critic.py:359-368, 428-430):
// input -> operand planes, the 1-wide output layer and its backward.  The hidden layers run on the tensor
// cores through scan_gemm_kernel<MODE_MLP_FWD / MODE_MLP_DX> and dw_gemm_kernel<true>.
#pragma once
#include "plan.cuh"
#include "umma.cuh"

namespace snn {

// x [B][K] fp32 (row stride xs floats; optional second addend x2: the sum of two gradient sources) -> hi/lo bf16 planes
// as swz64 image [KP/64][Bpad rows]; padding rows / columns are zero
__global__ void planes_from_f32_kernel(const float* __restrict__ x, long long xs, const float* __restrict__ x2, long long x2s,
                                       int B, int Bpad, int K, int KP, uint8_t* __restrict__ img, size_t plane) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int kp8 = KP / 8;
  if (idx >= Bpad * kp8) return;
  const int r = idx / kp8, c0 = (idx - r * kp8) * 8;
  float h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float v = (r < B && c < K) ? x[static_cast<size_t>(r) * xs + c] : 0.f;
    if (x2 != nullptr && r < B && c < K) v += x2[static_cast<size_t>(r) * x2s + c];
    h[j] = bf16_round(v);
    l[j] = v - h[j];
  }
  const size_t off = swz64_offset(r, c0, Bpad);
  *reinterpret_cast<uint4*>(img + off) =
      make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
  *reinterpret_cast<uint4*>(img + plane + off) =
      make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
}

// part[blk][c] = sum over the block's 256 rows of (x + x2)[r, c]   (bias gradient of a linear output layer), c < K <= 128
__global__ void __launch_bounds__(256) col_sum_kernel(const float* __restrict__ x, long long xs, const float* __restrict__ x2,
                                                     long long x2s, int B, int K, float* __restrict__ part) {
  __shared__ float sh[8][128];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  for (int c = cl; c < K; c += 32) {
    float acc = 0.f;
    for (int i = rl; i < 256; i += 8) {
      const int r = blockIdx.x * 256 + i;
      if (r < B) {
        acc += x[static_cast<size_t>(r) * xs + c];
        if (x2 != nullptr) acc += x2[static_cast<size_t>(r) * x2s + c];
      }
    }
    sh[rl][c] = acc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < K; c += 256) {
    float t = 0.f;
    for (int y = 0; y < 8; ++y) t += sh[y][c];
    part[static_cast<size_t>(blockIdx.x) * K + c] = t;
  }
}

__device__ __forceinline__ void unpack8(const uint4 hi, const uint4 lo, float (&x)[8]) {
  const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    x[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
    x[2 * j + 1] = __uint_as_float(hw[j] & 0xFFFF0000u) + __uint_as_float(lw[j] & 0xFFFF0000u);
  }
}

// value[b] = sum_n h[b,n] w[n] + bias.  One warp per row, lane = 8-column piece (NP <= 256).
__global__ void __launch_bounds__(256) value_head_fwd_kernel(const uint8_t* __restrict__ h_img, size_t h_plane, int Bpad,
                                                            int B, int N, int NP, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ value) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  float acc = 0.f;
  if (lane * 8 < NP) {
    const size_t off = swz64_offset(b, lane * 8, Bpad);
    float h[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h_img + off)), __ldg(reinterpret_cast<const uint4*>(h_img + h_plane + off)), h);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = lane * 8 + j;
      if (n < N) acc += h[j] * w[n];
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) value[b] = acc + bias[0];
}

// dz[b,n] = g_v[b] w[n] ELU'(h[b,n]) -> planes; partials per block of 64 rows:
//   part[blk][0..NP)      = sum_b g_v[b] h[b,n]      (d w_out)
//   part[blk][NP..2NP)    = sum_b dz[b,n]            (d bias of the last hidden layer)
//   part[blk][2NP]        = sum_b g_v[b]             (d bias_out)
#ifndef SNN_VH_AHEAD
#define SNN_VH_AHEAD 4
#endif
constexpr int kVhAhead = SNN_VH_AHEAD;      // 1, 2, 4 or 8 (12.28 / - / 12.18 / 12.22 ms per update on one box: profiles/r2zK_vh_ab.txt)
__global__ void __launch_bounds__(256) value_head_bwd_kernel(const uint8_t* __restrict__ h_img, size_t h_plane, int Bpad,
                                                            int B, int N, int NP, const float* __restrict__ w,
                                                            const float* __restrict__ g_v, uint8_t* __restrict__ dz_img,
                                                            size_t dz_plane, float* __restrict__ part) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool col_live = lane * 8 < NP;
  float wl[8], dw[8], dbh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = lane * 8 + j;
    wl[j] = (col_live && n < N) ? w[n] : 0.f;
    dw[j] = 0.f; dbh[j] = 0.f;
  }
  float dbo = 0.f;
  // kVhAhead rows of this warp are requested before the first of them is used (eight dependent round trips otherwise:
  // the kernel is pure load latency)
#pragma unroll
  for (int c = 0; c < 8; c += kVhAhead) {
  uint4 qh[kVhAhead], ql[kVhAhead];
  float gq[kVhAhead];
#pragma unroll
  for (int i = 0; i < kVhAhead; ++i) {
    const int b = blockIdx.x * 64 + (c + i) * 8 + wid;
    const bool on = b < B && col_live;
    const size_t off = on ? swz64_offset(b, lane * 8, Bpad) : 0;
    qh[i] = on ? __ldg(reinterpret_cast<const uint4*>(h_img + off)) : make_uint4(0, 0, 0, 0);
    ql[i] = on ? __ldg(reinterpret_cast<const uint4*>(h_img + h_plane + off)) : make_uint4(0, 0, 0, 0);
    gq[i] = b < B ? __ldg(g_v + b) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < kVhAhead; ++i) {
    const int b = blockIdx.x * 64 + (c + i) * 8 + wid;
    if (b >= Bpad) break;
    const float g = gq[i];
    if (lane == 0) dbo += g;
    if (col_live) {
      const size_t off = swz64_offset(b, lane * 8, Bpad);
      float h[8], dz[8];
      unpack8(qh[i], ql[i], h);      // rows >= B: zeros
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dz[j] = g * wl[j] * (h[j] > 0.f ? 1.f : h[j] + 1.f);
        dw[j] += g * h[j];
        dbh[j] += dz[j];
      }
      uint32_t hp[4], lp[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {      // hi = bf16(dz) two columns per cvt, lo = bf16(dz - hi)
        hp[j] = pack_bf16x2(dz[2 * j], dz[2 * j + 1]);
        lp[j] = pack_bf16x2(dz[2 * j] - __uint_as_float(hp[j] << 16), dz[2 * j + 1] - __uint_as_float(hp[j] & 0xFFFF0000u));
      }
      *reinterpret_cast<uint4*>(dz_img + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      *reinterpret_cast<uint4*>(dz_img + dz_plane + off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    }
  }
  }
  __shared__ float sh[8][2 * 256 + 1];
  if (col_live) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { sh[wid][lane * 8 + j] = dw[j]; sh[wid][256 + lane * 8 + j] = dbh[j]; }
  }
  if (lane == 0) sh[wid][512] = dbo;
  __syncthreads();
  float* out = part + static_cast<size_t>(blockIdx.x) * (2 * NP + 1);
  for (int i = threadIdx.x; i < 2 * NP + 1; i += 256) {
    const int src = i < NP ? i : (i < 2 * NP ? 256 + (i - NP) : 512);
    float t = 0.f;
    for (int y = 0; y < 8; ++y) t += sh[y][src];
    out[i] = t;
  }
}

}  // namespace snn
