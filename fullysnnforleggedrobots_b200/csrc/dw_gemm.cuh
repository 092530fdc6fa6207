// dw_gemm.cuh — weight-gradient GEMM  dW[n,k] = sum_r g_c[r,n] * x[r,k]   (r = (t, b) rows)
// (the "dW += g_c^T x_t" line of SURVEY.md §8a, i.e. what autograd's mm-backward of nn.Linear does
// for AMP/.../modules/actor_critic.py:228 summed over the T spike steps).
//
// The contraction runs over batch rows, so both operands are MN-major for the tensor core:
//   A = g_c hi/lo bf16 planes, read straight from their swz64 images (rows = r = MMA-K,
//       cols = n = MMA-M) with 1-D bulk copies,
//   B = input spikes x (exact in bf16), expanded in shared memory from the bit-packed trace
//       (rows = r = MMA-K, cols = k = MMA-N).
// One CTA owns one [128 x <=256] fp32 output tile in TMEM and a contiguous slice of r (split-K over
// the batch); partial tiles go to a workspace and are summed by multi_reduce_kernel — a fixed
// order, so gradients are bit-reproducible run to run.
#pragma once
#include "plan.cuh"
#include "umma.cuh"

namespace snn {

struct DwGemmParams {
  const uint8_t* g_img;    // 2 planes, swz64 image [NP/64][R rows]
  size_t g_plane;
  const uint32_t* x_bits;  // [KP/32][R]           (spiking layers: binary input, expanded in shared memory)
  const uint8_t* x_img;    // 2 planes, swz64 image [KP/64][R rows]   (dense layers: real-valued input)
  size_t x_plane;
  int NP, KP;
  int64_t R;
  int n_tiles;             // ceil(KP / 256)
  int splits;              // CTAs per output tile
  int rblocks;             // R / 64
  float* partial;          // [gridDim.x][128][256]
  unsigned long long* dbg_out;   // optional [gridDim.x][16] wait-cycle counters (tools/role_cycles.py)
};

constexpr int kDwStages = 3;
constexpr int kDwStageBytes = 65536;   // 32 KiB A (2 planes x 2 n-chunks x 8 KiB) + 32 KiB B (4 k-chunks x 8 KiB)
constexpr int kDwBitStages = 8;        // ring of spike-word stages: 8 runs of 64 rows x 4 B
constexpr int kDwBitStageBytes = 2048;
constexpr size_t kDwSmemBytes = 1024 + kDwStages * kDwStageBytes + kDwBitStages * kDwBitStageBytes + 512;

// DENSE = false: B operand = spike bits expanded in smem (3 stages of 64 KiB).
// DENSE = true : B operand = hi/lo planes bulk-copied like A (2 stages of 96 KiB), three products.
template <bool DENSE>
__device__ __forceinline__ void dw_gemm_body(const DwGemmParams& p, const int cta) {
  constexpr int kStages = DENSE ? 2 : 3;
  constexpr int kStageBytes = DENSE ? 98304 : 65536;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* bits_ring = ring + kStages * kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(bits_ring + kDwBitStages * kDwBitStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* full_bits = acc_full + 1;
  uint64_t* empty_bits = full_bits + kDwBitStages;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(empty_bits + kDwBitStages);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef SNN_ABLATE
  unsigned long long wcyc[3] = {0, 0, 0};
  const long long t_start = clock64();
#endif
  const int tile = cta / p.splits, split = cta - tile * p.splits;
  const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
  const int kc0 = nt * 4;                                   // first 64-wide k chunk of this tile
  const int nck = min(4, p.KP / 64 - kc0);                  // k chunks in this tile
  const int per = (p.rblocks + p.splits - 1) / p.splits;
  const int rb0 = split * per, rb1 = min(p.rblocks, rb0 + per);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], DENSE ? 1 : 3); mbar_init(&empty[s], 1); }
    for (int s = 0; s < kDwBitStages; ++s) { mbar_init(&full_bits[s], 1); mbar_init(&empty_bits[s], 2); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // operand producer: warp-uniform loop, one elected lane issues the bulk copies of the g_c (and dense x) tiles
    uint32_t i = 0;
    for (int rb = rb0; rb < rb1; ++rb, ++i) {
      const uint32_t s = i % kStages;
      SNN_TWAIT(&empty[s], ((i / kStages) & 1) ^ 1, 1);
      if (elect_one()) {
        mbar_expect_tx(&full[s], DENSE ? static_cast<uint32_t>(32768 + nck * 16384) : 32768u);
        uint8_t* dst = ring + s * kStageBytes;
        if (DENSE) {
          for (int pl = 0; pl < 2; ++pl)
            for (int kcl = 0; kcl < nck; ++kcl)
              bulk_g2s(dst + 32768 + pl * 32768 + kcl * 8192,
                       p.x_img + pl * p.x_plane + (static_cast<size_t>(kc0 + kcl) * p.R + static_cast<size_t>(rb) * 64) * 128,
                       8192u, &full[s]);
        }
        for (int pl = 0; pl < 2; ++pl)
          for (int nc = 0; nc < 2; ++nc)
            bulk_g2s(dst + pl * 16384 + nc * 8192,
                     p.g_img + pl * p.g_plane + (static_cast<size_t>(mt * 2 + nc) * p.R + static_cast<size_t>(rb) * 64) * 128,
                     8192u, &full[s]);
      }
      __syncwarp();
    }
  } else if (warp == 2 && !DENSE) {
    // spike-word producer, decoupled from the operand producer so that it runs the full depth of its ring ahead
    uint32_t i = 0;
    for (int rb = rb0; rb < rb1; ++rb, ++i) {
      const uint32_t sbit = i % kDwBitStages;
      SNN_TWAIT(&empty_bits[sbit], ((i / kDwBitStages) & 1) ^ 1, 0);
      if (elect_one()) {   // for every k-chunk two runs of 64 rows x 4 B
        mbar_expect_tx(&full_bits[sbit], static_cast<uint32_t>(nck * 512));
        uint8_t* bd = bits_ring + sbit * kDwBitStageBytes;
        for (int w = 0; w < 2 * nck; ++w)
          bulk_g2s(bd + w * 256, p.x_bits + static_cast<size_t>(2 * kc0 + w) * p.R + static_cast<size_t>(rb) * 64, 256u,
                   &full_bits[sbit]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues
    const uint32_t idesc = make_idesc_bf16(128, static_cast<uint32_t>(nck * 64), 1, 1);
    const uint32_t dhi = smem_desc_hi(1024);
    const uint32_t ring_lo = smem_desc_lo(smem_u32(ring), 8192);
    uint32_t i = 0;
    for (int rb = rb0; rb < rb1; ++rb, ++i) {
      const uint32_t s = i % kStages;
      SNN_TWAIT(&full[s], (i / kStages) & 1, 0);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t a_lo = ring_lo + s * (kStageBytes >> 4);
        const uint32_t b_lo = a_lo + (32768 >> 4);
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          if (DENSE) {
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) {
              const uint32_t pa = pr == 2 ? 1 : 0, pb = pr == 1 ? 1 : 0;
              umma_bf16_lohi(tmem_base, a_lo + pa * (16384 >> 4) + k16 * (2048 >> 4), dhi,
                             b_lo + pb * (32768 >> 4) + k16 * (2048 >> 4), dhi, idesc, (i > 0 || k16 > 0 || pr > 0) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
              umma_bf16_lohi(tmem_base, a_lo + pl * (16384 >> 4) + k16 * (2048 >> 4), dhi, b_lo + k16 * (2048 >> 4), dhi, idesc,
                             (i > 0 || k16 > 0 || pl > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
        if (rb == rb1 - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
    if (rb1 <= rb0 && elect_one()) umma_commit(acc_full);
  } else if (warp >= 4) {
    // spike-bit expanders: warps 4 .. 4 + 2 * kStages - 1.  Ring slot ew is owned by the warp PAIR (ew, ew + kStages),
    // each expanding 32 of the stage's 64 sample-steps (all nck k-chunks).  A fixed owner per slot keeps each slot's
    // uses in order: a warp can never wait on an mbarrier phase that is two phases ahead (parity aliasing).  The
    // expansion is the slowest role of this kernel, hence two warps per slot.  Warps 4-7 run the epilogue afterwards.
    const int ew = (warp - 4) % kStages, hsel = (warp - 4) / kStages;
    const bool expands = !DENSE && warp < 4 + 2 * kStages;
    const int sw = lane & 7;
    uint32_t i = 0;
    for (int rb = rb0; rb < (expands ? rb1 : rb0); ++rb, ++i) {
      if (static_cast<int>(i % kStages) != ew) continue;
      const uint32_t sbit = i % kDwBitStages;
      SNN_TWAIT(&full_bits[sbit], (i / kDwBitStages) & 1, 0);
      const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(bits_ring + sbit * kDwBitStageBytes);
      const int row = lane + 32 * hsel;
      uint32_t w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = (u < 2 * nck) ? wsrc[u * 64 + row] : 0u;
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bits[sbit]);
      const uint32_t s = i % kStages;
      SNN_TWAIT(&empty[s], ((i / kStages) & 1) ^ 1, 1);
      uint8_t* rbase = ring + s * kStageBytes + 32768 + (row >> 3) * 1024 + sw * 128;
#pragma unroll
      for (int kcl = 0; kcl < 4; ++kcl) {
        if (kcl < nck) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t byte = ((j < 4 ? w[2 * kcl] : w[2 * kcl + 1]) >> ((j & 3) * 8)) & 0xFFu;
            *reinterpret_cast<uint4*>(rbase + kcl * 8192 + ((j ^ sw) << 4)) = expand_spike_byte(byte);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
    if (warp < 8) {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      float* out = p.partial + (static_cast<size_t>(cta) * 128 + row) * 256;
      mbar_wait(acc_full, 0);
      tc_fence_after_sync();
      const bool any = rb1 > rb0;
      for (int g = 0; g < nck * 4; ++g) {
        uint32_t x[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g * 16), x);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(out + g * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = any ? make_float4(__uint_as_float(x[4 * j]), __uint_as_float(x[4 * j + 1]), __uint_as_float(x[4 * j + 2]),
                                     __uint_as_float(x[4 * j + 3]))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
#ifdef SNN_ABLATE
  if (p.dbg_out != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 4)) {
    const int role = warp == 0 ? 0 : (warp == 1 ? 1 : 3);
    unsigned long long* o = p.dbg_out + static_cast<size_t>(cta) * 16 + role * 4;
    o[0] = wcyc[0]; o[1] = wcyc[1]; o[2] = wcyc[2];
    o[3] = static_cast<unsigned long long>(clock64() - t_start);
  }
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
}

template <bool DENSE>
__global__ void __launch_bounds__(384, 1) dw_gemm_kernel(const DwGemmParams p) {
  dw_gemm_body<DENSE>(p, static_cast<int>(blockIdx.x));
}

// The weight gradients of ALL spiking layers in one launch: CTA c belongs to the job (layer) whose range contains
// it.  Every dW launch pays ~14 us of fixed cost (pipeline fill, accumulator drain, 128 KiB partial tile store per
// CTA) that dominates the narrow upper layers; one launch pays it once and balances the layers over the SMs.
struct DwJobs { DwGemmParams j[SNN_MAX_LAYERS]; int first_cta[SNN_MAX_LAYERS + 1]; int count; };
__global__ void __launch_bounds__(384, 1) dw_gemm_multi_kernel(const DwJobs jobs) {
  int l = 0;
  while (l + 1 < jobs.count && static_cast<int>(blockIdx.x) >= jobs.first_cta[l + 1]) ++l;
  dw_gemm_body<false>(jobs.j[l], static_cast<int>(blockIdx.x) - jobs.first_cta[l]);
}

}  // namespace snn
