// scan_gemm.cuh — the "T-accumulator GEMM + time scan" kernel family (tcgen05 / TMEM / bulk-TMA).
//
// One work item = (128 batch rows) x (one chunk of <= NCmax output neurons).  For such an item the
// kernel keeps T accumulators [128 x NC] fp32 in TMEM (T * NCmax <= 512 columns), one per spike
// timestep, fills them with tcgen05.mma over the whole contraction dimension, and then the
// epilogue warps run the sequential-in-time recurrence with all neuron state in registers:
//
//   MODE_FWD     A = spikes of the layer below (bit-packed in HBM, staged by bulk copies into a
//                shared-memory ring and expanded there to bf16 0/1 operand tiles), B = bf16
//                hi/mid/lo planes of W.   Epilogue = LIF update
//                (AMP/.../modules/actor_critic.py:216-232): c = .5c + (Wx+b); v = .75 v (1-s) + c;
//                s = v > .5.   Emits spike bits (+ fp32 voltage trace when training).
//   MODE_DX      A = bf16 hi/lo planes of g_c of the layer above, B = hi/lo planes of W^T.
//                Epilogue = BPTT recurrence of SURVEY.md §8a for the layer below (surrogate
//                gradient of PseudoSpikeRect, actor_critic.py:154-167).  Emits g_c planes + db.
//   MODE_DX_ENC  same GEMM into the encoder neurons; epilogue = straight-through encoder
//                recurrence (actor_critic.py:50-59,116-119): Gamma_t = g_t + 0.001 Gamma_{t+1}.
//                Emits g_a = sum_t Gamma_t.
//
// Warp roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer + TMEM owner, warps 4-7 and 8-11 =
// two epilogue groups (TMEM lane quarter = warp % 4; the groups split the 16-column groups of a
// chunk), warps 12-15 = spike-bit expanders (FWD only).
//
// Voltage traces use a column-group-major layout  v[t][n / 16][Bpad][16]  so that the 32 rows a warp
// owns are one contiguous 2 KiB run for every 16-column group (coalesced writes here, coalesced
// reads in the backward epilogue).
#pragma once
#include "plan.cuh"
#include "umma.cuh"

namespace snn {

enum { MODE_FWD = 0, MODE_DX = 1, MODE_DX_ENC = 2, MODE_MLP_FWD = 3, MODE_MLP_DX = 4, MODE_FWD_TA = 5 };

struct ScanGemmParams {
  // A operand
  const uint32_t* a_bits;   // FWD: [KP/32][R]
  const uint8_t* a_img;     // DX: 2 planes of swz64 image [KP/64][R rows]
  size_t a_plane;           // bytes between A planes
  // B operand: planes of swz64 image [KP/64][NPout rows]
  const uint8_t* b_img;
  size_t b_plane;
  int b_planes;             // FWD: 1..3 ; DX: 2
  int KP;                   // contraction length (padded, % 64 == 0)
  int NPout;                // output neurons (padded, % 32 == 0)
  int NCmax;                // chunk width, % 32 == 0, T * NCmax <= 512
  int T;
  int B, Bpad;
  int64_t R;                // T * Bpad
  int nitems, nchunks;
  unsigned long long* dbg_out;   // optional [gridDim.x][16] cycle counters of the barrier waits (tools/role_cycles.py)
  int sets;                 // accumulator sets actually used (<= ScanCfg::SETS); 0 = ScanCfg::SETS
  int dbg;                  // SNN_DBG ablation mask (timing experiments only): 1 = no MMA, 2 = no scan, 4 = no expand
  // FWD
  const float* bias;        // [NPout]
  uint32_t* out_bits;       // [NPout/32][R]
  float* v_out;             // [T][NPout/16][Bpad][16] or null (inference)
  // DX
  const float* v_in;        // [T][NPout/16][Bpad][16] voltages of the layer whose gradient is produced
  uint8_t* g_img;           // 2 planes of swz64 image [NPout/64][R rows]
  size_t g_plane;
  float* db_part;           // [gridDim.x][NPout]
  // DX_ENC
  float* g_a;               // [Bpad][NPout]
  // MLP_FWD: writes ELU(acc + bias) as hi/lo planes into g_img.  MLP_DX: reads the layer's ELU output planes
  const uint8_t* h_img;     // 2 planes of swz64 image [NPout/64][R rows]
  size_t h_plane;
};

template <int MODE> struct ScanCfg;
template <> struct ScanCfg<MODE_FWD> { static constexpr int A_STAGE = 16384, NA = 6, NB = 2, BP = 3, NBITS = 16, SETS = 1, THREADS = 512; };
template <> struct ScanCfg<MODE_DX> { static constexpr int A_STAGE = 32768, NA = 4, NB = 2, BP = 2, NBITS = 0, SETS = 2, THREADS = 384; };
template <> struct ScanCfg<MODE_DX_ENC> { static constexpr int A_STAGE = 32768, NA = 4, NB = 2, BP = 2, NBITS = 0, SETS = 1, THREADS = 384; };
template <> struct ScanCfg<MODE_MLP_FWD> { static constexpr int A_STAGE = 32768, NA = 4, NB = 2, BP = 2, NBITS = 0, SETS = 2, THREADS = 384; };
// forward with the spike operand in TENSOR memory (T * NCmax <= 320 accumulator columns + 6 A stages of 32 columns)
template <> struct ScanCfg<MODE_FWD_TA> { static constexpr int A_STAGE = 0, NA = 6, NB = 4, BP = 3, NBITS = 16, SETS = 1, THREADS = 512; };
template <> struct ScanCfg<MODE_MLP_DX> { static constexpr int A_STAGE = 32768, NA = 4, NB = 2, BP = 2, NBITS = 0, SETS = 2, THREADS = 384; };

constexpr int kScanMaxNP = 2048;   // floats of per-column smem scratch (FWD: bias, DX: db); layer widths <= 2048
constexpr int kTaAccCols = 320;    // MODE_FWD_TA: TMEM columns [0, 320) = accumulators, [320, 512) = 6 A stages
constexpr int kBitsStage = 1024;   // one A stage worth of spike words: 2 x (128 rows x 4 B)

// 16 consecutive columns of one row -> hi/lo bf16 planes of a swz64 image (two 16-byte pieces per plane)
__device__ __forceinline__ void store_planes16(uint8_t* img, size_t plane, int64_t R, size_t r, int col0, const float (&x)[16]) {
  float hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { hi[j] = bf16_round(x[j]); lo[j] = x[j] - hi[j]; }
  const size_t base = static_cast<size_t>(col0 >> 6) * R * 128 + (r >> 3) * 1024 + (r & 7) * 128;
  const int p0 = (col0 & 63) >> 3;
  const int sw = static_cast<int>(r & 7);
  const uint4 h0 = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]), pack_bf16x2(hi[6], hi[7]));
  const uint4 h1 = make_uint4(pack_bf16x2(hi[8], hi[9]), pack_bf16x2(hi[10], hi[11]), pack_bf16x2(hi[12], hi[13]), pack_bf16x2(hi[14], hi[15]));
  const uint4 l0 = make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]), pack_bf16x2(lo[6], lo[7]));
  const uint4 l1 = make_uint4(pack_bf16x2(lo[8], lo[9]), pack_bf16x2(lo[10], lo[11]), pack_bf16x2(lo[12], lo[13]), pack_bf16x2(lo[14], lo[15]));
  *reinterpret_cast<uint4*>(img + base + ((p0 ^ sw) << 4)) = h0;
  *reinterpret_cast<uint4*>(img + base + (((p0 + 1) ^ sw) << 4)) = h1;
  *reinterpret_cast<uint4*>(img + plane + base + ((p0 ^ sw) << 4)) = l0;
  *reinterpret_cast<uint4*>(img + plane + base + (((p0 + 1) ^ sw) << 4)) = l1;
}
// inverse: hi + lo of 16 consecutive columns of one row
__device__ __forceinline__ void load_planes16(const uint8_t* img, size_t plane, int64_t R, size_t r, int col0, float (&x)[16]) {
  const size_t base = static_cast<size_t>(col0 >> 6) * R * 128 + (r >> 3) * 1024 + (r & 7) * 128;
  const int p0 = (col0 & 63) >> 3;
  const int sw = static_cast<int>(r & 7);
  const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(img + base + ((p0 ^ sw) << 4)));
  const uint4 h1 = __ldg(reinterpret_cast<const uint4*>(img + base + (((p0 + 1) ^ sw) << 4)));
  const uint4 l0 = __ldg(reinterpret_cast<const uint4*>(img + plane + base + ((p0 ^ sw) << 4)));
  const uint4 l1 = __ldg(reinterpret_cast<const uint4*>(img + plane + base + (((p0 + 1) ^ sw) << 4)));
  const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    x[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
    x[2 * j + 1] = __uint_as_float(hw[j] & 0xFFFF0000u) + __uint_as_float(lw[j] & 0xFFFF0000u);
  }
}

template <int MODE>
__host__ __device__ constexpr size_t scan_gemm_smem_bytes(int NCmax) {
  return 1024 /*align slack*/ + static_cast<size_t>(ScanCfg<MODE>::NA) * ScanCfg<MODE>::A_STAGE +
         static_cast<size_t>(ScanCfg<MODE>::NB) * ScanCfg<MODE>::BP * NCmax * 128 + 4096 /*LUT*/ +
         static_cast<size_t>(ScanCfg<MODE>::NBITS) * kBitsStage + kScanMaxNP * 4 + 512;
}

// x[16] per lane (16 columns of this lane's row) -> sum over the warp's 32 rows; lanes 2c and 2c+1 both
// return column c's total
__device__ __forceinline__ float warp_transpose_reduce16(float (&x)[16], int lane) {
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

#ifndef SNN_TWAIT
#define SNN_TWAIT(bar, par, slot)                                   \
  do {                                                              \
    const long long t0__ = clock64();                               \
    mbar_wait(bar, par);                                            \
    wcyc[slot] += static_cast<unsigned long long>(clock64() - t0__); \
  } while (0)
#endif

template <int MODE>
__global__ void __launch_bounds__(ScanCfg<MODE>::THREADS, 1) scan_gemm_kernel(const ScanGemmParams p) {
  using C = ScanCfg<MODE>;
  constexpr bool IS_FWD = MODE == MODE_FWD || MODE == MODE_FWD_TA;
  constexpr bool TA = MODE == MODE_FWD_TA;
  constexpr uint32_t NBM = C::NBITS > 0 ? C::NBITS : 1;   // modulus of the spike-word ring (unused in DX modes)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage = C::BP * p.NCmax * 128;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = a_ring + C::NA * C::A_STAGE;
  uint4* lut = reinterpret_cast<uint4*>(b_ring + C::NB * b_stage);
  uint8_t* bits_ring = reinterpret_cast<uint8_t*>(lut) + 4096;
  float* s_col = reinterpret_cast<float*>(bits_ring + C::NBITS * kBitsStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_col + kScanMaxNP);
  uint64_t* full_a = bars;                 // [NA]
  uint64_t* empty_a = full_a + C::NA;      // [NA]
  uint64_t* full_b = empty_a + C::NA;      // [NB]
  uint64_t* empty_b = full_b + C::NB;      // [NB]
  uint64_t* acc_full = empty_b + C::NB;    // [SETS]: SETS == 2 double-buffers the accumulators (T * NCmax <= 256 per set)
  uint64_t* acc_empty = acc_full + C::SETS;  // [SETS]
  uint64_t* full_bits = acc_empty + C::SETS;     // [NBITS]
  uint64_t* empty_bits = full_bits + C::NBITS;   // [NBITS]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(empty_bits + C::NBITS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = p.KP / 64;
  const uint32_t nsets = (C::SETS == 2 && p.sets != 1) ? 2u : 1u;
  unsigned long long wcyc[3] = {0, 0, 0};
  const long long t_start = clock64();

  if (tid == 0) {
    for (int s = 0; s < C::NA; ++s) {
      mbar_init(&full_a[s], TA ? 4 : 1);
      mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < C::NB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], TA ? 2 : 1); }
    for (int s = 0; s < C::NBITS; ++s) { mbar_init(&full_bits[s], 1); mbar_init(&empty_bits[s], TA ? 4 : 1); }
    for (int s = 0; s < C::SETS; ++s) { mbar_init(&acc_full[s], TA ? 2 : 1); mbar_init(&acc_empty[s], 256); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  if (IS_FWD) {
    // LUT: byte of 8 spike bits -> eight bf16 {0, 1}
    for (int i = tid; i < 256; i += blockDim.x) {
      uint4 v;
      v.x = ((i & 1) ? 0x3F80u : 0u) | ((i & 2) ? 0x3F800000u : 0u);
      v.y = ((i & 4) ? 0x3F80u : 0u) | ((i & 8) ? 0x3F800000u : 0u);
      v.z = ((i & 16) ? 0x3F80u : 0u) | ((i & 32) ? 0x3F800000u : 0u);
      v.w = ((i & 64) ? 0x3F80u : 0u) | ((i & 128) ? 0x3F800000u : 0u);
      lut[i] = v;
    }
    for (int i = tid; i < p.NPout; i += blockDim.x) s_col[i] = p.bias[i];
  } else if (MODE == MODE_MLP_FWD) {
    for (int i = tid; i < p.NPout; i += blockDim.x) s_col[i] = p.bias[i];
  } else if (MODE == MODE_DX || MODE == MODE_MLP_DX) {
    for (int i = tid; i < p.NPout; i += blockDim.x) s_col[i] = 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================================================== producer (warp-uniform loop, one elected lane issues)
    {
      uint32_t ib = 0, ia = 0;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        const int m = item / p.nchunks, c = item - m * p.nchunks;
        const int n0 = c * p.NCmax;
        const int ncw = min(p.NCmax, p.NPout - n0);
        for (int kc = 0; kc < KC; ++kc) {
          const uint32_t sb = ib % C::NB;
          SNN_TWAIT(&empty_b[sb], ((ib / C::NB) & 1) ^ 1, 0);
          if (elect_one()) {
            mbar_expect_tx(&full_b[sb], static_cast<uint32_t>(p.b_planes * ncw * 128));
            for (int pl = 0; pl < p.b_planes; ++pl)
              bulk_g2s(b_ring + sb * b_stage + pl * p.NCmax * 128,
                       p.b_img + pl * p.b_plane + (static_cast<size_t>(kc) * p.NPout + n0) * 128,
                       static_cast<uint32_t>(ncw * 128), &full_b[sb]);
          }
          __syncwarp();
          ++ib;
          for (int t = 0; t < p.T; ++t, ++ia) {
            const size_t row0 = static_cast<size_t>(t) * p.Bpad + static_cast<size_t>(m) * 128;
            if (IS_FWD) {
              // spike words of this A stage: two 512-byte runs (128 rows x one 32-column word each)
              const uint32_t s = ia % NBM;
              SNN_TWAIT(&empty_bits[s], ((ia / NBM) & 1) ^ 1, 1);
              if (elect_one()) {
                if (p.dbg & 8) {
                  mbar_arrive(&full_bits[s]);
                } else {
                  mbar_expect_tx(&full_bits[s], static_cast<uint32_t>(kBitsStage));
                  uint8_t* dst = bits_ring + s * kBitsStage;
                  bulk_g2s(dst, p.a_bits + static_cast<size_t>(2 * kc) * p.R + row0, 512u, &full_bits[s]);
                  bulk_g2s(dst + 512, p.a_bits + static_cast<size_t>(2 * kc + 1) * p.R + row0, 512u, &full_bits[s]);
                }
              }
              __syncwarp();
            } else {
              const uint32_t sa = ia % C::NA;
              SNN_TWAIT(&empty_a[sa], ((ia / C::NA) & 1) ^ 1, 1);
              if (elect_one()) {
                mbar_expect_tx(&full_a[sa], 32768u);
                for (int pl = 0; pl < 2; ++pl)
                  bulk_g2s(a_ring + sa * C::A_STAGE + pl * 16384,
                           p.a_img + pl * p.a_plane + (static_cast<size_t>(kc) * p.R + row0) * 128, 16384u,
                           &full_a[sa]);
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp == 1 || (TA && warp == 2)) {
    // ===================================================== MMA issuer (warp-uniform loop, one elected lane issues)
    // MODE_FWD_TA runs TWO issuer warps (even / odd timesteps: independent accumulators)
    {
      const int issuer = warp - 1;
      uint32_t ib = 0, ia = 0, it = 0;
      const uint32_t dhi = smem_desc_hi(1024);
      const uint32_t a_ring_lo = smem_desc_lo(smem_u32(a_ring), 16);
      const uint32_t b_ring_lo = smem_desc_lo(smem_u32(b_ring), 16);
      const uint32_t b_plane_lo = static_cast<uint32_t>(p.NCmax * 128) >> 4;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it) {
        const int c = item % p.nchunks;
        const int n0 = c * p.NCmax;
        const int ncw = min(p.NCmax, p.NPout - n0);
        const uint32_t idesc = make_idesc_bf16(128, static_cast<uint32_t>(ncw), 0, 0);
        const uint32_t aset = it % nsets;
        SNN_TWAIT(&acc_empty[aset], ((it / nsets) & 1) ^ 1, 0);
        tc_fence_after_sync();
        for (int kc = 0; kc < KC; ++kc) {
          const uint32_t sb = ib % C::NB;
          SNN_TWAIT(&full_b[sb], (ib / C::NB) & 1, 1);
          const uint32_t b_lo = b_ring_lo + sb * (static_cast<uint32_t>(b_stage) >> 4);
          for (int t = 0; t < p.T; ++t, ++ia) {
            if (TA && (t & 1) != issuer) continue;         // TA: two issuer warps split the timesteps
            const uint32_t sa = ia % C::NA;
            SNN_TWAIT(&full_a[sa], (ia / C::NA) & 1, 2);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t a_lo = a_ring_lo + sa * (C::A_STAGE >> 4);
              const uint32_t d = tmem_base + aset * 256 + static_cast<uint32_t>(t * p.NCmax);
              uint32_t acc = kc > 0 ? 1u : 0u;
              if (p.dbg & 1) {
              } else if (TA) {
                const uint32_t a_t = tmem_base + kTaAccCols + sa * 32;
                for (int pl = 0; pl < p.b_planes; ++pl) {
#pragma unroll
                  for (int k16 = 0; k16 < 4; ++k16) {
                    umma_bf16_ts(d, a_t + k16 * 8, b_lo + pl * b_plane_lo + k16 * 2, dhi, idesc, acc);
                    acc = 1u;
                  }
                }
              } else if (MODE == MODE_FWD) {
                for (int pl = 0; pl < p.b_planes; ++pl) {
#pragma unroll
                  for (int k16 = 0; k16 < 4; ++k16) {
                    umma_bf16_lohi(d, a_lo + k16 * 2, dhi, b_lo + pl * b_plane_lo + k16 * 2, dhi, idesc, acc);
                    acc = 1u;
                  }
                }
              } else {
                // g_c ~ (a_hi + a_lo), W^T ~ (b_hi + b_lo): hi*hi + hi*lo + lo*hi (lo*lo < 2^-16 relative)
#pragma unroll
                for (int pr = 0; pr < 3; ++pr) {
                  const uint32_t pa = pr == 2 ? 1 : 0, pb = pr == 1 ? 1 : 0;
#pragma unroll
                  for (int k16 = 0; k16 < 4; ++k16) {
                    umma_bf16_lohi(d, a_lo + pa * (16384 >> 4) + k16 * 2, dhi, b_lo + pb * b_plane_lo + k16 * 2, dhi,
                                   idesc, acc);
                    acc = 1u;
                  }
                }
              }
              if (p.dbg & 16) mbar_arrive(&empty_a[sa]); else umma_commit(&empty_a[sa]);
            }
            __syncwarp();
          }
          // this issuer's MMAs of the k-chunk are all issued: release the B stage / publish the accumulators
          if (elect_one()) {
            umma_commit(&empty_b[sb]);
            if (kc == KC - 1) umma_commit(&acc_full[aset]);
          }
          __syncwarp();
          ++ib;
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================================== epilogue: the time scan (two groups of 4 warps)
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;          // 0: even 16-column groups, 1: odd ones
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t it = 0, gctr = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it) {
      const int m = item / p.nchunks, c = item - m * p.nchunks;
      const int n0 = c * p.NCmax;
      const int ncw = min(p.NCmax, p.NPout - n0);
      const int b = m * 128 + row;
      const bool valid = b < p.B;
      if (MODE == MODE_DX && valid) {
        // the scan below reads this row's voltages [t][g] (64 B each) straight from HBM-resident traces:
        // pull them into L2 now, while the tensor core is still busy with this item's MMAs
        const size_t vstep = static_cast<size_t>(p.NPout >> 4) * p.Bpad * 16;
        for (int g = static_cast<int>((half + gctr) & 1u); g < ncw / 16; g += 2) {
          const float* vb = p.v_in + (static_cast<size_t>((n0 >> 4) + g) * p.Bpad + b) * 16;
          for (int t = 0; t < p.T; ++t)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + static_cast<size_t>(t) * vstep));
        }
      }
      const uint32_t aset = it % nsets;
      SNN_TWAIT(&acc_full[aset], (it / nsets) & 1, 0);
      tc_fence_after_sync();
      const int ng = (p.dbg & 2) ? 0 : ncw / 16;
      for (int g = static_cast<int>((half + gctr) & 1u); g < ng; g += 2) {
        const int col0 = n0 + g * 16;
        const uint32_t t_col = t_lane + aset * 256 + static_cast<uint32_t>(g * 16);
        if (IS_FWD) {
          float cur[16], vol[16], bia[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { cur[j] = 0.f; vol[j] = 0.f; bia[j] = s_col[col0 + j]; }
          uint32_t sprev = 0;
          uint32_t x[16], xn[16];
          tmem_ld16(t_col, x);
          tmem_ld_wait();
          uint16_t* bits16 = reinterpret_cast<uint16_t*>(p.out_bits) + (static_cast<size_t>(col0 >> 5) * p.R) * 2 + ((col0 >> 4) & 1);
          for (int t = 0; t < p.T; ++t) {
            if (t + 1 < p.T) tmem_ld16(t_col + static_cast<uint32_t>((t + 1) * p.NCmax), xn);
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float syn = __fadd_rn(__uint_as_float(x[j]), bia[j]);
              const float cj = __fadd_rn(__fmul_rn(cur[j], 0.5f), syn);
              const float vj = ((sprev >> j) & 1u) ? cj : __fadd_rn(__fmul_rn(vol[j], 0.75f), cj);
              cur[j] = cj;
              vol[j] = vj;
              word |= (vj > 0.5f ? 1u : 0u) << j;
            }
            const size_t r = static_cast<size_t>(t) * p.Bpad + b;
            bits16[r * 2] = static_cast<uint16_t>(valid ? word : 0u);
            if (p.v_out != nullptr && valid) {
              float4* dst = reinterpret_cast<float4*>(
                  p.v_out + ((static_cast<size_t>(t) * (p.NPout >> 4) + (col0 >> 4)) * p.Bpad + b) * 16);
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[j] = make_float4(vol[4 * j], vol[4 * j + 1], vol[4 * j + 2], vol[4 * j + 3]);
            }
            sprev = word;
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = xn[j];
          }
        } else if (MODE == MODE_DX) {
          float gvn[16], gcn[16], dbacc[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { gvn[j] = 0.f; gcn[j] = 0.f; dbacc[j] = 0.f; }
          const float* vbase = p.v_in + (static_cast<size_t>(col0 >> 4) * p.Bpad + b) * 16;
          const size_t vstep = static_cast<size_t>(p.NPout >> 4) * p.Bpad * 16;     // floats between timesteps
          float4 vn[4];
          if (valid) {
            const float4* src = reinterpret_cast<const float4*>(vbase + static_cast<size_t>(p.T - 1) * vstep);
#pragma unroll
            for (int j = 0; j < 4; ++j) vn[j] = __ldg(src + j);
          }
          for (int t = p.T - 1; t >= 0; --t) {
            uint32_t x[16];
            tmem_ld16(t_col + static_cast<uint32_t>(t * p.NCmax), x);
            float v[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[4 * j] = valid ? vn[j].x : 0.f; v[4 * j + 1] = valid ? vn[j].y : 0.f;
              v[4 * j + 2] = valid ? vn[j].z : 0.f; v[4 * j + 3] = valid ? vn[j].w : 0.f;
            }
            if (valid && t > 0) {   // prefetch the next (earlier) timestep's voltages
              const float4* src = reinterpret_cast<const float4*>(vbase + static_cast<size_t>(t - 1) * vstep);
#pragma unroll
              for (int j = 0; j < 4; ++j) vn[j] = __ldg(src + j);
            }
            tmem_ld_wait();
            const size_t r = static_cast<size_t>(t) * p.Bpad + b;
            float hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float gup = valid ? __uint_as_float(x[j]) : 0.f;
              const float vj = v[j];
              const float gs = gup - gvn[j] * (0.75f * vj);
              const float win = fabsf(__fsub_rn(vj, 0.5f)) < 0.5f ? 1.f : 0.f;
              const float keep = vj > 0.5f ? 0.f : 0.75f;
              const float gv = gs * win + gvn[j] * keep;
              const float gc = gv + 0.5f * gcn[j];
              gvn[j] = gv;
              gcn[j] = gc;
              dbacc[j] += gc;
              hi[j] = bf16_round(gc);
              lo[j] = gc - hi[j];
            }
            const size_t base = static_cast<size_t>(col0 >> 6) * p.R * 128 + (r >> 3) * 1024 + (r & 7) * 128;
            const int p0 = (col0 & 63) >> 3;
            const int sw = static_cast<int>(r & 7);
            const uint4 h0 = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]), pack_bf16x2(hi[6], hi[7]));
            const uint4 h1 = make_uint4(pack_bf16x2(hi[8], hi[9]), pack_bf16x2(hi[10], hi[11]), pack_bf16x2(hi[12], hi[13]), pack_bf16x2(hi[14], hi[15]));
            const uint4 l0 = make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]), pack_bf16x2(lo[6], lo[7]));
            const uint4 l1 = make_uint4(pack_bf16x2(lo[8], lo[9]), pack_bf16x2(lo[10], lo[11]), pack_bf16x2(lo[12], lo[13]), pack_bf16x2(lo[14], lo[15]));
            *reinterpret_cast<uint4*>(p.g_img + base + ((p0 ^ sw) << 4)) = h0;
            *reinterpret_cast<uint4*>(p.g_img + base + (((p0 + 1) ^ sw) << 4)) = h1;
            *reinterpret_cast<uint4*>(p.g_img + p.g_plane + base + ((p0 ^ sw) << 4)) = l0;
            *reinterpret_cast<uint4*>(p.g_img + p.g_plane + base + (((p0 + 1) ^ sw) << 4)) = l1;
          }
          const float tot = warp_transpose_reduce16(dbacc, lane);
          if ((lane & 1) == 0) atomicAdd(&s_col[col0 + (lane >> 1)], tot);
        } else if (MODE == MODE_MLP_FWD) {
          // dense layer: h = ELU(acc + bias) (torch: x > 0 ? x : exp(x) - 1), kept as hi/lo planes for the next GEMMs
          uint32_t x[16];
          tmem_ld16(t_col, x);
          tmem_ld_wait();
          float h[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float z = __uint_as_float(x[j]) + s_col[col0 + j];
            h[j] = valid ? (z > 0.f ? z : expf(z) - 1.f) : 0.f;
          }
          store_planes16(p.g_img, p.g_plane, p.R, static_cast<size_t>(b), col0, h);
        } else if (MODE == MODE_MLP_DX) {
          // dz = (dz_above . W) * ELU'(z), ELU'(z) = 1 for z > 0 else exp(z) = h + 1
          uint32_t x[16];
          tmem_ld16(t_col, x);
          float h[16];
          load_planes16(p.h_img, p.h_plane, p.R, static_cast<size_t>(b), col0, h);
          tmem_ld_wait();
          float dz[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) dz[j] = valid ? __uint_as_float(x[j]) * (h[j] > 0.f ? 1.f : h[j] + 1.f) : 0.f;
          store_planes16(p.g_img, p.g_plane, p.R, static_cast<size_t>(b), col0, dz);
          const float tot = warp_transpose_reduce16(dz, lane);
          if ((lane & 1) == 0) atomicAdd(&s_col[col0 + (lane >> 1)], tot);
        } else {   // MODE_DX_ENC
          float gam[16], ga[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { gam[j] = 0.f; ga[j] = 0.f; }
          for (int t = p.T - 1; t >= 0; --t) {
            uint32_t x[16];
            tmem_ld16(t_col + static_cast<uint32_t>(t * p.NCmax), x);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              gam[j] = __uint_as_float(x[j]) + 0.001f * gam[j];
              ga[j] += gam[j];
            }
          }
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(p.g_a + static_cast<size_t>(b) * p.NPout + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(ga[4 * j], ga[4 * j + 1], ga[4 * j + 2], ga[4 * j + 3]);
          }
        }
      }
      gctr += static_cast<uint32_t>(ncw / 16);
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[aset]);
    }
  } else if (TA && warp >= 12) {
    // ===================================================== spike-bit expanders, tensor-memory variant
    // thread = batch row (TMEM lane); per A stage: 64 spike bits -> 32 packed bf16x2 words -> tcgen05.st
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kTaAccCols;
    uint32_t ia = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      for (int i = 0; i < KC * p.T; ++i, ++ia) {
        const uint32_t s = ia % NBM;
        SNN_TWAIT(&full_bits[s], (ia / NBM) & 1, 0);
        const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(bits_ring + s * kBitsStage);
        const uint32_t w0 = wsrc[row], w1 = wsrc[128 + row];
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bits[s]);
        uint32_t r[32];
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const uint4 v = lut[((c8 < 4 ? w0 : w1) >> ((c8 & 3) * 8)) & 0xFFu];
          r[4 * c8] = v.x; r[4 * c8 + 1] = v.y; r[4 * c8 + 2] = v.z; r[4 * c8 + 3] = v.w;
        }
        const uint32_t sa = ia % C::NA;
        SNN_TWAIT(&empty_a[sa], ((ia / C::NA) & 1) ^ 1, 1);
        tc_fence_after_sync();
        if (!(p.dbg & 4)) {
          tmem_st32(t_row + sa * 32, r);
          tmem_st_wait();
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[sa]);
      }
    }
  } else if (MODE == MODE_FWD && warp >= 12) {
    // ===================================================== spike-bit expanders (smem words -> bf16 A tile)
    // Each of the four warps owns every fourth A stage and expands the whole 128-row tile of it (4 rows per
    // lane), so four stages are in flight and the per-stage barrier / proxy-fence cost is paid once per warp.
    const int ew = warp - 12;
    uint32_t ia = 0;
    const int sw = lane & 7;
    uint8_t* const lane_base = a_ring + (lane >> 3) * 1024 + sw * 128;     // row = lane + 32 j -> + j * 4096
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      for (int i = 0; i < KC * p.T; ++i, ++ia) {
        if (static_cast<int>(ia & 3) != ew) continue;
        const uint32_t s = ia % NBM;
        SNN_TWAIT(&full_bits[s], (ia / NBM) & 1, 0);
        const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(bits_ring + s * kBitsStage);
        uint32_t w0[4], w1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { w0[j] = wsrc[lane + 32 * j]; w1[j] = wsrc[128 + lane + 32 * j]; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bits[s]);
        const uint32_t sa = ia % C::NA;
        SNN_TWAIT(&empty_a[sa], ((ia / C::NA) & 1) ^ 1, 1);
        uint8_t* dst = lane_base + sa * C::A_STAGE;
        if (!(p.dbg & 4)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              const uint32_t byte = ((c8 < 4 ? w0[j] : w1[j]) >> ((c8 & 3) * 8)) & 0xFFu;
              *reinterpret_cast<uint4*>(dst + j * 4096 + ((c8 ^ sw) << 4)) = lut[byte];
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[sa]);
      }
    }
  }

  if (p.dbg_out != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 12)) {
    const int role = warp == 0 ? 0 : (warp == 1 ? 1 : (warp == 4 ? 2 : 3));
    unsigned long long* o = p.dbg_out + static_cast<size_t>(blockIdx.x) * 16 + role * 4;
    o[0] = wcyc[0]; o[1] = wcyc[1]; o[2] = wcyc[2];
    o[3] = static_cast<unsigned long long>(clock64() - t_start);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (MODE == MODE_DX || MODE == MODE_MLP_DX) {
    for (int i = tid; i < p.NPout; i += blockDim.x) p.db_part[static_cast<size_t>(blockIdx.x) * p.NPout + i] = s_col[i];
  }
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace snn
