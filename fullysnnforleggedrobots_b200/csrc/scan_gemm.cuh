// scan_gemm.cuh — dense-layer GEMMs of the critic MLP on the tensor cores (tcgen05 / TMEM / bulk-TMA).
// (The actor's spiking layers use the neuron-major kernels of nm_gemm.cuh; this file keeps the batch-major kernel
// the critic was built on: rows of the batch are the MMA's M, a chunk of <= 128 output features its N.)
//
// One work item = (128 batch rows) x (one chunk of <= NCmax output features); the accumulator [128 x NC] fp32 lives
// in TMEM (two sets: the MMA warp fills one while the epilogue drains the other).
//
//   MODE_MLP_FWD  A = hi/lo bf16 planes of the layer input, B = hi/lo planes of W (three products hi*hi + hi*lo +
//                 lo*hi); epilogue h = ELU(acc + bias) written back as hi/lo planes for the next GEMMs.
//   MODE_MLP_DX   A = hi/lo planes of dz of the layer above, B = hi/lo planes of W^T; epilogue dz = acc * ELU'(z)
//                 (from the stored h) as hi/lo planes, plus the bias gradient.
//
// Warp roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer + TMEM owner, warps 4-7 and 8-11 = two epilogue
// groups (TMEM lane quarter = warp % 4; the groups split the 16-column groups of a chunk).
#pragma once
#include "plan.cuh"
#include "umma.cuh"

namespace snn {

enum { MODE_MLP_FWD = 3, MODE_MLP_DX = 4 };

struct ScanGemmParams {
  // A operand: 2 planes of swz64 image [KP/64][R rows]
  const uint8_t* a_img;
  size_t a_plane;           // bytes between A planes
  // B operand: planes of swz64 image [KP/64][NPout rows]
  const uint8_t* b_img;
  size_t b_plane;
  int b_planes;             // 2
  int KP;                   // contraction length (padded, % 64 == 0)
  int NPout;                // output neurons (padded, % 32 == 0)
  int NCmax;                // chunk width, % 32 == 0, <= 128
  int T;                    // always 1 (kept for the launch code's symmetry with the actor kernels)
  int B, Bpad;
  int64_t R;                // rows of the images (= Bpad)
  int nitems, nchunks;
  unsigned long long* dbg_out;   // optional [gridDim.x][16] cycle counters of the barrier waits (tools/role_cycles.py)
  int sets;                 // accumulator sets actually used (<= ScanCfg::SETS); 0 = ScanCfg::SETS
  int dbg;                  // unused
  const float* bias;        // MLP_FWD: [NPout]
  uint8_t* g_img;           // output: 2 planes of swz64 image [NPout/64][R rows]
  size_t g_plane;
  float* db_part;           // MLP_DX: [4 * Bpad / 128][NPout]: one row per (row tile, lane quarter)
  // MLP_FWD: writes ELU(acc + bias) as hi/lo planes into g_img.  MLP_DX: reads the layer's ELU output planes
  const uint8_t* h_img;     // 2 planes of swz64 image [NPout/64][R rows]
  size_t h_plane;
  // Linear ends of a general MLP (RMA env_factor_encoder / adaptation_module; the critic's input gradient): when f32_out
  // is set the epilogue writes plain fp32 rows instead of operand planes and applies no activation —
  //   MLP_FWD: y[b, col] = acc + bias (the linear output layer);  MLP_DX: d_x[b, col] = acc (gradient into the input).
  float* f32_out;           // [B rows][f32_stride floats], columns < f32_cols are written
  long long f32_stride;
  int f32_cols;
};

template <int MODE> struct ScanCfg;
template <> struct ScanCfg<MODE_MLP_FWD> { static constexpr int A_STAGE = 32768, NA = 4, NB = 2, BP = 2, NBITS = 0, SETS = 2, THREADS = 384; };
template <> struct ScanCfg<MODE_MLP_DX> { static constexpr int A_STAGE = 32768, NA = 4, NB = 2, BP = 2, NBITS = 0, SETS = 2, THREADS = 384; };

constexpr int kScanMaxNP = 2048;   // floats of per-column smem scratch (MLP_FWD: bias, MLP_DX: db); layer widths <= 2048

// 16 consecutive columns of one row -> hi/lo bf16 planes of a swz64 image (two 16-byte pieces per plane)
__device__ __forceinline__ void store_planes16(uint8_t* img, size_t plane, int64_t R, size_t r, int col0, const float (&x)[16]) {
  // packed conversions: hi = bf16(x) for two columns per cvt, lo = bf16(x - hi) (the same values as one cvt per element)
  uint32_t hp[8], lp[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hp[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
    lp[j] = pack_bf16x2(x[2 * j] - __uint_as_float(hp[j] << 16), x[2 * j + 1] - __uint_as_float(hp[j] & 0xFFFF0000u));
  }
  const size_t base = static_cast<size_t>(col0 >> 6) * R * 128 + (r >> 3) * 1024 + (r & 7) * 128;
  const int p0 = (col0 & 63) >> 3;
  const int sw = static_cast<int>(r & 7);
  const uint4 h0 = make_uint4(hp[0], hp[1], hp[2], hp[3]), h1 = make_uint4(hp[4], hp[5], hp[6], hp[7]);
  const uint4 l0 = make_uint4(lp[0], lp[1], lp[2], lp[3]), l1 = make_uint4(lp[4], lp[5], lp[6], lp[7]);
  *reinterpret_cast<uint4*>(img + base + ((p0 ^ sw) << 4)) = h0;
  *reinterpret_cast<uint4*>(img + base + (((p0 + 1) ^ sw) << 4)) = h1;
  *reinterpret_cast<uint4*>(img + plane + base + ((p0 ^ sw) << 4)) = l0;
  *reinterpret_cast<uint4*>(img + plane + base + (((p0 + 1) ^ sw) << 4)) = l1;
}
// inverse, in two halves so that the loads can be issued well ahead of their use: the four 16-byte pieces (hi, lo) of 16
// consecutive columns of one row ...
struct Planes16 { uint4 h0, h1, l0, l1; };
__device__ __forceinline__ Planes16 load_planes16_raw(const uint8_t* img, size_t plane, int64_t R, size_t r, int col0) {
  const size_t base = static_cast<size_t>(col0 >> 6) * R * 128 + (r >> 3) * 1024 + (r & 7) * 128;
  const int p0 = (col0 & 63) >> 3;
  const int sw = static_cast<int>(r & 7);
  Planes16 q;
  q.h0 = __ldg(reinterpret_cast<const uint4*>(img + base + ((p0 ^ sw) << 4)));
  q.h1 = __ldg(reinterpret_cast<const uint4*>(img + base + (((p0 + 1) ^ sw) << 4)));
  q.l0 = __ldg(reinterpret_cast<const uint4*>(img + plane + base + ((p0 ^ sw) << 4)));
  q.l1 = __ldg(reinterpret_cast<const uint4*>(img + plane + base + (((p0 + 1) ^ sw) << 4)));
  return q;
}
// ... and hi + lo as fp32
__device__ __forceinline__ void planes16_to_f32(const Planes16& q, float (&x)[16]) {
  const uint32_t hw[8] = {q.h0.x, q.h0.y, q.h0.z, q.h0.w, q.h1.x, q.h1.y, q.h1.z, q.h1.w};
  const uint32_t lw[8] = {q.l0.x, q.l0.y, q.l0.z, q.l0.w, q.l1.x, q.l1.y, q.l1.z, q.l1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    x[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
    x[2 * j + 1] = __uint_as_float(hw[j] & 0xFFFF0000u) + __uint_as_float(lw[j] & 0xFFFF0000u);
  }
}

template <int MODE>
__host__ __device__ constexpr size_t scan_gemm_smem_bytes(int NCmax) {
  return 1024 /*align slack*/ + static_cast<size_t>(ScanCfg<MODE>::NA) * ScanCfg<MODE>::A_STAGE +
         static_cast<size_t>(ScanCfg<MODE>::NB) * ScanCfg<MODE>::BP * NCmax * 128 + kScanMaxNP * 4 + 512;
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

template <int MODE>
__global__ void __launch_bounds__(ScanCfg<MODE>::THREADS, 1) scan_gemm_kernel(const ScanGemmParams p) {
  using C = ScanCfg<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage = C::BP * p.NCmax * 128;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = a_ring + C::NA * C::A_STAGE;
  float* s_col = reinterpret_cast<float*>(b_ring + C::NB * b_stage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_col + kScanMaxNP);
  uint64_t* full_a = bars;                 // [NA]
  uint64_t* empty_a = full_a + C::NA;      // [NA]
  uint64_t* full_b = empty_a + C::NA;      // [NB]
  uint64_t* empty_b = full_b + C::NB;      // [NB]
  uint64_t* acc_full = empty_b + C::NB;    // [SETS]: two accumulator sets (NCmax <= 256 columns each)
  uint64_t* acc_empty = acc_full + C::SETS;  // [SETS]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + C::SETS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = p.KP / 64;
  const uint32_t nsets = (C::SETS == 2 && p.sets != 1) ? 2u : 1u;
#ifdef SNN_ABLATE
  unsigned long long wcyc[3] = {0, 0, 0};
  const long long t_start = clock64();
#endif

  if (tid == 0) {
    for (int s = 0; s < C::NA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < C::NB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < C::SETS; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 256); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  if (MODE == MODE_MLP_FWD) {
    for (int i = tid; i < p.NPout; i += blockDim.x) s_col[i] = p.bias[i];
  } else {
    for (int i = tid; i < p.NPout; i += blockDim.x) s_col[i] = 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================================================== producer (warp-uniform loop, one elected lane issues)
    uint32_t ib = 0, ia = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int m = item / p.nchunks, c = item - m * p.nchunks;
      const int n0 = c * p.NCmax;
      const int ncw = min(p.NCmax, p.NPout - n0);
      for (int kc = 0; kc < KC; ++kc, ++ib, ++ia) {
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
        const uint32_t sa = ia % C::NA;
        SNN_TWAIT(&empty_a[sa], ((ia / C::NA) & 1) ^ 1, 1);
        if (elect_one()) {
          mbar_expect_tx(&full_a[sa], 32768u);
          for (int pl = 0; pl < 2; ++pl)
            bulk_g2s(a_ring + sa * C::A_STAGE + pl * 16384,
                     p.a_img + pl * p.a_plane + (static_cast<size_t>(kc) * p.R + static_cast<size_t>(m) * 128) * 128, 16384u,
                     &full_a[sa]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp-uniform loop, one elected lane issues)
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
      for (int kc = 0; kc < KC; ++kc, ++ib, ++ia) {
        const uint32_t sb = ib % C::NB;
        SNN_TWAIT(&full_b[sb], (ib / C::NB) & 1, 1);
        const uint32_t b_lo = b_ring_lo + sb * (static_cast<uint32_t>(b_stage) >> 4);
        const uint32_t sa = ia % C::NA;
        SNN_TWAIT(&full_a[sa], (ia / C::NA) & 1, 2);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_lo = a_ring_lo + sa * (C::A_STAGE >> 4);
          const uint32_t d = tmem_base + aset * 256;
          uint32_t acc = kc > 0 ? 1u : 0u;
          // x ~ (a_hi + a_lo), W ~ (b_hi + b_lo): hi*hi + hi*lo + lo*hi (lo*lo < 2^-16 relative)
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {
            const uint32_t pa = pr == 2 ? 1 : 0, pb = pr == 1 ? 1 : 0;
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              umma_bf16_lohi(d, a_lo + pa * (16384 >> 4) + k16 * 2, dhi, b_lo + pb * b_plane_lo + k16 * 2, dhi, idesc, acc);
              acc = 1u;
            }
          }
          umma_commit(&empty_a[sa]);
          umma_commit(&empty_b[sb]);
          if (kc == KC - 1) umma_commit(&acc_full[aset]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================================== epilogue (two groups of 4 warps)
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
      const uint32_t aset = it % nsets;
      const int ng = ncw / 16;
      const int g_first = static_cast<int>((half + gctr) & 1u);
      // MLP_DX: the layer's stored ELU outputs do not depend on the accumulator: the first group's planes are requested
      // before the wait for the MMAs, every further group's while the group before it is processed (the loads were the
      // epilogue's largest stall: thread = batch row, every 16-byte piece in a different line)
      Planes16 hq{};
      const bool dx_planes = MODE == MODE_MLP_DX && p.f32_out == nullptr;
      if (dx_planes && g_first < ng) hq = load_planes16_raw(p.h_img, p.h_plane, p.R, static_cast<size_t>(b), n0 + g_first * 16);
      SNN_TWAIT(&acc_full[aset], (it / nsets) & 1, 0);
      tc_fence_after_sync();
      for (int g = g_first; g < ng; g += 2) {
        const int col0 = n0 + g * 16;
        const uint32_t t_col = t_lane + aset * 256 + static_cast<uint32_t>(g * 16);
        if (MODE == MODE_MLP_FWD) {
          // dense layer: h = ELU(acc + bias) (torch: x > 0 ? x : exp(x) - 1), kept as hi/lo planes for the next GEMMs
          uint32_t x[16];
          tmem_ld16(t_col, x);
          tmem_ld_wait();
          float h[16];
          if (p.f32_out != nullptr) {
            if (valid) {
              float* dst = p.f32_out + static_cast<size_t>(b) * p.f32_stride;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j < p.f32_cols) dst[col0 + j] = __uint_as_float(x[j]) + s_col[col0 + j];
            }
            continue;
          }
          // branch-free ELU (the per-element branch around expf compiled to sixteen divergent regions per group):
          // exp of min(z, 0) is always evaluated, the select keeps z where it is positive — same values as torch's
          // x > 0 ? x : exp(x) - 1
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float z = __uint_as_float(x[j]) + s_col[col0 + j];
            const float e = expf(fminf(z, 0.f)) - 1.f;
            h[j] = valid ? (z > 0.f ? z : e) : 0.f;
          }
          store_planes16(p.g_img, p.g_plane, p.R, static_cast<size_t>(b), col0, h);
        } else {
          if (p.f32_out != nullptr) {
            uint32_t x[16];
            tmem_ld16(t_col, x);
            tmem_ld_wait();
            if (valid) {
              float* dst = p.f32_out + static_cast<size_t>(b) * p.f32_stride;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j < p.f32_cols) dst[col0 + j] = __uint_as_float(x[j]);
            }
            continue;
          }
          // dz = (dz_above . W) * ELU'(z), ELU'(z) = 1 for z > 0 else exp(z) = h + 1
          uint32_t x[16];
          tmem_ld16(t_col, x);
          float h[16];
          planes16_to_f32(hq, h);
          if (g + 2 < ng) hq = load_planes16_raw(p.h_img, p.h_plane, p.R, static_cast<size_t>(b), col0 + 32);
          tmem_ld_wait();
          float dz[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) dz[j] = valid ? __uint_as_float(x[j]) * (h[j] > 0.f ? 1.f : h[j] + 1.f) : 0.f;
          store_planes16(p.g_img, p.g_plane, p.R, static_cast<size_t>(b), col0, dz);
          // bias gradient: the column sums of this warp's 32 rows go to their own partial row (row tile, lane quarter) —
          // every (row, column) of db_part is written exactly once and the deferred row reduction sums them in a fixed
          // order (r2: shared-memory float atomics before, the one source of run-to-run differences of an update)
          const float tot = warp_transpose_reduce16(dz, lane);
          if ((lane & 1) == 0) p.db_part[(static_cast<size_t>(m) * 4 + q) * p.NPout + col0 + (lane >> 1)] = tot;
        }
      }
      gctr += static_cast<uint32_t>(ncw / 16);
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[aset]);
    }
  }

#ifdef SNN_ABLATE
  if (p.dbg_out != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 4)) {
    const int role = warp == 0 ? 0 : (warp == 1 ? 1 : 2);
    unsigned long long* o = p.dbg_out + static_cast<size_t>(blockIdx.x) * 16 + role * 4;
    o[0] = wcyc[0]; o[1] = wcyc[1]; o[2] = wcyc[2];
    o[3] = static_cast<unsigned long long>(clock64() - t_start);
  }
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace snn
