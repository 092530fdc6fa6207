// nm_gemm.cuh — the actor's "neuron-major" scan-GEMMs (tcgen05 / TMEM / bulk-TMA).
//
// One work item = (128 output neurons) x (one batch tile of Bt samples, all T spike steps).  The tensor core
// computes D[neuron][sample-step] : the 128 neurons are the MMA's M (= the TMEM lanes), the tile's T * Bt
// sample-steps are its N (= the TMEM columns; 240 for T = 5), so ONE tcgen05.mma per 16 contraction elements and
// weight plane covers all T timesteps.  tools/mma_rate.cu measures an M = 128 MMA at max(N/2, 57 + N/8) cycles:
// N = 240 runs at the tensor core's full rate, while "one N <= 96 accumulator per timestep" (the first design of
// this path) was capped at 70 % (N = 96) / 39 % (N = 48) by the fixed issue cost.
// After the GEMM the epilogue warps run the sequential-in-time recurrence with thread = neuron, registers = the
// 16 samples x T steps it is working on:
//
//   NM_FWD     A = bf16 hi/mid/lo planes of W (K-major), B = spikes of the layer below: bit-packed in HBM, staged
//              by bulk copies into a shared-memory ring and expanded there to bf16 0/1 operand tiles, MN-major
//              ([input neuron][sample-step], a warp-wide 32 x 32 bit transpose per block of spike words), written
//              straight into the operand slot by the expander warps.
//              Epilogue = LIF update (AMP/.../modules/actor_critic.py:216-232): c = .5c + (Wx+b);
//              v = .75 v (1-s) + c; s = v > .5.  Spike words come out of a warp bit transpose (lane = neuron); emits the
//              16-bit voltage-code trace (umma.cuh: volt_code2) when training.
//   NM_DX      A = hi/lo planes of W^T (K-major), B = hi/lo planes of g_c of the layer above (swz64 image with rows =
//              sample-steps: K-major, one bulk copy per plane); three products hi*hi + lo*hi + hi*lo.  Epilogue = BPTT recurrence of
//              SURVEY.md §8a for the layer below (surrogate gradient of PseudoSpikeRect, actor_critic.py:154-167).
//              Emits g_c planes (swz64 image) and the bias gradient (a per-thread sum: no shuffles, no atomics).
//   NM_DX_ENC  same GEMM into the encoder neurons; epilogue = straight-through encoder recurrence
//              (actor_critic.py:50-59,116-119): Gamma_t = g_t + 0.001 Gamma_{t+1}; emits g_a = sum_t Gamma_t.
//
// Accumulators: TB = T * Bt <= 256 columns -> two sets (set = item parity): the MMA warp fills one while an
// epilogue group (4 warps = 128 TMEM lanes) drains the other.  TB > 256 (T > 16): one set, the two epilogue
// groups split its 16-sample groups.
// A CTA keeps ONE neuron tile (mt = blockIdx.x % nmt) and takes batch tiles from a scheduler: its first tile is
// static, every further one comes from a global counter (one atomicAdd per tile by warp 2, drawn one tile ahead and
// handed to the other warps through a four-entry shared-memory ring).  CTAs that start late — the critic's
// kernels hold some SMs when an actor kernel launches inside the PPO step — or run on a slower SM simply take fewer
// tiles.  Per-neuron constants (bias) live in registers for the whole kernel; per-neuron sums (bias / encoder
// gradients) are written per TILE, so their summation order does not depend on which CTA ran which tile (bitwise
// run-to-run reproducibility, tests/test_gpu_properties.py).
//
// Warp roles: 0 (and 3 in FWD) = weight producers, 1 = MMA issuer + TMEM owner, 2 = B-operand producer (FWD: spike
// words, dX: g_c units) + tile scheduler, 4-7 and 8-11 = epilogue groups, 12-15 = spike-bit expanders (FWD) / third
// epilogue group (dX).
#pragma once
#include "plan.cuh"
#include "umma.cuh"

namespace snn {

enum { NM_FWD = 0, NM_DX = 1, NM_DX_ENC = 2 };
constexpr int kSchedInts = 128;     // ints per scheduler slot: one ticket counter per neuron tile (K0P <= 8192: 64 tiles), [127] = CTAs done
constexpr int kNSch = 4;            // entries of the shared-memory tile ring

struct NmParams {
  // batch-tile geometry (SnnPlan)
  int T, Bt, TB, CT, sets, ntile, B;
  int64_t Rt;
  int KP;                   // contraction length (padded, % 64 == 0)
  int MP;                   // output neurons (padded, % 128 == 0) = TMEM lanes over all neuron tiles
  int nmt;                  // MP / 128; gridDim.x % nmt == 0
  // A operand: planes of a swz64 image [KP/64][MP rows][128 B]   (FWD: W hi/mid/lo, DX: W^T hi/lo)
  const uint8_t* a_img;
  size_t a_plane;
  int a_planes;             // FWD: 1..3
  // B operand
  const uint32_t* b_bits;   // FWD: spike words [KP/32][Rt]
  const uint8_t* b_img;     // DX: 2 planes of the swz64 image [KP/64][Rt rows][128 B] of g_c (rows = sample-steps)
  size_t b_plane;
  // FWD epilogue
  const float* bias;        // [MP]
  uint32_t* out_bits;       // [MP/32][Rt]
  uint16_t* v_out;          // voltage codes [MP/32][Rt/8][32 neurons][8 sample-steps] or null (inference)
  // DX epilogue
  const uint16_t* v_in;     // [MP/32][Rt/8][32][8] voltage codes of the layer whose gradient is produced
  uint8_t* g_img;           // 2 planes of the swz64 image [MP/64][Rt rows][128 B]
  size_t g_plane;
  float* db_part;           // [3 * ntile][MP]   (one row per batch tile and epilogue group: a fixed summation order
                            // whatever CTA the scheduler hands a tile to)
  // DX_ENC epilogue: g_a = d loss / d pop_act (written only when the caller needs d obs) and, fused, the encoder
  // parameter gradients (actor_critic.py:105 under autograd):  a = exp(-1/2 d^2 / var), d = obs - mean,
  //   d mean = sum_b g_a a d / var ;  d std = sum_b g_a a d^2 std / var^2     (per-thread sums, one row per group)
  float* g_a;               // [Bcap/16][MP][16] or null
  const float* enc_obs;     // [B][enc_O]
  const float* enc_mean;    // [enc_K0]
  const float* enc_std;     // [enc_K0]
  int enc_O, enc_P, enc_K0;
  float enc_eps;
  float* enc_part;          // [3 * ntile][2][MP]
  int s_stage;              // FWD: bytes of one spike sub-tile (operand slot): 64-column blocks x 8 KiB
  int* sched;               // tile scheduler: [mt] tiles handed out per neuron tile, [kSchedInts - 1] CTAs that ran out of
                            // work; all zero at launch, re-armed by the last CTA (one slot per launch, api.cu)
  int rev;                  // walk the batch tiles last to first: consecutive kernels of the backward chain alternate, so
                            // a kernel starts on the tiles its predecessor wrote last (still in L2)
  int dbg;                  // SNN_DBG ablation mask (timing experiments only): 1 = no MMA, 2 = no scan, 4 = no expand,
                            // 16 = dX: no g_c stores, 32 = dX: no voltage loads, 128 = FWD: no voltage stores, 256 = FWD: no spike-word stores
  unsigned long long* dbg_out;   // optional [gridDim.x][16] cycle counters of the barrier waits (tools/role_cycles.py)
};

template <int MODE> struct NmCfg;
// FWD: a ring of 7 weight-plane units of 16 KiB (hi/mid/lo of a 64-wide chunk, released plane by plane so the
// reloads start early), 3 operand slots of one spike sub-tile each (64 input neurons x <= 256 sample-steps, MN-major:
// [64-column block][64 rows][128 B]; the size is a launch parameter), 2 spike-word stages
#ifndef SNN_FWD_NW
#define SNN_FWD_NW 7
#define SNN_FWD_NS 3
#endif
template <> struct NmCfg<NM_FWD> {
  static constexpr int NW = SNN_FWD_NW, W_STAGE = 16384, NS = SNN_FWD_NS, S_STAGE = 0, NBITS = 2, BITS_STAGE = 2048, THREADS = 512;
};
// DX: 2 weight stages of 2 x 16 KiB planes, 4 g_c units (one plane of one 64-neuron chunk of one sub-tile:
// <= 256 sample-steps x 128 B)
#ifndef SNN_DX_NS
#define SNN_DX_NS 4
#endif
template <> struct NmCfg<NM_DX> {
  static constexpr int NW = 2, W_STAGE = 32768, NS = SNN_DX_NS, S_STAGE = 32768, NBITS = 0, BITS_STAGE = 0, THREADS = 512;
};
template <> struct NmCfg<NM_DX_ENC> : NmCfg<NM_DX> {};

// the two bf16 halves of v to [addr] and [addr + 1024] (the same column of two rows 8 apart in a swz64 image)
__device__ __forceinline__ void st_global_u16x2_rows(unsigned long long addr, uint32_t v) {
  asm volatile(
      "{\n\t.reg .b16 lo, hi;\n\t"
      "mov.b32 {lo, hi}, %1;\n\t"
      "st.global.u16 [%0], lo;\n\t"
      "st.global.u16 [%0+1024], hi;\n\t}" ::"l"(addr), "r"(v)
      : "memory");
}

template <int MODE>
__host__ __device__ constexpr size_t nm_gemm_smem_bytes() {      // dX kernels
  return 1024 /*align slack*/ + static_cast<size_t>(NmCfg<MODE>::NW) * NmCfg<MODE>::W_STAGE +
         static_cast<size_t>(NmCfg<MODE>::NS) * NmCfg<MODE>::S_STAGE + 512;
}
constexpr size_t kSmemOptIn = 232448;    // 227 KiB per CTA
// forward kernel: bytes of one spike sub-tile (operand slot) for a batch-tile geometry
inline int nm_fwd_s_stage(int TB) { return ((TB < 256 ? TB : 256) + 63) / 64 * 8192; }
inline size_t nm_fwd_smem_bytes(int s_stage) {
  return 1024 + static_cast<size_t>(NmCfg<NM_FWD>::NW) * NmCfg<NM_FWD>::W_STAGE +
         static_cast<size_t>(NmCfg<NM_FWD>::NS) * s_stage +
         static_cast<size_t>(NmCfg<NM_FWD>::NBITS) * NmCfg<NM_FWD>::BITS_STAGE + 512;
}

// CTAs to launch: a multiple of the neuron tiles, at most one per SM, at most one per item
inline int nm_gemm_grid(int nmt, int ntile, int num_sms) {
  int per = num_sms / nmt;
  if (per < 1) per = 1;
  if (per > ntile) per = ntile;
  return per * nmt;
}

template <int MODE>
__global__ void __launch_bounds__(NmCfg<MODE>::THREADS, 1) nm_gemm_kernel(const NmParams p) {
  using C = NmCfg<MODE>;
  constexpr bool IS_FWD = MODE == NM_FWD;
  constexpr uint32_t NBM = C::NBITS > 0 ? C::NBITS : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S_STAGE = IS_FWD ? p.s_stage : C::S_STAGE;
  uint8_t* w_ring = smem;
  uint8_t* s_ring = w_ring + C::NW * C::W_STAGE;
  uint8_t* bits_ring = s_ring + C::NS * S_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bits_ring + C::NBITS * C::BITS_STAGE);
  uint64_t* w_full = bars;                      // [<= 8]
  uint64_t* w_empty = w_full + 8;               // [<= 8]
  uint64_t* s_full = w_empty + 8;               // [<= 8]
  uint64_t* s_empty = s_full + 8;               // [<= 8]
  uint64_t* acc_full = s_empty + 8;             // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint64_t* bits_full = acc_empty + 2;          // [NBITS]
  uint64_t* bits_empty = bits_full + C::NBITS;  // [NBITS]
  uint64_t* sch_full = bits_empty + C::NBITS;   // [kNSch]  tile ring entry published by the scheduler warp
  uint64_t* sch_empty = sch_full + kNSch;       // [kNSch]  ... read by every consumer warp
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sch_empty + kNSch);
  volatile int* s_tiles = reinterpret_cast<volatile int*>(s_tmem + 1);      // [kNSch]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = p.KP / 64;
  const uint32_t nsets = p.sets == 2 ? 2u : 1u;
  // epilogue groups of 4 warps: FWD has two (warps 12-15 expand spikes) and, with two accumulator sets, binds group g
  // to set g; the dX kernels have three, every group works on every item and the 16-sample groups go round-robin
  constexpr int NGRP = IS_FWD ? 2 : 3;
  const bool by_set = IS_FWD && p.sets == 2;
  const int nsub = (p.TB + 255) / 256;                 // sample-step sub-tiles of <= 256 columns (2 only when T > 16)
  const int mt = blockIdx.x % p.nmt;                   // this CTA's neuron tile
  const int tile0 = blockIdx.x / p.nmt, tstep = gridDim.x / p.nmt;
#ifdef SNN_ABLATE
  unsigned long long wcyc[3] = {0, 0, 0};
  const long long t_start = clock64();
  const unsigned long long g_start = globaltimer_ns();
#endif

  if (tid == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    // FWD: an operand slot is full once the four expander warps have written it
    for (int s = 0; s < 8; ++s) { mbar_init(&s_full[s], IS_FWD ? 4 : 1); mbar_init(&s_empty[s], 1); }
    for (int s = 0; s < C::NBITS; ++s) { mbar_init(&bits_full[s], 1); mbar_init(&bits_empty[s], 4); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], by_set ? 128 : NGRP * 128); }
    // consumers of the tile ring: every warp but the scheduling warp 2 (and, in the dX kernels, the idle warp 3)
    for (int s = 0; s < kNSch; ++s) { mbar_init(&sch_full[s], 1); mbar_init(&sch_empty[s], IS_FWD ? 15 : 14); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  // entry k of the tile ring: this CTA's k-th batch tile, or -1 when the batch is exhausted (every consumer warp calls
  // this once per entry)
  auto next_tile = [&](uint32_t& k) -> int {
    const uint32_t sl = k % kNSch;
    mbar_wait(&sch_full[sl], (k / kNSch) & 1);
    const int t = s_tiles[sl];
    __syncwarp();
    if (lane == 0) mbar_arrive(&sch_empty[sl]);
    ++k;
    return t;
  };

  if (warp == 2) {
    // ===================================================== B-operand producer + tile scheduler
    // FWD: the spike words of the layer below, per (chunk, sub-tile) two runs of words; dX: the g_c units of the layer
    // above.  This warp is the first to need a tile's index, so it also hands the tiles out: the first one is static, the
    // ticket of the next one is drawn (one atomicAdd) BEFORE this tile's copies are issued and published to the other
    // warps after them, so the round trip to the counter never shows.
    uint32_t u = 0, k = 0;
    int tile = p.rev ? p.ntile - 1 - tile0 : tile0;
    for (;;) {
      {   // publish entry k
        const uint32_t sl = k % kNSch;
        mbar_wait(&sch_empty[sl], ((k / kNSch) & 1) ^ 1);
        if (lane == 0) {
          s_tiles[sl] = tile;
          mbar_arrive(&sch_full[sl]);      // release: the entry is visible to whoever sees the phase complete
        }
        __syncwarp();
        ++k;
      }
      if (tile < 0) break;
      int ticket = 0;
      if (lane == 0) ticket = tstep + atomicAdd(p.sched + mt, 1);
      for (int kc = 0; kc < KC; ++kc) {
        if (IS_FWD) {
          for (int sub = 0; sub < nsub; ++sub, ++u) {
            const uint32_t s = u % NBM;
            SNN_TWAIT(&bits_empty[s], ((u / NBM) & 1) ^ 1, 0);
            if (elect_one()) {
              const int ncols = min(256, p.TB - sub * 256);
              mbar_expect_tx(&bits_full[s], static_cast<uint32_t>(2 * ncols * 4));
              const size_t c0 = static_cast<size_t>(tile) * p.CT + static_cast<size_t>(sub) * 256;
              uint8_t* dst = bits_ring + s * C::BITS_STAGE;
              bulk_g2s(dst, p.b_bits + static_cast<size_t>(2 * kc) * p.Rt + c0, static_cast<uint32_t>(ncols * 4), &bits_full[s]);
              bulk_g2s(dst + 1024, p.b_bits + static_cast<size_t>(2 * kc + 1) * p.Rt + c0, static_cast<uint32_t>(ncols * 4),
                       &bits_full[s]);
            }
            __syncwarp();
          }
        } else {
          // g_c units of this 64-neuron chunk: (plane) x (sub-tile), each one contiguous run of rows of the image
          for (int pl = 0; pl < 2; ++pl)
            for (int sub = 0; sub < nsub; ++sub, ++u) {
              const uint32_t gs = u % C::NS;
              SNN_TWAIT(&s_empty[gs], ((u / C::NS) & 1) ^ 1, 1);
              if (elect_one()) {
                const uint32_t bytes = static_cast<uint32_t>(min(256, p.TB - sub * 256) * 128);
                mbar_expect_tx(&s_full[gs], bytes);
                bulk_g2s(s_ring + gs * S_STAGE,
                         p.b_img + pl * p.b_plane +
                             (static_cast<size_t>(kc) * p.Rt + static_cast<size_t>(tile) * p.CT + static_cast<size_t>(sub) * 256) * 128,
                         bytes, &s_full[gs]);
              }
              __syncwarp();
            }
        }
      }
      ticket = __shfl_sync(0xffffffffu, ticket, 0);
      tile = ticket >= p.ntile ? -1 : (p.rev ? p.ntile - 1 - ticket : ticket);
    }
    // every CTA has drawn its last ticket before it counts itself done: the last one re-arms the slot for the next launch
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(p.sched + kSchedInts - 1, 1) == static_cast<int>(gridDim.x) - 1) {
        for (int i = 0; i < p.nmt; ++i) p.sched[i] = 0;
        p.sched[kSchedInts - 1] = 0;
        __threadfence();
      }
    }
  } else if (warp == 0 || (IS_FWD && warp == 3)) {
    // ===================================================== weight producer(s) (warp-uniform loop, one elected lane issues)
    // One thread issues a bulk copy every ~250 cycles at best (tools/l2_bulk_rate.cu: 65 B/cycle/SM with 16 KiB copies
    // from one warp whatever the ring depth, 77 - 110 from three), which starved the forward kernel's three 16 KiB plane
    // units per 64-wide chunk: there warps 0 and 3 take alternate units.
    const uint32_t wsel = warp == 0 ? 0u : 1u;
    uint32_t iw = 0, sk = 0;
    for (;;) {
      if (next_tile(sk) < 0) break;
      for (int kc = 0; kc < KC; ++kc) {
        if (IS_FWD) {
          // one ring unit per weight plane
          for (int pl = 0; pl < p.a_planes; ++pl, ++iw) {
            if ((iw & 1u) != wsel) continue;
            const uint32_t ws = iw % C::NW;
            SNN_TWAIT(&w_empty[ws], ((iw / C::NW) & 1) ^ 1, 0);
            if (elect_one()) {
              mbar_expect_tx(&w_full[ws], 16384u);
              bulk_g2s(w_ring + ws * C::W_STAGE,
                       p.a_img + pl * p.a_plane + (static_cast<size_t>(kc) * p.MP + static_cast<size_t>(mt) * 128) * 128, 16384u,
                       &w_full[ws]);
            }
            __syncwarp();
          }
        } else {
          const uint32_t ws = iw % C::NW;
          SNN_TWAIT(&w_empty[ws], ((iw / C::NW) & 1) ^ 1, 0);
          if (elect_one()) {
            mbar_expect_tx(&w_full[ws], 32768u);
            for (int pl = 0; pl < 2; ++pl)
              bulk_g2s(w_ring + ws * C::W_STAGE + pl * 16384,
                       p.a_img + pl * p.a_plane + (static_cast<size_t>(kc) * p.MP + static_cast<size_t>(mt) * 128) * 128, 16384u,
                       &w_full[ws]);
          }
          __syncwarp();
          ++iw;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp-uniform loop, one elected lane issues)
    uint32_t iw = 0, u = 0, it = 0, sk = 0;
    const uint32_t dhi = smem_desc_hi(1024);
    const uint32_t w_ring_lo = smem_desc_lo(smem_u32(w_ring), 16);
    // FWD: MN-major spike tile, leading-dimension byte offset = one 64-column block; dX: K-major g_c tile
    const uint32_t s_ring_lo = smem_desc_lo(smem_u32(s_ring), IS_FWD ? 8192 : 16);
    for (;; ++it) {
      if (next_tile(sk) < 0) break;
      const uint32_t aset = it % nsets;
      SNN_TWAIT(&acc_empty[aset], ((it / nsets) & 1) ^ 1, 0);
      tc_fence_after_sync();
      for (int kc = 0; kc < KC; ++kc) {
        if (IS_FWD) {
          const int npl = p.a_planes;
          if (nsub == 1) {
            // plane-major: each weight plane unit is released as soon as its four MMAs are issued
            const uint32_t ss = u % C::NS;
            SNN_TWAIT(&s_full[ss], (u / C::NS) & 1, 2);
            const uint32_t b_lo = s_ring_lo + ss * (S_STAGE >> 4);
            const uint32_t d = tmem_base + aset * 256;
            const uint32_t idesc = make_idesc_bf16(128, static_cast<uint32_t>(p.TB), 0, 1);
            for (int pl = 0; pl < npl; ++pl, ++iw) {
              const uint32_t ws = iw % C::NW;
              SNN_TWAIT(&w_full[ws], (iw / C::NW) & 1, 1);
              tc_fence_after_sync();
              if (elect_one()) {
                const uint32_t a_lo = w_ring_lo + ws * (C::W_STAGE >> 4);
                if (!(SNN_DBG_MASK(p) & 1)) {
#pragma unroll
                  for (int k16 = 0; k16 < 4; ++k16)
                    umma_bf16_lohi(d, a_lo + k16 * 2, dhi, b_lo + k16 * (2048 >> 4), dhi, idesc, (kc > 0 || pl > 0 || k16 > 0) ? 1u : 0u);
                }
                umma_commit(&w_empty[ws]);
                if (pl == npl - 1) {
                  umma_commit(&s_empty[ss]);
                  if (kc == KC - 1) umma_commit(&acc_full[aset]);
                }
              }
              __syncwarp();
            }
            ++u;
          } else {
            // two sub-tiles share the chunk's weight planes: wait for all of them, release them together
            const uint32_t iw0 = iw;
            for (int pl = 0; pl < npl; ++pl) SNN_TWAIT(&w_full[(iw0 + pl) % C::NW], ((iw0 + pl) / C::NW) & 1, 1);
            for (int sub = 0; sub < nsub; ++sub, ++u) {
              const uint32_t ss = u % C::NS;
              SNN_TWAIT(&s_full[ss], (u / C::NS) & 1, 2);
              tc_fence_after_sync();
              if (elect_one()) {
                const int ncols = min(256, p.TB - sub * 256);
                const uint32_t b_lo = s_ring_lo + ss * (S_STAGE >> 4);
                const uint32_t d = tmem_base + aset * 256 + static_cast<uint32_t>(sub * 256);
                const uint32_t idesc = make_idesc_bf16(128, static_cast<uint32_t>(ncols), 0, 1);
                uint32_t acc = kc > 0 ? 1u : 0u;
                if (!(SNN_DBG_MASK(p) & 1)) {
#pragma unroll
                  for (int k16 = 0; k16 < 4; ++k16)
                    for (int pl = 0; pl < npl; ++pl) {
                      umma_bf16_lohi(d, w_ring_lo + ((iw0 + pl) % C::NW) * (C::W_STAGE >> 4) + k16 * 2, dhi, b_lo + k16 * (2048 >> 4), dhi, idesc, acc);
                      acc = 1u;
                    }
                }
                umma_commit(&s_empty[ss]);
                if (sub == nsub - 1) {
                  for (int pl = 0; pl < npl; ++pl) umma_commit(&w_empty[(iw0 + pl) % C::NW]);
                  if (kc == KC - 1) umma_commit(&acc_full[aset]);
                }
              }
              __syncwarp();
            }
            iw += npl;
          }
        } else {
          const uint32_t ws = iw % C::NW;
          SNN_TWAIT(&w_full[ws], (iw / C::NW) & 1, 1);
          const uint32_t a_lo = w_ring_lo + ws * (C::W_STAGE >> 4);
          for (int pl = 0; pl < 2; ++pl)          // g_c plane of the unit
            for (int sub = 0; sub < nsub; ++sub, ++u) {
              const uint32_t ss = u % C::NS;
              SNN_TWAIT(&s_full[ss], (u / C::NS) & 1, 2);
              tc_fence_after_sync();
              if (elect_one()) {
                const int ncols = min(256, p.TB - sub * 256);
                const uint32_t b_lo = s_ring_lo + ss * (S_STAGE >> 4);
                const uint32_t d = tmem_base + aset * 256 + static_cast<uint32_t>(sub * 256);
                if (!(SNN_DBG_MASK(p) & 1)) {
                  // W^T ~ (a_hi + a_lo), g_c ~ (b_hi + b_lo): hi*hi + lo*hi on the g_c hi plane, then hi*lo on its lo plane
                  // (lo*lo < 2^-16 relative)
                  const uint32_t idesc = make_idesc_bf16(128, static_cast<uint32_t>(ncols), 0, 0);
                  uint32_t acc = (kc > 0 || pl > 0) ? 1u : 0u;
#pragma unroll
                  for (int k16 = 0; k16 < 4; ++k16) {
                    umma_bf16_lohi(d, a_lo + k16 * 2, dhi, b_lo + k16 * 2, dhi, idesc, acc);
                    acc = 1u;
                    if (pl == 0) umma_bf16_lohi(d, a_lo + (16384 >> 4) + k16 * 2, dhi, b_lo + k16 * 2, dhi, idesc, 1u);
                  }
                }
                umma_commit(&s_empty[ss]);
                if (pl == 1 && sub == nsub - 1) {
                  umma_commit(&w_empty[ws]);
                  if (kc == KC - 1) umma_commit(&acc_full[aset]);
                }
              }
              __syncwarp();
            }
          ++iw;
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + 4 * NGRP) {
    // ===================================================== epilogue: the time scan, thread = neuron
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int n = mt * 128 + q * 32 + lane;                      // this thread's neuron (row of the layer)
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nbg = p.Bt >> 4;                                    // 16-sample groups per tile
    const int bg0 = by_set ? 0 : grp, bgstep = by_set ? 1 : NGRP;
    const float bias = IS_FWD ? p.bias[n] : 0.f;
    float dbsum = 0.f;
    // DX_ENC: this thread's encoder neuron
    const bool enc_live = MODE == NM_DX_ENC && n < p.enc_K0;
    const float enc_mk = enc_live ? p.enc_mean[n] : 0.f, enc_sk = enc_live ? p.enc_std[n] : 1.f;
    const float enc_iv = 1.f / (enc_sk * enc_sk + p.enc_eps);
    const int enc_o = enc_live ? n / p.enc_P : 0;
    float enc_dm = 0.f, enc_ds = 0.f;
    uint32_t it = 0, sk = 0;
    for (;; ++it) {
      const int tile = next_tile(sk);
      if (tile < 0) break;
      dbsum = 0.f; enc_dm = 0.f; enc_ds = 0.f;      // per-tile partial rows
      const uint32_t aset = it % nsets;
      if (by_set && aset != static_cast<uint32_t>(grp)) continue;
      const size_t c_tile = static_cast<size_t>(tile) * p.CT;      // first sample-step of the tile
      const int b_tile = tile * p.Bt;
      if (MODE == NM_DX && lane < 8) {
        // the scan below reads this warp's voltage codes (64 B per sample-step, 1 KiB per 16-sample group and timestep)
        // from the HBM-resident trace: pull them into L2 while the tensor core is still busy with this item's MMAs
        // (lane = 128-byte line of the group).  Prefetching one tile further ahead changes nothing (profiles/r2i_dxab.txt).
        for (int bg = bg0; bg < nbg; bg += bgstep)
          for (int t = 0; t < p.T; ++t)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(
                p.v_in + (static_cast<size_t>(n >> 5) * p.Rt + c_tile + static_cast<size_t>(t * p.Bt + bg * 16 + lane * 2)) * 32));
      }
      SNN_TWAIT(&acc_full[aset], (it / nsets) & 1, 0);
      tc_fence_after_sync();
      const uint32_t t_set = t_lane + aset * 256;
      if (MODE == NM_DX) {
        // BPTT scan, software-pipelined over (16-sample group, timestep) steps: the loads of a step's voltage codes
        // (one contiguous 1 KiB block per 16 sample-steps and warp) run TWO steps ahead of the recurrence.
        // No validity masks: samples >= B have g_c == 0 in the layer above (the top scan writes zeros there), so
        // their accumulators are exactly 0, and with the zero initial state every g_c written here is 0 again.
        const int nb = bg0 < nbg ? (nbg - bg0 + bgstep - 1) / bgstep : 0;
        const int nsteps = (SNN_DBG_MASK(p) & 2) ? 0 : nb * p.T;
        // g_c image [MP/64][Rt rows][128 B]: this thread's neuron is bf16 column n % 64 of chunk n / 64; a warp's 32
        // neurons are 64 contiguous bytes of a row, so every store below is one coalesced 64-byte request.
        // xo[jj]: byte offset of this neuron inside row (8 i + jj) of an 8-row group (128B swizzle: piece ^= row % 8)
        uint8_t* const gcol = p.g_img + static_cast<size_t>(n >> 6) * p.Rt * 128;
        uint32_t xo[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) xo[jj] = static_cast<uint32_t>(jj * 128 + ((((n & 63) >> 3) ^ jj) << 4) + (n & 7) * 2);
        const uint4* const vcol = reinterpret_cast<const uint4*>(p.v_in + (static_cast<size_t>(n >> 5) * p.Rt + c_tile) * 32) + lane;
        // state of the recurrence as packed pairs of neighbouring sample-steps (2 i, 2 i + 1): FFMA2 / FADD2 / FMUL2
        f32x2 gvn[8], gcn[8], db2 = 0ull;
        const f32x2 k_nc75 = f2_splat(-0.75f * kVoltInv), k_o75 = f2_splat(0.375f * kVoltInv), k_half = f2_splat(0.5f),
                    k_mid = f2_splat(-32767.5f);
        // step s <-> (group index bi, timestep t): s = bi * T + (T - 1 - t); tracked incrementally (no divisions)
        int l_bi = 0, l_t = p.T - 1;          // cursor of the loads
        int s_bi = 0, s_t = p.T - 1;          // cursor of the recurrence
        auto loadv = [&](uint32_t (&buf)[8]) {
          if (l_bi < nb && !(SNN_DBG_MASK(p) & 32)) {
            const uint4* src = vcol + static_cast<size_t>(l_t * p.Bt + (bg0 + l_bi * bgstep) * 16) * 4;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint4 q4 = __ldg(src + j * 32);
              buf[4 * j] = q4.x; buf[4 * j + 1] = q4.y; buf[4 * j + 2] = q4.z; buf[4 * j + 3] = q4.w;
            }
          }
          if (--l_t < 0) { l_t = p.T - 1; ++l_bi; }
        };
        auto step = [&](const uint32_t (&v)[8]) {
          const int col0 = (bg0 + s_bi * bgstep) * 16;
          if (s_t == p.T - 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { gvn[i] = 0ull; gcn[i] = 0ull; }
          }
          uint32_t x[16];
          tmem_ld16(t_set + static_cast<uint32_t>(s_t * p.Bt + col0), x);
          tmem_ld_wait();
          const size_t c = c_tile + static_cast<size_t>(s_t * p.Bt + col0);          // % 16 == 0
          // The epilogue is bound by instruction issue on the fp32 pipe (profiles/r2e: the warps are selected / not
          // selected / in fixed-latency or math-pipe waits 3/4 of the time), so it is written for instruction count:
          // packed fp32 pairs for the recurrence; rows c .. c + 15 of this neuron's chunk are two 1 KiB row groups and
          // both plane bases are 1 KiB aligned (api.cu aligns the workspace), so the in-group offset xo[jj] is OR-ed
          // into the low address word instead of a 64-bit add per store, and elements jj and jj + 8 (same xo, rows 8
          // apart) share one address and one bf16x2 conversion per plane.
          const unsigned long long dst_hi = reinterpret_cast<unsigned long long>(gcol + (c >> 3) * 1024);
          const unsigned long long dst_lo = dst_hi + p.g_plane;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float gc4[4];       // elements 2 i, 2 i + 1, 2 i + 8, 2 i + 9
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int q = i + 4 * h;                                              // pair (2 q, 2 q + 1) = trace word v[q]
              const f32x2 t2 = volt_t2(v[q]);                                       // voltage codes (umma.cuh)
              const f32x2 nd2 = f2_fma(t2, k_nc75, k_o75);                          // -0.75 v
              const f32x2 u2 = f2_add(t2, k_mid);
              const f32x2 win2 = f2_pack(fabsf(f2_lo(u2)) < 32767.f ? 1.f : 0.f, fabsf(f2_hi(u2)) < 32767.f ? 1.f : 0.f);
              const f32x2 keep2 = f2_pack(f2_lo(u2) > 0.f ? 0.f : 0.75f, f2_hi(u2) > 0.f ? 0.f : 0.75f);
              const f32x2 gs2 = f2_fma(gvn[q], nd2, f2_pack(__uint_as_float(x[2 * q]), __uint_as_float(x[2 * q + 1])));
              const f32x2 gv2 = f2_fma(gvn[q], keep2, f2_mul(gs2, win2));
              const f32x2 gc2 = f2_fma(k_half, gcn[q], gv2);
              gvn[q] = gv2;
              gcn[q] = gc2;
              db2 = f2_add(db2, gc2);
              gc4[2 * h] = f2_lo(gc2);
              gc4[2 * h + 1] = f2_hi(gc2);
            }
            if (!(SNN_DBG_MASK(p) & 16)) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                // rows jj and jj + 8: hi = bf16(g_c), lo = bf16(g_c - hi)
                const int jj = 2 * i + e;
                const uint32_t hp = pack_bf16x2(gc4[e], gc4[2 + e]);
                const uint32_t lp = pack_bf16x2(gc4[e] - __uint_as_float(hp << 16), gc4[2 + e] - __uint_as_float(hp & 0xFFFF0000u));
                st_global_u16x2_rows(dst_hi | xo[jj], hp);
                st_global_u16x2_rows(dst_lo | xo[jj], lp);
              }
            }
          }
          if (--s_t < 0) { s_t = p.T - 1; ++s_bi; }
        };
        uint32_t vb0[8] = {}, vb1[8] = {};
        loadv(vb0); loadv(vb1);
        for (int sidx = 0; sidx < nsteps; sidx += 2) {
          step(vb0);
          loadv(vb0);
          if (sidx + 1 < nsteps) { step(vb1); loadv(vb1); }
        }
        dbsum = f2_lo(db2) + f2_hi(db2);
      }
      for (int bg = (MODE == NM_DX || (SNN_DBG_MASK(p) & 2)) ? nbg : bg0; bg < nbg; bg += bgstep) {
        const int col0 = bg * 16;
        // bit j: sample b_tile + col0 + j exists
        const int nlive = p.B - (b_tile + col0);
        const uint32_t vmask = nlive >= 16 ? 0xFFFFu : (nlive <= 0 ? 0u : ((1u << nlive) - 1u));
        if (IS_FWD) {
          // vst = 0.75-decayed part of the next step's voltage: v (1 - s).  The reset is applied when the spike is decided
          // (one select on the predicate just computed) instead of testing the previous step's spike bit again.  State
          // as packed pairs of neighbouring samples (2 i, 2 i + 1): FADD2 / FFMA2 / FMUL2, each lane rounded like the
          // scalar operation, so spikes and voltages are bit-identical to the scalar recurrence.
          f32x2 cur[8], vst[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { cur[i] = 0ull; vst[i] = 0ull; }
          uint32_t spend = 0;
          const f32x2 bias2 = f2_splat(bias), k_half = f2_splat(0.5f), k_075 = f2_splat(0.75f);
          const bool emit_v = p.v_out != nullptr && !(SNN_DBG_MASK(p) & 128);
          // one LIF step on the accumulators of timestep t (16 samples of this neuron); the next step's accumulators are
          // in flight while it runs (two register sets, ping-pong: no copies)
          auto lif_step = [&](const uint32_t (&x)[16], int t) {
            uint32_t scur = 0;
            uint32_t cw[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              // reference order (actor_critic.py:228-231): c = c * 0.5 + (Wx + b); v = v * 0.75 * (1 - s) + c.  c * 0.5 is
              // exact, so the fused multiply-add rounds once like the reference's separate add; 0.75 * 0 + c == c exactly
              const f32x2 syn2 = f2_add(f2_pack(__uint_as_float(x[2 * i]), __uint_as_float(x[2 * i + 1])), bias2);
              const f32x2 c2 = f2_fma(cur[i], k_half, syn2);
              const f32x2 v2 = f2_add(f2_mul(vst[i], k_075), c2);
              const float v0 = f2_lo(v2), v1 = f2_hi(v2);
              const bool s0 = v0 > 0.5f, s1 = v1 > 0.5f;
              cur[i] = c2;
              vst[i] = f2_pack(s0 ? 0.f : v0, s1 ? 0.f : v1);
              scur |= ((s0 ? 1u : 0u) | (s1 ? 2u : 0u)) << (2 * i);
              if (emit_v) cw[i] = volt_code2(v0, v1);
            }
            const size_t c = c_tile + static_cast<size_t>(t * p.Bt + col0);
            // spike words (bit = neuron of this warp, one word per sample-step): the thread holds a 16-bit mask over its
            // samples, the words are the transpose.  Two timesteps share one 32 x 32 warp transpose: lane L ends up with
            // the word of sample L % 16 at step t - 1 (L < 16) or t (L >= 16) — 5 shuffles instead of 32 ballots.
            if ((t & 1) == 0 && t + 1 < p.T) {
              spend = scur;
            } else {
              const bool pair = (t & 1) != 0;
              const uint32_t wd = warp_transpose32(pair ? (spend | (scur << 16)) : scur, lane);
              const int tt = pair ? t - 1 + (lane >> 4) : t;
              if ((pair || lane < 16) && !(SNN_DBG_MASK(p) & 256))
                p.out_bits[static_cast<size_t>(n >> 5) * p.Rt + c_tile + static_cast<size_t>(tt * p.Bt + col0 + (lane & 15))] =
                    ((vmask >> (lane & 15)) & 1u) ? wd : 0u;
            }
            if (emit_v) {
              // 16-bit voltage codes (umma.cuh: volt_code2), [neuron / 32][sample-step / 8][32 neurons][8 sample-steps]:
              // the thread's 16 sample-steps are two 16-byte pieces, the warp's 16 x 32 codes one contiguous 1 KiB block
              uint4* dst = reinterpret_cast<uint4*>(p.v_out + (static_cast<size_t>(n >> 5) * p.Rt + c) * 32) + lane;
              dst[0] = make_uint4(cw[0], cw[1], cw[2], cw[3]);
              dst[32] = make_uint4(cw[4], cw[5], cw[6], cw[7]);
            }
          };
          uint32_t xa[16], xb[16];
          tmem_ld16(t_set + static_cast<uint32_t>(col0), xa);
          tmem_ld_wait();
          for (int t = 0; t < p.T; t += 2) {
            if (t + 1 < p.T) tmem_ld16(t_set + static_cast<uint32_t>((t + 1) * p.Bt + col0), xb);
            lif_step(xa, t);
            tmem_ld_wait();
            if (t + 1 < p.T) {
              if (t + 2 < p.T) tmem_ld16(t_set + static_cast<uint32_t>((t + 2) * p.Bt + col0), xa);
              lif_step(xb, t + 1);
              tmem_ld_wait();
            }
          }
        } else if (MODE == NM_DX) {
          // handled by the software-pipelined loop above
        } else {   // NM_DX_ENC
          float gam[16], ga[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { gam[j] = 0.f; ga[j] = 0.f; }
          for (int t = p.T - 1; t >= 0; --t) {
            uint32_t x[16];
            tmem_ld16(t_set + static_cast<uint32_t>(t * p.Bt + col0), x);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              gam[j] = __uint_as_float(x[j]) + 0.001f * gam[j];
              ga[j] += gam[j];
            }
          }
          if (p.g_a != nullptr) {
            float4* dst = reinterpret_cast<float4*>(p.g_a + ((static_cast<size_t>(b_tile + col0) >> 4) * p.MP + n) * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_float4(((vmask >> (4 * j)) & 1u) ? ga[4 * j] : 0.f, ((vmask >> (4 * j + 1)) & 1u) ? ga[4 * j + 1] : 0.f,
                                   ((vmask >> (4 * j + 2)) & 1u) ? ga[4 * j + 2] : 0.f, ((vmask >> (4 * j + 3)) & 1u) ? ga[4 * j + 3] : 0.f);
          }
          if (enc_live) {
            // all 16 observation loads first (clamped for samples >= B), so their latencies overlap
            float xv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int b = min(b_tile + col0 + j, p.B - 1);
              xv[j] = __ldg(p.enc_obs + static_cast<size_t>(b) * p.enc_O + enc_o);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float d = xv[j] - enc_mk;
              const float gaa = ((vmask >> j) & 1u) ? ga[j] * __expf(-0.5f * d * d * enc_iv) : 0.f;
              enc_dm += gaa * d * enc_iv;
              enc_ds += gaa * d * d * enc_iv * enc_iv * enc_sk;
            }
          }
        }
      }
      if (MODE == NM_DX && grp == 0) {
        // the padding sample-steps [TB, CT) of the tile must read as zero in the dW GEMM
        uint8_t* const gcol = p.g_img + static_cast<size_t>(n >> 6) * p.Rt * 128 + (n & 7) * 2;
        const int cn = (n & 63) >> 3;
        for (int cc = p.TB; cc < p.CT; ++cc) {
          const size_t r = c_tile + cc;
          uint8_t* q = gcol + (r >> 3) * 1024 + (r & 7) * 128 + ((cn ^ static_cast<int>(r & 7)) << 4);
          *reinterpret_cast<uint16_t*>(q) = 0;
          *reinterpret_cast<uint16_t*>(q + p.g_plane) = 0;
        }
      }
      if (MODE == NM_DX)
        p.db_part[(static_cast<size_t>(tile) * NGRP + grp) * p.MP + n] = dbsum;      // set by the scan: sum of its packed pair
      if (MODE == NM_DX_ENC) {
        p.enc_part[((static_cast<size_t>(tile) * NGRP + grp) * 2) * p.MP + n] = enc_dm;
        p.enc_part[((static_cast<size_t>(tile) * NGRP + grp) * 2 + 1) * p.MP + n] = enc_ds;
      }
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[aset]);
    }
  } else if (IS_FWD && warp >= 12) {
    // ===================================================== spike-bit expanders (smem words -> bf16 B sub-tile)
    // The operand slot is MN-major: [64-column block][64 rows = input neurons of the chunk][128 B = 64 sample-steps],
    // 128-byte swizzle (piece ^= row % 8).  The words arrive as (32 neurons) x (one sample-step); a warp transposes a
    // 32 x 32 block (lane = sample-step -> lane = neuron, five shuffles), every lane then owns 32 consecutive
    // sample-steps of one operand row = four 16-byte pieces.  Generic stores -> fence.proxy.async -> mbarrier ->
    // tcgen05.mma, the pattern dw_gemm.cuh and act_fused.cuh use for their MN-major spike operands.  (The K-major form
    // of this tile, [sample-step][64 neurons], written the same way was NOT reliable: the tensor core now and then read
    // stale 16-byte pieces, and until r2 a shared->shared bulk copy from a staging buffer stood in)
    const int ew = warp - 12;
    uint32_t u = 0, sk = 0;
    for (;;) {
      if (next_tile(sk) < 0) break;
      for (int kc = 0; kc < KC; ++kc)
        for (int sub = 0; sub < nsub; ++sub, ++u) {
          const int ncols = min(256, p.TB - sub * 256);
          const int nblk = ((ncols + 31) >> 5) * 2;                 // (32-column block, 32-neuron half) pairs
          const uint32_t s = u % NBM;
          SNN_TWAIT(&bits_full[s], (u / NBM) & 1, 0);
          const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(bits_ring + s * C::BITS_STAGE);
          uint32_t wv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int blk = ew + 4 * i, r = (blk >> 1) * 32 + lane;
            wv[i] = (blk < nblk && r < ncols) ? wsrc[(blk & 1) * 256 + r] : 0u;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&bits_empty[s]);
          // the four transposes are independent shuffle chains: no branch between them, so they interleave (a lone warp
          // per scheduler runs at the latency of its dependent instructions)
#pragma unroll
          for (int i = 0; i < 4; ++i) wv[i] = warp_transpose32(wv[i], lane);      // bit j: this lane's neuron at sample-step 32 * rb + j
          const uint32_t ss = u % C::NS;
          SNN_TWAIT(&s_empty[ss], ((u / C::NS) & 1) ^ 1, 1);
          if (!(SNN_DBG_MASK(p) & 4)) {
            uint8_t* const slot = s_ring + ss * S_STAGE;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int blk = ew + 4 * i;
              const int kl = (blk & 1) * 32 + lane, rb = blk >> 1;
              uint8_t* row = slot + (rb >> 1) * 8192 + (kl >> 3) * 1024 + (kl & 7) * 128;
              const bool on = blk < nblk;                                            // predicated stores, no branch
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 v4 = expand_spike_byte((wv[i] >> (8 * j)) & 0xFFu);
                if (on) *reinterpret_cast<uint4*>(row + (((((rb & 1) << 2) + j) ^ (kl & 7)) << 4)) = v4;
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_full[ss]);
        }
    }
  }

#ifdef SNN_ABLATE
  if (p.dbg_out != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 12)) {
    const int role = warp == 0 ? 0 : (warp == 1 ? 1 : (warp == 4 ? 2 : 3));
    unsigned long long* o = p.dbg_out + static_cast<size_t>(blockIdx.x) * 16 + role * 4;
    o[0] = wcyc[0]; o[1] = wcyc[1]; o[2] = wcyc[2];
    o[3] = static_cast<unsigned long long>(clock64() - t_start);
    if (role == 2) { o[1] = g_start; o[2] = globaltimer_ns(); }      // wall-clock stamps (ns) of this CTA: first / last instruction
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
