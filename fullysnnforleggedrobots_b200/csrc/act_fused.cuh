// act_fused.cuh — the rollout-step forward of the spiking actor as ONE kernel (tcgen05 / TMEM / bulk-TMA).
//
// PopSpikeActor.forward under inference_mode (AMP/.../modules/actor_critic.py:94-122 encoder, :216-271 LIF layers,
// :140-150 decoder, :300-320 forward; called once per environment step from PPO.act, AMP/.../algorithms/ppo.py:90-102).
// One CTA owns one batch tile of Bt samples (all T spike steps: TB = T * Bt accumulator columns) and walks the whole
// network for it; spikes never leave shared memory:
//
//   layer l   D[neuron][sample-step] for ALL neuron tiles of the layer at once (nmt accumulators side by side in
//             tensor memory), chunk-major: per 64-wide contraction chunk, A = bf16 hi/mid/lo planes of W_l (K-major)
//             streamed from L2 by bulk copies, B = the spikes of the layer below as an MN-major bf16 tile
//             [input neuron][sample-step] that its PRODUCER wrote directly in operand layout:
//               layer 0    all SIMT warps encode the tile's observations into spike words in shared memory (no encoder
//                          launch), six warps then turn them chunk by chunk into operand tiles in a two-slot ring
//                          (a warp-wide 32 x 8 bit transpose per 16-byte operand piece) while the MMAs of the chunk
//                          before run;
//               layer > 0  the epilogue of layer l-1 (thread = neuron = one operand row) stores its 0/1 spikes into the
//                          tile-wide operand buffer X, in place over its own input (all MMAs of the layer are complete
//                          before any epilogue thread writes).
//             (MN-major B written with generic stores + fence.proxy.async is the pattern dw_gemm.cuh uses for its
//             spike operand; the K-major spike tile of nm_gemm_kernel<FWD> needed a staging copy)
//   epilogue  LIF scan (actor_critic.py:216-232) with thread = neuron, same rounding order as nm_gemm_kernel<FWD>
//   decoder   the output population's epilogue keeps spike counts; counts -> rates -> grouped-conv mu in decode_kernel's
//             arithmetic order (bit-identical mu), then, optionally, the sampling tail of PPO.act
//             (sample_logprob_kernel's arithmetic).
//
// Per launch: obs in, mu (+ actions / log-prob / sigma) out; the only other traffic is the weight planes from L2.
// The layered path (encode_kernel -> nm_gemm_kernel<FWD> x L -> decode_kernel) stays for training (it emits the BPTT
// traces) and for geometries this kernel does not take (see act_geometry in api.cu).
//
// Warp roles: 0 = weight producer, 1 = MMA issuer + TMEM owner, 2-3 and 12-15 = layer-0 operand chunks,
// 4-11 = epilogue (two groups of 128 TMEM lanes; work items (neuron tile, 16-sample group) alternate between them),
// then decoder + sampling tail.
#pragma once
#include "nm_gemm.cuh"
#include "simt_kernels.cuh"

namespace snn {

struct ActLayer {
  const uint8_t* a_img;     // planes of the swz64 image [KP/64][MP rows][128 B] of W_l
  size_t a_plane;
  const float* bias;        // [MP]
  int KP, MP, nmt;
};

struct ActParams {
  int L, T, Bt, TB, ntile, B;
  int a_planes;
  int nch;                  // 64-column blocks of the sample-step axis: ceil(TB / 64)
  int xrows;                // rows of the operand buffer X (max over layers > 0 of KP)
  int wd;                   // tensor-memory columns per neuron tile (power of two >= TB, nmt * wd <= 512)
  int nw;                   // weight ring units (16 KiB each), 2..8
  int K0P;                  // encoder neurons, padded (% 128 == 0)
  int K0, Pe;               // live encoder neurons, encoder population per observation
  uint32_t pe_magic;        // ceil(2^20 / Pe) (k / Pe == (k * pe_magic) >> 20 for k < 8192, Pe < 128), or 0: divide
  int NPlast;               // padded width of the output population
  int O;
  ActLayer layer[SNN_MAX_LAYERS];
  const float* obs;         // [B][O]
  const float4* enc_tab;    // [K0P] {mean, std^2 + eps, -0.5 / (std^2 + eps), obs index as int bits (-1: padding)}
  int A, P, dec_tanh;
  const float* dec_w;       // [A * P]
  const float* dec_b;       // [A]
  float* mu;                // [B][A]
  // sampling tail (all null: mu only)
  const float* noise;       // [B][A]
  const float* log_std;     // [A]
  float* actions;           // [B][A]
  float* logp;              // [B]
  float* sigma;             // [B][A]
  int dbg;                       // -DSNN_ABLATE tools build: 1 = no MMAs (timing experiments)
  unsigned long long* dbg_out;   // -DSNN_ABLATE tools build: [gridDim.x][32] wait-cycle counters and time stamps
};

constexpr int kActWStage = 16384;       // one plane of one 64-wide chunk of one neuron tile
constexpr int kActEncSlots = 2;
constexpr int kActThreads = 512;
constexpr int kActEncThreads = 192;     // warps 2, 3, 12-15

// shared memory: weight ring | encoder slots | X | spike counts of the output population | mu tile | tables | barriers
__host__ __device__ inline size_t act_x_bytes(int nch, int xrows, int K0P, int TB) {      // X also holds the tile's encoder spike words
  const size_t x = static_cast<size_t>(nch) * xrows * 128, w = static_cast<size_t>(K0P / 32) * TB * 4;
  return ((x > w ? x : w) + 1023) / 1024 * 1024;
}
// the encoder's staging area behind the words: obs tile [Bt][O] fp32 + (mean, -0.5 / var) [K0P]
__host__ __device__ inline size_t act_enc_stage_bytes(int K0P, int TB, int Bt, int O) {
  return static_cast<size_t>(K0P / 32) * TB * 4 + static_cast<size_t>((Bt * O + 1) & ~1) * 4 + static_cast<size_t>(K0P) * 8;
}
inline size_t act_smem_fixed(int nch, int xrows, int NPlast, int Bt, int A, int K0P, int TB) {
  return 1024 + static_cast<size_t>(kActEncSlots) * nch * 8192 + act_x_bytes(nch, xrows, K0P, TB) +
         static_cast<size_t>(NPlast) * Bt + static_cast<size_t>(Bt) * A * 4 + (kMaxT + 1) * 4 + 3 * 64 * 4 + 512;
}

__device__ __forceinline__ void act_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Encoder (actor_critic.py:94-122) of one batch tile into spike words [K0P/32][TB] (column = t * Bt + sample), the
// arithmetic of encode_kernel: thread = (sample, spike word); pass 1 finds the receptive fields that can fire at all,
// pass 2 runs the regular spike generator on those (fast exponent, correctly rounded fall-back near the threshold).
// Called by all 512 threads of the CTA (TT = 5: unrolled for the reference's spike_ts; 0: generic), which first stage the tile's observations and the per-neuron (mean,
// -0.5 / var) pairs in shared memory with coalesced loads (the CTA has ~14 warps: chains of global loads would be
// pure latency).  Layout inside X: words | obs tile [Bt][O] | (mean, -0.5/var) [K0P].
template <int TT>
__device__ __forceinline__ void act_encode_tile(const ActParams& p, uint8_t* __restrict__ xbuf, int b_tile, int e14) {
  uint32_t* bits = reinterpret_cast<uint32_t*>(xbuf);
  float* s_obs = reinterpret_cast<float*>(xbuf + static_cast<size_t>(p.K0P / 32) * p.TB * 4);
  float2* s_tab = reinterpret_cast<float2*>(s_obs + ((p.Bt * p.O + 1) & ~1));
  const int nlive = min(p.Bt, p.B - b_tile);
  for (int i = e14; i < nlive * p.O; i += 512) s_obs[i] = __ldg(p.obs + static_cast<size_t>(b_tile) * p.O + i);
  for (int i = e14; i < p.K0P; i += 512) {
    const float4 tb = __ldg(p.enc_tab + i);
    s_tab[i] = make_float2(tb.x, tb.z);
  }
  act_bar_sync(2, 512);
  const float q_min = __logf(0.8f * 0.999f / static_cast<float>(TT > 0 ? TT : p.T));
  for (int idx = e14; idx < (p.K0P / 32) * p.Bt; idx += 512) {
    const int kw = idx / p.Bt, bl = idx - kw * p.Bt;
    const float2* tab = s_tab + kw * 32;
    const float* xrow = s_obs + bl * p.O;
    // obs index of neuron kw * 32 + i: (kw * 32 + i) / P, tracked incrementally; neurons >= K0 are padding
    const int k0 = kw * 32;
    const int o0 = k0 / p.Pe, r0 = k0 - o0 * p.Pe;
    uint32_t active = 0;
    if (bl < nlive) {
      int o = o0, r = r0;
      float x = o < p.O ? xrow[o] : 0.f;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        if (k0 + i < p.K0) {
          const float2 tb = tab[i];
          const float d = __fsub_rn(x, tb.x);
          active |= (d * d * tb.y > q_min) ? (1u << i) : 0u;
        }
        if (++r == p.Pe) { r = 0; ++o; x = o < p.O ? xrow[o] : 0.f; }
      }
    }
    constexpr int NT = TT > 0 ? TT : 16;
    uint32_t w[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) w[t] = 0;
    while (active) {
      const int i = __ffs(active) - 1;
      active &= active - 1;
      const float2 tb = tab[i];
      const int k = k0 + i;
      const float x = xrow[p.pe_magic ? static_cast<int>((static_cast<uint32_t>(k) * p.pe_magic) >> 20) : k / p.Pe];
      const float d = __fsub_rn(x, tb.x);
      float margin;
      uint32_t sp = regular_spikes<TT>(__expf(d * d * tb.y), p.T, margin);
      if (margin < 4e-5f) {
        const float qq = __fdiv_rn(__fmul_rn(-0.5f, __fmul_rn(d, d)), __ldg(p.enc_tab + k).y);
        sp = regular_spikes<0>(static_cast<float>(exp(static_cast<double>(qq))), p.T, margin);
      }
#pragma unroll
      for (int t = 0; t < NT; ++t) w[t] |= ((sp >> t) & 1u) << i;
    }
    uint32_t* dst = bits + static_cast<size_t>(kw) * p.TB + bl;
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (TT > 0 || t < p.T) dst[t * p.Bt] = w[t];
  }
}

// every warp of the CTA takes part: once the operand buffer X of the tile before is free (tile_it > 0), stage + encode
__device__ __forceinline__ void act_encode_step(const ActParams& p, uint8_t* xbuf, int b_tile, int tid) {
  if (p.T == 5) act_encode_tile<5>(p, xbuf, b_tile, tid);
  else act_encode_tile<0>(p, xbuf, b_tile, tid);
  act_bar_sync(2, 512);
}

__global__ void __launch_bounds__(kActThreads, 1) act_fused_kernel(const ActParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_ring = smem;
  uint8_t* e_ring = w_ring + p.nw * kActWStage;                              // [kActEncSlots][nch][64 rows][128 B]
  uint8_t* xbuf = e_ring + kActEncSlots * p.nch * 8192;                      // [nch][xrows][128 B]
  uint8_t* cnt = xbuf + act_x_bytes(p.nch, p.xrows, p.K0P, p.TB);        // [NPlast][Bt] spike counts (uint8)
  float* mu_tile = reinterpret_cast<float*>(cnt + static_cast<size_t>(p.NPlast) * p.Bt);   // [Bt][A]
  float* s_rate = mu_tile + p.Bt * p.A;                                      // [kMaxT + 1]
  float* s_sig = s_rate + (kMaxT + 1);                                       // [3][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_sig + 3 * 64) + 7) & ~uintptr_t(7));
  uint64_t* w_full = bars;                       // [8]
  uint64_t* w_empty = w_full + 8;                // [8]
  uint64_t* e_full = w_empty + 8;                // [2]  encoder slot written
  uint64_t* e_empty = e_full + 2;                // [2]  encoder slot consumed by the MMAs
  uint64_t* acc_full = e_empty + 2;              // [1]  all MMAs of a layer complete
  uint64_t* layer_bar = acc_full + 1;            // [SNN_MAX_LAYERS]  layer l of the tile: accumulators read, X written
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(layer_bar + SNN_MAX_LAYERS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t nw = static_cast<uint32_t>(p.nw);
  const int KC0 = p.K0P / 64;
#ifdef SNN_ABLATE
  unsigned long long wcyc[3] = {0, 0, 0};
  const long long t_start = clock64();
#define ACT_STAMP(k) do { if (p.dbg_out != nullptr && tile_it == 0) p.dbg_out[static_cast<size_t>(blockIdx.x) * 32 + 20 + (k)] = static_cast<unsigned long long>(clock64() - t_start); } while (0)
#else
#define ACT_STAMP(k) do { } while (0)
#endif

  if (tid == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&e_full[s], kActEncThreads / 32); mbar_init(&e_empty[s], 1); }
    mbar_init(acc_full, 1);
    for (int l = 0; l < SNN_MAX_LAYERS; ++l) mbar_init(&layer_bar[l], 8);
    fence_mbar_init();
  }
  if (tid >= 32 && tid <= 32 + p.T) s_rate[tid - 32] = __fdiv_rn(static_cast<float>(tid - 32), static_cast<float>(p.T));
  if (p.noise != nullptr && tid >= 64 && tid < 64 + p.A) {
    const float sg = expf(p.log_std[tid - 64]);
    s_sig[tid - 64] = sg;
    s_sig[64 + tid - 64] = 2.f * (sg * sg);
    s_sig[128 + tid - 64] = logf(sg);
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================================================== weight producer: ring unit = (chunk, neuron tile, plane)
    // A lone warp: its issue rate is instruction latency, so the loop is kept to a wait, one elected bulk copy and
    // incremental indices (no division).  Per tile: fill the ring, help encode the tile, then stream the rest.
    uint32_t ws = 0, wph = 1, tile_it = 0;
    int units_tile = 0;
    for (int l = 0; l < p.L; ++l) units_tile += (p.layer[l].KP / 64) * p.layer[l].nmt * p.a_planes;
    const int pre = units_tile < static_cast<int>(nw) ? units_tile : static_cast<int>(nw);
    for (int tile = blockIdx.x; tile < p.ntile; tile += gridDim.x, ++tile_it) {
      int cnt = 0;
      for (int l = 0; l < p.L; ++l) {
        const ActLayer& ly = p.layer[l];
        const int KC = ly.KP / 64;
        for (int kc = 0; kc < KC; ++kc)
          for (int mt = 0; mt < ly.nmt; ++mt) {
            const uint8_t* src = ly.a_img + (static_cast<size_t>(kc) * ly.MP + static_cast<size_t>(mt) * 128) * 128;
            for (int pl = 0; pl < p.a_planes; ++pl, src += ly.a_plane) {
              SNN_TWAIT(&w_empty[ws], wph, 0);
              if (elect_one()) {
                mbar_expect_tx(&w_full[ws], 16384u);
                bulk_g2s(w_ring + ws * kActWStage, src, 16384u, &w_full[ws]);
              }
              __syncwarp();
              if (++ws == nw) { ws = 0; wph ^= 1u; }
              if (++cnt == pre) {
                if (tile_it > 0) SNN_TWAIT(&layer_bar[p.L - 1], (tile_it - 1) & 1, 1);
                if (tid == 0) ACT_STAMP(9);
                act_encode_step(p, xbuf, tile * p.Bt, tid);
                if (tid == 0) ACT_STAMP(10);
              }
            }
          }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    uint32_t ws = 0, wph = 0, ue = 0, tile_it = 0;      // weight ring slot / its phase parity, encoder chunk, tile count
    const int npl = p.a_planes;
    const uint32_t dhi = smem_desc_hi(1024);
    const uint32_t w_ring_lo = smem_desc_lo(smem_u32(w_ring), 16);
    const uint32_t e_ring_lo = smem_desc_lo(smem_u32(e_ring), 8192);                               // MN-major: LBO = 64-column block
    const uint32_t x_lo = smem_desc_lo(smem_u32(xbuf), static_cast<uint32_t>(p.xrows) * 128u);
    const uint32_t idesc = make_idesc_bf16(128, static_cast<uint32_t>(p.TB), 0, 1);
    for (int tile = blockIdx.x; tile < p.ntile; tile += gridDim.x, ++tile_it) {
      // the accumulators and X are free once the epilogue of the tile before is through; then every warp encodes
      if (tile_it > 0) SNN_TWAIT(&layer_bar[p.L - 1], (tile_it - 1) & 1, 0);
      act_encode_step(p, xbuf, tile * p.Bt, tid);
      for (int l = 0; l < p.L; ++l) {
        const int KC = p.layer[l].KP / 64, nmt = p.layer[l].nmt;
        // X holds this layer's input, and the accumulators are free, once the epilogue of the layer before is through
        if (l > 0) SNN_TWAIT(&layer_bar[l - 1], tile_it & 1, 0);
        tc_fence_after_sync();
        if (lane == 0) ACT_STAMP(4 + l);
        for (int kc = 0; kc < KC; ++kc) {
          uint32_t b_lo;
          const uint32_t es = ue % kActEncSlots;
          if (l == 0) {
            SNN_TWAIT(&e_full[es], (ue / kActEncSlots) & 1, 2);
            b_lo = e_ring_lo + es * (static_cast<uint32_t>(p.nch) * 8192u >> 4);
          } else {
            b_lo = x_lo + static_cast<uint32_t>(kc) * (8192u >> 4);
          }
          // per weight plane: wait, four MMAs, release the plane.  The issue loop is a lone warp (instruction latency):
          // ring index and phase are kept incrementally (no division)
          for (int mt = 0; mt < nmt; ++mt) {
            const uint32_t d = tmem_base + static_cast<uint32_t>(mt * p.wd);
            for (int pl = 0; pl < npl; ++pl) {
              SNN_TWAIT(&w_full[ws], wph, 1);
              tc_fence_after_sync();
              if (elect_one()) {
                const uint32_t a_lo = w_ring_lo + ws * (kActWStage >> 4);
                if (!(SNN_DBG_MASK(p) & 1)) {
                  umma_bf16_lohi(d, a_lo, dhi, b_lo, dhi, idesc, (kc > 0 || pl > 0) ? 1u : 0u);
                  umma_bf16_lohi(d, a_lo + 2, dhi, b_lo + (2048 >> 4), dhi, idesc, 1u);
                  umma_bf16_lohi(d, a_lo + 4, dhi, b_lo + 2 * (2048 >> 4), dhi, idesc, 1u);
                  umma_bf16_lohi(d, a_lo + 6, dhi, b_lo + 3 * (2048 >> 4), dhi, idesc, 1u);
                }
                umma_commit(&w_empty[ws]);
                if (mt == nmt - 1 && pl == npl - 1) {
                  if (l == 0) umma_commit(&e_empty[es]);
                  if (kc == KC - 1) umma_commit(acc_full);
                }
              }
              __syncwarp();
              if (++ws == nw) { ws = 0; wph ^= 1u; }
            }
          }
          if (l == 0) ++ue;
        }
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 12) {
    // ===================================================== layer-0 operand chunks from the tile's encoder spike words
    // (a 32 x 8 bit transpose per warp task: lane = operand row = encoder neuron, the 8 words of the piece's 8
    // sample-steps are broadcast loads, every lane extracts its own bit -> one 16-byte operand piece)
    const int ew = warp < 4 ? warp - 2 : warp - 10;                     // 0 .. 5
    const uint32_t* bits = reinterpret_cast<const uint32_t*>(xbuf);
    const int ntask = (p.TB >> 3) * 2;
    uint32_t ue = 0, tile_it = 0;
    for (int tile = blockIdx.x; tile < p.ntile; tile += gridDim.x, ++tile_it) {
      // the word buffer shares X: free once the last layer of the tile before has consumed its input
      if (tile_it > 0) SNN_TWAIT(&layer_bar[p.L - 1], (tile_it - 1) & 1, 1);
      act_encode_step(p, xbuf, tile * p.Bt, tid);
      for (int kc = 0; kc < KC0; ++kc, ++ue) {
        const uint32_t es = ue % kActEncSlots;
        SNN_TWAIT(&e_empty[es], ((ue / kActEncSlots) & 1) ^ 1, 0);
        uint8_t* slot = e_ring + static_cast<size_t>(es) * p.nch * 8192;
        for (int wt = ew; wt < ntask; wt += kActEncThreads / 32) {
          const int pc = wt >> 1, h = wt & 1;
          const int kl = 32 * h + lane;
          const uint4* wp = reinterpret_cast<const uint4*>(bits + static_cast<size_t>(2 * kc + h) * p.TB + 8 * pc);
          const uint4 wa = wp[0], wb = wp[1];
          const uint32_t byte = ((wa.x >> lane) & 1u) | (((wa.y >> lane) & 1u) << 1) | (((wa.z >> lane) & 1u) << 2) |
                                (((wa.w >> lane) & 1u) << 3) | (((wb.x >> lane) & 1u) << 4) | (((wb.y >> lane) & 1u) << 5) |
                                (((wb.z >> lane) & 1u) << 6) | (((wb.w >> lane) & 1u) << 7);
          *reinterpret_cast<uint4*>(slot + (pc >> 3) * 8192 + (kl >> 3) * 1024 + (kl & 7) * 128 + (((pc & 7) ^ (kl & 7)) << 4)) =
              expand_spike_byte(byte);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&e_full[es]);
      }
    }
  } else {
    // ===================================================== epilogue: LIF scan (actor_critic.py:216-232), thread = neuron
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int et = tid - 128;                                           // 0 .. 255
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nbg = p.Bt >> 4;
    uint32_t lc = 0, tile_it = 0;
    for (int tile = blockIdx.x; tile < p.ntile; tile += gridDim.x, ++tile_it) {
      const int b_tile = tile * p.Bt;
      // all warps encode the tile first (these warps have just finished the tile before: X is free)
      act_encode_step(p, xbuf, b_tile, tid);
      for (int l = 0; l < p.L; ++l, ++lc) {
        const ActLayer& ly = p.layer[l];
        const bool last = l == p.L - 1;
        SNN_TWAIT(acc_full, lc & 1, 0);
        tc_fence_after_sync();
        if (tid == 128) ACT_STAMP(l);
        const int nitems = ly.nmt * nbg;
        for (int idx = grp; idx < nitems; idx += 2) {
          const int mt = idx / nbg, bg = idx - mt * nbg;
          const int n = mt * 128 + q * 32 + lane;
          const float bias = __ldg(ly.bias + n);
          const int col0 = bg * 16;
          const int nlive = p.B - (b_tile + col0);
          const uint32_t vmask = nlive >= 16 ? 0xFFFFu : (nlive <= 0 ? 0u : ((1u << nlive) - 1u));
          const uint32_t t_acc = t_lane + static_cast<uint32_t>(mt * p.wd + col0);
          uint8_t* const xrow = xbuf + (n >> 3) * 1024 + (n & 7) * 128;
          // LIF state as packed fp32 pairs of neighbouring samples (FADD2 / FFMA2 / FMUL2, each lane rounded like the scalar
          // operation); vst = v (1 - s): the reset is applied when the spike is decided; c * 0.5 is exact, so the fused
          // multiply-add rounds once like the reference's separate add — the same code as the training forward
          // (nm_gemm.cuh), bit-identical spikes
          f32x2 cur[8], vst[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { cur[i] = 0ull; vst[i] = 0ull; }
          const f32x2 bias2 = f2_splat(bias), k_half = f2_splat(0.5f), k_075 = f2_splat(0.75f);
          uint32_t c4[4] = {0, 0, 0, 0};                                  // spike counts of the 16 samples, one byte each
          auto lif_step = [&](const uint32_t (&x)[16], int t) {
            uint32_t scur = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const f32x2 syn2 = f2_add(f2_pack(__uint_as_float(x[2 * i]), __uint_as_float(x[2 * i + 1])), bias2);
              const f32x2 c2 = f2_fma(cur[i], k_half, syn2);
              const f32x2 v2 = f2_add(f2_mul(vst[i], k_075), c2);
              const float v0 = f2_lo(v2), v1 = f2_hi(v2);
              const bool s0 = v0 > 0.5f, s1 = v1 > 0.5f;
              cur[i] = c2;
              vst[i] = f2_pack(s0 ? 0.f : v0, s1 ? 0.f : v1);
              scur |= ((s0 ? 1u : 0u) | (s1 ? 2u : 0u)) << (2 * i);
            }
            const uint32_t sm = scur & vmask;
            if (!last) {
              // this neuron's row of the next layer's operand: 16 sample-steps = two 16-byte pieces
              const int c = t * p.Bt + col0;
              uint8_t* dst = xrow + static_cast<size_t>(c >> 6) * p.xrows * 128;
              const int pc = (c & 63) >> 3;
              *reinterpret_cast<uint4*>(dst + ((pc ^ (n & 7)) << 4)) = expand_spike_byte(sm & 0xFFu);
              *reinterpret_cast<uint4*>(dst + (((pc + 1) ^ (n & 7)) << 4)) = expand_spike_byte(sm >> 8);
            } else {
#pragma unroll
              for (int w4 = 0; w4 < 4; ++w4) {
                const uint32_t nib = (sm >> (4 * w4)) & 0xFu;             // 4 samples -> 4 byte lanes
                c4[w4] += (nib * 0x00204081u) & 0x01010101u;
              }
            }
          };
          // the next step's accumulators are in flight while a step runs (two register sets, ping-pong: no copies)
          uint32_t xa[16], xb[16];
          tmem_ld16(t_acc, xa);
          tmem_ld_wait();
          for (int t = 0; t < p.T; t += 2) {
            if (t + 1 < p.T) tmem_ld16(t_acc + static_cast<uint32_t>((t + 1) * p.Bt), xb);
            lif_step(xa, t);
            tmem_ld_wait();
            if (t + 1 < p.T) {
              if (t + 2 < p.T) tmem_ld16(t_acc + static_cast<uint32_t>((t + 2) * p.Bt), xa);
              lif_step(xb, t + 1);
              tmem_ld_wait();
            }
          }
          if (last) *reinterpret_cast<uint4*>(cnt + static_cast<size_t>(n) * p.Bt + col0) = make_uint4(c4[0], c4[1], c4[2], c4[3]);
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&layer_bar[l]);
      }
      // ===================================================== decoder (actor_critic.py:140-150) + sampling tail
      act_bar_sync(1, 256);
      if (tid == 128) ACT_STAMP(p.L);
      for (int idx = et; idx < p.A * p.Bt; idx += 256) {
        const int a = idx / p.Bt, bl = idx - a * p.Bt;
        const int n0 = a * p.P;
        float acc = 0.f;
        for (int pp = 0; pp < p.P; ++pp)
          acc = __fmaf_rn(__ldg(p.dec_w + n0 + pp), s_rate[cnt[static_cast<size_t>(n0 + pp) * p.Bt + bl]], acc);
        acc = __fadd_rn(acc, __ldg(p.dec_b + a));
        const float m = p.dec_tanh ? tanhf(acc) : acc;
        mu_tile[bl * p.A + a] = m;
        if (b_tile + bl < p.B) p.mu[static_cast<size_t>(b_tile + bl) * p.A + a] = m;
      }
      if (p.noise != nullptr) {
        // PPO.act's tail (actor_critic.py:398-421): sigma, action = mu + sigma * noise, summed Normal log-prob
        act_bar_sync(1, 256);
        const int b = b_tile + et;
        if (et < p.Bt && b < p.B) {
          float lp = 0.f;
          for (int a = 0; a < p.A; ++a) {
            const size_t i = static_cast<size_t>(b) * p.A + a;
            const float m = mu_tile[et * p.A + a], sg = s_sig[a];
            const float x = __fadd_rn(m, __fmul_rn(sg, __ldg(p.noise + i)));
            const float dd = __fsub_rn(x, m);
            lp += __fsub_rn(__fsub_rn(__fdiv_rn(-(dd * dd), s_sig[64 + a]), s_sig[128 + a]), 0.91893853320467267f);
            p.actions[i] = x;
            p.sigma[i] = sg;
          }
          p.logp[b] = lp;
        }
      }
      act_bar_sync(1, 256);       // counts / mu tile are rewritten by the next tile
    }
  }

#ifdef SNN_ABLATE
  if (p.dbg_out != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 4 || warp == 8)) {
    const int role = warp == 0 ? 0 : (warp == 1 ? 1 : (warp == 2 ? 2 : (warp == 4 ? 3 : 4)));
    unsigned long long* o = p.dbg_out + static_cast<size_t>(blockIdx.x) * 32 + role * 4;
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
#undef ACT_STAMP

}  // namespace snn
