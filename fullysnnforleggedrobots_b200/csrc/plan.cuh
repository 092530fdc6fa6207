// plan.cuh — padded dimensions and byte offsets of every caller-owned buffer.
//
// Data layout in HBM:
//   * "swz64 image" of a bf16 matrix X[R rows][C cols] (R % 8 == 0, C % 64 == 0):
//       [C/64 column chunks][R/8 row groups][8 rows][128 bytes], the eight 16-byte pieces of a
//       128-byte row XOR-permuted by (row % 8).  A contiguous run of rows of one column chunk is
//       exactly a 128B-swizzled tcgen05 operand tile, so it is fetched with ONE 1-D bulk copy and
//       can be consumed K-major (rows = M/N, cols = K) or MN-major (rows = K, cols = M/N).
//   * spike bits: uint32 words [C/32][Rt]; word (c32, r) holds neurons 32*c32 .. +31 of sample-step r.
//   * sample-steps of the actor's time-stepped tensors are "tile-blocked": the batch is cut into tiles of Bt
//     samples (T * Bt <= 256 when that leaves Bt >= 16, so two accumulator sets fit tensor memory), and
//       r = tile * CT + t * Bt + (b % Bt),  tile = b / Bt,  CT = roundup(T * Bt, 64),  Rt = ntile * CT.
//     One tile's T * Bt sample-steps are therefore contiguous: they are the N dimension of ONE tcgen05.mma
//     (N = 240 for T = 5), which is what lets the scan-GEMMs run at the tensor core's full rate
//     (tools/mma_rate.cu: an M = 128 MMA costs max(N / 2, 57 + N / 8) cycles).
//   * g_c (time-stepped bf16 hi/lo planes): swz64 image with rows = sample-steps r, columns = neurons — the K-major
//     B operand of dX and the MN-major A operand of dW; the epilogue threads (= neurons) of a warp store 64
//     contiguous bytes of one row.
//   * voltages: 16-bit codes (umma.cuh: volt_code2 — spike flag exact, surrogate-window flag exact up to its edge, the
//     value inside the window to 2^-17) [NP/32][Rt/8][32 neurons][8 sample-steps]: an epilogue thread (= neuron) reads /
//     writes the 16 sample-steps of a group as two 16-byte pieces, its warp (32 consecutive neurons) one contiguous 1 KiB
//     block.
//   * the dense critic keeps plain rows r = b, Bpad = roundup(B, 128).
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/snn_b200.h"

namespace snn {

constexpr int kTileM = 128;       // batch rows per MMA tile (UMMA_M, cta_group::1)
constexpr int kChunkK = 64;       // bf16 elements per 128-byte swizzled row
constexpr int kMaxT = 32;

__host__ __device__ inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// byte offset of element (r, c) inside a swz64 image with R rows
__host__ __device__ inline size_t swz64_offset(int64_t r, int64_t c, int64_t R) {
  const int64_t chunk = c >> 6, cc = c & 63;
  return static_cast<size_t>(chunk) * static_cast<size_t>(R) * 128 + static_cast<size_t>(r >> 3) * 1024 +
         static_cast<size_t>(r & 7) * 128 + static_cast<size_t>(((cc >> 3) ^ (r & 7)) << 4) + (cc & 7) * 2;
}

struct LayerPlan {
  int K, N;          // logical fan-in / fan-out
  int KP, NP;        // padded: KP % 64 == 0, NP % 128 == 0
  size_t w_img;      // packed: forward operand, 3 planes of swz64 image of W  [NP rows][KP cols]
  size_t w_plane;    // bytes per plane of w_img
  size_t wt_img;     // packed: backward operand, 2 planes of swz64 image of W^T [KP rows][NP cols]
  size_t wt_plane;
  size_t bias_pad;   // packed: fp32 bias padded to NP
  size_t tr_bits;    // trace: output spike words [NP/32][Rt]
  size_t tr_volt;    // trace: 16-bit voltage codes [NP/32][Rt/8][32][8]
};

struct SnnPlan {
  int L;             // number of LIF layers (hidden + output population)
  int T;
  int O, A, Pe, Pd;
  int K0, K0P;       // encoder neurons, padded
  int Bt;            // samples per batch tile
  int TB;            // T * Bt: live sample-steps (accumulator columns) per tile
  int CT;            // roundup(TB, 64): sample-steps per tile incl. padding
  int sets;          // accumulator sets in tensor memory: 2 when TB <= 256
  int ntile;         // ceil(B / Bt)
  int64_t B, Bcap, Rt;   // Bcap = ntile * Bt, Rt = ntile * CT
  LayerPlan layer[SNN_MAX_LAYERS];
  size_t enc_tab;           // packed: per encoder neuron float4 {mean, std^2 + eps, -0.5 / (std^2 + eps), obs index bits}
  size_t packed_bytes;
  size_t tr_enc_bits;       // trace: encoder spike words [K0P/32][Rt]
  size_t trace_bytes;
  // backward workspace
  size_t ws_g[SNN_MAX_LAYERS];        // g_c of every layer (swz64 images [NP/64][Rt rows], 2 planes each): all of them
  size_t ws_g_plane[SNN_MAX_LAYERS];  // stay alive until the single dW launch at the end of the backward pass
  size_t ws_ga;             // fp32 [Bcap/16][K0P][16]  d loss / d pop_act
  size_t ws_partial, ws_partial_stride;   // fp32 split-K partials of dW, one region per layer
  size_t ws_dbpart, ws_dbpart_stride;     // fp32 partial bias grads, one region per layer
  size_t ws_small, ws_small_enc;          // partials of the decoder (ws_small) / encoder (ws_small_enc) gradients
  size_t ws_red, ws_red_bytes;            // scratch of the segmented row reductions (api.cu: Deferred)
  size_t ws_bwd_bytes;
  size_t ws_fwd_bytes;      // forward-only workspace == a trace without voltages
  int dw_max_ctas;
};

// returns 0 or SNN_EINVAL
inline int make_plan(const SnnDesc& d, int64_t B, int num_sms, SnnPlan* out) {
  SnnPlan p{};
  if (d.obs_dim <= 0 || d.act_dim <= 0 || d.en_pop <= 0 || d.de_pop <= 0) return SNN_EINVAL;
  if (d.num_hidden < 1 || d.num_hidden + 1 > SNN_MAX_LAYERS) return SNN_EINVAL;
  if (d.spike_ts < 1 || d.spike_ts > kMaxT) return SNN_EINVAL;
  if (d.fwd_planes < 1 || d.fwd_planes > 3) return SNN_EINVAL;
  if (B < 0) return SNN_EINVAL;
  p.L = d.num_hidden + 1;
  p.T = d.spike_ts;
  p.O = d.obs_dim; p.A = d.act_dim; p.Pe = d.en_pop; p.Pd = d.de_pop;
  p.K0 = d.obs_dim * d.en_pop;
  p.K0P = static_cast<int>(round_up(p.K0, 128));   // the encoder neurons are an M = 128 tile dimension in DX_ENC
  p.Bt = (256 / p.T) / 16 * 16;
  if (p.Bt > 128) p.Bt = 128;
  if (p.Bt < 16) p.Bt = 16;                           // T > 16: one accumulator set of T * 16 <= 512 columns
  if (B > 0 && p.Bt > 16) {
    // Small batches (rollout: B = num_envs): the widest tile leaves the persistent kernels with a ragged last wave
    // (4096 envs, T = 5: 86 tiles of 48 over 74 CTAs per neuron tile = 2 waves).  Pick the tile that minimises
    // waves x (tile + fixed per-item cost) for the first layer; large batches keep the widest tile.  The SM count is
    // a constant here because buffer-size queries and launches must agree on the geometry.
    const int nmt0 = static_cast<int>(round_up(d.hidden[0], 128)) / 128;
    const int ctas = 148 / nmt0 > 0 ? 148 / nmt0 : 1;
    int best = p.Bt;
    int64_t best_cost = -1;
    for (int bt = p.Bt; bt >= 16; bt -= 16) {
      const int64_t tiles = (B + bt - 1) / bt;
      const int64_t cost = ((tiles + ctas - 1) / ctas) * (bt + 8);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bt; }
    }
    p.Bt = best;
  }
  p.TB = p.T * p.Bt;
  p.CT = static_cast<int>(round_up(p.TB, 64));
  p.sets = p.TB <= 256 ? 2 : 1;
  p.B = B;
  p.ntile = static_cast<int>(((B > 0 ? B : 1) + p.Bt - 1) / p.Bt);
  p.Bcap = static_cast<int64_t>(p.ntile) * p.Bt;
  p.Rt = static_cast<int64_t>(p.ntile) * p.CT;
  size_t off = 0;
  auto take = [&off](size_t bytes) { size_t o = off; off += round_up(static_cast<int64_t>(bytes), 1024); return o; };
  int k = p.K0, kp = p.K0P;
  for (int l = 0; l < p.L; ++l) {
    LayerPlan& lp = p.layer[l];
    const int n = l < d.num_hidden ? d.hidden[l] : d.act_dim * d.de_pop;
    if (n <= 0 || n > 2048) return SNN_EINVAL;
    lp.K = k; lp.KP = kp; lp.N = n; lp.NP = static_cast<int>(round_up(n, 128));
    lp.w_plane = static_cast<size_t>(lp.KP / 64) * lp.NP * 128;
    lp.wt_plane = static_cast<size_t>(lp.NP / 64) * lp.KP * 128;
    lp.w_img = take(3 * lp.w_plane);
    lp.wt_img = take(2 * lp.wt_plane);
    lp.bias_pad = take(static_cast<size_t>(lp.NP) * 4);
    k = n; kp = lp.NP;
  }
  if (p.K0P > 8192) return SNN_EINVAL;
  p.enc_tab = take(static_cast<size_t>(p.K0P) * 16);
  p.packed_bytes = off;
  // trace
  off = 0;
  p.tr_enc_bits = take(static_cast<size_t>(p.K0P / 32) * p.Rt * 4);
  for (int l = 0; l < p.L; ++l) p.layer[l].tr_bits = take(static_cast<size_t>(p.layer[l].NP / 32) * p.Rt * 4);
  p.ws_fwd_bytes = off;
  for (int l = 0; l < p.L; ++l) p.layer[l].tr_volt = take(static_cast<size_t>(p.Rt) * p.layer[l].NP * 2);
  p.trace_bytes = off;
  // backward workspace
  off = 0;
  for (int l = 0; l < p.L; ++l) {
    p.ws_g_plane[l] = static_cast<size_t>(p.layer[l].NP) * p.Rt * 2;
    p.ws_g[l] = take(2 * p.ws_g_plane[l]);
  }
  p.ws_ga = take(static_cast<size_t>(p.Bcap) * p.K0P * 4);
  p.dw_max_ctas = num_sms > 0 ? 2 * num_sms : 296;
  const size_t row_parts = 3 * static_cast<size_t>(p.ntile);     // partial rows of the dX kernels: one per batch tile and epilogue group
  size_t part = 0;
  for (int l = 0; l < p.L; ++l) {
    // worst case: every CTA of the dW launch writes one 128 x 256 fp32 tile
    const size_t tiles = static_cast<size_t>(p.layer[l].NP / 128) * ((p.layer[l].KP + 255) / 256);
    const size_t ctas = tiles > static_cast<size_t>(p.dw_max_ctas) ? tiles : static_cast<size_t>(p.dw_max_ctas);
    const size_t b = ctas * 128 * 256 * 4;
    if (b > part) part = b;
  }
  p.ws_partial_stride = static_cast<size_t>(round_up(static_cast<int64_t>(part), 1024));
  p.ws_partial = take(p.ws_partial_stride * p.L);
  int npmax = p.K0P;
  for (int l = 0; l < p.L; ++l) if (p.layer[l].NP > npmax) npmax = p.layer[l].NP;
  size_t dbp = row_parts * npmax;
  const size_t dbp_top = static_cast<size_t>(p.ntile) * (p.Bt / 8) * p.layer[p.L - 1].NP;   // top scan: one row per 8-sample group
  if (dbp_top > dbp) dbp = dbp_top;
  p.ws_dbpart_stride = static_cast<size_t>(round_up(static_cast<int64_t>(dbp * 4), 1024));
  p.ws_dbpart = take(p.ws_dbpart_stride * p.L);
  p.ws_small = take(static_cast<size_t>(p.ntile) * (p.Bt / 8) * 2 * p.layer[p.L - 1].NP * 4 + 1024);      // decoder partials
  p.ws_small_enc = take(row_parts * 2 * p.K0P * 4 + 1024);          // encoder partials (3 rows per tile of DX_ENC)
  // scratch of the two-pass row reductions, 256 rows per segment: three jobs (bias, decoder weight, decoder bias) over
  // the top scan's partial rows, the bias rows of the layers below and the two encoder jobs over the dX kernels' rows
  size_t red_cols = 2 * static_cast<size_t>(p.K0P);
  for (int l = 0; l + 1 < p.L; ++l) red_cols += p.layer[l].NP;
  p.ws_red_bytes = 3 * ((static_cast<size_t>(p.ntile) * (p.Bt / 8) + 255) / 256) * p.layer[p.L - 1].NP * 4 +
                   ((row_parts + 255) / 256) * red_cols * 4 + 4096;
  p.ws_red = take(p.ws_red_bytes);
  p.ws_bwd_bytes = off + 1024;      // slack: snn_bwd aligns the caller's pointer up to 1 KiB (the dX epilogue ORs in-row-group
                                    // offsets into the g_c plane addresses)
  *out = p;
  return SNN_OK;
}

// ------------------------------------------------------------------ dense ELU MLP: the critic (one output, SIMT value
// head) and the RMA fork's env_factor_encoder / adaptation_module (out_dim outputs: the linear output layer runs on the
// tensor cores like a hidden layer, as plan entry layer[H])
struct MlpLayerPlan {
  int K, N, KP, NP;
  size_t w_img, w_plane;     // 3 planes (hi, mid, lo) of W [NP rows][KP cols]; the MLP kernels use hi + mid
  size_t wt_img, wt_plane;   // 2 planes of W^T [KP rows][NP cols]
  size_t bias_pad;
  size_t tr_h;               // trace: 2 planes of the layer's ELU output [NP/64][Bpad rows]
  size_t tr_h_plane;
};

struct MlpPlan {
  int H;                     // hidden (ELU) layers; out_dim == 1: the output layer [1 x hidden[H-1]] is handled by SIMT kernels
  int out_dim;               // >= 2: layer[H] is the linear output layer [out_dim x hidden[H-1]]
  int LT;                    // layers on the tensor cores: H (+ 1 when out_dim >= 2)
  int in_dim, K0P;
  int64_t B, Bpad;
  MlpLayerPlan layer[SNN_MAX_LAYERS + 1];
  size_t packed_bytes;
  size_t tr_x, tr_x_plane;   // trace: 2 planes of the input [K0P/64][Bpad rows]
  size_t trace_bytes;
  size_t ws_dz[2], ws_dz_plane[2];
  size_t ws_partial, ws_partial_stride, ws_dbpart, ws_dbpart_stride, ws_small;
  size_t ws_bytes;
};

inline int make_mlp_plan(int in_dim, int num_hidden, const int32_t* hidden, int out_dim, int64_t B, MlpPlan* out) {
  MlpPlan p{};
  if (in_dim <= 0 || num_hidden < 1 || num_hidden > SNN_MAX_LAYERS || B < 0) return SNN_EINVAL;
  if (out_dim <= 0) out_dim = 1;
  if (out_dim > 128) return SNN_EINVAL;
  p.H = num_hidden;
  p.out_dim = out_dim;
  p.LT = num_hidden + (out_dim >= 2 ? 1 : 0);
  p.in_dim = in_dim;
  p.K0P = static_cast<int>(round_up(in_dim, 64));
  p.B = B;
  p.Bpad = round_up(B > 0 ? B : 1, kTileM);
  size_t off = 0;
  auto take = [&off](size_t bytes) { size_t o = off; off += round_up(static_cast<int64_t>(bytes), 1024); return o; };
  int k = in_dim, kp = p.K0P;
  for (int l = 0; l < p.LT; ++l) {
    MlpLayerPlan& lp = p.layer[l];
    const int n = l < p.H ? hidden[l] : out_dim;
    if (n <= 0 || n > 2048) return SNN_EINVAL;
    lp.K = k; lp.KP = kp; lp.N = n; lp.NP = static_cast<int>(round_up(n, 128));
    lp.w_plane = static_cast<size_t>(lp.KP / 64) * lp.NP * 128;
    lp.wt_plane = static_cast<size_t>(lp.NP / 64) * lp.KP * 128;
    lp.w_img = take(3 * lp.w_plane);
    lp.wt_img = take(2 * lp.wt_plane);
    lp.bias_pad = take(static_cast<size_t>(lp.NP) * 4);
    k = lp.N; kp = lp.NP;
  }
  if (p.K0P > 16384) return SNN_EINVAL;        // contraction length only (RMA's observation history: 40 x 320 inputs)
  p.packed_bytes = off;
  off = 0;
  p.tr_x_plane = static_cast<size_t>(p.K0P / 64) * p.Bpad * 128;
  p.tr_x = take(2 * p.tr_x_plane);
  size_t dzmax = 0;
  int npmax = p.K0P;
  for (int l = 0; l < p.LT; ++l) {
    MlpLayerPlan& lp = p.layer[l];
    lp.tr_h_plane = static_cast<size_t>(lp.NP / 64) * p.Bpad * 128;
    if (l < p.H) lp.tr_h = take(2 * lp.tr_h_plane);       // the linear output is not an operand of anything: not kept
    if (lp.tr_h_plane > dzmax) dzmax = lp.tr_h_plane;
    if (lp.NP > npmax) npmax = lp.NP;
  }
  p.trace_bytes = off;
  off = 0;
  for (int s = 0; s < 2; ++s) { p.ws_dz_plane[s] = dzmax; p.ws_dz[s] = take(2 * dzmax); }
  size_t part = 0;
  for (int l = 0; l < p.LT; ++l) {
    const size_t tiles = static_cast<size_t>(p.layer[l].NP / 128) * ((p.layer[l].KP + 255) / 256);
    const size_t ctas = tiles > 296 ? tiles : 296;
    if (ctas * 128 * 256 * 4 > part) part = ctas * 128 * 256 * 4;
  }
  p.ws_partial_stride = static_cast<size_t>(round_up(static_cast<int64_t>(part), 1024));
  p.ws_partial = take(p.ws_partial_stride * p.LT);
  // bias-gradient partial rows: one per (128-row tile, lane quarter) of the dX epilogue = Bpad / 32
  const int64_t db_rows = p.Bpad / 32 > 296 ? p.Bpad / 32 : 296;
  p.ws_dbpart_stride = static_cast<size_t>(round_up(db_rows * npmax * 4, 1024));
  p.ws_dbpart = take(p.ws_dbpart_stride * p.LT);
  p.ws_small = take(static_cast<size_t>(p.Bpad / 64) * (2 * npmax + 8) * 4 + 1024);
  p.ws_bytes = off;
  *out = p;
  return SNN_OK;
}

}  // namespace snn
