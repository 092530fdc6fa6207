// api.cu — C-ABI entry points declared in include/snn_b200.h.  Host-side launch sequencing only:
// every buffer is caller-owned, nothing here allocates device memory or synchronises.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "act_fused.cuh"
#include "dw_gemm.cuh"
#include "mlp_kernels.cuh"
#include "nm_gemm.cuh"
#include "plan.cuh"
#include "ppo_kernels.cuh"
#include "scan_gemm.cuh"
#include "simt_kernels.cuh"

using namespace snn;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) return fail(SNN_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__));   \
  } while (0)

#define LAUNCH_CHECK(name)                                                                      \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) return fail(SNN_ECUDA, "launch %s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

// ---- optional per-launch timing (bench.py roofline): CUDA events on the launching stream, read back later
enum { CAT_FWD = 0, CAT_DX = 1, CAT_DW = 2, CAT_OTHER = 3, NUM_KIND = 4, NUM_CAT = 32 };   // slot = kind + 4 * layer
struct ProfRec { int cat; cudaEvent_t a, b; double flops; };
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  double ms[NUM_CAT] = {};
  double flops[NUM_CAT] = {};
  long long launches[NUM_CAT] = {};
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};
Profiler g_prof;
long long g_launch_count = 0;

struct ProfScope {
  cudaStream_t st; int idx = -1;
  ProfScope(int cat, double flops, cudaStream_t s) : st(s) {
    ++g_launch_count;
    if (!g_prof.on) return;
    ProfRec r{cat, g_prof.get(), g_prof.get(), flops};
    cudaEventRecord(r.a, st);
    g_prof.recs.push_back(r);
    idx = static_cast<int>(g_prof.recs.size()) - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(g_prof.recs[idx].b, st); }
};

struct DeviceInfo {
  int checked = 0;     // 0 = not yet, 1 = ok, -1 = wrong arch
  int num_sms = 0;
  int attrs_set = 0;
  int* sched = nullptr;      // this device's instance of g_sched_slots
  int* count = nullptr;      // ... of g_count_slots
};

// Tile-scheduler counters of the persistent scan-GEMM kernels (nm_gemm.cuh): zero at module load, and every kernel leaves
// its slot zeroed again.  Launches rotate through the slots, so kernels that run concurrently on different streams do
// not share one (a slot index is baked into a captured graph node: graphs that replay concurrently must not be
// captured 64 launches apart — this library's graphs replay one after the other on one stream).
constexpr int kSchedSlots = 64;
__device__ int g_sched_slots[kSchedSlots * kSchedInts];
unsigned g_sched_next = 0;
// A second pool with the same contract (zero at load, left zeroed by every user) for the SIMT kernels that count blocks:
// the per-job "last block" counters of multi_reduce_kernel and the grid barrier of adam_fused_kernel.  Separate from the
// scheduler slots so that a reduction of the critic's stream never shares a slot with an actor kernel running next to it.
constexpr int kCountInts = 16;
__device__ int g_count_slots[kSchedSlots * kCountInts];
unsigned g_count_next = 0;
DeviceInfo g_dev[64];

int device_ready(int* num_sms) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SNN_EINVAL, "device index out of range");
  DeviceInfo& di = g_dev[dev];
  if (di.checked == 0) {
    int major = 0, sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    di.num_sms = sms;
    di.checked = (major == 10) ? 1 : -1;
  }
  if (di.checked < 0) return fail(SNN_EARCH, "device is not compute capability 10.x (sm_100a kernels only, no fallback)");
  if (!di.attrs_set) {
    CUDA_TRY(cudaFuncSetAttribute(nm_gemm_kernel<NM_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kSmemOptIn)));
    CUDA_TRY(cudaFuncSetAttribute(act_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemOptIn)));
    CUDA_TRY(cudaFuncSetAttribute(nm_gemm_kernel<NM_DX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(nm_gemm_smem_bytes<NM_DX>())));
    CUDA_TRY(cudaFuncSetAttribute(nm_gemm_kernel<NM_DX_ENC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(nm_gemm_smem_bytes<NM_DX_ENC>())));
    CUDA_TRY(cudaFuncSetAttribute(dw_gemm_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kDwSmemBytes)));
    CUDA_TRY(cudaFuncSetAttribute(dw_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kDwSmemBytes)));
    CUDA_TRY(cudaFuncSetAttribute(scan_gemm_kernel<MODE_MLP_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(scan_gemm_smem_bytes<MODE_MLP_FWD>(128))));
    CUDA_TRY(cudaFuncSetAttribute(scan_gemm_kernel<MODE_MLP_DX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(scan_gemm_smem_bytes<MODE_MLP_DX>(128))));
    void* sp = nullptr;
    CUDA_TRY(cudaGetSymbolAddress(&sp, g_sched_slots));
    di.sched = static_cast<int*>(sp);
    CUDA_TRY(cudaGetSymbolAddress(&sp, g_count_slots));
    di.count = static_cast<int*>(sp);
    di.attrs_set = 1;
  }
  *num_sms = di.num_sms;
  return SNN_OK;
}

// SM count the buffer-size queries plan for: the current device's (so that a query and the launch that follows agree on
// row_parts / dw_max_ctas on any cc 10.x part), or 148 (B200) when no device is visible (CPU-only build box).
int plan_sms() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
    return sms;
  cudaGetLastError();
  return 148;
}

constexpr int kMlpMaxCtasDefault = 0;      // 0 = no cap
unsigned long long* g_dbg_out = nullptr;   // set by snn_debug_set_cycle_buffer (tools only)

// SMs the dense critic's kernels may occupy.  In the fused PPO step they run on a second stream next to the actor's
// persistent one-CTA-per-SM kernels: a critic grid as wide as the machine takes whole waves of SMs away from them.
int mlp_sms(int sms) {
  static int cap = -1;
  if (cap < 0) { const char* e = getenv("SNN_MLP_MAX_CTAS"); cap = e ? atoi(e) : kMlpMaxCtasDefault; }
  return cap > 0 ? cap : sms;      // cap > SM count: one item per CTA (short CTAs that free their SM early)
}

// Work-skipping ablation switches (SNN_DBG) exist only in the -DSNN_ABLATE tools build (a separate library file): the
// release binary that tests and bench.py load has no such code path.
int debug_mask() {
#ifdef SNN_ABLATE
  static int m = -1;
  if (m < 0) { const char* e = getenv("SNN_DBG"); m = e ? atoi(e) : 0; }
  return m;
#else
  return 0;
#endif
}

constexpr int kRowSeg = 256;          // rows per segment of the two-level row reductions (8 loads in flight per thread, one round)
int* count_slot();
struct Deferred {
  ReduceJobs jobs{};
  int max_cols = 0, max_z = 1;
  float* scratch = nullptr;      // optional: lets jobs with many partial rows be reduced in two levels
  size_t scratch_floats = 0, scratch_used = 0;
  void row(const float* src, int nparts, size_t stride, int n, float* dst, int col_step = 1) {
    const int width = (n - 1) * col_step + 1;
    const int nseg = (nparts + kRowSeg - 1) / kRowSeg;
    RowJob jb{src, dst, nullptr, nparts, n, col_step, 1, nparts, width, static_cast<unsigned long long>(stride)};
    int cols = n;
    if (nparts > 2 * kRowSeg && scratch != nullptr && scratch_used + static_cast<size_t>(nseg) * width <= scratch_floats) {
      jb.tmp = scratch + scratch_used;
      scratch_used += static_cast<size_t>(nseg) * width;
      jb.nseg = nseg; jb.seg_rows = kRowSeg;
      cols = width;
      if (nseg > max_z) max_z = nseg;
    }
    jobs.r[jobs.nr++] = jb;
    if (cols > max_cols) max_cols = cols;
  }
  void part(const float* partial, int splits, int n_tiles, int N, int K, float* dW) {
    jobs.p[jobs.np++] = PartJob{partial, dW, splits, n_tiles, N, K};
    if (K > max_cols) max_cols = K;
    if ((N + 31) / 32 > max_z) max_z = (N + 31) / 32;
  }
};

int launch_deferred(Deferred& df, cudaStream_t st) {
  if (df.jobs.nr + df.jobs.np == 0) return SNN_OK;
  df.jobs.counters = count_slot();      // one zero-initialised, self-re-arming counter per row job (<= 12 of kCountInts)
  dim3 grd((df.max_cols + 31) / 32, df.jobs.nr + df.jobs.np, df.max_z);
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  multi_reduce_kernel<<<grd, dim3(32, 32), 0, st>>>(df.jobs);
  }
  LAUNCH_CHECK("multi_reduce_kernel");
  return SNN_OK;
}

int plan_for(const SnnDesc* d, int64_t B, int num_sms, SnnPlan* p) {
  if (d == nullptr) return fail(SNN_EINVAL, "null descriptor");
  if (make_plan(*d, B, num_sms, p) != SNN_OK) return fail(SNN_EINVAL, "unsupported SnnDesc (dims, spike_ts 1..32, planes 1..3)");
  return SNN_OK;
}

inline uint8_t* at(void* base, size_t off) { return static_cast<uint8_t*>(base) + off; }
inline const uint8_t* at(const void* base, size_t off) { return static_cast<const uint8_t*>(base) + off; }

TileGeom geom_of(const SnnPlan& p) { return TileGeom{p.Bt, p.CT, p.ntile, static_cast<long long>(p.Rt)}; }

int* sched_slot() {      // device_ready() has run on the current device
  int dev = 0;
  cudaGetDevice(&dev);
  return g_dev[dev].sched + static_cast<size_t>(g_sched_next++ % kSchedSlots) * kSchedInts;
}

int* count_slot() {      // device_ready() has run on the current device
  int dev = 0;
  cudaGetDevice(&dev);
  return g_dev[dev].count + static_cast<size_t>(g_count_next++ % kSchedSlots) * kCountInts;
}

void nm_geometry(const SnnPlan& p, int64_t B, NmParams* q) {
  q->T = p.T; q->Bt = p.Bt; q->TB = p.TB; q->CT = p.CT; q->sets = p.sets; q->ntile = p.ntile;
  q->B = static_cast<int>(B); q->Rt = p.Rt;
  q->sched = sched_slot();
  q->dbg = debug_mask();
}

int forward(const SnnDesc* d, const SnnParams* prm, const void* packed, const float* obs, int64_t B, float* mu,
            void* bits_base, void* volt_base /* nullable */, const SnnPlan& p, int num_sms, cudaStream_t st) {
  if (B == 0) return SNN_OK;
  const TileGeom tg = geom_of(p);
  uint32_t* enc_bits = reinterpret_cast<uint32_t*>(at(bits_base, p.tr_enc_bits));
  {
    dim3 blk(256), grd(static_cast<unsigned>((p.Bcap + 255) / 256), p.K0P / 32);
    { ProfScope prof__(CAT_OTHER, 0.0, st);
    if (p.T == 5)
      encode_kernel<5><<<grd, blk, 0, st>>>(obs, prm->enc_mean, prm->enc_std, static_cast<int>(B), tg, p.O, p.Pe, p.K0, p.T,
                                            d->enc_eps, enc_bits);
    else
      encode_kernel<0><<<grd, blk, 0, st>>>(obs, prm->enc_mean, prm->enc_std, static_cast<int>(B), tg, p.O, p.Pe, p.K0, p.T,
                                            d->enc_eps, enc_bits);
    }
    LAUNCH_CHECK("encode_kernel");
  }
  if (debug_mask() & 64) return SNN_OK;      // timing experiments: encoder only
  const uint32_t* in_bits = enc_bits;
  for (int l = 0; l < p.L; ++l) {
    const LayerPlan& lp = p.layer[l];
    NmParams q{};
    nm_geometry(p, B, &q);
    q.KP = lp.KP; q.MP = lp.NP; q.nmt = lp.NP / 128;
    q.a_img = at(packed, lp.w_img); q.a_plane = lp.w_plane; q.a_planes = d->fwd_planes;
    q.b_bits = in_bits;
    q.dbg_out = g_dbg_out ? g_dbg_out + static_cast<size_t>(l) * 148 * 16 : nullptr;
    q.bias = reinterpret_cast<const float*>(at(packed, lp.bias_pad));
    q.out_bits = reinterpret_cast<uint32_t*>(at(bits_base, lp.tr_bits));
    q.v_out = volt_base ? reinterpret_cast<uint16_t*>(at(volt_base, lp.tr_volt)) : nullptr;
    const int grid = nm_gemm_grid(q.nmt, p.ntile, num_sms);
    q.s_stage = nm_fwd_s_stage(p.TB);
    { ProfScope prof__(CAT_FWD + 4 * l, 2.0 * B * lp.K * lp.N * p.T, st);
    nm_gemm_kernel<NM_FWD><<<grid, NmCfg<NM_FWD>::THREADS, nm_fwd_smem_bytes(q.s_stage), st>>>(q);
    }
    LAUNCH_CHECK("nm_gemm_kernel<FWD>");
    in_bits = q.out_bits;
  }
  if (mu != nullptr) {      // training forward of the fused PPO step: the decoder runs inside actor_head_kernel (snn_bwd_ppo_top)
    const int n = static_cast<int>(B) * p.A;
    { ProfScope prof__(CAT_OTHER, 0.0, st);
    decode_kernel<<<(n + 255) / 256, 256, 0, st>>>(in_bits, tg, static_cast<int>(B), p.A, p.Pd, p.T, prm->dec_w, prm->dec_b,
                                                   d->dec_tanh, mu);
    }
    LAUNCH_CHECK("decode_kernel");
  }
  return SNN_OK;
}

// ---- geometry of the fused rollout kernel (act_fused.cuh): batch-tile width, weight-ring depth, grid.  The kernel keeps
// nothing in HBM between layers, so its tile width is free of the layered path's plan: it minimises
// waves x (max(MMA issue time, weight streaming) + epilogue scans) under the shared-memory budget.
struct ActGeom { int Bt, TB, ntile, nch, xrows, wd, nw, grid; size_t smem; };
bool act_geometry(const SnnDesc& d, const SnnPlan& p, int64_t B, int sms, ActGeom* g) {
  if (B <= 0) return false;
  int nmt_max = 0, xrows = 64;
  double nmma = 0;
  for (int l = 0; l < p.L; ++l) {
    const int nmt = p.layer[l].NP / 128;
    if (nmt > nmt_max) nmt_max = nmt;
    if (l > 0 && p.layer[l].KP > xrows) xrows = p.layer[l].KP;
    nmma += static_cast<double>(nmt) * (p.layer[l].KP / 64) * d.fwd_planes * 4;
  }
  if (nmt_max > 4) return false;
  double best = -1;
  for (int bt = 16; bt * p.T <= 256; bt += 16) {
    const int TB = bt * p.T;
    const int wd = TB <= 128 ? 128 : 256;
    if (nmt_max * wd > 512) break;
    const int nch = (TB + 63) / 64;
    const size_t fixed = act_smem_fixed(nch, xrows, p.layer[p.L - 1].NP, bt, p.A, p.K0P, TB);
    if (fixed + 3 * kActWStage > kSmemOptIn) break;
    if (act_enc_stage_bytes(p.K0P, TB, bt, p.O) > act_x_bytes(nch, xrows, p.K0P, TB)) continue;     // encoder staging lives in X
    int nw = static_cast<int>((kSmemOptIn - fixed) / kActWStage);
    if (nw > 8) nw = 8;
    const int64_t tiles = (B + bt - 1) / bt;
    const double cyc = TB / 2.0 > 57.0 + TB / 8.0 ? TB / 2.0 : 57.0 + TB / 8.0;
    const double w_rate = nw * 16384.0 / 1500.0 < 80.0 ? nw * 16384.0 / 1500.0 : 80.0;       // bytes per cycle: ring depth / L2 latency
    const double t_mma = nmma * cyc, t_w = nmma / 4 * 16384.0 / w_rate;
    double t_epi = 0;
    for (int l = 0; l < p.L; ++l) t_epi += ((p.layer[l].NP / 128) * (bt / 16) + 1) / 2 * (600.0 * p.T);
    const double per_tile = (t_mma > t_w ? t_mma : t_w) + t_epi + 3000.0;
    const double cost = static_cast<double>((tiles + sms - 1) / sms) * per_tile;
    if (best < 0 || cost < best) {
      best = cost;
      g->Bt = bt; g->TB = TB; g->ntile = static_cast<int>(tiles); g->nch = nch; g->xrows = xrows; g->wd = wd; g->nw = nw;
      g->smem = fixed + static_cast<size_t>(nw) * kActWStage; g->grid = static_cast<int>(tiles < sms ? tiles : sms);
    }
  }
  return best >= 0;
}

}  // namespace

extern "C" {

const char* snn_last_error(void) { return g_err; }
int snn_abi_version(void) { return 1; }
int snn_build_flags(void) {
#ifdef SNN_ABLATE
  return 1;
#else
  return 0;
#endif
}

size_t snn_packed_weights_bytes(const SnnDesc* d) {
  SnnPlan p;
  if (d == nullptr || make_plan(*d, 1, plan_sms(), &p) != SNN_OK) return 0;
  return p.packed_bytes;
}
size_t snn_trace_bytes(const SnnDesc* d, int64_t B) {
  SnnPlan p;
  if (d == nullptr || make_plan(*d, B, plan_sms(), &p) != SNN_OK) return 0;
  return p.trace_bytes;
}
size_t snn_workspace_bytes(const SnnDesc* d, int64_t B, int backward) {
  SnnPlan p;
  if (d == nullptr || make_plan(*d, B, plan_sms(), &p) != SNN_OK) return 0;
  return backward ? p.ws_bwd_bytes : p.ws_fwd_bytes;
}

int snn_pack_weights(const SnnDesc* d, const SnnParams* prm, void* packed, void* stream) {
  int sms = 0;
  int rc = device_ready(&sms);
  if (rc) return rc;
  SnnPlan p;
  if ((rc = plan_for(d, 1, sms, &p))) return rc;
  if (prm == nullptr || packed == nullptr) return fail(SNN_EINVAL, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PackJobs jobs{};
  int max_total = 0;
  for (int l = 0; l < p.L; ++l) {
    const LayerPlan& lp = p.layer[l];
    if (prm->weight[l] == nullptr || prm->bias[l] == nullptr) return fail(SNN_EINVAL, "null weight/bias pointer");
    jobs.j[l] = PackJob{prm->weight[l], prm->bias[l], lp.N, lp.K, lp.NP, lp.KP, at(packed, lp.w_img), lp.w_plane,
                        at(packed, lp.wt_img), lp.wt_plane, reinterpret_cast<float*>(at(packed, lp.bias_pad))};
    if (lp.NP * lp.KP / 8 > max_total) max_total = lp.NP * lp.KP / 8;
  }
  if (prm->enc_mean == nullptr || prm->enc_std == nullptr) return fail(SNN_EINVAL, "null encoder mean/std pointer");
  jobs.enc_row = p.L;
  jobs.enc_mean = prm->enc_mean; jobs.enc_std = prm->enc_std;
  jobs.enc_tab = reinterpret_cast<float4*>(at(packed, p.enc_tab));
  jobs.K0 = p.K0; jobs.K0P = p.K0P; jobs.P = p.Pe; jobs.eps = d->enc_eps;
  if (p.K0P > max_total) max_total = p.K0P;
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  pack_w_kernel<<<dim3((max_total + 255) / 256, p.L + 1), 256, 0, st>>>(jobs);
  }
  LAUNCH_CHECK("pack_w_kernel");
  return SNN_OK;
}

int snn_fwd_infer(const SnnDesc* d, const SnnParams* prm, const void* packed, const float* obs, int64_t B, float* mu,
                  void* workspace, size_t workspace_bytes, void* stream) {
  int sms = 0;
  int rc = device_ready(&sms);
  if (rc) return rc;
  SnnPlan p;
  if ((rc = plan_for(d, B, sms, &p))) return rc;
  if (!prm || !packed || !obs || !mu || !workspace) return fail(SNN_EINVAL, "null pointer");
  if (workspace_bytes < p.ws_fwd_bytes) return fail(SNN_ENOMEM, "workspace too small");
  return forward(d, prm, packed, obs, B, mu, workspace, nullptr, p, sms, static_cast<cudaStream_t>(stream));
}

int snn_fwd_act_supported(const SnnDesc* d, int64_t B) {
  SnnPlan p;
  ActGeom g;
  if (d == nullptr || B <= 0 || make_plan(*d, B, plan_sms(), &p) != SNN_OK) return 0;
  return act_geometry(*d, p, B, plan_sms(), &g) ? 1 : 0;
}

int snn_fwd_act_preferred(const SnnDesc* d, int64_t B) {
  // One CTA walks a whole tile through the network without overlapping the MMAs of one layer with the scan of the one
  // before, so beyond ~two waves of tiles the layered kernels (persistent, double-buffered accumulators) win: measured
  // at C1 dims 46.8 vs 67.7 us at B = 4096, 86 vs 98 us at 8192, 165 vs 143 us at 16384 (tools/time_fwd.py).
  SnnPlan p;
  ActGeom g;
  const int sms = plan_sms();
  if (d == nullptr || B <= 0 || make_plan(*d, B, sms, &p) != SNN_OK) return 0;
  if (!act_geometry(*d, p, B, sms, &g)) return 0;
  return g.ntile <= 2 * sms ? 1 : 0;
}

int snn_fwd_act(const SnnDesc* d, const SnnParams* prm, const void* packed, const float* obs, int64_t B, float* mu,
                const float* noise, const float* log_std, float* actions, float* logp, float* sigma, void* stream) {
  int sms = 0;
  int rc = device_ready(&sms);
  if (rc) return rc;
  SnnPlan p;
  if ((rc = plan_for(d, B, sms, &p))) return rc;
  if (!prm || !packed || !obs || !mu) return fail(SNN_EINVAL, "null pointer");
  const bool sample = noise != nullptr;
  if (sample && (!log_std || !actions || !logp || !sigma)) return fail(SNN_EINVAL, "snn_fwd_act: the sampling tail needs log_std, actions, logp and sigma");
  if (sample && p.A > 64) return fail(SNN_EINVAL, "snn_fwd_act: sampling tail supports act_dim <= 64");
  if (B == 0) return SNN_OK;
  ActGeom g;
  if (!act_geometry(*d, p, B, sms, &g)) return fail(SNN_EINVAL, "snn_fwd_act: geometry not supported by the fused kernel (spike_ts > 16, a layer wider than 512, or shared memory); use snn_fwd_infer");
  ActParams q{};
  q.L = p.L; q.T = p.T; q.Bt = g.Bt; q.TB = g.TB; q.ntile = g.ntile; q.B = static_cast<int>(B);
  q.a_planes = d->fwd_planes; q.nch = g.nch; q.xrows = g.xrows; q.wd = g.wd; q.nw = g.nw; q.K0P = p.K0P;
  q.NPlast = p.layer[p.L - 1].NP; q.O = p.O; q.K0 = p.K0; q.Pe = p.Pe;
  q.pe_magic = p.Pe < 128 ? static_cast<uint32_t>(((1u << 20) + p.Pe - 1) / p.Pe) : 0u;
  for (int l = 0; l < p.L; ++l) {
    const LayerPlan& lp = p.layer[l];
    q.layer[l] = ActLayer{at(packed, lp.w_img), lp.w_plane, reinterpret_cast<const float*>(at(packed, lp.bias_pad)), lp.KP, lp.NP, lp.NP / 128};
  }
  q.obs = obs;
  q.enc_tab = reinterpret_cast<const float4*>(at(packed, p.enc_tab));
  q.A = p.A; q.P = p.Pd; q.dec_tanh = d->dec_tanh; q.dec_w = prm->dec_w; q.dec_b = prm->dec_b; q.mu = mu;
  q.noise = noise; q.log_std = log_std; q.actions = actions; q.logp = logp; q.sigma = sigma;
  q.dbg_out = g_dbg_out; q.dbg = debug_mask();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double flops = 0;
  for (int l = 0; l < p.L; ++l) flops += 2.0 * B * p.layer[l].K * p.layer[l].N * p.T;
  { ProfScope prof__(CAT_FWD, flops, st);
  act_fused_kernel<<<g.grid, kActThreads, g.smem, st>>>(q);
  }
  LAUNCH_CHECK("act_fused_kernel");
  return SNN_OK;
}

int snn_fwd_train(const SnnDesc* d, const SnnParams* prm, const void* packed, const float* obs, int64_t B, float* mu,
                  void* trace, size_t trace_bytes, void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  int sms = 0;
  int rc = device_ready(&sms);
  if (rc) return rc;
  SnnPlan p;
  if ((rc = plan_for(d, B, sms, &p))) return rc;
  if (!prm || !packed || !obs || !trace) return fail(SNN_EINVAL, "null pointer");      // mu == NULL: no decoder pass
  if (trace_bytes < p.trace_bytes) return fail(SNN_ENOMEM, "trace buffer too small");
  return forward(d, prm, packed, obs, B, mu, trace, trace, p, sms, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

namespace {
// stages: 1 = the output layer's scan (top_scan_rows_kernel on a given g_mu, preceded by actor_head_kernel when `head` carries the
// PPO actor head's inputs), 2 = everything below it + dW + the deferred reductions.  snn_bwd runs both; the fused PPO step
// calls them separately so that the caller can fork its finalize off between them.
int backward_impl(const SnnDesc* d, const SnnParams* prm, const void* packed, const float* obs, int64_t B, const float* mu,
                  const float* g_mu, const SnnPpoActorHead* head, int stages, const void* trace, const SnnGrads* g,
                  void* workspace, size_t workspace_bytes, void* stream) {
  int sms = 0;
  int rc = device_ready(&sms);
  if (rc) return rc;
  SnnPlan p;
  if ((rc = plan_for(d, B, sms, &p))) return rc;
  if (!prm || !packed || !obs || !trace || !workspace) return fail(SNN_EINVAL, "null pointer");
  if ((stages & 2) && !g) return fail(SNN_EINVAL, "null gradient struct");
  if ((stages & 1) && head == nullptr && (!mu || !g_mu)) return fail(SNN_EINVAL, "null pointer");
  if (workspace_bytes < p.ws_bwd_bytes) return fail(SNN_ENOMEM, "workspace too small");
  if (B == 0) return fail(SNN_EINVAL, "empty batch has no gradient");
  // every region of the plan is a multiple of 1 KiB from the base: make the base itself 1 KiB aligned (ws_bwd_bytes
  // includes the slack), so the g_c planes are too
  workspace = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~static_cast<uintptr_t>(1023));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Bi = static_cast<int>(B);
  const TileGeom tg = geom_of(p);
  auto dbpart_of = [&](int l) { return reinterpret_cast<float*>(at(workspace, p.ws_dbpart + static_cast<size_t>(l) * p.ws_dbpart_stride)); };
  auto partial_of = [&](int l) { return reinterpret_cast<float*>(at(workspace, p.ws_partial + static_cast<size_t>(l) * p.ws_partial_stride)); };
  float* small = reinterpret_cast<float*>(at(workspace, p.ws_small));
  float* small_enc = reinterpret_cast<float*>(at(workspace, p.ws_small_enc));
  Deferred df;
  df.scratch = reinterpret_cast<float*>(at(workspace, p.ws_red));
  df.scratch_floats = p.ws_red_bytes / 4;
  float* g_a = reinterpret_cast<float*>(at(workspace, p.ws_ga));
  const uint32_t* enc_bits = reinterpret_cast<const uint32_t*>(at(trace, p.tr_enc_bits));

  // ---- output population layer: scan with g_up from the decoder (also yields the decoder's gradients)
  const int top = p.L - 1;
  if (stages & 1) {
    const LayerPlan& lp = p.layer[top];
    const int nrows = static_cast<int>(p.Bcap / kTopRowsSPB);
    const uint16_t* volt = reinterpret_cast<const uint16_t*>(at(trace, lp.tr_volt));
    if (head != nullptr) {
      // decoder + policy terms of the PPO head (actor_head_kernel) write mu and d loss / d mu; the scan follows directly
      if (p.A > 64) return fail(SNN_EINVAL, "snn_bwd_ppo_top: act_dim <= 64");
      if (!head->log_std || !head->actions || !head->old_logp || !head->adv || !head->old_mu || !head->old_sigma || !head->part ||
          !head->mu_out || !head->g_mu_out)
        return fail(SNN_EINVAL, "snn_bwd_ppo_top: null head pointer");
      const int hblk = static_cast<int>((B + kHeadSPB - 1) / kHeadSPB);
      if (head->part_bytes < static_cast<size_t>(hblk) * (4 + p.A) * 4) return fail(SNN_ENOMEM, "snn_bwd_ppo_top: part buffer too small");
      ActorHeadParams q{};
      q.bits = reinterpret_cast<const uint32_t*>(at(trace, lp.tr_bits));
      q.dec_w = prm->dec_w; q.dec_b = prm->dec_b; q.dec_tanh = d->dec_tanh;
      q.log_std = head->log_std; q.actions = head->actions; q.old_logp = head->old_logp; q.adv = head->adv;
      q.old_mu = head->old_mu; q.old_sigma = head->old_sigma; q.clip = head->clip; q.inv_count = 1.0f / static_cast<float>(B);
      q.mu = head->mu_out; q.g_mu = head->g_mu_out; q.part = head->part;
      q.B = Bi; q.A = p.A; q.P = p.Pd; q.T = p.T; q.tg = tg;
      { ProfScope prof__(CAT_OTHER, 0.0, st);
      actor_head_kernel<<<hblk, 256, 0, st>>>(q);
      }
      LAUNCH_CHECK("actor_head_kernel");
      mu = head->mu_out;
      g_mu = head->g_mu_out;
    }
    {
      // n / P as a multiply-shift: exact while n (P - 1) < 2^20 (n < NP <= 2048)
      const uint32_t p_magic = static_cast<int64_t>(lp.NP) * (p.Pd - 1) < (1 << 20) ? ((1u << 20) + p.Pd - 1) / p.Pd : 0u;
      { ProfScope prof__(CAT_OTHER, 0.0, st);
      const dim3 grd(nrows, lp.NP / 64);
      if (p.T == 5)
        top_scan_rows_kernel<5><<<grd, 32 * kTopRowsSPB, 0, st>>>(g_mu, mu, d->dec_tanh, prm->dec_w, volt, Bi, tg, p.A, p.Pd, lp.NP, p.T,
                                                    at(workspace, p.ws_g[top]), p.ws_g_plane[top], dbpart_of(top), small, p_magic);
      else
        top_scan_rows_kernel<0><<<grd, 32 * kTopRowsSPB, 0, st>>>(g_mu, mu, d->dec_tanh, prm->dec_w, volt, Bi, tg, p.A, p.Pd, lp.NP, p.T,
                                                    at(workspace, p.ws_g[top]), p.ws_g_plane[top], dbpart_of(top), small, p_magic);
      }
      LAUNCH_CHECK("top_scan_rows_kernel");
    }
  }
  if (!(stages & 2)) return SNN_OK;
  {
    const LayerPlan& lp = p.layer[top];
    const int nrows = static_cast<int>(p.Bcap / kTopRowsSPB);
    df.row(dbpart_of(top), nrows, lp.NP, lp.N, g->bias[top]);
    df.row(small, nrows, 2 * static_cast<size_t>(lp.NP), p.A * p.Pd, g->dec_w);
    df.row(small + lp.NP, nrows, 2 * static_cast<size_t>(lp.NP), p.A, g->dec_b, p.Pd);
  }
  // ---- walk down the layers: dX_l + scan of the layer below; the weight gradients follow in one launch
  for (int l = top; l >= 0; --l) {
    const LayerPlan& lp = p.layer[l];
    const uint8_t* G = at(workspace, p.ws_g[l]);
    const size_t Gplane = p.ws_g_plane[l];
    {   // dX_l + scan of the layer below (or of the encoder)
      NmParams q{};
      nm_geometry(p, B, &q);
      q.KP = lp.NP; q.MP = lp.KP; q.nmt = lp.KP / 128;
      q.a_img = at(packed, lp.wt_img); q.a_plane = lp.wt_plane; q.a_planes = 2;
      q.b_img = G; q.b_plane = Gplane;
      q.rev = (top - l) & 1;      // top scan: last to first; dX(top): first to last; then alternating
      q.dbg_out = g_dbg_out ? g_dbg_out + static_cast<size_t>(8 + l) * 148 * 16 : nullptr;
      const int grid = nm_gemm_grid(q.nmt, p.ntile, sms);
      if (l > 0) {
        const LayerPlan& lo = p.layer[l - 1];
        q.v_in = reinterpret_cast<const uint16_t*>(at(trace, lo.tr_volt));
        q.g_img = at(workspace, p.ws_g[l - 1]);
        q.g_plane = p.ws_g_plane[l - 1];
        q.db_part = dbpart_of(l - 1);
        { ProfScope prof__(CAT_DX + 4 * l, 2.0 * B * lp.K * lp.N * p.T, st);
        nm_gemm_kernel<NM_DX><<<grid, NmCfg<NM_DX>::THREADS, nm_gemm_smem_bytes<NM_DX>(), st>>>(q);
        }
        LAUNCH_CHECK("nm_gemm_kernel<DX>");
        df.row(dbpart_of(l - 1), 3 * p.ntile, lo.NP, lo.N, g->bias[l - 1]);
      } else {
        q.g_a = g->obs != nullptr ? g_a : nullptr;
        q.enc_obs = obs; q.enc_mean = prm->enc_mean; q.enc_std = prm->enc_std;
        q.enc_O = p.O; q.enc_P = p.Pe; q.enc_K0 = p.K0; q.enc_eps = d->enc_eps;
        q.enc_part = small_enc;
        df.row(small_enc, 3 * p.ntile, 2 * static_cast<size_t>(p.K0P), p.K0, g->enc_mean);
        df.row(small_enc + p.K0P, 3 * p.ntile, 2 * static_cast<size_t>(p.K0P), p.K0, g->enc_std);
        { ProfScope prof__(CAT_DX + 4 * l, 2.0 * B * lp.K * lp.N * p.T, st);
        nm_gemm_kernel<NM_DX_ENC><<<grid, NmCfg<NM_DX_ENC>::THREADS, nm_gemm_smem_bytes<NM_DX_ENC>(), st>>>(q);
        }
        LAUNCH_CHECK("nm_gemm_kernel<DX_ENC>");
      }
    }
  }
  // ---- dW of all layers: one launch, CTAs split over the layers in proportion to their work
  {
    DwJobs jobs{};
    jobs.count = p.L;
    double work[SNN_MAX_LAYERS], total = 0.0;
    for (int l = 0; l < p.L; ++l) { work[l] = static_cast<double>(p.layer[l].NP / 128) * (p.layer[l].KP / 64); total += work[l]; }
    int cta = 0;
    double flops = 0.0;
    for (int l = 0; l < p.L; ++l) {
      const LayerPlan& lp = p.layer[l];
      DwGemmParams& q = jobs.j[l];
      q.g_img = at(workspace, p.ws_g[l]); q.g_plane = p.ws_g_plane[l];
      q.x_bits = l == 0 ? enc_bits : reinterpret_cast<const uint32_t*>(at(trace, p.layer[l - 1].tr_bits));
      q.NP = lp.NP; q.KP = lp.KP; q.R = p.Rt;
      q.n_tiles = (lp.KP + 255) / 256;
      q.rblocks = static_cast<int>(p.Rt / 64);
      const int tiles = (lp.NP / 128) * q.n_tiles;
      int splits = static_cast<int>(sms * (work[l] / total)) / tiles;
      if (splits < 1) splits = 1;
      if (splits > q.rblocks) splits = q.rblocks;
      q.splits = splits;
      q.partial = partial_of(l);
      q.dbg_out = g_dbg_out ? g_dbg_out + static_cast<size_t>(4 + l) * 148 * 16 : nullptr;
      jobs.first_cta[l] = cta;
      cta += tiles * splits;
      flops += 2.0 * B * lp.K * lp.N * p.T;
      df.part(partial_of(l), splits, q.n_tiles, lp.N, lp.K, g->weight[l]);
    }
    jobs.first_cta[p.L] = cta;
    { ProfScope prof__(CAT_DW, flops, st);
    dw_gemm_multi_kernel<<<cta, 384, kDwSmemBytes, st>>>(jobs);
    }
    LAUNCH_CHECK("dw_gemm_multi_kernel");
  }
  // ---- d obs (the encoder parameter gradients came out of the DX_ENC epilogue, the decoder's out of top_scan_rows_kernel)
  {
    if (g->obs != nullptr) {
      const int n = Bi * p.O;
      { ProfScope prof__(CAT_OTHER, 0.0, st);
      enc_dobs_kernel<<<(n + 255) / 256, 256, 0, st>>>(obs, prm->enc_mean, prm->enc_std, g_a, Bi, p.O, p.Pe, p.K0P,
                                                       d->enc_eps, g->obs);
      }
      LAUNCH_CHECK("enc_dobs_kernel");
    }
  }
  return launch_deferred(df, st);
}
}  // namespace

extern "C" {

int snn_bwd(const SnnDesc* d, const SnnParams* prm, const void* packed, const float* obs, int64_t B, const float* mu,
            const float* g_mu, const void* trace, const SnnGrads* g, void* workspace, size_t workspace_bytes,
            void* stream) {
  if (!mu || !g_mu || !g) return fail(SNN_EINVAL, "null pointer");
  return backward_impl(d, prm, packed, obs, B, mu, g_mu, nullptr, 3, trace, g, workspace, workspace_bytes, stream);
}

int snn_bwd_ppo_top(const SnnDesc* d, const SnnParams* prm, const void* packed, const float* obs, int64_t B,
                    const SnnPpoActorHead* head, const void* trace, void* workspace, size_t workspace_bytes, void* stream) {
  if (!head) return fail(SNN_EINVAL, "null pointer");
  return backward_impl(d, prm, packed, obs, B, nullptr, nullptr, head, 1, trace, nullptr, workspace, workspace_bytes, stream);
}

int snn_bwd_rest(const SnnDesc* d, const SnnParams* prm, const void* packed, const float* obs, int64_t B, const void* trace,
                 const SnnGrads* g, void* workspace, size_t workspace_bytes, void* stream) {
  if (!g) return fail(SNN_EINVAL, "null pointer");
  return backward_impl(d, prm, packed, obs, B, nullptr, nullptr, nullptr, 2, trace, g, workspace, workspace_bytes, stream);
}

int snn_ppo_value_loss(const float* value, const float* returns, const float* target_v, int64_t B, float clip, float value_coef,
                       int clipped_value, float* g_value, void* part, size_t part_bytes, void* stream) {
  if (!value || !returns || !target_v || !g_value || !part) return fail(SNN_EINVAL, "null pointer");
  if (B <= 0) return fail(SNN_EINVAL, "empty batch");
  const int nblk = static_cast<int>((B + 255) / 256);
  if (part_bytes < static_cast<size_t>(nblk) * 4) return fail(SNN_ENOMEM, "value-loss part buffer too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  value_loss_kernel<<<nblk, 256, 0, st>>>(value, returns, target_v, static_cast<int>(B), clip, value_coef, clipped_value,
                                          1.0f / static_cast<float>(B), g_value, static_cast<float*>(part));
  }
  LAUNCH_CHECK("value_loss_kernel");
  return SNN_OK;
}

size_t snn_ppo_head_part_bytes(const SnnDesc* d, int64_t B) {
  SnnPlan p;
  if (d == nullptr || B <= 0 || make_plan(*d, B, plan_sms(), &p) != SNN_OK) return 0;
  return static_cast<size_t>((B + kHeadSPB - 1) / kHeadSPB) * (4 + p.A) * 4;
}

int snn_ppo_finalize(const SnnDesc* d, int64_t B, const void* head_part, const void* value_part, const float* log_std,
                     float ent_coef, float* g_log_std, float* stats, float* kl_out, void* stream) {
  if (!head_part || !value_part || !log_std || !g_log_std || !stats) return fail(SNN_EINVAL, "null pointer");
  SnnPlan p;
  if (d == nullptr || B <= 0 || make_plan(*d, B, plan_sms(), &p) != SNN_OK) return fail(SNN_EINVAL, "bad descriptor / batch");
  const int A = p.A;
  if (A > kHeadMaxA) return fail(SNN_EINVAL, "act_dim <= 64");
  // rows of head_part: one per block of actor_head_kernel (kHeadSPB samples); rows of value_part: one per 256 samples
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  ppo_finalize_kernel<<<1, 1024, 0, st>>>(
      static_cast<const float*>(head_part), static_cast<int>((B + kHeadSPB - 1) / kHeadSPB), static_cast<const float*>(value_part),
      static_cast<int>((B + 255) / 256), A, static_cast<int>(B), log_std, ent_coef, 1.0f, g_log_std, stats, kl_out);
  }
  LAUNCH_CHECK("ppo_finalize_kernel");
  return SNN_OK;
}

int snn_gae(const float* rewards, const uint8_t* dones, const float* values, const float* last_values, int64_t S,
            int64_t N, float gamma, float lam, float* returns, float* advantages, int normalize, void* scratch,
            void* stream) {
  if (!rewards || !dones || !values || !last_values || !returns || !advantages || !scratch)
    return fail(SNN_EINVAL, "null pointer");
  if (S <= 0 || N <= 0) return fail(SNN_EINVAL, "empty rollout");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = static_cast<int>((N + 127) / 128);
  if (nb > 1024) return fail(SNN_EINVAL, "more than 131072 environments per rank not supported by snn_gae");
  double* part_sum = static_cast<double*>(scratch);        // [1024]
  double* part_sq = part_sum + 1024;                       // [256]
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  gae_scan_kernel<<<nb, 128, 0, st>>>(rewards, dones, values, last_values, static_cast<int>(S), static_cast<int>(N), gamma,
                                      lam, returns, advantages, part_sum);
  }
  LAUNCH_CHECK("gae_scan_kernel");
  if (normalize) {
    const int64_t count = S * N;
    { ProfScope prof__(CAT_OTHER, 0.0, st);
    gae_var_kernel<<<256, 256, 0, st>>>(advantages, count, part_sum, nb, part_sq);
    }
    { ProfScope prof__(CAT_OTHER, 0.0, st);
    gae_norm_kernel<<<256, 256, 0, st>>>(advantages, count, part_sum, nb, part_sq, 256);
    }
    LAUNCH_CHECK("gae_norm_kernel");
  }
  return SNN_OK;
}

int snn_gather_rows(int count, const void* const* src, void* const* dst, const int32_t* row_bytes, const int64_t* idx,
                    int64_t n, void* stream) {
  if (!src || !dst || !row_bytes || !idx) return fail(SNN_EINVAL, "null pointer");
  if (count < 1 || count > 12 || n < 0) return fail(SNN_EINVAL, "snn_gather_rows: 1..12 fields");
  if (n == 0) return SNN_OK;
  GatherJobs jobs{};
  jobs.count = count;
  for (int k = 0; k < count; ++k) {
    if (!src[k] || !dst[k] || row_bytes[k] <= 0 || row_bytes[k] % 4) return fail(SNN_EINVAL, "snn_gather_rows: rows must be whole 4-byte words");
    jobs.j[k] = GatherJob{static_cast<const uint32_t*>(src[k]), static_cast<uint32_t*>(dst[k]), row_bytes[k] / 4};
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  gather_rows_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, st>>>(jobs, reinterpret_cast<const long long*>(idx), n);
  }
  LAUNCH_CHECK("gather_rows_kernel");
  return SNN_OK;
}

// ------------------------------------------------------------------ dense ELU MLPs on the tensor cores
// out_dim == 1: the critic (SIMT value head).  out_dim >= 2: RMA's env_factor_encoder / adaptation_module
// (RMA-rl/modules/actor_critic.py:394-418): the linear output layer is one more scan-GEMM with a plain fp32 epilogue.
namespace {
int mlp_plan_of(const SnnMlpDesc* d, int64_t B, MlpPlan* p) {
  if (!d || make_mlp_plan(d->in_dim, d->num_hidden, d->hidden, d->out_dim, B, p) != SNN_OK) return fail(SNN_EINVAL, "bad MLP descriptor");
  return SNN_OK;
}
int prof_layer(int l) { return 4 + (l < 3 ? l : 3); }

int mlp_forward(const SnnMlpDesc* d, const SnnMlpParams* prm, const void* packed, const float* x, long long xs, int64_t B,
                float* y, long long ys, void* trace, size_t trace_bytes, void* stream) {
  int sms = 0;
  int rc = device_ready(&sms);
  if (rc) return rc;
  MlpPlan p;
  if ((rc = mlp_plan_of(d, B, &p))) return rc;
  if (!prm || !packed || !x || !y || !trace) return fail(SNN_EINVAL, "null pointer");
  if (trace_bytes < p.trace_bytes) return fail(SNN_ENOMEM, "mlp trace too small");
  if (p.out_dim == 1 && p.layer[p.H - 1].NP > 256) return fail(SNN_EINVAL, "last hidden layer wider than 256 is not supported by the value head");
  if (p.out_dim == 1 && ys != 1) return fail(SNN_EINVAL, "the 1-wide head writes a contiguous [B] vector");
  if (xs < p.in_dim || ys < p.out_dim) return fail(SNN_EINVAL, "row stride smaller than the row");
  if (B == 0) return SNN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Bi = static_cast<int>(B), Bp = static_cast<int>(p.Bpad);
  {
    const int total = Bp * p.K0P / 8;
    { ProfScope prof__(CAT_OTHER, 0.0, st);
    planes_from_f32_kernel<<<(total + 255) / 256, 256, 0, st>>>(x, xs, nullptr, 0, Bi, Bp, p.in_dim, p.K0P, at(trace, p.tr_x), p.tr_x_plane);
    }
    LAUNCH_CHECK("planes_from_f32_kernel");
  }
  const uint8_t* a_img = at(trace, p.tr_x);
  size_t a_plane = p.tr_x_plane;
  for (int l = 0; l < p.LT; ++l) {
    const MlpLayerPlan& lp = p.layer[l];
    ScanGemmParams q{};
    q.a_img = a_img; q.a_plane = a_plane;
    q.b_img = at(packed, lp.w_img); q.b_plane = lp.w_plane; q.b_planes = 2;
    q.KP = lp.KP; q.NPout = lp.NP; q.NCmax = 128; q.T = 1;
    q.B = Bi; q.Bpad = Bp; q.R = p.Bpad;
    q.nchunks = lp.NP / 128;
    q.nitems = (Bp / 128) * q.nchunks;
    q.dbg = 0;
    q.bias = reinterpret_cast<const float*>(at(packed, lp.bias_pad));
    if (l < p.H) {
      q.g_img = at(trace, lp.tr_h); q.g_plane = lp.tr_h_plane;
    } else {                      // linear output layer: fp32 rows straight into the caller's buffer
      q.f32_out = y; q.f32_stride = ys; q.f32_cols = p.out_dim;
    }
    const int grid = q.nitems < mlp_sms(sms) ? q.nitems : mlp_sms(sms);
    { ProfScope prof__(CAT_FWD + 4 * prof_layer(l), 2.0 * B * lp.K * lp.N, st);
    scan_gemm_kernel<MODE_MLP_FWD><<<grid, ScanCfg<MODE_MLP_FWD>::THREADS, scan_gemm_smem_bytes<MODE_MLP_FWD>(128), st>>>(q);
    }
    LAUNCH_CHECK("scan_gemm_kernel<MLP_FWD>");
    if (l < p.H) { a_img = at(trace, lp.tr_h); a_plane = lp.tr_h_plane; }
  }
  if (p.out_dim == 1) {
    const MlpLayerPlan& lp = p.layer[p.H - 1];
    { ProfScope prof__(CAT_OTHER, 0.0, st);
    value_head_fwd_kernel<<<(Bi + 7) / 8, 256, 0, st>>>(a_img, a_plane, Bp, Bi, lp.N, lp.NP, prm->weight[p.H], prm->bias[p.H], y);
    }
    LAUNCH_CHECK("value_head_fwd_kernel");
  }
  return SNN_OK;
}

int mlp_backward(const SnnMlpDesc* d, const SnnMlpParams* prm, const void* packed, int64_t B, const float* g_y, long long gs,
                 const float* g_y2, long long g2s, const void* trace, const SnnMlpGrads* g, float* d_x, long long dxs,
                 void* workspace, size_t workspace_bytes, void* stream) {
  int sms = 0;
  int rc = device_ready(&sms);
  if (rc) return rc;
  MlpPlan p;
  if ((rc = mlp_plan_of(d, B, &p))) return rc;
  if (!prm || !packed || !g_y || !trace || !g || !workspace) return fail(SNN_EINVAL, "null pointer");
  if (workspace_bytes < p.ws_bytes) return fail(SNN_ENOMEM, "mlp workspace too small");
  if (B == 0) return fail(SNN_EINVAL, "empty batch has no gradient");
  if (p.out_dim == 1 && (gs != 1 || g_y2 != nullptr)) return fail(SNN_EINVAL, "the 1-wide head takes one contiguous [B] gradient");
  if (gs < p.out_dim || (g_y2 && g2s < p.out_dim) || (d_x && dxs < p.in_dim)) return fail(SNN_EINVAL, "row stride smaller than the row");
  if (d_x && p.K0P > kScanMaxNP) return fail(SNN_EINVAL, "gradient into the input is limited to 2048 input features");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Bi = static_cast<int>(B), Bp = static_cast<int>(p.Bpad);
  auto dbpart_of = [&](int l) { return reinterpret_cast<float*>(at(workspace, p.ws_dbpart + static_cast<size_t>(l) * p.ws_dbpart_stride)); };
  auto partial_of = [&](int l) { return reinterpret_cast<float*>(at(workspace, p.ws_partial + static_cast<size_t>(l) * p.ws_partial_stride)); };
  float* small = reinterpret_cast<float*>(at(workspace, p.ws_small));
  Deferred df;
  const int top = p.LT - 1;       // highest layer whose dz lives in operand planes
  if (p.out_dim == 1) {   // output layer backward + dz of the last hidden layer
    const MlpLayerPlan& lp = p.layer[top];
    const int nblk = Bp / 64;
    { ProfScope prof__(CAT_OTHER, 0.0, st);
    value_head_bwd_kernel<<<nblk, 256, 0, st>>>(at(trace, lp.tr_h), lp.tr_h_plane, Bp, Bi, lp.N, lp.NP, prm->weight[p.H], g_y,
                                                at(workspace, p.ws_dz[0]), p.ws_dz_plane[0], small);
    }
    LAUNCH_CHECK("value_head_bwd_kernel");
    const size_t stride = 2 * static_cast<size_t>(lp.NP) + 1;
    df.row(small, nblk, stride, lp.N, g->weight[p.H]);
    df.row(small + lp.NP, nblk, stride, lp.N, g->bias[top]);
    df.row(small + 2 * lp.NP, nblk, stride, 1, g->bias[p.H]);
  } else {                // dz of the linear output layer = the incoming gradient (sum of up to two sources)
    const MlpLayerPlan& lp = p.layer[top];
    const int total = Bp * lp.NP / 8;
    { ProfScope prof__(CAT_OTHER, 0.0, st);
    planes_from_f32_kernel<<<(total + 255) / 256, 256, 0, st>>>(g_y, gs, g_y2, g2s, Bi, Bp, p.out_dim, lp.NP,
                                                                 at(workspace, p.ws_dz[0]), p.ws_dz_plane[0]);
    }
    LAUNCH_CHECK("planes_from_f32_kernel(dz)");
    const int nblk = (Bi + 255) / 256;
    { ProfScope prof__(CAT_OTHER, 0.0, st);
    col_sum_kernel<<<nblk, 256, 0, st>>>(g_y, gs, g_y2, g2s, Bi, p.out_dim, small);
    }
    LAUNCH_CHECK("col_sum_kernel");
    df.row(small, nblk, static_cast<size_t>(p.out_dim), p.out_dim, g->bias[top]);
  }
  for (int l = top; l >= 0; --l) {
    const MlpLayerPlan& lp = p.layer[l];
    const int slot = (top - l) & 1;
    const uint8_t* G = at(workspace, p.ws_dz[slot]);
    const size_t Gplane = p.ws_dz_plane[slot];
    {   // dW_l = dz_l^T . input_l
      DwGemmParams q{};
      q.g_img = G; q.g_plane = Gplane;
      q.x_img = l == 0 ? at(trace, p.tr_x) : at(trace, p.layer[l - 1].tr_h);
      q.x_plane = l == 0 ? p.tr_x_plane : p.layer[l - 1].tr_h_plane;
      q.NP = lp.NP; q.KP = lp.KP; q.R = p.Bpad;
      q.n_tiles = (lp.KP + 255) / 256;
      q.rblocks = Bp / 64;
      const int tiles = (lp.NP / 128) * q.n_tiles;
      int splits = (mlp_sms(sms) < sms ? mlp_sms(sms) : sms) / tiles;
      { // half as many, twice as long split-K slices: half the partial tiles for the critic's reduction kernels, which
        // run next to the actor's dW launch (measured at C2: 13.75 -> 13.55 ms per update; a quarter: 13.9 ms).
        static int div = -1;      // SNN_MLP_DW_DIV overrides
        if (div < 0) { const char* e = getenv("SNN_MLP_DW_DIV"); div = e ? atoi(e) : 2; if (div < 1) div = 1; }
        splits /= div; }
      if (splits < 1) splits = 1;
      if (splits > q.rblocks) splits = q.rblocks;
      q.splits = splits;
      q.partial = partial_of(l);
      { ProfScope prof__(CAT_DW + 4 * prof_layer(l), 2.0 * B * lp.K * lp.N, st);
      dw_gemm_kernel<true><<<tiles * splits, 384, kDwSmemBytes, st>>>(q);
      }
      LAUNCH_CHECK("dw_gemm_kernel<dense>");
      df.part(partial_of(l), splits, q.n_tiles, lp.N, lp.K, g->weight[l]);
    }
    if (l > 0 || d_x != nullptr) {   // dz_{l-1} = (dz_l . W_l) * ELU'(z_{l-1});  l == 0: d x = dz_0 . W_0 (no activation below)
      ScanGemmParams q{};
      q.a_img = G; q.a_plane = Gplane;
      q.b_img = at(packed, lp.wt_img); q.b_plane = lp.wt_plane; q.b_planes = 2;
      q.KP = lp.NP; q.NPout = lp.KP; q.NCmax = 128; q.T = 1;
      q.B = Bi; q.Bpad = Bp; q.R = p.Bpad;
      q.nchunks = (lp.KP + 127) / 128;
      q.nitems = (Bp / 128) * q.nchunks;
      q.dbg = 0;
      const int grid = q.nitems < mlp_sms(sms) ? q.nitems : mlp_sms(sms);
      if (l > 0) {
        const MlpLayerPlan& lo = p.layer[l - 1];
        q.h_img = at(trace, lo.tr_h); q.h_plane = lo.tr_h_plane;
        q.g_img = at(workspace, p.ws_dz[slot ^ 1]); q.g_plane = p.ws_dz_plane[slot ^ 1];
        q.db_part = dbpart_of(l - 1);
        df.row(dbpart_of(l - 1), static_cast<int>(Bp / 32), lo.NP, lo.N, g->bias[l - 1]);
      } else {
        q.f32_out = d_x; q.f32_stride = dxs; q.f32_cols = p.in_dim;
      }
      { ProfScope prof__(CAT_DX + 4 * prof_layer(l), 2.0 * B * lp.K * lp.N, st);
      scan_gemm_kernel<MODE_MLP_DX><<<grid, ScanCfg<MODE_MLP_DX>::THREADS, scan_gemm_smem_bytes<MODE_MLP_DX>(128), st>>>(q);
      }
      LAUNCH_CHECK("scan_gemm_kernel<MLP_DX>");
    }
  }
  return launch_deferred(df, st);
}
}  // namespace

size_t snn_mlp_packed_bytes(const SnnMlpDesc* d) {
  MlpPlan p;
  if (!d || make_mlp_plan(d->in_dim, d->num_hidden, d->hidden, d->out_dim, 1, &p) != SNN_OK) return 0;
  return p.packed_bytes;
}
size_t snn_mlp_trace_bytes(const SnnMlpDesc* d, int64_t B) {
  MlpPlan p;
  if (!d || make_mlp_plan(d->in_dim, d->num_hidden, d->hidden, d->out_dim, B, &p) != SNN_OK) return 0;
  return p.trace_bytes;
}
size_t snn_mlp_workspace_bytes(const SnnMlpDesc* d, int64_t B) {
  MlpPlan p;
  if (!d || make_mlp_plan(d->in_dim, d->num_hidden, d->hidden, d->out_dim, B, &p) != SNN_OK) return 0;
  return p.ws_bytes;
}

int snn_mlp_pack(const SnnMlpDesc* d, const SnnMlpParams* prm, void* packed, void* stream) {
  int sms = 0;
  int rc = device_ready(&sms);
  if (rc) return rc;
  MlpPlan p;
  if (!prm || !packed) return fail(SNN_EINVAL, "null pointer");
  if ((rc = mlp_plan_of(d, 1, &p))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PackJobs jobs{};
  jobs.enc_row = -1;
  int max_total = 0;
  for (int l = 0; l < p.LT; ++l) {
    const MlpLayerPlan& lp = p.layer[l];
    if (!prm->weight[l] || !prm->bias[l]) return fail(SNN_EINVAL, "null weight/bias pointer");
    jobs.j[l] = PackJob{prm->weight[l], prm->bias[l], lp.N, lp.K, lp.NP, lp.KP, at(packed, lp.w_img), lp.w_plane,
                        at(packed, lp.wt_img), lp.wt_plane, reinterpret_cast<float*>(at(packed, lp.bias_pad))};
    if (lp.NP * lp.KP / 8 > max_total) max_total = lp.NP * lp.KP / 8;
  }
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  pack_w_kernel<<<dim3((max_total + 255) / 256, p.LT), 256, 0, st>>>(jobs);
  }
  LAUNCH_CHECK("pack_w_kernel(mlp)");
  return SNN_OK;
}

int snn_mlp_fwd(const SnnMlpDesc* d, const SnnMlpParams* prm, const void* packed, const float* x, int64_t B,
                float* value, void* trace, size_t trace_bytes, void* stream) {
  if (d && d->out_dim > 1) return fail(SNN_EINVAL, "snn_mlp_fwd is the 1-wide (critic) form: use snn_mlp_fwd_ex");
  return mlp_forward(d, prm, packed, x, d ? d->in_dim : 0, B, value, 1, trace, trace_bytes, stream);
}

int snn_mlp_bwd(const SnnMlpDesc* d, const SnnMlpParams* prm, const void* packed, int64_t B, const float* g_value,
                const void* trace, const SnnMlpGrads* g, void* workspace, size_t workspace_bytes, void* stream) {
  if (d && d->out_dim > 1) return fail(SNN_EINVAL, "snn_mlp_bwd is the 1-wide (critic) form: use snn_mlp_bwd_ex");
  return mlp_backward(d, prm, packed, B, g_value, 1, nullptr, 0, trace, g, nullptr, 0, workspace, workspace_bytes, stream);
}

int snn_mlp_fwd_ex(const SnnMlpDesc* d, const SnnMlpParams* prm, const void* packed, const float* x, int64_t x_stride,
                   int64_t B, float* y, int64_t y_stride, void* trace, size_t trace_bytes, void* stream) {
  return mlp_forward(d, prm, packed, x, x_stride, B, y, y_stride, trace, trace_bytes, stream);
}

int snn_mlp_bwd_ex(const SnnMlpDesc* d, const SnnMlpParams* prm, const void* packed, int64_t B, const float* g_y,
                   int64_t g_stride, const float* g_y2, int64_t g2_stride, const void* trace, const SnnMlpGrads* g,
                   float* d_x, int64_t dx_stride, void* workspace, size_t workspace_bytes, void* stream) {
  return mlp_backward(d, prm, packed, B, g_y, g_stride, g_y2, g2_stride, trace, g, d_x, dx_stride, workspace,
                      workspace_bytes, stream);
}

int snn_ppo_head(const float* mu, const float* log_std, const float* value, const float* actions,
                 const float* old_logp, const float* adv, const float* returns, const float* target_v,
                 const float* old_mu, const float* old_sigma, int64_t B, int A, float clip, float value_coef,
                 float ent_coef, int clipped_value, float* g_mu, float* g_value, float* g_log_std, float* stats,
                 void* scratch, size_t scratch_bytes, void* stream) {
  if (!mu || !log_std || !value || !actions || !old_logp || !adv || !returns || !target_v || !old_mu || !old_sigma ||
      !g_mu || !g_value || !g_log_std || !stats || !scratch)
    return fail(SNN_EINVAL, "null pointer");
  if (B <= 0 || A <= 0 || A > kHeadMaxA) return fail(SNN_EINVAL, "bad B / act_dim (act_dim <= 64)");
  const int nblk = static_cast<int>((B + 255) / 256);
  if (scratch_bytes < static_cast<size_t>(nblk) * (4 + A) * 4) return fail(SNN_ENOMEM, "ppo head scratch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PpoHeadParams q{};
  q.mu = mu; q.log_std = log_std; q.value = value; q.actions = actions; q.old_logp = old_logp; q.adv = adv;
  q.returns = returns; q.target_v = target_v; q.old_mu = old_mu; q.old_sigma = old_sigma;
  q.B = static_cast<int>(B); q.A = A; q.clip = clip; q.value_coef = value_coef; q.ent_coef = ent_coef;
  q.clipped_value = clipped_value; q.inv_count = 1.0f / static_cast<float>(B);
  q.g_mu = g_mu; q.g_value = g_value; q.part = static_cast<float*>(scratch);
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  ppo_head_kernel<<<nblk, 256, 0, st>>>(q);
  }
  LAUNCH_CHECK("ppo_head_kernel");
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  ppo_head_finalize_kernel<<<1, 32 * (4 + A) <= 1024 ? 32 * (4 + A) : 1024, 0, st>>>(q.part, nblk, A, q.B, log_std, ent_coef, 1.0f, g_log_std, stats);
  }
  LAUNCH_CHECK("ppo_head_finalize_kernel");
  return SNN_OK;
}

int snn_sample_logprob(const float* mu, const float* log_std, const float* noise, int64_t B, int A, float* actions,
                       float* logp, float* sigma, void* stream) {
  if (!mu || !log_std || !noise || !actions || !logp || !sigma) return fail(SNN_EINVAL, "null pointer");
  if (A < 1 || A > 64 || B < 0) return fail(SNN_EINVAL, "snn_sample_logprob: 1 <= act_dim <= 64");
  if (B == 0) return SNN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  sample_logprob_kernel<<<static_cast<unsigned>((B + 127) / 128), 128, 0, st>>>(mu, log_std, noise, static_cast<int>(B), A,
                                                                                actions, logp, sigma);
  }
  LAUNCH_CHECK("sample_logprob_kernel");
  return SNN_OK;
}

int snn_lr_adapt(const float* kl, float kl_scale, float desired_kl, float* lr, void* stream) {
  if (!kl || !lr) return fail(SNN_EINVAL, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  lr_adapt_kernel<<<1, 32, 0, st>>>(kl, kl_scale, desired_kl, lr);
  }
  LAUNCH_CHECK("lr_adapt_kernel");
  return SNN_OK;
}

int snn_clip_adam_kl(float* params, const float* grads, float* m, float* v, int64_t n, float* lr, int* step,
                     double beta1, double beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                     const float* kl, float kl_scale, float desired_kl, float* norm_out, void* scratch, void* stream) {
  if (!params || !grads || !m || !v || !lr || !step || !scratch) return fail(SNN_EINVAL, "null pointer");
  if (n <= 0) return fail(SNN_EINVAL, "empty parameter buffer");
  int sms = 0;
  int rc = device_ready(&sms);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // at most one block per SM: every block is resident, which the kernel's grid barrier relies on
  int nb = static_cast<int>((n + 256 * 8 - 1) / (256 * 8));
  if (nb > sms) nb = sms;
  if (nb > kAdamCoefs) nb = kAdamCoefs;
  float* partial = static_cast<float*>(scratch);     // <= 384 partial sums + Adam's two step coefficients at kAdamCoefs
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  adam_fused_kernel<<<nb, 256, 0, st>>>(params, grads, m, v, n, partial, step, lr, kl, kl_scale, desired_kl, beta1, beta2,
                                        eps, weight_decay, max_norm, grad_scale, norm_out, count_slot());
  }
  LAUNCH_CHECK("adam_fused_kernel");
  return SNN_OK;
}

// ---- peer-memory buffers of the data-parallel optimizer kernel (one process per GPU, same node: CUDA IPC over NVLink)
int snn_p2p_alloc(size_t bytes, void** ptr, void* handle64) {
  if (!ptr || !handle64 || bytes == 0) return fail(SNN_EINVAL, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, bytes));
  CUDA_TRY(cudaMemset(p, 0, bytes));
  CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  *ptr = p;
  return SNN_OK;
}
int snn_p2p_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return fail(SNN_EINVAL, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));      // maps the peer's allocation into this device
  *ptr = p;
  return SNN_OK;
}
int snn_p2p_close(void* ptr) { if (ptr) CUDA_TRY(cudaIpcCloseMemHandle(ptr)); return SNN_OK; }
int snn_p2p_free(void* ptr) { if (ptr) CUDA_TRY(cudaFree(ptr)); return SNN_OK; }

int snn_clip_adam_kl_p2p(float* params, float* grads, float* m, float* v, int64_t n, int64_t n_reduce, float* lr, int* step,
                         double beta1, double beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                         int64_t kl_index, float kl_scale, float desired_kl, float* norm_out, void* scratch,
                         void* const* stage, void* const* flags, int rank, int world, int64_t stage_stride, int* epoch,
                         void* stream) {
  if (!params || !grads || !m || !v || !lr || !step || !scratch || !stage || !flags || !epoch) return fail(SNN_EINVAL, "null pointer");
  if (n <= 0 || n_reduce < n || n_reduce > stage_stride) return fail(SNN_EINVAL, "bad sizes (n <= n_reduce <= stage_stride)");
  if (world < 2 || world > kP2PMaxWorld || rank < 0 || rank >= world) return fail(SNN_EINVAL, "world size 2..8");
  if (kl_index >= n_reduce) return fail(SNN_EINVAL, "kl_index outside the reduced range");
  int sms = 0;
  int rc = device_ready(&sms);
  if (rc) return rc;
  P2PPeers peers{};
  for (int q = 0; q < world; ++q) {
    if (!stage[q] || !flags[q]) return fail(SNN_EINVAL, "null peer pointer");
    peers.stage[q] = static_cast<float*>(stage[q]);
    peers.flags[q] = static_cast<int*>(flags[q]);
  }
  peers.rank = rank; peers.world = world; peers.stage_stride = stage_stride;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static const int one_shot = [] { const char* e = getenv("SNN_P2P_ONE_SHOT"); return e && e[0] == '1' ? 1 : 0; }();   // validation switch
  // at most one block per SM: every block is resident, which the kernel's grid barriers rely on
  int nb = static_cast<int>((n_reduce + 256 * 8 - 1) / (256 * 8));
  if (nb > sms) nb = sms;
  if (nb > kAdamCoefs) nb = kAdamCoefs;
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  p2p_adam_kernel<<<nb, 256, 0, st>>>(params, grads, m, v, n, n_reduce, static_cast<float*>(scratch), step, lr, kl_index, kl_scale,
                                      desired_kl, beta1, beta2, eps, weight_decay, max_norm, grad_scale, norm_out, count_slot(),
                                      epoch, peers, one_shot);
  }
  LAUNCH_CHECK("p2p_adam_kernel");
  return SNN_OK;
}

int snn_clip_adam(float* params, const float* grads, float* m, float* v, int64_t n, const float* lr, int* step,
                  double beta1, double beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                  float* norm_out, void* scratch, void* stream) {
  return snn_clip_adam_kl(params, grads, m, v, n, const_cast<float*>(lr), step, beta1, beta2, eps, weight_decay, max_norm,
                          grad_scale, nullptr, 0.f, 0.f, norm_out, scratch, stream);
}

int snn_mse_grad(const float* pred, const float* target, int64_t n, float* g, float* loss_accum, void* scratch,
                 void* stream) {
  if (!pred || !target || !g || !loss_accum || !scratch) return fail(SNN_EINVAL, "null pointer");
  if (n <= 0) return fail(SNN_EINVAL, "empty tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int nb = static_cast<int>((n + 1023) / 1024);
  if (nb > 256) nb = 256;
  float* partial = static_cast<float*>(scratch);     // >= 256 floats
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  mse_grad_kernel<<<nb, 256, 0, st>>>(pred, target, n, 2.0f / static_cast<float>(n), g, partial);
  }
  LAUNCH_CHECK("mse_grad_kernel");
  { ProfScope prof__(CAT_OTHER, 0.0, st);
  mse_finalize_kernel<<<1, 32, 0, st>>>(partial, nb, 1.0 / static_cast<double>(n), loss_accum);
  }
  LAUNCH_CHECK("mse_finalize_kernel");
  return SNN_OK;
}

int snn_profile_enable(int on) {
  for (auto& r : g_prof.recs) { g_prof.pool.push_back(r.a); g_prof.pool.push_back(r.b); }
  g_prof.recs.clear();
  for (int c = 0; c < NUM_CAT; ++c) { g_prof.ms[c] = 0; g_prof.flops[c] = 0; g_prof.launches[c] = 0; }
  g_prof.on = on != 0;
  return SNN_OK;
}

int snn_profile_read(double* ms, double* flops, int64_t* launches) {
  if (!ms || !flops || !launches) return fail(SNN_EINVAL, "null pointer");
  for (auto& r : g_prof.recs) {
    CUDA_TRY(cudaEventSynchronize(r.b));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, r.a, r.b));
    g_prof.ms[r.cat] += t; g_prof.flops[r.cat] += r.flops; g_prof.launches[r.cat] += 1;
    g_prof.pool.push_back(r.a); g_prof.pool.push_back(r.b);
  }
  g_prof.recs.clear();
  for (int c = 0; c < NUM_CAT; ++c) { ms[c] = g_prof.ms[c]; flops[c] = g_prof.flops[c]; launches[c] = g_prof.launches[c]; }
  return NUM_CAT;
}

int64_t snn_launch_count(void) { return g_launch_count; }

int snn_debug_set_cycle_buffer(void* buf) {
#ifdef SNN_ABLATE
  g_dbg_out = static_cast<unsigned long long*>(buf);
  return SNN_OK;
#else
  (void)buf;
  return fail(SNN_EINVAL, "cycle counters exist only in the -DSNN_ABLATE tools build");
#endif
}

int snn_debug_trace_layout(const SnnDesc* d, int64_t B, int64_t* out, int n_out) {
  SnnPlan p;
  if (d == nullptr || out == nullptr || make_plan(*d, B, plan_sms(), &p) != SNN_OK) return fail(SNN_EINVAL, "bad descriptor");
  // [0]=L [1]=T [2]=Bt [3]=CT [4]=ntile [5]=Rt [6]=K0P [7]=enc_bits_off then per layer: NP, KP, bits_off, volt_off
  const int need = 8 + 4 * p.L;
  if (n_out < need) return fail(SNN_EINVAL, "output array too small");
  out[0] = p.L; out[1] = p.T; out[2] = p.Bt; out[3] = p.CT; out[4] = p.ntile; out[5] = p.Rt; out[6] = p.K0P;
  out[7] = static_cast<int64_t>(p.tr_enc_bits);
  for (int l = 0; l < p.L; ++l) {
    out[8 + 4 * l] = p.layer[l].NP;
    out[9 + 4 * l] = p.layer[l].KP;
    out[10 + 4 * l] = static_cast<int64_t>(p.layer[l].tr_bits);
    out[11 + 4 * l] = static_cast<int64_t>(p.layer[l].tr_volt);
  }
  return need;
}

}  // extern "C"
