// Microbenchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16, SS mode, 128B swizzle) as a function of N.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fullysnnforleggedrobots_b200/csrc tools/mma_rate.cu -o gpurun_out/mma_rate
// One CTA per SM, one issuing warp; operands are whatever is in shared memory (zeros) — only the issue rate matters.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace snn;

struct Args {
  int N;          // MMA N
  int nacc;       // number of distinct accumulators cycled through (nacc * N <= 512)
  int a_tiles;    // number of distinct A tiles cycled through (1 = same A every MMA)
  int b_tiles;    // number of distinct B tiles
  int iters;
  int commit_every;   // > 0: tcgen05.commit to a scratch mbarrier after every commit_every groups of 4 MMAs
  int b_mn;       // 1: B operand MN-major (rows = K, 64-column blocks 8 KiB apart), as act_fused.cuh / dw_gemm.cuh
  long long* out; // [gridDim.x]
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[8];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  // A tiles: 16 KB each (128 rows x 64 K), B tiles: 32 KB each (256 rows x 64 K)
  uint8_t* sA = smem;
  uint8_t* sB = smem + 4 * 16384;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bar2[i], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, a.N, 0, a.b_mn);
    const uint32_t hi = smem_desc_hi(1024);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      int acc = 0, at = 0, bt = 0;
      for (int it = 0; it < a.iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t alo = smem_desc_lo(smem_u32(sA + at * 16384) + k * 32, 0);
          const uint32_t blo = a.b_mn ? smem_desc_lo(smem_u32(sB + bt * 32768) + k * 2048, 8192)
                                      : smem_desc_lo(smem_u32(sB + bt * 32768) + k * 32, 0);
          umma_bf16_lohi(tmem + acc * a.N, alo, hi, blo, hi, idesc, 1);
          if (++acc == a.nacc) acc = 0;
          if (++at == a.a_tiles) at = 0;
          if (++bt == a.b_tiles) bt = 0;
        }
        if (a.commit_every > 0 && (it % a.commit_every) == 0) umma_commit(&bar2[it & 7]);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one()) a.out[blockIdx.x] = t1 - t0;
    long long tt0 = __shfl_sync(0xffffffffu, t0, 0);
    (void)tt0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* out;
  cudaMalloc(&out, sms * sizeof(long long));
  const size_t smem = 4 * 16384 + 4 * 32768 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const int Ns[] = {16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 224, 240, 256};
  std::printf("# N nacc a_tiles b_tiles cycles_per_mma mac_per_cycle\n");
  for (int cfg = 0; cfg < 8; ++cfg) {
    for (int N : Ns) {
      Args a{};
      a.N = N;
      a.nacc = (cfg == 0) ? 1 : (512 / N);
      if (a.nacc > 8) a.nacc = 8;
      a.a_tiles = (cfg >= 2) ? 4 : 1;
      a.b_tiles = (cfg == 3 || cfg == 5) ? 4 : 1;
      a.b_mn = (cfg == 4 || cfg == 5) ? 1 : 0;
      a.commit_every = cfg == 6 ? 1 : (cfg == 7 ? 4 : 0);
      a.iters = 2000;
      a.out = out;
      for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<<<sms, 128, smem>>>(a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { std::printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[256];
      cudaMemcpy(h, out, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < sms; ++i) mean += static_cast<double>(h[i]);
      mean /= sms;
      const double per = mean / (a.iters * 4.0);
      std::printf("%d %d %d %d%s %.1f %.0f\n", N, a.nacc, a.a_tiles, a.b_tiles, a.b_mn ? " mn" : (a.commit_every ? (a.commit_every == 1 ? " c1" : " c4") : ""), per, 128.0 * N * 16 / per);
    }
  }
  return 0;
}
