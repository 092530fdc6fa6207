// Microbenchmark: aggregate L2 -> shared-memory bandwidth of 1-D bulk copies (cp.async.bulk) issued by every SM
// against an L2-resident buffer.  Sets the ceiling for operand staging in the scan-GEMMs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fullysnnforleggedrobots_b200/csrc tools/l2_bulk_rate.cu -o build/l2_bulk_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace snn;

#ifndef KSLOTS
#define KSLOTS 3
#define KWARPS 4
#endif
constexpr int kSlots = KSLOTS;      // per issuing warp
constexpr int kWarps = KWARPS;

__global__ void __launch_bounds__(32 * kWarps, 1) l2_rate_kernel(const uint8_t* src, size_t region, int chunk, int iters, int shared_stream,
                                                       long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_all[kSlots * kWarps];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots * kWarps; ++s) mbar_init(&full_all[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    uint64_t* full = full_all + w * kSlots;
    smem += w * kSlots * 16384;
    // shared_stream = 1: every CTA walks the SAME addresses (best case for L2 / multicast-like reuse);
    // 0: CTA-private offsets spread over the whole region
    const size_t nchunks = region / chunk;
    size_t pos = (shared_stream ? static_cast<size_t>(w) * 131 : (static_cast<size_t>(blockIdx.x) * 977 + w * 131)) % nchunks;
    const long long t0 = clock64();
    for (int i = 0; i < iters + kSlots; ++i) {
      const int s = i % kSlots;
      if (i >= kSlots) mbar_wait(&full[s], ((i / kSlots) - 1) & 1);
      if (i < iters) {
        mbar_expect_tx(&full[s], static_cast<uint32_t>(chunk));
        bulk_g2s(smem + s * chunk, src + pos * chunk, static_cast<uint32_t>(chunk), &full[s]);
        pos = pos + 1 == nchunks ? 0 : pos + 1;
      }
    }
    if (w == 0) out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const size_t region = 48u << 20;          // 48 MiB: L2-resident on both dies' slices
  uint8_t* src;
  cudaMalloc(&src, region);
  cudaMemset(src, 1, region);
  long long* out;
  cudaMalloc(&out, 1024 * sizeof(long long));
  cudaFuncSetAttribute(l2_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * kSlots * 16384 + 1024);
  std::printf("# chunk_bytes shared_stream ctas  B/cycle/SM  TB/s(at %d MHz)  event_TB/s\n", khz / 1000);
  const int chunks[] = {2048, 4096, 8192, 16384};
  for (int shared_stream = 0; shared_stream < 2; ++shared_stream)
    for (int chunk : chunks) {
      const int iters = 4096;
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      l2_rate_kernel<<<sms, 32 * kWarps, kWarps * kSlots * 16384 + 1024>>>(src, region, chunk, iters, shared_stream, out);   // warm L2
      cudaEventRecord(a);
      l2_rate_kernel<<<sms, 32 * kWarps, kWarps * kSlots * 16384 + 1024>>>(src, region, chunk, iters, shared_stream, out);
      cudaEventRecord(b);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { std::printf("error %s\n", cudaGetErrorString(e)); return 1; }
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      long long h[1024];
      cudaMemcpy(h, out, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < sms; ++i) mean += static_cast<double>(h[i]);
      mean /= sms;
      const double bpc = static_cast<double>(iters) * chunk * kWarps / mean;
      std::printf("%d %d %d %.1f %.2f %.2f\n", chunk, shared_stream, sms, bpc, bpc * sms * khz * 1e3 / 1e12,
                  static_cast<double>(iters) * chunk * kWarps * sms / (ms * 1e-3) / 1e12);
    }
  return 0;
}
