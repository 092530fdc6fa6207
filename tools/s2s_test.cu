// Does cp.async.bulk shared::cta -> shared::cluster (own CTA) work in a plain (non-cluster) launch?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace snn;

// 1-D shared -> shared copy by the bulk-copy engine (destination: this CTA's own shared memory, addressed in the
// shared::cluster window), completion signalled on an mbarrier.  (Lived in umma.cuh while the forward kernel staged its
// spike operand through such a copy, r1; nothing in the library uses it since r2.)
__device__ __forceinline__ void bulk_s2s(void* dst_smem, const void* src_smem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "r"(smem_u32(src_smem)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(128, 1) s2s_kernel(uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  uint32_t* src = reinterpret_cast<uint32_t*>(smem);
  uint32_t* dst = reinterpret_cast<uint32_t*>(smem + 32768);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) { src[i] = i * 2654435761u + blockIdx.x; dst[i] = 0; }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 30720u);
    bulk_s2s(dst, src, 30720u, &bar);
  }
  mbar_wait(&bar, 0);
  // timing: one 30 KiB copy vs the same bytes as 8 / 30 smaller copies issued back to back
  __shared__ long long cyc[4];
  if (threadIdx.x == 0) {
    const int splits[3] = {1, 8, 30};
    uint32_t ph = 1;
    for (int v = 0; v < 3; ++v) {
      const int n = splits[v];
      const uint32_t each = 30720u / n;
      long long best = 1 << 30;
      for (int rep = 0; rep < 5; ++rep, ph ^= 1) {
        const long long t0 = clock64();
        mbar_expect_tx(&bar, 30720u);
        for (int i = 0; i < n; ++i) bulk_s2s(reinterpret_cast<uint8_t*>(dst) + i * each, reinterpret_cast<uint8_t*>(src) + i * each, each, &bar);
        mbar_wait(&bar, ph);
        const long long dt = clock64() - t0;
        best = dt < best ? dt : best;
      }
      cyc[v] = best;
    }
    if (blockIdx.x == 0) printf("30 KiB shared->shared: 1 copy %lld cyc, 8 copies %lld cyc, 30 copies %lld cyc\n", cyc[0], cyc[1], cyc[2]);
  }
  __syncthreads();
  uint32_t bad = 0;
  for (int i = threadIdx.x; i < 7680; i += blockDim.x) bad += dst[i] != src[i];
  atomicAdd(out, bad);
}

int main_cluster();
int main(int argc, char**) {
  if (argc > 1) return main_cluster();
  uint32_t* out;
  cudaMalloc(&out, 4);
  cudaMemset(out, 0, 4);
  cudaFuncSetAttribute(s2s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  s2s_kernel<<<148, 128, 65536 + 1024>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  uint32_t h = 123;
  cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
  std::printf("plain launch: %s, mismatches %u\n", cudaGetErrorString(e), h);
  return 0;
}

int main_cluster() {
  uint32_t* out;
  cudaMalloc(&out, 4);
  cudaMemset(out, 0, 4);
  cudaFuncSetAttribute(s2s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 65536 + 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, s2s_kernel, out);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  uint32_t h = 123;
  cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
  std::printf("cluster(1) launch: %s, mismatches %u\n", cudaGetErrorString(e), h);
  return 0;
}
