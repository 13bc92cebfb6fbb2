// probe: can 8 clusters of 16 CTAs (384 threads, ~216 KB smem) be co-resident on a B200?
// also times barrier.cluster and a st.async + mbarrier all-to-all exchange inside a 16-CTA cluster
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok != 0;
}
template <int CS>
__global__ void __launch_bounds__(384, 1) probe(int* out, unsigned long long* tim, int iters) {
  extern __shared__ __align__(1024) uint8_t sm[];
  float2* stat = reinterpret_cast<float2*>(sm);              // [2][CS][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * CS * 128 * 8);
  const uint32_t rank = ctarank();
  const int t = threadIdx.x;
  if (t == 0) {
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + b)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    out[blockIdx.x * 2] = (int)smid(); out[blockIdx.x * 2 + 1] = (int)rank;
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  // (1) plain cluster barriers
  unsigned long long t0 = gt();
  for (int i = 0; i < iters; ++i) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  unsigned long long t1 = gt();
  // (2) st.async exchange: threads 128..255 (row = t-128) push a float2 to every peer, all of
  // threads 128..383 wait on the local mbarrier
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    const int b = i & 1;
    const uint32_t bar = smem_u32(bars + b);
    if (t == 128) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(CS * 128 * 8) : "memory");
    if (t >= 128 && t < 256) {
      const int row = t - 128;
      const uint32_t dst = smem_u32(&stat[(b * CS + rank) * 128 + row]);
      const float vx = (float)(i + row), vy = (float)rank;
#pragma unroll
      for (int r = 0; r < CS; ++r) {
        uint32_t pd, pb;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(pd) : "r"(dst), "r"(r));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(pb) : "r"(bar), "r"(r));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(pd), "f"(vx), "f"(vy), "r"(pb) : "memory");
      }
    }
    if (t >= 128) {
      while (!try_wait(bar, (i >> 1) & 1)) {}
      const int row = (t - 128) & 127;
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < CS; ++r) { const float2 v = stat[(b * CS + r) * 128 + row]; s += v.x + v.y; }
      // check: sum_r (i+row + r) = CS*(i+row) + CS(CS-1)/2
      if (s != (float)(CS * (i + row) + CS * (CS - 1) / 2)) acc += 1.f;
    }
  }
  unsigned long long t2 = gt();
  if (t == 128) { tim[blockIdx.x * 3] = t1 - t0; tim[blockIdx.x * 3 + 1] = t2 - t1; }
  if (acc != 0.f) atomicAdd(out + 4096, 1);
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CS>
int run(int smem_kb) {
  auto k = probe<CS>;
  size_t smem = (size_t)smem_kb * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (CS > 8) cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(8 * CS); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int m = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&m, k, &cfg);
  printf("CS=%d smem=%dKB maxActiveClusters=%d (%s)\n", CS, smem_kb, m, cudaGetErrorString(e));
  int* out; unsigned long long* tim;
  cudaMalloc(&out, 4097 * 4 * 2); cudaMemset(out, 0, 4097 * 4 * 2);
  cudaMalloc(&tim, 8 * CS * 3 * 8); cudaMemset(tim, 0, 8 * CS * 3 * 8);
  const int iters = 200;
  for (int rep = 0; rep < 2; ++rep) {
    e = cudaLaunchKernelEx(&cfg, k, out, tim, iters);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sync failed: %s\n", cudaGetErrorString(e)); return 1; }
  }
  static int h[4097 * 2]; static unsigned long long ht[16 * 8 * 3];
  cudaMemcpy(h, out, sizeof(int) * 4097 * 2, cudaMemcpyDeviceToHost);
  cudaMemcpy(ht, tim, 8 * CS * 3 * 8, cudaMemcpyDeviceToHost);
  printf("  mismatches=%d\n", h[4096]);
  for (int c = 0; c < 8; ++c) {
    printf("  cluster %d: sms", c);
    for (int r = 0; r < CS; ++r) printf(" %d", h[(c * CS + r) * 2]);
    printf(" | barrier %.0f ns, st.async exchange %.0f ns per iter\n", (double)ht[c * CS * 3] / iters, (double)ht[c * CS * 3 + 1] / iters);
  }
  return 0;
}
int main() {
  run<16>(216); run<16>(100); run<8>(216);
  return 0;
}
