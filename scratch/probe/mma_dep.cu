// probe: cost per tcgen05.mma (kind::tf32, M=128, K=8) as a function of N and of whether consecutive
// MMAs accumulate into the SAME TMEM columns (dependent chain) or rotate over NACC accumulators.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int N, int NACC, bool ATMEM, int COMMITS = 0>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[4];
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 1.0f;
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    for (int b = 0; b < 4; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&dummy[b])), "r"(1 << 20) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = slot;
  if (t == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t au = (smem_u32(sm) >> 4) & 0x3FFF, bu = au + (16384 >> 4);
    const uint64_t hi = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k8 = i & 3;
      const uint64_t adesc = hi | (uint64_t)((au + k8 * 2) | (1u << 16));
      const uint64_t bdesc = hi | (uint64_t)((bu + k8 * 2) | (1u << 16));
      const uint32_t acc = tb + (i % NACC) * N;
      if (ATMEM) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(acc), "r"(tb + 448 + k8 * 8), "l"(bdesc), "r"(idesc), "r"(1) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(acc), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1) : "memory");
      }
      if (COMMITS > 0 && (i & 7) == 7) {
#pragma unroll
        for (int c = 0; c < COMMITS; ++c)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[c])) : "memory");
      }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
template <int N, int NACC, bool ATMEM, int COMMITS = 0>
void run(long long* d) {
  auto k = probe<N, NACC, ATMEM, COMMITS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 512;
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    k<<<1, 128, 64 * 1024>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d NACC=%d: %s\n", N, NACC, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("commits/8=%d N=%3d accumulators=%d A=%s: issue %.1f clk/MMA, complete %.1f clk/MMA\n", COMMITS, N, NACC, ATMEM ? "tmem" : "smem",
         (double)h[0] / iters, (double)h[1] / iters);
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<32, 1, true, 0>(d); run<32, 1, true, 1>(d); run<32, 1, true, 2>(d); run<32, 1, true, 3>(d);
  run<64, 1, true, 0>(d); run<64, 1, true, 1>(d); run<64, 1, true, 3>(d);
  run<64, 1, false, 0>(d); run<64, 1, false, 1>(d); run<64, 1, false, 2>(d);
  return 0;
}
