// probe: tcgen05.mma kind::tf32 with the A operand in TMEM (written by tcgen05.st, lane = row,
// column = k) and B in shared memory (K-major, 128B swizzle written by hand).  D = A B^T, M=128,N=32,K=32.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cmath>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D) {
  __shared__ __align__(1024) float sB[32 * 32];     // 32 rows (n) x 32 k, 128 B per row, swizzled
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int i = t; i < 32 * 8; i += 128) {           // 16-byte chunks
    const int r = i >> 3, c = i & 7;
    const float4 v = *reinterpret_cast<const float4*>(B + r * 32 + 4 * c);
    *reinterpret_cast<float4*>(sB + r * 32 + 4 * (c ^ (r & 7))) = v;
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = slot;
  // A row (32 k) of this thread's row -> TMEM columns [64, 96) of lane = row
  uint32_t a[32];
  const int row = warp * 32 + lane;
  for (int k = 0; k < 32; ++k) a[k] = __float_as_uint(A[row * 32 + k]);
  const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + 64;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]),
        "r"(a[16]), "r"(a[17]), "r"(a[18]), "r"(a[19]), "r"(a[20]), "r"(a[21]), "r"(a[22]), "r"(a[23]),
        "r"(a[24]), "r"(a[25]), "r"(a[26]), "r"(a[27]), "r"(a[28]), "r"(a[29]), "r"(a[30]), "r"(a[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (t == 0) {
    // c=F32 (bit 4), a=b=TF32 (2<<7, 2<<10), N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t bu = (smem_u32(sB) >> 4) & 0x3FFF;
    for (int k8 = 0; k8 < 4; ++k8) {
      const uint64_t bdesc = ((uint64_t)(64u | (1u << 14) | (2u << 29)) << 32) | (uint64_t)((bu + k8 * 2) | (1u << 16));
      const uint32_t acc = k8 != 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
          ::"r"(tb), "r"(tb + 64 + k8 * 8), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(tb + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 32; ++n) D[row * 32 + n] = __uint_as_float(v[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tb) : "memory");
}
int main() {
  static float hA[128 * 32], hB[32 * 32], hD[128 * 32];
  for (int i = 0; i < 128 * 32; ++i) hA[i] = (float)((i * 7 + 3) % 17 - 8);
  for (int i = 0; i < 32 * 32; ++i) hB[i] = (float)((i * 5 + 1) % 13 - 6);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  probe<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      float r = 0.f;
      for (int k = 0; k < 32; ++k) r += hA[m * 32 + k] * hB[n * 32 + k];
      if (r != hD[m * 32 + n]) { if (bad < 8) printf("mismatch m=%d n=%d got %g want %g\n", m, n, hD[m * 32 + n], r); ++bad; }
    }
  printf("mismatches: %d of %d\n", bad, 128 * 32);
  return 0;
}
