// Device helpers shared by the persistent (cooperative, grid-barrier) observe kernels.
#pragma once
#include "dv3_common.cuh"

namespace dv3 {

constexpr int PO_THREADS = 256;
constexpr int PO_WARPS = PO_THREADS / 32;
constexpr int PO_ROWS = 16;
constexpr int PO_CP = 6;                    // output columns per GEMV pass
constexpr int PO_NV = PO_ROWS * PO_CP;      // 96 partial sums per thread

// Grid barrier on a monotonically increasing arrival counter: barrier number n is passed once
// the counter reaches n * nblocks.  One release-atomic per CTA and an acquire spin on the same
// word -- no reset / flag second hop.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks, unsigned& gen) {
  __syncthreads();
  const unsigned target = (gen + 1) * nblocks;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  gen += 1;
  __syncthreads();
}

template <bool GLOBAL_IN>
__device__ __forceinline__ float4 load4(const float* p) {
  if (GLOBAL_IN) return __ldcg(reinterpret_cast<const float4*>(p));
  return *reinterpret_cast<const float4*>(p);
}

// out[m][c] = sum_k Ws[c*K + k] * in[m][k] for m < B <= 16, c < ncols; in = [in1 (K1) | in2];
// K split over all threads, weights from smem, 96 partials folded by butterfly + smem.
template <bool GLOBAL_IN, typename Epi>
__device__ __forceinline__ void gemv16(const float* Ws, int ncols, int K, const float* in1, int ld1,
                                       int K1, const float* in2, int ld2, int B, float* part,
                                       Epi epi) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool once = (K >> 2) <= PO_THREADS;      // one k-quad per thread: load the inputs once
  float4 a[PO_ROWS];
  if (once) {
    const int k = tid << 2;
    const bool live = k < K;
    const float* src = live ? ((k < K1) ? in1 + k : in2 + (k - K1)) : in1;
    const int ld = (k < K1) ? ld1 : ld2;
#pragma unroll
    for (int m = 0; m < PO_ROWS; ++m)
      a[m] = (live && m < B) ? load4<GLOBAL_IN>(src + (size_t)m * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int cb = 0; cb < ncols; cb += PO_CP) {
    float acc[PO_NV];
#pragma unroll
    for (int i = 0; i < PO_NV; ++i) acc[i] = 0.f;
#pragma unroll 1
    for (int q = tid; q < (K >> 2); q += PO_THREADS) {
      const int k = q << 2;
      if (!once) {
        const float* src;
        int ld;
        if (k < K1) { src = in1 + k; ld = ld1; } else { src = in2 + (k - K1); ld = ld2; }
#pragma unroll
        for (int m = 0; m < PO_ROWS; ++m)
          a[m] = (m < B) ? load4<GLOBAL_IN>(src + (size_t)m * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int c = 0; c < PO_CP; ++c) {
        const float4 w = (cb + c < ncols)
                             ? *reinterpret_cast<const float4*>(Ws + (size_t)(cb + c) * K + k)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int m = 0; m < PO_ROWS; ++m) {
          float s = acc[m * PO_CP + c];
          s = fmaf(a[m].x, w.x, s);
          s = fmaf(a[m].y, w.y, s);
          s = fmaf(a[m].z, w.z, s);
          s = fmaf(a[m].w, w.w, s);
          acc[m * PO_CP + c] = s;
        }
      }
    }
    // butterfly fold 96 -> 3 per lane (lane L ends with indices 3L .. 3L+2)
#pragma unroll
    for (int off = 16, nv = PO_NV; off > 0; off >>= 1, nv >>= 1) {
      const bool up = (lane & off) != 0;
      const int half = nv >> 1;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float send = up ? acc[i] : acc[i + half];
        const float keep = up ? acc[i + half] : acc[i];
        acc[i] = keep + __shfl_xor_sync(FULL, send, off);
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) part[warp * PO_NV + lane * 3 + j] = acc[j];
    __syncthreads();
    if (tid < PO_NV) {
      float r = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < PO_WARPS; ++w2) r += part[w2 * PO_NV + tid];
      const int m = tid / PO_CP, c = cb + tid % PO_CP;
      if (m < B && c < ncols) epi(m, c, r);
    }
    __syncthreads();
  }
}

}  // namespace dv3
