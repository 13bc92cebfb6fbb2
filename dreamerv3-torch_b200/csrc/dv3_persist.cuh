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
// K split over all threads (k-quads of 4), weights from smem, the 16*CP partial sums of a column
// pass folded by a shuffle butterfly + shared memory.
//   CP: columns per pass (6 -> 96 accumulators; 4 when few columns are owned)
//   NQ: k-quads a thread keeps in registers; when K/4 <= NQ*256 the inputs are loaded from L2
//       exactly once (all loads in flight together) and reused by every column pass -- otherwise
//       they are re-read per pass and per quad, one L2 round trip each.
template <bool GLOBAL_IN, int CP, int NQ, typename Epi>
__device__ __forceinline__ void gemv16(const float* Ws, int ncols, int K, const float* in1, int ld1,
                                       int K1, const float* in2, int ld2, int B, float* part,
                                       Epi epi) {
  constexpr int NV = PO_ROWS * CP, PL = NV / 32;     // partials per thread / per lane after the fold
  static_assert(NV % 32 == 0 && NV <= PO_NV, "gemv16: CP must be 2, 4 or 6");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nq = K >> 2;
  const bool once = nq <= NQ * PO_THREADS;
  float4 a[NQ][PO_ROWS];
  auto load_quad = [&](int q, float4 (&dst)[PO_ROWS]) {
    const int k = q << 2;
    const bool live = q < nq;
    const float* src = live ? ((k < K1) ? in1 + k : in2 + (k - K1)) : in1;
    const int ld = (k < K1) ? ld1 : ld2;
#pragma unroll
    for (int m = 0; m < PO_ROWS; ++m)
      dst[m] = (live && m < B) ? load4<GLOBAL_IN>(src + (size_t)m * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (once) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) load_quad(tid + j * PO_THREADS, a[j]);
  }
  for (int cb = 0; cb < ncols; cb += CP) {
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    auto fma_quad = [&](int q, const float4 (&av)[PO_ROWS]) {
      const int k = q << 2;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const float4 w = (cb + c < ncols && q < nq)
                             ? *reinterpret_cast<const float4*>(Ws + (size_t)(cb + c) * K + k)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int m = 0; m < PO_ROWS; ++m) {
          float s = acc[m * CP + c];
          s = fmaf(av[m].x, w.x, s);
          s = fmaf(av[m].y, w.y, s);
          s = fmaf(av[m].z, w.z, s);
          s = fmaf(av[m].w, w.w, s);
          acc[m * CP + c] = s;
        }
      }
    };
    if (once) {
#pragma unroll
      for (int j = 0; j < NQ; ++j)
        if (tid + j * PO_THREADS < nq) fma_quad(tid + j * PO_THREADS, a[j]);
    } else {
#pragma unroll 1
      for (int q = tid; q < nq; q += PO_THREADS) {
        load_quad(q, a[0]);
        fma_quad(q, a[0]);
      }
    }
    // butterfly fold NV -> PL per lane (lane L ends with indices PL*L .. PL*L + PL-1)
#pragma unroll
    for (int off = 16, nv = NV; off > 0; off >>= 1, nv >>= 1) {
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
    for (int j = 0; j < PL; ++j) part[warp * NV + lane * PL + j] = acc[j];
    __syncthreads();
    if (tid < NV) {
      float r = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < PO_WARPS; ++w2) r += part[w2 * NV + tid];
      const int m = tid / CP, c = cb + tid % CP;
      if (m < B && c < ncols) epi(m, c, r);
    }
    __syncthreads();
  }
}

}  // namespace dv3
