// fp32 "Linear" kernels:  C = [A1|A2] W^T + bias + addend   (W is PyTorch [out,in] row-major)
//
// Two shapes of problem appear on the path:
//   * M = 16 rows (observe, B=16): a GEMV-like, weight-bandwidth/latency-bound product.
//     -> linear_skinny_kernel: 8 output columns per CTA, K split over all 256 threads, every
//        128-bit weight / activation load of the CTA in flight at once, butterfly + smem fold.
//   * M = 1024+ rows (imagination, bulk backward): a real contraction.
//     -> linear_tiled_kernel: 128x64x16 register-tiled SIMT GEMM (8x4 per thread), register
//        staged double buffering.  Plain fp32 FMA path: used below 64 rows and when K % 4 != 0;
//        everything else goes to the tcgen05 3xTF32 kernels (dv3_umma2.cu / dv3_umma2x.cu).
#include "dv3_common.cuh"

namespace dv3 {


// ------------------------------------------------------------------------------------------
// skinny: M-tile of 16 rows, 8 output columns per CTA, K split over all 256 threads.
// Every thread issues its 8 weight + 16 activation 128-bit loads up front (one L2 round trip for
// the whole CTA), does 16x8x4 FMAs per float4 of K, then the 128 partial sums per thread are
// folded with a shuffle butterfly (124 shuffles) and an 8-way shared-memory add.
// ------------------------------------------------------------------------------------------
constexpr int SK_ROWS = 16;
constexpr int SK_COLS = 8;
constexpr int SK_THREADS = 256;
constexpr int SK_NV = SK_ROWS * SK_COLS;   // 128 accumulators per thread

__global__ void __launch_bounds__(SK_THREADS)
linear_skinny_kernel(LinearArgs g) {
  __shared__ float part[SK_THREADS / 32][SK_NV];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.y * SK_ROWS, n0 = blockIdx.x * SK_COLS;
  const int K1 = g.A[0] ? g.K[0] : 0, K2 = g.A[1] ? g.K[1] : 0;
  const int Q1 = K1 >> 2, Q = (K1 + K2) >> 2;

  float acc[SK_NV];
#pragma unroll
  for (int i = 0; i < SK_NV; ++i) acc[i] = 0.f;

#pragma unroll 1
  for (int q = tid; q < Q; q += SK_THREADS) {
    const int seg = q >= Q1;
    const int k = (seg ? q - Q1 : q) << 2;
    const float* __restrict__ A = g.A[seg] + k;
    const float* __restrict__ W = g.W[seg] + k;
    const int lda = g.lda[seg], ldw = g.ldw[seg];
    float4 w[SK_COLS], a[SK_ROWS];
#pragma unroll
    for (int c = 0; c < SK_COLS; ++c)
      w[c] = (n0 + c < g.N) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + c) * ldw))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < SK_ROWS; ++m)
      a[m] = (m0 + m < g.M) ? *reinterpret_cast<const float4*>(A + (size_t)(m0 + m) * lda)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < SK_ROWS; ++m)
#pragma unroll
      for (int c = 0; c < SK_COLS; ++c) {
        float s = acc[m * SK_COLS + c];
        s = fmaf(a[m].x, w[c].x, s);
        s = fmaf(a[m].y, w[c].y, s);
        s = fmaf(a[m].z, w[c].z, s);
        s = fmaf(a[m].w, w[c].w, s);
        acc[m * SK_COLS + c] = s;
      }
  }
  // butterfly fold: 128 values x 32 lanes -> lane L ends with values 4L .. 4L+3
#pragma unroll
  for (int off = 16, nv = SK_NV; off > 0; off >>= 1, nv >>= 1) {
    const bool up = (lane & off) != 0;
    const int half = nv >> 1;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? acc[i] : acc[i + half];
      const float keep = up ? acc[i + half] : acc[i];
      acc[i] = keep + __shfl_xor_sync(FULL, send, off);
    }
  }
  // after the fold lane L holds, in acc[0..3], the values whose index has bit pattern
  // (lane bit4, bit3, bit2, bit1, bit0) as its top five bits: index = L*4 + j
#pragma unroll
  for (int j = 0; j < 4; ++j) part[warp][lane * 4 + j] = acc[j];
  __syncthreads();
  if (tid < SK_NV) {
    float r = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < SK_THREADS / 32; ++w2) r += part[w2][tid];
    const int m = m0 + tid / SK_COLS, n = n0 + tid % SK_COLS;
    if (m < g.M && n < g.N) {
      if (g.bias) r += g.bias[n];
      if (g.addend) r += g.addend[(size_t)m * g.ldadd + n];
      float* c = g.C + (size_t)m * g.ldc + n;
      if (g.accumulate) r += *c;
      *c = r;
    }
  }
}

// ------------------------------------------------------------------------------------------
// tiled: 128 x 64 x 16, 256 threads, 8x4 micro-tile
// ------------------------------------------------------------------------------------------
constexpr int TB_M = 128, TB_N = 64, TB_K = 16, TB_PAD = 4;

template <bool VEC>
__device__ __forceinline__ float4 load4_guard(const float* base, int row, int nrows, int ld, int k,
                                              int K) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < nrows) {
    const float* p = base + (size_t)row * ld + k;
    if (VEC) {
      if (k < K) v = *reinterpret_cast<const float4*>(p);
    } else {
      if (k + 0 < K) v.x = p[0];
      if (k + 1 < K) v.y = p[1];
      if (k + 2 < K) v.z = p[2];
      if (k + 3 < K) v.w = p[3];
    }
  }
  return v;
}

template <bool VEC>
__global__ void __launch_bounds__(256)
linear_tiled_kernel(LinearArgs g) {
  __shared__ __align__(16) float As[2][TB_K][TB_M + TB_PAD];
  __shared__ __align__(16) float Bs[2][TB_K][TB_N + TB_PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TB_M, n0 = blockIdx.x * TB_N;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;

  const int nk0 = (g.A[0] ? (g.K[0] + TB_K - 1) / TB_K : 0);
  const int nk1 = (g.A[1] ? (g.K[1] + TB_K - 1) / TB_K : 0);
  const int nk = nk0 + nk1;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra0, ra1, rb;
  auto gload = [&](int t) {
    const int seg = (t < nk0) ? 0 : 1;
    const int k = ((t < nk0) ? t : t - nk0) * TB_K + lk;
    ra0 = load4_guard<VEC>(g.A[seg], m0 + lrow, g.M, g.lda[seg], k, g.K[seg]);
    ra1 = load4_guard<VEC>(g.A[seg], m0 + lrow + 64, g.M, g.lda[seg], k, g.K[seg]);
    rb = load4_guard<VEC>(g.W[seg], n0 + lrow, g.N, g.ldw[seg], k, g.K[seg]);
  };
  auto sstore = [&](int buf) {
    As[buf][lk + 0][lrow] = ra0.x; As[buf][lk + 1][lrow] = ra0.y;
    As[buf][lk + 2][lrow] = ra0.z; As[buf][lk + 3][lrow] = ra0.w;
    As[buf][lk + 0][lrow + 64] = ra1.x; As[buf][lk + 1][lrow + 64] = ra1.y;
    As[buf][lk + 2][lrow + 64] = ra1.z; As[buf][lk + 3][lrow + 64] = ra1.w;
    Bs[buf][lk + 0][lrow] = rb.x; Bs[buf][lk + 1][lrow] = rb.y;
    Bs[buf][lk + 2][lrow] = rb.z; Bs[buf][lk + 3][lrow] = rb.w;
  };

  if (nk > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int t = 0; t < nk; ++t) {
    const int buf = t & 1;
    if (t + 1 < nk) gload(t + 1);
#pragma unroll
    for (int k = 0; k < TB_K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float r = acc[i][j];
      if (g.bias) r += g.bias[n];
      if (g.addend) r += g.addend[(size_t)m * g.ldadd + n];
      float* c = g.C + (size_t)m * g.ldc + n;
      if (g.accumulate) r += *c;
      *c = r;
    }
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int launch_linear(const LinearArgs& g, cudaStream_t st) {
  DV3_REQUIRE(g.M > 0 && g.N > 0, DV3_ERR_BAD_SHAPE, "linear: M=%d N=%d", g.M, g.N);
  DV3_REQUIRE(g.A[0] && g.W[0] && g.C, DV3_ERR_NULL, "linear: null operand");
  bool vec = true;
  for (int s = 0; s < 2; ++s) {
    if (!g.A[s]) continue;
    vec = vec && (g.K[s] % 4 == 0) && (g.lda[s] % 4 == 0) && (g.ldw[s] % 4 == 0) &&
          aligned16(g.A[s]) && aligned16(g.W[s]);
  }
  const bool prof = prof_on();
  const double flops = 2.0 * g.M * g.N * ((g.A[0] ? g.K[0] : 0) + (g.A[1] ? g.K[1] : 0));
  if (g.M <= 2 * SK_ROWS && vec) {
    dim3 grid((g.N + SK_COLS - 1) / SK_COLS, (g.M + SK_ROWS - 1) / SK_ROWS);
    if (prof) prof_begin(st);
    linear_skinny_kernel<<<grid, SK_THREADS, 0, st>>>(g);
    if (prof) prof_end(st, 0, flops);
    DV3_CHECK_LAUNCH("linear_skinny_kernel");
    return 0;
  }
  dim3 grid((g.N + TB_N - 1) / TB_N, (g.M + TB_M - 1) / TB_M);
  if (prof) prof_begin(st);
  if (vec)
    linear_tiled_kernel<true><<<grid, 256, 0, st>>>(g);
  else
    linear_tiled_kernel<false><<<grid, 256, 0, st>>>(g);
  if (prof) prof_end(st, 1, flops);
  DV3_CHECK_LAUNCH("linear_tiled_kernel");
  return 0;
}

// convenience: C = A W^T (+bias) with a single K segment
int linear1(const float* A, int lda, const float* W, int ldw, int K, const float* bias, float* C,
            int ldc, int M, int N, int accumulate, cudaStream_t st) {
  LinearArgs g{};
  g.A[0] = A; g.lda[0] = lda; g.W[0] = W; g.ldw[0] = ldw; g.K[0] = K;
  g.bias = bias; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.accumulate = accumulate;
  return launch_linear(g, st);
}

// ------------------------------------------------------------------------------------------
// transpose (weights are transposed once per backward call so that every in-loop product is
// the same K-contiguous "NT" form)
// ------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ in, int ld, int R, int C,
                                 float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = in[(size_t)r * ld + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

int launch_transpose(const float* in, int ld, int R, int C, float* out, cudaStream_t st) {
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(in, ld, R, C, out);
  DV3_CHECK_LAUNCH("transpose_kernel");
  return 0;
}

}  // namespace dv3

extern "C" int dv3_linear_fwd(const float* A1, int32_t lda1, const float* W1, int32_t ldw1,
                              int32_t K1, const float* A2, int32_t lda2, const float* W2,
                              int32_t ldw2, int32_t K2, const float* bias, const float* addend,
                              int32_t ldadd, float* C, int32_t ldc, int32_t M, int32_t N,
                              int32_t accumulate, void* stream) {
  dv3::LinearArgs g{};
  g.A[0] = A1; g.lda[0] = lda1; g.W[0] = W1; g.ldw[0] = ldw1; g.K[0] = K1;
  g.A[1] = A2; g.lda[1] = lda2; g.W[1] = W2; g.ldw[1] = ldw2; g.K[1] = A2 ? K2 : 0;
  g.bias = bias; g.addend = addend; g.ldadd = ldadd; g.C = C; g.ldc = ldc; g.M = M; g.N = N;
  g.accumulate = accumulate;
  return dv3::launch_linear(g, static_cast<cudaStream_t>(stream));
}

extern "C" int dv3_transpose(const float* in, int32_t ld, int32_t R, int32_t C, float* out,
                             void* stream) {
  DV3_REQUIRE(in && out, DV3_ERR_NULL, "transpose: null pointer");
  return dv3::launch_transpose(in, ld, R, C, out, static_cast<cudaStream_t>(stream));
}
