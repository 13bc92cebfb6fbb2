// Row-wise kernels of the RSSM step: LayerNorm+SiLU, the LayerNorm-GRU gate block, the one-hot
// gather that replaces "one_hot(stoch) @ W", and the unimix categorical draw with its
// straight-through backward.  One CTA per row (LayerNorm needs the whole row), one warp per
// categorical group (C <= 32 classes: one class per lane, reductions are shuffles).
//
// Reference semantics (paths relative to the reference tree):
//   Linear->LayerNorm(eps 1e-3)->SiLU blocks      networks.py:48-78, 623-632
//   GRUCell.forward                                networks.py:760-768
//   OneHotDist.__init__/sample/mode                tools.py:436-460 (+ torch Categorical)
#include <cstdlib>
#include "dv3_common.cuh"

namespace dv3 {

constexpr int ROW_THREADS = 256;

// optional tf32 hi/lo planes of a kernel's output (the A operand of the next tensor-core GEMM):
// hi = x with the 13 low mantissa bits cleared, lo = x - hi.  Written next to the fp32 output so
// no separate split pass (and its extra read of the tensor) is needed.
__device__ __forceinline__ void put_split(const SplitOut& so, size_t row, int col, float x) {
  if (so.hi) {
    const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    so.hi[row * so.ld + col] = h;
    so.lo[row * so.ld + col] = x - h;
  }
}

// mean and 1/sqrt(var+eps) of a row (two-pass, biased variance, like ATen's layer_norm)
__device__ __forceinline__ void row_stats(const float* __restrict__ row, int n, float eps,
                                          float* red, float& mean, float& rstd) {
  float s[1] = {0.f};
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[0] += row[i];
  block_sum<1>(s, red);
  mean = s[0] / (float)n;
  float v[1] = {0.f};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = row[i] - mean;
    v[0] = fmaf(d, d, v[0]);
  }
  block_sum<1>(v, red);
  rstd = 1.f / sqrtf(v[0] / (float)n + eps);
}

// ------------------------------------------------------------------------------------------
// SiLU(LayerNorm(pre))
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS)
ln_silu_fwd_kernel(const float* __restrict__ pre, int ld, const float* __restrict__ g,
                   const float* __restrict__ b, float eps, int n, float* __restrict__ out,
                   int ldo, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[4 * 32];
  const float* row = pre + (size_t)blockIdx.x * ld;
  float mean, rstd;
  row_stats(row, n, eps, red, mean, rstd);
  float* o = out + (size_t)blockIdx.x * ldo;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float y = siluf_(fmaf((row[i] - mean) * rstd, g[i], b[i]));
    o[i] = y;
    put_split(so, blockIdx.x, i, y);
  }
}

// ---- bulk rows (M >= bulk_rows()): one warp per row, the row lives in registers ---------------
// The block-per-row kernels above are latency-bound (three passes over the row, two block
// reductions): 35 us for 15360 x 512 where a copy takes 7.  With thousands of rows a warp per row
// fills the machine; each lane holds NV float4 of the row, reads it once, reduces with shuffles and
// writes 16-byte vectors.  Needs n % 4 == 0, n <= 128 NV and 16-byte aligned rows.
constexpr int BULK_WARPS = 8;
static int bulk_rows() {               // row count from which the warp-per-row kernels take over
  const char* e = DV3_ENV("DV3_BULK_ROWS");
  const int v = e ? atoi(e) : 4096;
  return v < 1 ? 1 : v;
}

__device__ __forceinline__ float4 split_hi4(float4 v) {
  return make_float4(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u),
                     __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                     __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u),
                     __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
}
__device__ __forceinline__ void put_split4(const SplitOut& so, size_t row, int col, float4 v) {
  if (so.hi) {
    const float4 h = split_hi4(v);
    *reinterpret_cast<float4*>(so.hi + row * so.ld + col) = h;
    *reinterpret_cast<float4*>(so.lo + row * so.ld + col) =
        make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  }
}

// mean and 1/sqrt(var+eps) of a register-resident row (two-pass, biased variance)
template <int NV>
__device__ __forceinline__ void warp_row_stats(const float4 (&x)[NV], const bool (&ok)[NV], int n,
                                               float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) s += (x[j].x + x[j].y) + (x[j].z + x[j].w);
  mean = warp_sum(s) / (float)n;
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (ok[j]) {
      const float a = x[j].x - mean, b = x[j].y - mean, c = x[j].z - mean, d = x[j].w - mean;
      v = fmaf(a, a, v); v = fmaf(b, b, v); v = fmaf(c, c, v); v = fmaf(d, d, v);
    }
  }
  rstd = 1.f / sqrtf(warp_sum(v) / (float)n + eps);
}

template <int NV>
__global__ void __launch_bounds__(BULK_WARPS * 32)
ln_silu_fwd_bulk_kernel(const float* __restrict__ pre, int ld, const float* __restrict__ g,
                        const float* __restrict__ b, float eps, int M, int n,
                        float* __restrict__ out, int ldo, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * BULK_WARPS + (threadIdx.x >> 5);
  if (r >= M) return;
  const float* row = pre + (size_t)r * ld;
  float4 x[NV];
  bool ok[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = 4 * (j * 32 + lane);
    ok[j] = c < n;
    x[j] = ok[j] ? *reinterpret_cast<const float4*>(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float mean, rstd;
  warp_row_stats<NV>(x, ok, n, eps, mean, rstd);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (ok[j]) {
      const int c = 4 * (j * 32 + lane);
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
      float4 y;
      y.x = siluf_(fmaf((x[j].x - mean) * rstd, gg.x, bb.x));
      y.y = siluf_(fmaf((x[j].y - mean) * rstd, gg.y, bb.y));
      y.z = siluf_(fmaf((x[j].z - mean) * rstd, gg.z, bb.z));
      y.w = siluf_(fmaf((x[j].w - mean) * rstd, gg.w, bb.w));
      *reinterpret_cast<float4*>(out + (size_t)r * ldo + c) = y;
      put_split4(so, r, c, y);
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(BULK_WARPS * 32)
ln_silu_bwd_bulk_kernel(const float* __restrict__ pre, int ld, const float* __restrict__ g,
                        const float* __restrict__ b, float eps, const float* __restrict__ d_out,
                        int ldd, int M, int n, float* __restrict__ d_pre, int ldp,
                        float* __restrict__ d_ln, int ldl, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * BULK_WARPS + (threadIdx.x >> 5);
  if (r >= M) return;
  const float* row = pre + (size_t)r * ld;
  const float* dor = d_out + (size_t)r * ldd;
  float4 x[NV], dx[NV];
  bool ok[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = 4 * (j * 32 + lane);
    ok[j] = c < n;
    x[j] = ok[j] ? *reinterpret_cast<const float4*>(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    dx[j] = ok[j] ? *reinterpret_cast<const float4*>(dor + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float mean, rstd;
  warp_row_stats<NV>(x, ok, n, eps, mean, rstd);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (ok[j]) {
      const int c = 4 * (j * 32 + lane);
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
      // x <- xhat, dx <- d_ln * gamma
      float4 dv;
      x[j].x = (x[j].x - mean) * rstd; dv.x = dx[j].x * silu_grad(fmaf(x[j].x, gg.x, bb.x));
      x[j].y = (x[j].y - mean) * rstd; dv.y = dx[j].y * silu_grad(fmaf(x[j].y, gg.y, bb.y));
      x[j].z = (x[j].z - mean) * rstd; dv.z = dx[j].z * silu_grad(fmaf(x[j].z, gg.z, bb.z));
      x[j].w = (x[j].w - mean) * rstd; dv.w = dx[j].w * silu_grad(fmaf(x[j].w, gg.w, bb.w));
      if (d_ln) *reinterpret_cast<float4*>(d_ln + (size_t)r * ldl + c) = dv;
      dx[j] = make_float4(dv.x * gg.x, dv.y * gg.y, dv.z * gg.z, dv.w * gg.w);
      a0 += (dx[j].x + dx[j].y) + (dx[j].z + dx[j].w);
      a1 = fmaf(dx[j].x, x[j].x, a1); a1 = fmaf(dx[j].y, x[j].y, a1);
      a1 = fmaf(dx[j].z, x[j].z, a1); a1 = fmaf(dx[j].w, x[j].w, a1);
    }
  }
  const float m1 = warp_sum(a0) / (float)n, m2 = warp_sum(a1) / (float)n;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (ok[j]) {
      const int c = 4 * (j * 32 + lane);
      float4 dp;
      dp.x = rstd * (dx[j].x - m1 - x[j].x * m2);
      dp.y = rstd * (dx[j].y - m1 - x[j].y * m2);
      dp.z = rstd * (dx[j].z - m1 - x[j].z * m2);
      dp.w = rstd * (dx[j].w - m1 - x[j].w * m2);
      *reinterpret_cast<float4*>(d_pre + (size_t)r * ldp + c) = dp;
      put_split4(so, r, c, dp);
    }
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int ln_silu_fwd(const float* pre, int ld, const float* g, const float* b, float eps, int M, int n,
                float* out, int ldo, cudaStream_t st, SplitOut so) {
  if (M <= 0) return 0;
  if (M >= bulk_rows() && n % 4 == 0 && n <= 1024 && ld % 4 == 0 && ldo % 4 == 0 && al16(pre) &&
      al16(out) && al16(g) && al16(b) && (!so.hi || (so.ld % 4 == 0 && al16(so.hi) && al16(so.lo)))) {
    const dim3 grid((M + BULK_WARPS - 1) / BULK_WARPS), block(BULK_WARPS * 32);
    if (n <= 512)
      DV3_CHECK_CUDA(launch_pdl(ln_silu_fwd_bulk_kernel<4>, grid, block, 0, st, pre, ld, g, b, eps, M,
                                n, out, ldo, so));
    else
      DV3_CHECK_CUDA(launch_pdl(ln_silu_fwd_bulk_kernel<8>, grid, block, 0, st, pre, ld, g, b, eps, M,
                                n, out, ldo, so));
    DV3_CHECK_LAUNCH("ln_silu_fwd_bulk_kernel");
    return 0;
  }
  DV3_CHECK_CUDA(launch_pdl(ln_silu_fwd_kernel, dim3(M), dim3(ROW_THREADS), 0, st, pre, ld, g, b, eps,
                            n, out, ldo, so));
  DV3_CHECK_LAUNCH("ln_silu_fwd_kernel");
  return 0;
}

// d_ln  = d_out * silu'(v)                 (gradient w.r.t. the LayerNorm affine output v)
// d_pre = rstd * (dx - mean(dx) - xhat * mean(dx*xhat)),  dx = d_ln * gamma
__global__ void __launch_bounds__(ROW_THREADS)
ln_silu_bwd_kernel(const float* __restrict__ pre, int ld, const float* __restrict__ g,
                   const float* __restrict__ b, float eps, const float* __restrict__ d_out,
                   int ldd, int n, float* __restrict__ d_pre, int ldp, float* __restrict__ d_ln,
                   int ldl, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[4 * 32];
  const float* row = pre + (size_t)blockIdx.x * ld;
  const float* dor = d_out + (size_t)blockIdx.x * ldd;
  float mean, rstd;
  row_stats(row, n, eps, red, mean, rstd);
  float acc[2] = {0.f, 0.f};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float xh = (row[i] - mean) * rstd;
    const float v = fmaf(xh, g[i], b[i]);
    const float dv = dor[i] * silu_grad(v);
    const float dx = dv * g[i];
    acc[0] += dx;
    acc[1] = fmaf(dx, xh, acc[1]);
    if (d_ln) d_ln[(size_t)blockIdx.x * ldl + i] = dv;
  }
  block_sum<2>(acc, red);
  const float m1 = acc[0] / (float)n, m2 = acc[1] / (float)n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float xh = (row[i] - mean) * rstd;
    const float v = fmaf(xh, g[i], b[i]);
    const float dx = dor[i] * silu_grad(v) * g[i];
    const float dp = rstd * (dx - m1 - xh * m2);
    d_pre[(size_t)blockIdx.x * ldp + i] = dp;
    put_split(so, blockIdx.x, i, dp);
  }
}

int ln_silu_bwd(const float* pre, int ld, const float* g, const float* b, float eps,
                const float* d_out, int ldd, int M, int n, float* d_pre, int ldp, float* d_ln,
                int ldl, cudaStream_t st, SplitOut so) {
  if (M <= 0) return 0;
  // the backward form already wins at the 1024 rows of the in-loop layers (5.5 vs 7.1 us at
  // 1024 x 512); the forward form only from a few thousand rows (4.7 vs 4.3 us at 1024)
  if (M >= (bulk_rows() < 1024 ? bulk_rows() : 1024) && n % 4 == 0 && n <= 1024 && ld % 4 == 0 &&
      ldd % 4 == 0 && ldp % 4 == 0 &&
      al16(pre) && al16(d_out) && al16(d_pre) && al16(g) && al16(b) &&
      (!d_ln || (ldl % 4 == 0 && al16(d_ln))) &&
      (!so.hi || (so.ld % 4 == 0 && al16(so.hi) && al16(so.lo)))) {
    const dim3 grid((M + BULK_WARPS - 1) / BULK_WARPS), block(BULK_WARPS * 32);
    if (n <= 512)
      DV3_CHECK_CUDA(launch_pdl(ln_silu_bwd_bulk_kernel<4>, grid, block, 0, st, pre, ld, g, b, eps,
                                d_out, ldd, M, n, d_pre, ldp, d_ln, ldl, so));
    else
      DV3_CHECK_CUDA(launch_pdl(ln_silu_bwd_bulk_kernel<8>, grid, block, 0, st, pre, ld, g, b, eps,
                                d_out, ldd, M, n, d_pre, ldp, d_ln, ldl, so));
    DV3_CHECK_LAUNCH("ln_silu_bwd_bulk_kernel");
    return 0;
  }
  DV3_CHECK_CUDA(launch_pdl(ln_silu_bwd_kernel, dim3(M), dim3(ROW_THREADS), 0, st, pre, ld, g, b, eps,
                            d_out, ldd, n, d_pre, ldp, d_ln, ldl, so));
  DV3_CHECK_LAUNCH("ln_silu_bwd_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// one-hot "Linear": the stoch input is one-hot per group, so  W [s, a]  is a sum of S rows of
// W^T plus a tiny dense part for the action.  Fused with LayerNorm+SiLU.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS)
gather_ln_silu_kernel(const int32_t* __restrict__ idx, int ldi, int S, int C,
                      const float* __restrict__ act, int lda, int A, const float* __restrict__ WT,
                      const float* __restrict__ addend, int ldadd, const float* __restrict__ g,
                      const float* __restrict__ b, float eps, int n, float* __restrict__ pre,
                      int ldp, float* __restrict__ out, int ldo, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sm[];
  float* rowbuf = sm;                                   // n
  float* red = sm + n;                                  // 128
  int* sidx = reinterpret_cast<int*>(red + 4 * 32);     // S
  float* sact = reinterpret_cast<float*>(sidx + S);     // A
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i < S; i += blockDim.x) sidx[i] = idx[(size_t)r * ldi + i] + i * C;
  for (int i = threadIdx.x; i < A; i += blockDim.x) sact[i] = act ? act[(size_t)r * lda + i] : 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float acc = addend ? addend[(size_t)r * ldadd + i] : 0.f;
    for (int s = 0; s < S; ++s) acc += WT[(size_t)sidx[s] * n + i];
    const float* wa = WT + (size_t)S * C * n + i;
    for (int a = 0; a < A; ++a) acc = fmaf(sact[a], wa[(size_t)a * n], acc);
    rowbuf[i] = acc;
    pre[(size_t)r * ldp + i] = acc;
  }
  __syncthreads();
  float mean, rstd;
  row_stats(rowbuf, n, eps, red, mean, rstd);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float y = siluf_(fmaf((rowbuf[i] - mean) * rstd, g[i], b[i]));
    out[(size_t)r * ldo + i] = y;
    put_split(so, r, i, y);
  }
}

// warp-per-row form (S, A <= 32, n = 128 NV <= 1024): the accumulation runs in the same order as
// above (addend, the S gathered rows in order, then the action terms), so `pre` is bit-identical;
// every lane keeps NV float4 of the row, the S row indices travel by shuffle.
template <int NV>
__global__ void __launch_bounds__(8 * 32)
gather_ln_silu_warp_kernel(const int32_t* __restrict__ idx, int ldi, int S, int C,
                           const float* __restrict__ act, int lda, int A,
                           const float* __restrict__ WT, const float* __restrict__ addend, int ldadd,
                           const float* __restrict__ g, const float* __restrict__ b, float eps, int M,
                           int n, float* __restrict__ pre, int ldp, float* __restrict__ out, int ldo,
                           SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= M) return;
  const int my_row = lane < S ? idx[(size_t)r * ldi + lane] + lane * C : 0;
  const float my_act = (act && lane < A) ? act[(size_t)r * lda + lane] : 0.f;
  float4 x[NV];
  bool ok[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = 4 * (j * 32 + lane);
    ok[j] = c < n;
    x[j] = (ok[j] && addend) ? *reinterpret_cast<const float4*>(addend + (size_t)r * ldadd + c)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll 4
  for (int s = 0; s < S; ++s) {
    const float* wr = WT + (size_t)__shfl_sync(FULL, my_row, s) * n;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (ok[j]) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wr + 4 * (j * 32 + lane)));
        x[j].x += w.x; x[j].y += w.y; x[j].z += w.z; x[j].w += w.w;
      }
    }
  }
  for (int a = 0; a < A; ++a) {
    const float sa = __shfl_sync(FULL, my_act, a);
    const float* wr = WT + ((size_t)S * C + a) * n;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (ok[j]) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wr + 4 * (j * 32 + lane)));
        x[j].x = fmaf(sa, w.x, x[j].x); x[j].y = fmaf(sa, w.y, x[j].y);
        x[j].z = fmaf(sa, w.z, x[j].z); x[j].w = fmaf(sa, w.w, x[j].w);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (ok[j]) *reinterpret_cast<float4*>(pre + (size_t)r * ldp + 4 * (j * 32 + lane)) = x[j];
  float mean, rstd;
  warp_row_stats<NV>(x, ok, n, eps, mean, rstd);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (ok[j]) {
      const int c = 4 * (j * 32 + lane);
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
      float4 y;
      y.x = siluf_(fmaf((x[j].x - mean) * rstd, gg.x, bb.x));
      y.y = siluf_(fmaf((x[j].y - mean) * rstd, gg.y, bb.y));
      y.z = siluf_(fmaf((x[j].z - mean) * rstd, gg.z, bb.z));
      y.w = siluf_(fmaf((x[j].w - mean) * rstd, gg.w, bb.w));
      *reinterpret_cast<float4*>(out + (size_t)r * ldo + c) = y;
      put_split4(so, r, c, y);
    }
  }
}

int gather_ln_silu(const int32_t* idx, int ldi, int S, int C, const float* act, int lda, int A,
                   const float* WT, const float* addend, int ldadd, const float* g, const float* b,
                   float eps, int M, int n, float* pre, int ldp, float* out, int ldo,
                   cudaStream_t st, SplitOut so) {
  if (M <= 0) return 0;
  {
    // opt-in ("1"): measured 0.5 % slower over the train step than the block-per-row kernel
    // (same-box A/B, 80.5 vs 80.9 steps/s) -- 128 blocks of 8 warps leave the gathers latency-bound
    const char* wf = DV3_ENV("DV3_GATHER_WARP");
    const bool ok = (wf && wf[0] == '1') && M >= 512 && S <= 32 && A <= 32 && n % 4 == 0 &&
                    n <= 1024 && ldp % 4 == 0 && ldo % 4 == 0 && al16(WT) && al16(pre) && al16(out) &&
                    al16(g) && al16(b) && (!addend || (ldadd % 4 == 0 && al16(addend))) &&
                    (!so.hi || (so.ld % 4 == 0 && al16(so.hi) && al16(so.lo)));
    if (ok) {
      const dim3 grid((M + 7) / 8), block(256);
      if (n <= 512)
        DV3_CHECK_CUDA(launch_pdl(gather_ln_silu_warp_kernel<4>, grid, block, 0, st, idx, ldi, S, C, act,
                                  lda, A, WT, addend, ldadd, g, b, eps, M, n, pre, ldp, out, ldo, so));
      else
        DV3_CHECK_CUDA(launch_pdl(gather_ln_silu_warp_kernel<8>, grid, block, 0, st, idx, ldi, S, C, act,
                                  lda, A, WT, addend, ldadd, g, b, eps, M, n, pre, ldp, out, ldo, so));
      DV3_CHECK_LAUNCH("gather_ln_silu_warp_kernel");
      return 0;
    }
  }
  const size_t smem = (size_t)(n + 4 * 32 + S + A) * 4;
  DV3_REQUIRE(smem <= 48 * 1024, DV3_ERR_BAD_SHAPE, "gather_ln_silu: row of %d too wide", n);
  DV3_CHECK_CUDA(launch_pdl(gather_ln_silu_kernel, dim3(M), dim3(ROW_THREADS), smem, st, idx, ldi, S,
                            C, act, lda, A, WT, addend, ldadd, g, b, eps, n, pre, ldp, out, ldo,
                            so));
  DV3_CHECK_LAUNCH("gather_ln_silu_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// LayerNorm-GRU gate block.  parts = LN_{3D}(g_pre); r=sig(p0); c=tanh(r*p1); u=sig(p2-1);
// h' = u*c + (1-u)*h.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS)
gru_gates_fwd_kernel(const float* __restrict__ g_pre, int ldg, const float* __restrict__ g,
                     const float* __restrict__ b, float eps, const float* __restrict__ h, int ldh,
                     int D, float* __restrict__ h_new, int ldn, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[4 * 32];
  const float* row = g_pre + (size_t)blockIdx.x * ldg;
  float mean, rstd;
  row_stats(row, 3 * D, eps, red, mean, rstd);
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const float pr = fmaf((row[j] - mean) * rstd, g[j], b[j]);
    const float pc = fmaf((row[D + j] - mean) * rstd, g[D + j], b[D + j]);
    const float pu = fmaf((row[2 * D + j] - mean) * rstd, g[2 * D + j], b[2 * D + j]);
    const float r = sigmoidf_(pr);
    const float c = tanhf(r * pc);
    const float u = sigmoidf_(pu - 1.f);
    const float hp = h[(size_t)blockIdx.x * ldh + j];
    const float hn = u * c + (1.f - u) * hp;
    h_new[(size_t)blockIdx.x * ldn + j] = hn;
    put_split(so, blockIdx.x, j, hn);
  }
}

// warp-per-row form (D = 128 DV <= 1024): row in registers, one read, shuffle reductions
template <int DV>
__global__ void __launch_bounds__(8 * 32)
gru_gates_fwd_warp_kernel(const float* __restrict__ g_pre, int ldg, const float* __restrict__ g,
                          const float* __restrict__ b, float eps, const float* __restrict__ h,
                          int ldh, int M, int D, float* __restrict__ h_new, int ldn, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= M) return;
  const float* row = g_pre + (size_t)r * ldg;
  float4 x[3][DV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < DV; ++j) {
      x[k][j] = *reinterpret_cast<const float4*>(row + k * D + 4 * (j * 32 + lane));
      s += (x[k][j].x + x[k][j].y) + (x[k][j].z + x[k][j].w);
    }
  const float n3 = (float)(3 * D);
  const float mean = warp_sum(s) / n3;
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < DV; ++j) {
      const float a = x[k][j].x - mean, c = x[k][j].y - mean, d = x[k][j].z - mean,
                  e = x[k][j].w - mean;
      v = fmaf(a, a, v); v = fmaf(c, c, v); v = fmaf(d, d, v); v = fmaf(e, e, v);
    }
  const float rstd = 1.f / sqrtf(warp_sum(v) / n3 + eps);
  auto gate = [&](float xr, float xc, float xu, float gr, float gc, float gu, float br, float bc,
                  float bu, float hp) {
    const float pr = fmaf((xr - mean) * rstd, gr, br);
    const float pc = fmaf((xc - mean) * rstd, gc, bc);
    const float pu = fmaf((xu - mean) * rstd, gu, bu);
    const float rg = sigmoidf_(pr);
    const float cc = tanhf(rg * pc);
    const float u = sigmoidf_(pu - 1.f);
    return u * cc + (1.f - u) * hp;
  };
#pragma unroll
  for (int j = 0; j < DV; ++j) {
    const int c = 4 * (j * 32 + lane);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + D + c));
    const float4 g2 = __ldg(reinterpret_cast<const float4*>(g + 2 * D + c));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(b + D + c));
    const float4 b2 = __ldg(reinterpret_cast<const float4*>(b + 2 * D + c));
    const float4 hp = *reinterpret_cast<const float4*>(h + (size_t)r * ldh + c);
    float4 hn;
    hn.x = gate(x[0][j].x, x[1][j].x, x[2][j].x, g0.x, g1.x, g2.x, b0.x, b1.x, b2.x, hp.x);
    hn.y = gate(x[0][j].y, x[1][j].y, x[2][j].y, g0.y, g1.y, g2.y, b0.y, b1.y, b2.y, hp.y);
    hn.z = gate(x[0][j].z, x[1][j].z, x[2][j].z, g0.z, g1.z, g2.z, b0.z, b1.z, b2.z, hp.z);
    hn.w = gate(x[0][j].w, x[1][j].w, x[2][j].w, g0.w, g1.w, g2.w, b0.w, b1.w, b2.w, hp.w);
    *reinterpret_cast<float4*>(h_new + (size_t)r * ldn + c) = hn;
    put_split4(so, r, c, hn);
  }
}

// block-per-row form for wide states (D = 1024 Q, Q in {2, 4, 8}: dyn_deter 2048 / 4096 / 8192): the
// 3 D row is read once (float4) and stays in registers, two block reductions give the LayerNorm
// statistics.  The scalar kernel above reads the row three times: 66 us per 1024 x 12288 launch of
// the large imagination config (ncu, profiles/ncu_large_rollout_r02.json).
template <int Q>
__global__ void __launch_bounds__(256)
gru_gates_fwd_block_kernel(const float* __restrict__ g_pre, int ldg, const float* __restrict__ g,
                           const float* __restrict__ b, float eps, const float* __restrict__ h,
                           int ldh, int D, float* __restrict__ h_new, int ldn, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[4 * 32];
  const int r = blockIdx.x, t = threadIdx.x;
  const float* row = g_pre + (size_t)r * ldg;
  float4 x[3][Q];
  float s[1] = {0.f};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      x[k][q] = *reinterpret_cast<const float4*>(row + k * D + 4 * (t + 256 * q));
      s[0] += (x[k][q].x + x[k][q].y) + (x[k][q].z + x[k][q].w);
    }
  block_sum<1>(s, red);
  const float n3 = (float)(3 * D);
  const float mean = s[0] / n3;
  float v[1] = {0.f};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float a = x[k][q].x - mean, c = x[k][q].y - mean, d = x[k][q].z - mean,
                  e = x[k][q].w - mean;
      v[0] = fmaf(a, a, v[0]); v[0] = fmaf(c, c, v[0]); v[0] = fmaf(d, d, v[0]); v[0] = fmaf(e, e, v[0]);
    }
  block_sum<1>(v, red);
  const float rstd = 1.f / sqrtf(v[0] / n3 + eps);
  auto gate = [&](float xr, float xc, float xu, float gr, float gc, float gu, float br, float bc,
                  float bu, float hp) {
    const float pr = fmaf((xr - mean) * rstd, gr, br);
    const float pc = fmaf((xc - mean) * rstd, gc, bc);
    const float pu = fmaf((xu - mean) * rstd, gu, bu);
    const float rg = sigmoidf_(pr);
    const float cc = tanhf(rg * pc);
    const float u = sigmoidf_(pu - 1.f);
    return u * cc + (1.f - u) * hp;
  };
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int c = 4 * (t + 256 * q);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + D + c));
    const float4 g2 = __ldg(reinterpret_cast<const float4*>(g + 2 * D + c));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(b + D + c));
    const float4 b2 = __ldg(reinterpret_cast<const float4*>(b + 2 * D + c));
    const float4 hp = *reinterpret_cast<const float4*>(h + (size_t)r * ldh + c);
    float4 hn;
    hn.x = gate(x[0][q].x, x[1][q].x, x[2][q].x, g0.x, g1.x, g2.x, b0.x, b1.x, b2.x, hp.x);
    hn.y = gate(x[0][q].y, x[1][q].y, x[2][q].y, g0.y, g1.y, g2.y, b0.y, b1.y, b2.y, hp.y);
    hn.z = gate(x[0][q].z, x[1][q].z, x[2][q].z, g0.z, g1.z, g2.z, b0.z, b1.z, b2.z, hp.z);
    hn.w = gate(x[0][q].w, x[1][q].w, x[2][q].w, g0.w, g1.w, g2.w, b0.w, b1.w, b2.w, hp.w);
    *reinterpret_cast<float4*>(h_new + (size_t)r * ldn + c) = hn;
    put_split4(so, r, c, hn);
  }
}

int gru_gates_fwd(const float* g_pre, int ldg, const float* g, const float* b, float eps,
                  const float* h, int ldh, int M, int D, float* h_new, int ldn, cudaStream_t st,
                  SplitOut so) {
  if (M <= 0) return 0;
  {
    const bool vec = ldg % 4 == 0 && ldh % 4 == 0 && ldn % 4 == 0 && al16(g_pre) && al16(g) &&
                     al16(b) && al16(h) && al16(h_new) &&
                     (!so.hi || (so.ld % 4 == 0 && al16(so.hi) && al16(so.lo)));
    if (vec && (D == 2048 || D == 4096 || D == 8192)) {
      const dim3 grid(M), block(256);
      if (D == 2048)
        DV3_CHECK_CUDA(launch_pdl(gru_gates_fwd_block_kernel<2>, grid, block, 0, st, g_pre, ldg, g, b,
                                  eps, h, ldh, D, h_new, ldn, so));
      else if (D == 4096)
        DV3_CHECK_CUDA(launch_pdl(gru_gates_fwd_block_kernel<4>, grid, block, 0, st, g_pre, ldg, g, b,
                                  eps, h, ldh, D, h_new, ldn, so));
      else
        DV3_CHECK_CUDA(launch_pdl(gru_gates_fwd_block_kernel<8>, grid, block, 0, st, g_pre, ldg, g, b,
                                  eps, h, ldh, D, h_new, ldn, so));
      DV3_CHECK_LAUNCH("gru_gates_fwd_block_kernel");
      return 0;
    }
  }
  {
    const char* wf = DV3_ENV("DV3_GRU_WARP_FWD");         // "0": keep the block-per-row kernel
    const bool ok = !(wf && wf[0] == '0') && M >= 512 && (D == 512 || D == 1024) && ldg % 4 == 0 &&
                    ldh % 4 == 0 && ldn % 4 == 0 && al16(g_pre) && al16(g) && al16(b) && al16(h) &&
                    al16(h_new) && (!so.hi || (so.ld % 4 == 0 && al16(so.hi) && al16(so.lo)));
    if (ok) {
      const dim3 grid((M + 7) / 8), block(256);
      if (D == 512)
        DV3_CHECK_CUDA(launch_pdl(gru_gates_fwd_warp_kernel<4>, grid, block, 0, st, g_pre, ldg, g, b,
                                  eps, h, ldh, M, D, h_new, ldn, so));
      else
        DV3_CHECK_CUDA(launch_pdl(gru_gates_fwd_warp_kernel<8>, grid, block, 0, st, g_pre, ldg, g, b,
                                  eps, h, ldh, M, D, h_new, ldn, so));
      DV3_CHECK_LAUNCH("gru_gates_fwd_warp_kernel");
      return 0;
    }
  }
  DV3_CHECK_CUDA(launch_pdl(gru_gates_fwd_kernel, dim3(M), dim3(ROW_THREADS), 0, st, g_pre, ldg, g, b,
                            eps, h, ldh, D, h_new, ldn, so));
  DV3_CHECK_LAUNCH("gru_gates_fwd_kernel");
  return 0;
}

struct DhIn {
  const float* p[4];
  int ld[4];
};

__global__ void __launch_bounds__(ROW_THREADS)
gru_gates_bwd_kernel(const float* __restrict__ g_pre, int ldg, const float* __restrict__ g,
                     const float* __restrict__ b, float eps, const float* __restrict__ h, int ldh,
                     DhIn dh, int D, float* __restrict__ d_g_pre, int ldp,
                     float* __restrict__ d_g_ln, int ldl, float* __restrict__ dh_direct, int ldd,
                     SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sm[];
  float* dparts = sm;             // 3D: gradient w.r.t. the LN affine output
  float* red = sm + 3 * D;        // 128
  const int r = blockIdx.x;
  const float* row = g_pre + (size_t)r * ldg;
  float mean, rstd;
  row_stats(row, 3 * D, eps, red, mean, rstd);
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const float pr = fmaf((row[j] - mean) * rstd, g[j], b[j]);
    const float pc = fmaf((row[D + j] - mean) * rstd, g[D + j], b[D + j]);
    const float pu = fmaf((row[2 * D + j] - mean) * rstd, g[2 * D + j], b[2 * D + j]);
    const float rg = sigmoidf_(pr);
    const float c = tanhf(rg * pc);
    const float u = sigmoidf_(pu - 1.f);
    const float hp = h[(size_t)r * ldh + j];
    float d = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (dh.p[q]) d += dh.p[q][(size_t)r * dh.ld[q] + j];
    const float du = d * (c - hp);
    const float dc = d * u;
    const float drc = dc * (1.f - c * c);
    dparts[j] = drc * pc * rg * (1.f - rg);
    dparts[D + j] = drc * rg;
    dparts[2 * D + j] = du * u * (1.f - u);
    dh_direct[(size_t)r * ldd + j] = d * (1.f - u);
  }
  __syncthreads();
  float acc[2] = {0.f, 0.f};
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
    const float xh = (row[i] - mean) * rstd;
    const float dx = dparts[i] * g[i];
    acc[0] += dx;
    acc[1] = fmaf(dx, xh, acc[1]);
    if (d_g_ln) d_g_ln[(size_t)r * ldl + i] = dparts[i];
  }
  block_sum<2>(acc, red);
  const float m1 = acc[0] / (float)(3 * D), m2 = acc[1] / (float)(3 * D);
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
    const float xh = (row[i] - mean) * rstd;
    const float dp = rstd * (dparts[i] * g[i] - m1 - xh * m2);
    d_g_pre[(size_t)r * ldp + i] = dp;
    put_split(so, r, i, dp);
  }
}

// Warp-per-row form of the gate backward (D = 128 DV <= 1024): the 3D-wide row, the gate
// derivatives and the LayerNorm backward all stay in registers; one read of every input, shuffle
// reductions, 16-byte stores.  The block-per-row kernel above took 21.8 us for 1024 x 1536.
template <int DV>
__global__ void __launch_bounds__(BULK_WARPS * 32)
gru_gates_bwd_warp_kernel(const float* __restrict__ g_pre, int ldg, const float* __restrict__ g,
                          const float* __restrict__ b, float eps, const float* __restrict__ h,
                          int ldh, DhIn dh, int M, int D, float* __restrict__ d_g_pre, int ldp,
                          float* __restrict__ d_g_ln, int ldl, float* __restrict__ dh_direct,
                          int ldd, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * BULK_WARPS + (threadIdx.x >> 5);
  if (r >= M) return;
  const float* row = g_pre + (size_t)r * ldg;
  float x[3][DV][4];                    // g_pre -> xhat
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < DV; ++j) {
      const float4 t = *reinterpret_cast<const float4*>(row + k * D + 4 * (j * 32 + lane));
      x[k][j][0] = t.x; x[k][j][1] = t.y; x[k][j][2] = t.z; x[k][j][3] = t.w;
      s += (t.x + t.y) + (t.z + t.w);
    }
  const float n3 = (float)(3 * D);
  const float mean = warp_sum(s) / n3;
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < DV; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = x[k][j][e] - mean;
        v = fmaf(d, d, v);
      }
  const float rstd = 1.f / sqrtf(warp_sum(v) / n3 + eps);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int j = 0; j < DV; ++j) {
    const int c = 4 * (j * 32 + lane);
    float gg[3][4], bb[3][4], hp[4], dsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(g + k * D + c));
      const float4 w = __ldg(reinterpret_cast<const float4*>(b + k * D + c));
      gg[k][0] = t.x; gg[k][1] = t.y; gg[k][2] = t.z; gg[k][3] = t.w;
      bb[k][0] = w.x; bb[k][1] = w.y; bb[k][2] = w.z; bb[k][3] = w.w;
    }
    {
      const float4 t = *reinterpret_cast<const float4*>(h + (size_t)r * ldh + c);
      hp[0] = t.x; hp[1] = t.y; hp[2] = t.z; hp[3] = t.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (dh.p[q]) {
        const float4 t = *reinterpret_cast<const float4*>(dh.p[q] + (size_t)r * dh.ld[q] + c);
        dsum[0] += t.x; dsum[1] += t.y; dsum[2] += t.z; dsum[3] += t.w;
      }
    float dpar[3][4], dhd[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
#pragma unroll
      for (int k = 0; k < 3; ++k) x[k][j][e] = (x[k][j][e] - mean) * rstd;
      const float pr = fmaf(x[0][j][e], gg[0][e], bb[0][e]);
      const float pc = fmaf(x[1][j][e], gg[1][e], bb[1][e]);
      const float pu = fmaf(x[2][j][e], gg[2][e], bb[2][e]);
      const float rg = sigmoidf_(pr);
      const float cc = tanhf(rg * pc);
      const float u = sigmoidf_(pu - 1.f);
      const float d = dsum[e];
      const float du = d * (cc - hp[e]);
      const float drc = d * u * (1.f - cc * cc);
      dpar[0][e] = drc * pc * rg * (1.f - rg);
      dpar[1][e] = drc * rg;
      dpar[2][e] = du * u * (1.f - u);
      dhd[e] = d * (1.f - u);
    }
    *reinterpret_cast<float4*>(dh_direct + (size_t)r * ldd + c) = make_float4(dhd[0], dhd[1], dhd[2], dhd[3]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (d_g_ln)
        *reinterpret_cast<float4*>(d_g_ln + (size_t)r * ldl + k * D + c) =
            make_float4(dpar[k][0], dpar[k][1], dpar[k][2], dpar[k][3]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float dx = dpar[k][e] * gg[k][e];
        a0 += dx;
        a1 = fmaf(dx, x[k][j][e], a1);
        gg[k][e] = dx;
      }
    }
    // dx (now in gg) is needed again after the row reductions: park it in this lane's slice of
    // the d_g_pre row and re-read it below instead of holding another 12 DV registers
#pragma unroll
    for (int k = 0; k < 3; ++k)
      *reinterpret_cast<float4*>(d_g_pre + (size_t)r * ldp + k * D + c) =
          make_float4(gg[k][0], gg[k][1], gg[k][2], gg[k][3]);
  }
  const float m1 = warp_sum(a0) / n3, m2 = warp_sum(a1) / n3;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < DV; ++j) {
      const int c = k * D + 4 * (j * 32 + lane);
      float4* dst = reinterpret_cast<float4*>(d_g_pre + (size_t)r * ldp + c);
      const float4 dx = *dst;            // this lane's own earlier store
      float4 dp;
      dp.x = rstd * (dx.x - m1 - x[k][j][0] * m2);
      dp.y = rstd * (dx.y - m1 - x[k][j][1] * m2);
      dp.z = rstd * (dx.z - m1 - x[k][j][2] * m2);
      dp.w = rstd * (dx.w - m1 - x[k][j][3] * m2);
      *dst = dp;
      put_split4(so, r, c, dp);
    }
}

int gru_gates_bwd(const float* g_pre, int ldg, const float* g, const float* b, float eps,
                  const float* h, int ldh, const float* const dh_in[4], const int ld_in[4], int M,
                  int D, float* d_g_pre, int ldp, float* d_g_ln, int ldl, float* dh_direct,
                  int ldd, cudaStream_t st, SplitOut so) {
  if (M <= 0) return 0;
  DhIn dh;
  for (int q = 0; q < 4; ++q) { dh.p[q] = dh_in[q]; dh.ld[q] = ld_in[q]; }
  {
    const char* wf = DV3_ENV("DV3_GRU_WARP");             // "0": keep the block-per-row kernel
    bool ok = !(wf && wf[0] == '0') && M >= 512 && (D == 512 || D == 1024) && ldg % 4 == 0 &&
              ldh % 4 == 0 && ldp % 4 == 0 && ldd % 4 == 0 && al16(g_pre) && al16(g) && al16(b) &&
              al16(h) && al16(d_g_pre) && al16(dh_direct) &&
              (!d_g_ln || (ldl % 4 == 0 && al16(d_g_ln))) &&
              (!so.hi || (so.ld % 4 == 0 && al16(so.hi) && al16(so.lo)));
    for (int q = 0; q < 4; ++q)
      if (dh.p[q] && (dh.ld[q] % 4 != 0 || !al16(dh.p[q]))) ok = false;
    if (ok) {
      const dim3 grid((M + BULK_WARPS - 1) / BULK_WARPS), block(BULK_WARPS * 32);
      if (D == 512)
        DV3_CHECK_CUDA(launch_pdl(gru_gates_bwd_warp_kernel<4>, grid, block, 0, st, g_pre, ldg, g, b,
                                  eps, h, ldh, dh, M, D, d_g_pre, ldp, d_g_ln, ldl, dh_direct, ldd, so));
      else
        DV3_CHECK_CUDA(launch_pdl(gru_gates_bwd_warp_kernel<8>, grid, block, 0, st, g_pre, ldg, g, b,
                                  eps, h, ldh, dh, M, D, d_g_pre, ldp, d_g_ln, ldl, dh_direct, ldd, so));
      DV3_CHECK_LAUNCH("gru_gates_bwd_warp_kernel");
      return 0;
    }
  }
  const size_t smem = (size_t)(3 * D + 4 * 32) * 4;
  static DeviceOnce attr_set;
  if (smem > 48 * 1024 && attr_set.need())
    DV3_CHECK_CUDA(cudaFuncSetAttribute(gru_gates_bwd_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  DV3_REQUIRE(smem <= 200 * 1024, DV3_ERR_BAD_SHAPE, "gru_gates_bwd: deter %d too wide", D);
  DV3_CHECK_CUDA(launch_pdl(gru_gates_bwd_kernel, dim3(M), dim3(ROW_THREADS), smem, st, g_pre, ldg, g,
                            b, eps, h, ldh, dh, D, d_g_pre, ldp, d_g_ln, ldl, dh_direct, ldd, so));
  DV3_CHECK_LAUNCH("gru_gates_bwd_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// unimix categorical: one warp per (row, group), lane k = class k.
// probs are built exactly along the reference's chain so the supplied-uniform argmax matches:
//   p = softmax(l); p = (1-r)p + r/C; l' = log p            tools.py:439-441
//   norm = l' - logsumexp(l'); probs = softmax(norm)        torch Categorical (logits -> probs)
//   idx = argmax_k probs_k / (-log u_k)                     ATen multinomial n=1
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
onehot_sample_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ u,
                     int ldu, int permT, int permB, float unimix, int M, int S, int C,
                     int32_t* __restrict__ idx, int ldi, float* __restrict__ onehot, int ldo) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)M * S) return;
  const int r = (int)(w / S), s = (int)(w % S);
  const bool valid = lane < C;
  const float l = valid ? logits[(size_t)r * ldl + s * C + lane] : 0.f;
  const Unimix um = unimix_probs(l, valid, C, unimix);
  float score;
  if (u) {
    const int ur = permT > 0 ? (r % permT) * permB + r / permT : r;
    const float uu = valid ? u[(size_t)ur * ldu + s * C + lane] : 1.f;
    score = um.probs / (-logf(uu));
  } else {
    score = um.norm;
  }
  const int k = warp_argmax(score, valid, lane);
  if (idx && lane == 0) idx[(size_t)r * ldi + s] = k;
  if (onehot && valid) onehot[(size_t)r * ldo + s * C + lane] = (lane == k) ? 1.f : 0.f;
}

// ---- C == 32: four lanes per (row, group) -------------------------------------------------------
// The lane-per-class form above spends its time in shuffles (six 5-step butterflies and a 10-shuffle
// argmax per group, one warp-shuffle per clock and SM): 15 us for 1024 x 32 groups.  Here lane q of
// a quad keeps classes q, q+4, .., q+28 in registers, so the butterfly's first three steps
// (i ^ 16, i ^ 8, i ^ 4) are adds between registers and only the last two cross lanes.  The
// association order is the butterfly's, so every value -- and every sampled index -- is
// bit-identical to the lane-per-class kernel (tests pin one against the other).
__device__ __forceinline__ float quad_sum(float (&a)[8]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] += a[j + 4];       // classes i, i + 16
  a[0] += a[2]; a[1] += a[3];                           // i, i + 8
  float v = a[0] + a[1];                                // i, i + 4
  v += __shfl_xor_sync(FULL, v, 2);
  v += __shfl_xor_sync(FULL, v, 1);
  return v;
}
__device__ __forceinline__ float quad_max(const float (&a)[8]) {
  float m = a[0];
#pragma unroll
  for (int j = 1; j < 8; ++j) m = fmaxf(m, a[j]);
  m = fmaxf(m, __shfl_xor_sync(FULL, m, 2));
  return fmaxf(m, __shfl_xor_sync(FULL, m, 1));
}

__global__ void __launch_bounds__(256)
onehot_sample_group_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ u,
                           int ldu, int permT, int permB, float unimix, int M, int S,
                           int32_t* __restrict__ idx, int ldi, float* __restrict__ onehot, int ldo) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int C = 32;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = t & 3;
  int w = t >> 2;                                       // (row, group)
  const bool live = w < M * S;                          // dead quads still take part in shuffles
  if (!live) w = 0;
  const int r = w / S, s = w - r * S;
  const float* lrow = logits + (size_t)r * ldl + s * C + q;
  float l[8], e[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) l[j] = __ldg(lrow + 4 * j);
  const float m = quad_max(l);
#pragma unroll
  for (int j = 0; j < 8; ++j) { l[j] = expf(l[j] - m); e[j] = l[j]; }
  const float s1 = quad_sum(e);
  if (unimix > 0.f) {
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = logf((l[j] / s1) * (1.f - unimix) + unimix / (float)C);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = __ldg(lrow + 4 * j);
  }
  const float m2 = quad_max(l);                         // l == lp
#pragma unroll
  for (int j = 0; j < 8; ++j) e[j] = expf(l[j] - m2);
  const float lse = m2 + logf(quad_sum(e));
#pragma unroll
  for (int j = 0; j < 8; ++j) l[j] = l[j] - lse;        // l <- norm
  if (u) {
    const float m3 = quad_max(l);
#pragma unroll
    for (int j = 0; j < 8; ++j) { l[j] = expf(l[j] - m3); e[j] = l[j]; }
    const float s3 = quad_sum(e);
    const int ur = permT > 0 ? (r % permT) * permB + r / permT : r;
    const float* urow = u + (size_t)ur * ldu + s * C + q;
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = (l[j] / s3) / (-logf(__ldg(urow + 4 * j)));   // l <- score
  }
  float bv = (l[0] != l[0]) ? -INFINITY : l[0];         // NaN never wins
  int k = q;
#pragma unroll
  for (int j = 1; j < 8; ++j) {
    const float v = (l[j] != l[j]) ? -INFINITY : l[j];
    if (v > bv) { bv = v; k = q + 4 * j; }              // first index wins ties
  }
#pragma unroll
  for (int o = 2; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(FULL, bv, o);
    const int oi = __shfl_xor_sync(FULL, k, o);
    if (ov > bv || (ov == bv && oi < k)) { bv = ov; k = oi; }
  }
  if (!live) return;
  if (idx && q == 0) idx[(size_t)r * ldi + s] = k;
  // lane q writes classes [8q, 8q + 8) of the one-hot row
  float4* o4 = reinterpret_cast<float4*>(onehot + (size_t)r * ldo + s * C + 8 * q);
  const int k0 = k - 8 * q;
  o4[0] = make_float4(k0 == 0 ? 1.f : 0.f, k0 == 1 ? 1.f : 0.f, k0 == 2 ? 1.f : 0.f, k0 == 3 ? 1.f : 0.f);
  o4[1] = make_float4(k0 == 4 ? 1.f : 0.f, k0 == 5 ? 1.f : 0.f, k0 == 6 ? 1.f : 0.f, k0 == 7 ? 1.f : 0.f);
}

int onehot_sample(const float* logits, int ldl, const float* u, int ldu, int permT, int permB,
                  float unimix, int M, int S, int C, int32_t* idx, int ldi, float* onehot, int ldo,
                  cudaStream_t st) {
  if (M <= 0) return 0;
  DV3_REQUIRE(C >= 1 && C <= 32, DV3_ERR_BAD_SHAPE, "onehot_sample: classes=%d (max 32)", C);
  const char* gf = DV3_ENV("DV3_SAMPLE_GROUP");         // "0": keep the lane-per-class kernel
  const bool group_form = !(gf && gf[0] == '0');
  if (group_form && C == 32 && onehot && (long long)M * S >= 4096 && (long long)M * S < (1ll << 28) &&
      ldl % 4 == 0 && ldo % 4 == 0 && al16(logits) && al16(onehot) && (!u || (ldu % 4 == 0 && al16(u)))) {
    const long long lanes = 4ll * M * S;
    DV3_CHECK_CUDA(launch_pdl(onehot_sample_group_kernel, dim3((unsigned)((lanes + 255) / 256)), dim3(256), 0, st,
                              logits, ldl, u, ldu, permT, permB, unimix, M, S, idx, ldi, onehot, ldo));
    DV3_CHECK_LAUNCH("onehot_sample_group_kernel");
    return 0;
  }
  const long long warps = (long long)M * S;
  const int grid = (int)((warps + 7) / 8);
  DV3_CHECK_CUDA(launch_pdl(onehot_sample_kernel, dim3(grid), dim3(256), 0, st, logits, ldl, u, ldu,
                            permT, permB, unimix, M, S, C, idx, ldi, onehot, ldo));
  DV3_CHECK_LAUNCH("onehot_sample_kernel");
  return 0;
}

// straight-through backward of sample = hard + probs - sg(probs):
//   q = (1-r) softmax(l) + r/C  (== probs);  gq = g - <g,q>;  d l = (1-r) p (gq - <gq,p>) + ext
__global__ void __launch_bounds__(256)
onehot_st_bwd_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ g1,
                     int ldg1, const float* __restrict__ g2, int ldg2,
                     const float* __restrict__ ext, int lde, float unimix, int M, int S, int C,
                     float* __restrict__ d_logits, int ldd, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)M * S) return;
  const int r = (int)(w / S), s = (int)(w % S);
  const bool valid = lane < C;
  const int col = s * C + lane;
  const float l = valid ? logits[(size_t)r * ldl + col] : 0.f;
  float g = 0.f;
  if (valid && g1) g += g1[(size_t)r * ldg1 + col];
  if (valid && g2) g += g2[(size_t)r * ldg2 + col];
  const float m = warp_max(valid ? l : -INFINITY);
  const float e = valid ? expf(l - m) : 0.f;
  const float p = e / warp_sum(e);
  const float q = valid ? (p * (1.f - unimix) + unimix / (float)C) : 0.f;
  const float gq = g - warp_sum(g * q);
  const float dot = warp_sum(valid ? gq * p : 0.f);
  float d = (1.f - unimix) * p * (gq - dot);
  if (valid) {
    if (ext) d += ext[(size_t)r * lde + col];
    d_logits[(size_t)r * ldd + col] = d;
    put_split(so, r, col, d);
  }
}

// four lanes per (row, group), as in onehot_sample_group_kernel (C == 32)
__global__ void __launch_bounds__(256)
onehot_st_bwd_group_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ g1,
                           int ldg1, const float* __restrict__ g2, int ldg2,
                           const float* __restrict__ ext, int lde, float unimix, int M, int S,
                           float* __restrict__ d_logits, int ldd, SplitOut so) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int C = 32;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = t & 3;
  int w = t >> 2;
  const bool live = w < M * S;
  if (!live) w = 0;
  const int r = w / S, s = w - r * S;
  const int col0 = s * C + q;
  float l[8], gq[8], tmp[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    l[j] = __ldg(logits + (size_t)r * ldl + col0 + 4 * j);
    float gg = 0.f;
    if (g1) gg += __ldg(g1 + (size_t)r * ldg1 + col0 + 4 * j);
    if (g2) gg += __ldg(g2 + (size_t)r * ldg2 + col0 + 4 * j);
    gq[j] = gg;
  }
  const float m = quad_max(l);
#pragma unroll
  for (int j = 0; j < 8; ++j) { l[j] = expf(l[j] - m); tmp[j] = l[j]; }
  const float se = quad_sum(tmp);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    l[j] = l[j] / se;                                             // p
    tmp[j] = gq[j] * (l[j] * (1.f - unimix) + unimix / (float)C); // g * q
  }
  const float gdotq = quad_sum(tmp);
#pragma unroll
  for (int j = 0; j < 8; ++j) { gq[j] = gq[j] - gdotq; tmp[j] = gq[j] * l[j]; }
  const float dot = quad_sum(tmp);
  if (!live) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = col0 + 4 * j;
    float d = (1.f - unimix) * l[j] * (gq[j] - dot);
    if (ext) d += __ldg(ext + (size_t)r * lde + col);
    d_logits[(size_t)r * ldd + col] = d;
    put_split(so, r, col, d);
  }
}

int onehot_st_bwd(const float* logits, int ldl, const float* g1, int ldg1, const float* g2,
                  int ldg2, const float* ext, int lde, float unimix, int M, int S, int C,
                  float* d_logits, int ldd, cudaStream_t st, SplitOut so) {
  if (M <= 0) return 0;
  DV3_REQUIRE(C >= 1 && C <= 32, DV3_ERR_BAD_SHAPE, "onehot_st_bwd: classes=%d (max 32)", C);
  {
    const char* gf = DV3_ENV("DV3_STBWD_GROUP");          // "0": keep the lane-per-class kernel
    if (!(gf && gf[0] == '0') && C == 32 && (long long)M * S >= 4096 &&
        (long long)M * S < (1ll << 28)) {
      const long long lanes = 4ll * M * S;
      DV3_CHECK_CUDA(launch_pdl(onehot_st_bwd_group_kernel, dim3((unsigned)((lanes + 255) / 256)),
                                dim3(256), 0, st, logits, ldl, g1, ldg1, g2, ldg2, ext, lde, unimix,
                                M, S, d_logits, ldd, so));
      DV3_CHECK_LAUNCH("onehot_st_bwd_group_kernel");
      return 0;
    }
  }
  const long long warps = (long long)M * S;
  const int grid = (int)((warps + 7) / 8);
  DV3_CHECK_CUDA(launch_pdl(onehot_st_bwd_kernel, dim3(grid), dim3(256), 0, st, logits, ldl, g1, ldg1,
                            g2, ldg2, ext, lde, unimix, M, S, C, d_logits, ldd, so));
  DV3_CHECK_LAUNCH("onehot_st_bwd_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// small data movers
// ------------------------------------------------------------------------------------------
__global__ void idx_to_onehot_kernel(const int32_t* __restrict__ idx, int ldi, int M, int S, int C,
                                     float* __restrict__ out, int ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long tot = (long long)M * S * C;
  if (i >= tot) return;
  const int c = (int)(i % C);
  const int s = (int)((i / C) % S);
  const int r = (int)(i / ((long long)S * C));
  out[(size_t)r * ld + s * C + c] = (idx[(size_t)r * ldi + s] == c) ? 1.f : 0.f;
}

int idx_to_onehot(const int32_t* idx, int ldi, int M, int S, int C, float* out, int ld,
                  cudaStream_t st) {
  const long long tot = (long long)M * S * C;
  if (tot <= 0) return 0;
  idx_to_onehot_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(idx, ldi, M, S, C, out, ld);
  DV3_CHECK_LAUNCH("idx_to_onehot_kernel");
  return 0;
}

template <typename T>
__global__ void copy_rows_kernel(const T* __restrict__ in, int ldi, int M, int n,
                                 T* __restrict__ out, int ldo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * n) return;
  const int r = (int)(i / n), c = (int)(i % n);
  out[(size_t)r * ldo + c] = in[(size_t)r * ldi + c];
}

int copy_rows(const float* in, int ldi, int M, int n, float* out, int ldo, cudaStream_t st) {
  const long long tot = (long long)M * n;
  if (tot <= 0) return 0;
  copy_rows_kernel<float><<<(int)((tot + 255) / 256), 256, 0, st>>>(in, ldi, M, n, out, ldo);
  DV3_CHECK_LAUNCH("copy_rows_kernel");
  return 0;
}

int copy_rows_i32(const int32_t* in, int ldi, int M, int n, int32_t* out, int ldo,
                  cudaStream_t st) {
  const long long tot = (long long)M * n;
  if (tot <= 0) return 0;
  copy_rows_kernel<int32_t><<<(int)((tot + 255) / 256), 256, 0, st>>>(in, ldi, M, n, out, ldo);
  DV3_CHECK_LAUNCH("copy_rows_kernel");
  return 0;
}

__global__ void tanh_kernel(const float* __restrict__ in, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tanhf(in[i]);
}

int tanh_vec(const float* in, int n, float* out, cudaStream_t st) {
  if (n <= 0) return 0;
  tanh_kernel<<<(n + 255) / 256, 256, 0, st>>>(in, n, out);
  DV3_CHECK_LAUNCH("tanh_kernel");
  return 0;
}

int fill_zero(void* p, size_t bytes, cudaStream_t st) {
  if (!p || bytes == 0) return 0;
  DV3_CHECK_CUDA(cudaMemsetAsync(p, 0, bytes, st));
  return 0;
}

}  // namespace dv3

// ------------------------------------------------------------------------------------------
// C ABI: the building blocks, individually testable against the oracle
// ------------------------------------------------------------------------------------------
using namespace dv3;
#define ST(s) static_cast<cudaStream_t>(s)

// ------------------------------------------------------------------------------------------
// LayerNorm parameter gradients over all rows:  dg[j] = sum_r d_ln[r,j] * xhat[r,j],
// db[j] = sum_r d_ln[r,j]   (xhat recomputed from the saved pre-LN rows, two-pass statistics).
// One warp per row, a lane owns columns lane + 32 i and keeps their partial sums in registers
// over all the rows its warp visits; the warps of a CTA fold through shared memory and the CTA
// issues one atomicAdd per column.  dg/db must be zeroed by the caller.
// ------------------------------------------------------------------------------------------
namespace dv3 {

constexpr int LG_THREADS = 256;

template <int NPL>   // columns per lane
__global__ void __launch_bounds__(LG_THREADS)
ln_param_grads_kernel(const float* __restrict__ pre, int ld, const float* __restrict__ d_ln,
                      int ldl, float eps, int M, int n, float* __restrict__ dg,
                      float* __restrict__ db) {
  extern __shared__ float lg_sm[];               // [2][n]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = LG_THREADS / 32;
  float ag[NPL], ab[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) ag[i] = ab[i] = 0.f;
  for (int r = blockIdx.x * nw + warp; r < M; r += gridDim.x * nw) {
    const float* row = pre + (size_t)r * ld;
    const float* dr = d_ln + (size_t)r * ldl;
    float v[NPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int j = lane + 32 * i;
      v[i] = j < n ? row[j] : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) / (float)n;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int j = lane + 32 * i;
      const float d = j < n ? v[i] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float rstd = 1.f / sqrtf(warp_sum(q) / (float)n + eps);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int j = lane + 32 * i;
      if (j < n) {
        const float d = dr[j];
        ag[i] = fmaf(d, (v[i] - mean) * rstd, ag[i]);
        ab[i] += d;
      }
    }
  }
  for (int i = threadIdx.x; i < 2 * n; i += LG_THREADS) lg_sm[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int j = lane + 32 * i;
    if (j < n) {
      atomicAdd(&lg_sm[j], ag[i]);
      atomicAdd(&lg_sm[n + j], ab[i]);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += LG_THREADS) {
    atomicAdd(dg + j, lg_sm[j]);
    atomicAdd(db + j, lg_sm[n + j]);
  }
}

template <int NPL>
static int launch_ln_param_grads(const float* pre, int ld, const float* d_ln, int ldl, float eps,
                                 int M, int n, float* dg, float* db, cudaStream_t st) {
  const int nw = LG_THREADS / 32;
  int grid = (M + nw - 1) / nw;
  if (grid > 2 * 148) grid = 2 * 148;
  ln_param_grads_kernel<NPL><<<grid, LG_THREADS, (size_t)2 * n * 4, st>>>(pre, ld, d_ln, ldl, eps,
                                                                           M, n, dg, db);
  DV3_CHECK_LAUNCH("ln_param_grads_kernel");
  return 0;
}

}  // namespace dv3

static int ln_param_grads_any(const float* pre, int32_t ld, const float* d_ln, int32_t ldl,
                              float eps, int32_t M, int32_t n, float* dg, float* db, bool zero,
                              void* stream) {
  using namespace dv3;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DV3_REQUIRE(M >= 0 && n > 0 && n <= 2048, DV3_ERR_BAD_SHAPE, "ln_param_grads: M=%d n=%d", M, n);
  DV3_REQUIRE(pre && d_ln && dg && db, DV3_ERR_NULL, "ln_param_grads: null pointer");
  if (zero) {
    DV3_CHECK_CUDA(cudaMemsetAsync(dg, 0, (size_t)n * 4, st));
    DV3_CHECK_CUDA(cudaMemsetAsync(db, 0, (size_t)n * 4, st));
  }
  if (M == 0) return 0;
  const int npl = (n + 31) / 32;
  if (npl <= 4) return launch_ln_param_grads<4>(pre, ld, d_ln, ldl, eps, M, n, dg, db, st);
  if (npl <= 16) return launch_ln_param_grads<16>(pre, ld, d_ln, ldl, eps, M, n, dg, db, st);
  if (npl <= 32) return launch_ln_param_grads<32>(pre, ld, d_ln, ldl, eps, M, n, dg, db, st);
  if (npl <= 48) return launch_ln_param_grads<48>(pre, ld, d_ln, ldl, eps, M, n, dg, db, st);
  return launch_ln_param_grads<64>(pre, ld, d_ln, ldl, eps, M, n, dg, db, st);
}

extern "C" int dv3_ln_param_grads(const float* pre, int32_t ld, const float* d_ln, int32_t ldl,
                                  float eps, int32_t M, int32_t n, float* dg, float* db,
                                  void* stream) {
  return ln_param_grads_any(pre, ld, d_ln, ldl, eps, M, n, dg, db, true, stream);
}

extern "C" int dv3_ln_param_grads_acc(const float* pre, int32_t ld, const float* d_ln, int32_t ldl,
                                      float eps, int32_t M, int32_t n, float* dg, float* db,
                                      void* stream) {
  return ln_param_grads_any(pre, ld, d_ln, ldl, eps, M, n, dg, db, false, stream);
}

extern "C" int dv3_ln_silu_fwd(const float* pre, int32_t ld, const float* g, const float* b,
                               float eps, int32_t M, int32_t n, float* out, int32_t ldo,
                               void* stream) {
  DV3_REQUIRE(pre && g && b && out, DV3_ERR_NULL, "ln_silu_fwd: null pointer");
  return ln_silu_fwd(pre, ld, g, b, eps, M, n, out, ldo, ST(stream));
}

extern "C" int dv3_ln_silu_fwd_split(const float* pre, int32_t ld, const float* g, const float* b,
                                     float eps, int32_t M, int32_t n, float* out, int32_t ldo,
                                     float* hi, float* lo, int32_t lds, void* stream) {
  DV3_REQUIRE(pre && g && b && out && hi && lo && lds >= n, DV3_ERR_NULL,
              "ln_silu_fwd_split: null pointer / short plane pitch");
  SplitOut so;
  so.hi = hi; so.lo = lo; so.ld = lds;
  return ln_silu_fwd(pre, ld, g, b, eps, M, n, out, ldo, ST(stream), so);
}
extern "C" int dv3_ln_silu_bwd_split(const float* pre, int32_t ld, const float* g, const float* b,
                                     float eps, const float* d_out, int32_t ldd, int32_t M,
                                     int32_t n, float* d_pre, float* d_ln, int32_t ldp, float* hi,
                                     float* lo, int32_t lds, void* stream) {
  DV3_REQUIRE(pre && g && b && d_out && d_pre && hi && lo && lds >= n, DV3_ERR_NULL,
              "ln_silu_bwd_split: null pointer / short plane pitch");
  SplitOut so;
  so.hi = hi; so.lo = lo; so.ld = lds;
  return ln_silu_bwd(pre, ld, g, b, eps, d_out, ldd, M, n, d_pre, ldp, d_ln, ldp, ST(stream), so);
}
extern "C" int dv3_ln_silu_bwd(const float* pre, int32_t ld, const float* g, const float* b,
                               float eps, const float* d_out, int32_t ldd, int32_t M, int32_t n,
                               float* d_pre, float* d_ln, int32_t ldp, void* stream) {
  DV3_REQUIRE(pre && g && b && d_out && d_pre, DV3_ERR_NULL, "ln_silu_bwd: null pointer");
  return ln_silu_bwd(pre, ld, g, b, eps, d_out, ldd, M, n, d_pre, ldp, d_ln, ldp, ST(stream));
}

extern "C" int dv3_gru_gates_fwd(const float* g_pre, int32_t ldg, const float* g, const float* b,
                                 float eps, const float* h, int32_t ldh, int32_t M, int32_t D,
                                 float* h_new, int32_t ldn, void* stream) {
  DV3_REQUIRE(g_pre && g && b && h && h_new, DV3_ERR_NULL, "gru_gates_fwd: null pointer");
  return gru_gates_fwd(g_pre, ldg, g, b, eps, h, ldh, M, D, h_new, ldn, ST(stream));
}

extern "C" int dv3_gru_gates_bwd(const float* g_pre, int32_t ldg, const float* g, const float* b,
                                 float eps, const float* h, int32_t ldh, const float* d_h_new,
                                 int32_t ldd, int32_t M, int32_t D, float* d_g_pre, float* d_g_ln,
                                 int32_t ldp, float* d_h, int32_t ldo, void* stream) {
  DV3_REQUIRE(g_pre && g && b && h && d_h_new && d_g_pre && d_h, DV3_ERR_NULL,
              "gru_gates_bwd: null pointer");
  const float* in[4] = {d_h_new, nullptr, nullptr, nullptr};
  const int lds[4] = {ldd, 0, 0, 0};
  return gru_gates_bwd(g_pre, ldg, g, b, eps, h, ldh, in, lds, M, D, d_g_pre, ldp, d_g_ln, ldp,
                       d_h, ldo, ST(stream));
}

extern "C" int dv3_onehot_linear_ln_silu(const int32_t* idx, int32_t S, int32_t C,
                                         const float* act, int32_t A, const float* WT,
                                         const float* addend, const float* g, const float* b,
                                         float eps, int32_t M, int32_t n, float* pre, float* out,
                                         void* stream) {
  DV3_REQUIRE(idx && WT && g && b && pre && out, DV3_ERR_NULL, "onehot_linear: null pointer");
  return gather_ln_silu(idx, S, S, C, act, A, act ? A : 0, WT, addend, n, g, b, eps, M, n, pre, n,
                        out, n, ST(stream));
}

extern "C" int dv3_onehot_sample(const float* logits, const float* u, float unimix, int32_t M,
                                 int32_t S, int32_t C, int32_t* idx, float* onehot,
                                 int32_t ld_onehot, void* stream) {
  DV3_REQUIRE(logits && (idx || onehot), DV3_ERR_NULL, "onehot_sample: null pointer");
  return onehot_sample(logits, S * C, u, S * C, 0, 0, unimix, M, S, C, idx, S, onehot, ld_onehot,
                       ST(stream));
}

extern "C" int dv3_onehot_st_bwd(const float* logits, const float* g_sample, const float* ext,
                                 float unimix, int32_t M, int32_t S, int32_t C, float* d_logits,
                                 void* stream) {
  DV3_REQUIRE(logits && d_logits, DV3_ERR_NULL, "onehot_st_bwd: null pointer");
  return onehot_st_bwd(logits, S * C, g_sample, S * C, nullptr, 0, ext, S * C, unimix, M, S, C,
                       d_logits, S * C, ST(stream));
}

extern "C" int dv3_idx_to_onehot(const int32_t* idx, int32_t M, int32_t S, int32_t C, float* out,
                                 int32_t ld, void* stream) {
  DV3_REQUIRE(idx && out, DV3_ERR_NULL, "idx_to_onehot: null pointer");
  return idx_to_onehot(idx, S, M, S, C, out, ld, ST(stream));
}
