// tf32 hi/lo split kernels and the scratch-based convenience entry of the tensor-core Linear.
// The GEMM itself is the persistent kernel in dv3_umma2.cu.
//
// 3xTF32 split: every fp32 operand x is written as hi + lo with hi = x with the 13 low mantissa
// bits cleared (exactly representable in tf32) and lo = x - hi (exact in fp32, |lo| < 2^-10 |x|).
//   A W^T  ~=  A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T          (dropped term ~2^-20 relative)
// Single-pass TF32 (2^-11) would break the 1e-4 / bit-exact-sample contract; this keeps the error
// at fp32-reordering level.
//
// Accumulation note (measured on B200): the tensor core adds into the fp32 TMEM accumulator with
// truncation, so a long accumulation chain drifts (error grows linearly in K, ~1e-5 at K=1024).
// Remedy ("promotion"), implemented in dv3_umma2.cu: the two small cross terms go to their own
// accumulator (they are 2^-10 of the result, their truncation is negligible); the hi*hi term is
// accumulated in TMEM only over chunks of 4 k-blocks (128 k), two chunk accumulators ping-pong,
// and the epilogue warps add the finished chunks in fp32 registers (round-to-nearest) while the
// next chunk is being computed.
#include "dv3_tc.cuh"

namespace dv3 {

// hi = x with the low 13 mantissa bits cleared, lo = x - hi; two row segments packed side by side
// (output rows have pitch Kp >= K1+K2, zero padded, so that the TMA row pitch is 16B-aligned)
__global__ void split_tf32_kernel(const float* __restrict__ a1, int ld1, int K1,
                                  const float* __restrict__ a2, int ld2, int K2, int M, int Kp,
                                  float* __restrict__ hi, float* __restrict__ lo) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * Kp) return;
  const int r = (int)(i / Kp), c = (int)(i % Kp);
  float x = 0.f;
  if (c < K1) x = a1[(size_t)r * ld1 + c];
  else if (c < K1 + K2) x = a2[(size_t)r * ld2 + (c - K1)];
  const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  hi[i] = h;
  lo[i] = x - h;
}

// the same for one segment with every pitch a multiple of 4 floats and 16-byte aligned bases:
// a thread moves float4s, threadIdx.x walks the row, threadIdx.y the rows (no per-element
// integer division; 15360 x 512 took 34 us in the scalar form, a copy takes 7)
__global__ void __launch_bounds__(256)
split_tf32_vec_kernel(const float* __restrict__ a, int ld, int K, int M, int Kp, int rpt,
                      float* __restrict__ hi, float* __restrict__ lo) {
  pdl_wait();
  pdl_launch_dependents();
  const int r0 = (blockIdx.x * blockDim.y + threadIdx.y) * rpt;   // rpt rows per thread-row
  for (int rr = 0; rr < rpt; ++rr) {
    const int r = r0 + rr;
    if (r >= M) return;
    for (int c = 4 * threadIdx.x; c < Kp; c += 4 * blockDim.x) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c + 4 <= K) {
        x = *reinterpret_cast<const float4*>(a + (size_t)r * ld + c);
      } else if (c < K) {
        const float* p = a + (size_t)r * ld + c;
        x.x = p[0];
        if (c + 1 < K) x.y = p[1];
        if (c + 2 < K) x.z = p[2];
      }
      float4 h;
      h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
      h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
      h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
      h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
      *reinterpret_cast<float4*>(hi + (size_t)r * Kp + c) = h;
      *reinterpret_cast<float4*>(lo + (size_t)r * Kp + c) =
          make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
    }
  }
}

// transposing split: in is stored [C, R] (row stride ld); hi/lo are the dense [R, C] split of in^T
__global__ void split_tf32_transpose_kernel(const float* __restrict__ in, int ld, int R, int C,
                                            int Cp, float* __restrict__ hi,
                                            float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;   // r indexes output rows = input columns
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && r < R) ? in[(size_t)c * ld + r] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cp) {
      const float x = tile[threadIdx.x][i];
      const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
      hi[(size_t)r * Cp + c] = h;
      lo[(size_t)r * Cp + c] = x - h;
    }
  }
}

int tc_split_t(const float* in, int ld, int R, int C, float* hi, float* lo, cudaStream_t st,
               int Cp) {
  if (R <= 0 || C <= 0) return 0;
  if (Cp < C) Cp = C;
  dim3 grid((Cp + 31) / 32, (R + 31) / 32), block(32, 8);
  split_tf32_transpose_kernel<<<grid, block, 0, st>>>(in, ld, R, C, Cp, hi, lo);
  DV3_CHECK_LAUNCH("split_tf32_transpose_kernel");
  return 0;
}

int tc_split(const float* a1, int ld1, int K1, const float* a2, int ld2, int K2, int M, float* hi,
             float* lo, cudaStream_t st, int Kp) {
  const int K = K1 + (a2 ? K2 : 0);
  if (Kp < K) Kp = K;
  const long long tot = (long long)M * Kp;
  if (tot <= 0) return 0;
  if (!a2 && Kp % 4 == 0 && ld1 % 4 == 0 && (reinterpret_cast<uintptr_t>(a1) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(lo) & 15) == 0) {
    // threadIdx.x covers a row in float4s (up to 64 lanes), the rest of the block takes rows
    int tx = 8;
    while (tx < 64 && tx * 4 < Kp) tx *= 2;
    const int ty = 256 / tx;
    const int rpt = (tot / 1024 >= 8192) ? 4 : 1;        // small tensors: spread over more blocks
    const int rows_pb = ty * rpt;
    DV3_CHECK_CUDA(launch_pdl(split_tf32_vec_kernel, dim3((unsigned)((M + rows_pb - 1) / rows_pb)),
                              dim3(tx, ty), 0, st, a1, ld1, K1, M, Kp, rpt, hi, lo));
    DV3_CHECK_LAUNCH("split_tf32_vec_kernel");
    return 0;
  }
  DV3_CHECK_CUDA(launch_pdl(split_tf32_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, st,
                            a1, ld1, K1, a2, ld2, a2 ? K2 : 0, M, Kp, hi, lo));
  DV3_CHECK_LAUNCH("split_tf32_kernel");
  return 0;
}

// C = A W^T from pre-split dense operands ([M,K] and [N,K], K % 4 == 0)
int tc_gemm(const float* Ah, const float* Al, const float* Wh, const float* Wl, const float* bias,
            const float* addend, int ldadd, float* C, int ldc, int M, int N, int K, int accumulate,
            cudaStream_t st) {
  DV3_REQUIRE(K % 4 == 0 && K > 0, DV3_ERR_BAD_SHAPE, "tc_gemm: K=%d must be a multiple of 4", K);
  TcOperand a{Ah, Al, K, false}, b{Wh, Wl, K, false};
  return tc_gemm_ops(a, K, nullptr, 0, b, bias, addend, ldadd, C, ldc, M, N, accumulate, st);
}

}  // namespace dv3

// C ABI: self-contained tensor-core Linear (splits both operands into the caller's scratch).
// scratch: 2*(M+N)*K floats.
extern "C" size_t dv3_linear_tc_scratch_bytes(int32_t M, int32_t N, int32_t K) {
  const size_t Kp = ((size_t)K + 3) & ~size_t(3);
  return (size_t)2 * ((size_t)M + N) * Kp * sizeof(float) + 1024;
}

extern "C" int dv3_linear_tc_fwd(const float* A, int32_t lda, int32_t transA, const float* W,
                                 int32_t ldw, int32_t transW, const float* bias,
                                 const float* addend, int32_t ldadd, float* C, int32_t ldc,
                                 int32_t M, int32_t N, int32_t K, void* scratch,
                                 size_t scratch_bytes, void* stream) {
  using namespace dv3;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DV3_REQUIRE(M >= 0 && N >= 0 && K > 0, DV3_ERR_BAD_SHAPE, "linear_tc_fwd: M=%d N=%d K=%d", M, N,
              K);
  if (M == 0 || N == 0) return 0;
  const int Kp = (K + 3) & ~3;   // split operands are zero padded to a 16-byte row pitch
  DV3_REQUIRE(A && W && C && scratch, DV3_ERR_NULL, "linear_tc_fwd: null pointer");
  DV3_REQUIRE(scratch_bytes >= dv3_linear_tc_scratch_bytes(M, N, K), DV3_ERR_WORKSPACE,
              "linear_tc_fwd: scratch too small");
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255));
  float* Ah = base;
  float* Al = Ah + (size_t)M * Kp;
  float* Wh = Al + (size_t)M * Kp;
  float* Wl = Wh + (size_t)N * Kp;
  if (transA) DV3_TRY(tc_split_t(A, lda, M, K, Ah, Al, st, Kp));
  else DV3_TRY(tc_split(A, lda, K, nullptr, 0, 0, M, Ah, Al, st, Kp));
  if (transW) DV3_TRY(tc_split_t(W, ldw, N, K, Wh, Wl, st, Kp));
  else DV3_TRY(tc_split(W, ldw, K, nullptr, 0, 0, N, Wh, Wl, st, Kp));
  return tc_gemm(Ah, Al, Wh, Wl, bias, addend, ldadd, C, ldc, M, N, Kp, 0, st);
}
