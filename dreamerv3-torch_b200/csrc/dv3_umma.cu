// fp32-accurate GEMM on the 5th-gen tensor cores:  C[M,N] = A[M,K] W[N,K]^T (+bias +addend)
//
// 3xTF32 split: every fp32 operand x is written as hi + lo with hi = x with the 13 low mantissa
// bits cleared (exactly representable in tf32) and lo = x - hi (exact in fp32, |lo| < 2^-10 |x|).
//   A W^T  ~=  A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T          (dropped term ~2^-20 relative)
// accumulated in fp32 in TMEM.  Single-pass TF32 (2^-11) would break the 1e-4 / bit-exact-sample
// contract; this keeps the error at fp32-reordering level.
//
// Kernel: one CTA per 128 x BN output tile, 192 threads, warp-specialised:
//   warp 0  TMA producer  (cp.async.bulk.tensor 2D, 128B swizzle, 4 operand tiles per stage)
//   warp 1  TMEM allocator + tcgen05.mma issuer (one elected lane, kind::tf32, M=128, N=BN, K=8)
//   warps 2-5 epilogue    (tcgen05.ld 32x32b -> registers -> bias/addend -> global)
// 3-stage smem ring with full/empty mbarriers; tcgen05.commit releases stages and signals the
// epilogue.  sm_100a only.
#include "dv3_tc.cuh"

namespace dv3 {

constexpr int UM_BM = 128, UM_BK = 32;            // BK floats = 128 bytes = one swizzle row
constexpr int UM_STAGES = 3;
constexpr int UM_THREADS = 192;

struct UmmaEpi {
  float* C;
  const float* bias;
  const float* addend;
  int ldc, ldadd, M, N, K, accumulate;
};

// The tensor core adds into the fp32 accumulator with truncation, so a long accumulation chain
// drifts (measured: error grows linearly in K, ~1e-5 at K=1024).  Remedy ("promotion"):
//   * the two small cross terms go to their own accumulator (they are 2^-10 of the result, their
//     truncation is negligible),
//   * the hi*hi term is accumulated in TMEM only over chunks of UM_CH k-blocks (16 MMAs), two
//     chunk accumulators ping-pong, and the epilogue warps add the finished chunks in fp32
//     registers (round-to-nearest) while the next chunk is being computed.
constexpr int UM_CH = 4;   // k-blocks (of 32 floats) per chunk accumulator

template <int BN>
__global__ void __launch_bounds__(UM_THREADS, 1)
umma_gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap mAh,
                        const __grid_constant__ CUtensorMap mAl,
                        const __grid_constant__ CUtensorMap mBh,
                        const __grid_constant__ CUtensorMap mBl, UmmaEpi e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t A_BYTES = UM_BM * UM_BK * 4;       // 16 KB
  constexpr uint32_t B_BYTES = BN * UM_BK * 4;
  constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = (3 * BN <= 256) ? 256 : 512;   // acc0 | acc1 | accLo
  // dynamic smem is only guaranteed 16B-aligned: round up to the 1024B the swizzle needs
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + UM_STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * UM_STAGES + 4);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + UM_STAGES),
                 tfull0 = smem_u32(bars + 2 * UM_STAGES), tempty0 = smem_u32(bars + 2 * UM_STAGES + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * UM_BM, n0 = blockIdx.x * BN;
  const int nk = (e.K + UM_BK - 1) / UM_BK;
  const int nchunks = (nk + UM_CH - 1) / UM_CH;

  if (threadIdx.x == 0) {
    for (int s = 0; s < UM_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 4);      // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % UM_STAGES;
        const uint32_t ph = (kb / UM_STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
        mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
        tma_load_2d(base, &mAh, full0 + 8 * s, kb * UM_BK, m0);
        tma_load_2d(base + A_BYTES, &mAl, full0 + 8 * s, kb * UM_BK, m0);
        tma_load_2d(base + 2 * A_BYTES, &mBh, full0 + 8 * s, kb * UM_BK, n0);
        tma_load_2d(base + 2 * A_BYTES + B_BYTES, &mBl, full0 + 8 * s, kb * UM_BK, n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // c=F32, a=b=TF32, K-major both, N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(UM_BM >> 4) << 24);
      const uint32_t acc_lo = tmem_base + 2 * BN;
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % UM_STAGES;
        const uint32_t ph = (kb / UM_STAGES) & 1;
        const int c = kb / UM_CH, buf = c & 1, kin = kb % UM_CH;
        if (kin == 0) {   // this chunk accumulator must have been drained by the epilogue
          mbar_wait(tempty0 + 8 * buf, ((c >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        mbar_wait(full0 + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t acc_hi = tmem_base + buf * BN;
#pragma unroll
        for (int k = 0; k < UM_BK / 8; ++k) {
          const uint64_t ah = umma_desc(base + k * 32);
          const uint64_t al = umma_desc(base + A_BYTES + k * 32);
          const uint64_t bh = umma_desc(base + 2 * A_BYTES + k * 32);
          const uint64_t bl = umma_desc(base + 2 * A_BYTES + B_BYTES + k * 32);
          umma_tf32(acc_lo, al, bh, idesc, (kb | k) != 0);
          umma_tf32(acc_lo, ah, bl, idesc, 1);
          umma_tf32(acc_hi, ah, bh, idesc, (kin | k) != 0);
        }
        umma_commit(empty0 + 8 * s);   // stage free once these MMAs have read it
        if (kin == UM_CH - 1 || kb == nk - 1) umma_commit(tfull0 + 8 * buf);   // chunk complete
      }
    }
    __syncwarp();
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; thread = one output row
    const int lane_base = (warp & 3) * 32;
    const int row = m0 + lane_base + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)lane_base << 16);
    float sum[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) sum[j] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(tfull0 + 8 * buf, (c >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        DV3_TMEM_LD32(v, lane_addr + (uint32_t)(buf * BN + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
    // the last chunk commit also covers the cross-term accumulator
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      DV3_TMEM_LD32(v, lane_addr + (uint32_t)(2 * BN + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(v[j]);
    }
    if (row < e.M) {
      float* crow = e.C + (size_t)row * e.ldc;
      const float* arow = e.addend ? e.addend + (size_t)row * e.ldadd : nullptr;
      const bool vec = ((e.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.C) & 15) == 0) &&
                       (n0 + BN <= e.N) && !e.accumulate && !arow;
      if (vec) {
#pragma unroll
        for (int j = 0; j < BN; j += 4) {
          float4 o = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
          if (e.bias) {
            o.x += e.bias[n0 + j]; o.y += e.bias[n0 + j + 1];
            o.z += e.bias[n0 + j + 2]; o.w += e.bias[n0 + j + 3];
          }
          *reinterpret_cast<float4*>(crow + n0 + j) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < BN; ++j) {
          const int col = n0 + j;
          if (col < e.N) {
            float r = sum[j];
            if (e.bias) r += e.bias[col];
            if (arow) r += arow[col];
            if (e.accumulate) r += crow[col];
            crow[col] = r;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_base), "n"(TMEM_COLS)
                 : "memory");
  }
}

// hi = x with the low 13 mantissa bits cleared, lo = x - hi; two row segments packed side by side
// (output rows have pitch Kp >= K1+K2, zero padded, so that the TMA row pitch is 16B-aligned)
__global__ void split_tf32_kernel(const float* __restrict__ a1, int ld1, int K1,
                                  const float* __restrict__ a2, int ld2, int K2, int M, int Kp,
                                  float* __restrict__ hi, float* __restrict__ lo) {
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
  split_tf32_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(a1, ld1, K1, a2, ld2, a2 ? K2 : 0, M,
                                                              Kp, hi, lo);
  DV3_CHECK_LAUNCH("split_tf32_kernel");
  return 0;
}

// ---- host: tensor maps through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_map(CUtensorMap* m, const float* base, int rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  DV3_REQUIRE(fn, DV3_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {(cuuint32_t)UM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DV3_REQUIRE(r == CUDA_SUCCESS, DV3_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d",
              (int)r, rows, K);
  return 0;
}

template <int BN>
static int launch_umma(const float* Ah, const float* Al, const float* Bh, const float* Bl,
                       const UmmaEpi& e, cudaStream_t st) {
  CUtensorMap mAh, mAl, mBh, mBl;
  DV3_TRY(make_map(&mAh, Ah, e.M, e.K, UM_BM));
  DV3_TRY(make_map(&mAl, Al, e.M, e.K, UM_BM));
  DV3_TRY(make_map(&mBh, Bh, e.N, e.K, BN));
  DV3_TRY(make_map(&mBl, Bl, e.N, e.K, BN));
  constexpr size_t smem = (size_t)UM_STAGES * (2 * UM_BM + 2 * BN) * UM_BK * 4 + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    DV3_CHECK_CUDA(cudaFuncSetAttribute(umma_gemm_3xtf32_kernel<BN>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid((e.N + BN - 1) / BN, (e.M + UM_BM - 1) / UM_BM);
  const bool prof = prof_on();
  if (prof) prof_begin(st);
  umma_gemm_3xtf32_kernel<BN><<<grid, UM_THREADS, smem, st>>>(mAh, mAl, mBh, mBl, e);
  if (prof) prof_end(st, 1, 2.0 * e.M * e.N * e.K);
  DV3_CHECK_LAUNCH("umma_gemm_3xtf32_kernel");
  return 0;
}

// C = A W^T from pre-split dense operands ([M,K] and [N,K], K % 4 == 0)
int tc_gemm(const float* Ah, const float* Al, const float* Wh, const float* Wl, const float* bias,
            const float* addend, int ldadd, float* C, int ldc, int M, int N, int K, int accumulate,
            cudaStream_t st) {
  DV3_REQUIRE(K % 4 == 0 && K > 0, DV3_ERR_BAD_SHAPE, "tc_gemm: K=%d must be a multiple of 4", K);
  UmmaEpi e{C, bias, addend, ldc, ldadd, M, N, K, accumulate};
  // narrow N tiles when the grid would not cover the machine
  const int tiles128 = ((N + 127) / 128) * ((M + 127) / 128);
  if (tiles128 >= 96 || N <= 64) {
    if (N <= 64) return launch_umma<64>(Ah, Al, Wh, Wl, e, st);
    return launch_umma<128>(Ah, Al, Wh, Wl, e, st);
  }
  return launch_umma<64>(Ah, Al, Wh, Wl, e, st);
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
