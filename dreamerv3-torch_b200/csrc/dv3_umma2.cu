// fp32-accurate GEMM on the tcgen05 tensor cores, raw-operand version:
//     C[M,N] = [A1 | A2][M, K1+K2] W[N, K1+K2]^T  (+bias) (+addend) (+C)
// A1/A2/W are plain fp32 matrices with arbitrary (16-byte aligned) row strides; nothing is
// pre-processed on the host side of the launch.
//
// 3xTF32: x = hi + lo with hi = x & 0xFFFFE000 (exact in tf32) and lo = x - hi (exact in fp32);
//     A W^T ~= A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T      (dropped lo*lo term ~2^-22 relative)
// The split happens inside the SM: TMA brings each fp32 operand tile from L2 exactly once
// (128B-swizzled), four converter warps rewrite it in place as hi and write lo beside it, then the
// MMA warp issues the three kind::tf32 products.  Compared with loading pre-split operands this
// halves the L2 -> SM traffic, which is what bounds the 128 x BN tile (measured: the pre-split
// kernel moved 10.5 TB/s through L2 at 60 % tensor-pipe activity).
//
// Persistent: grid = min(tiles, SMs); every role loops over tile = blockIdx.x + i * gridDim.x.
//   warp 0      TMA producer (one lane)
//   warp 1      TMEM owner + tcgen05.mma issuer (one lane)
//   warps 4-11  epilogue: warp w drains TMEM lanes 32*(w%4).., column half (w-4)/4
//   warps 12-15 converters
// Accumulation (see dv3_umma.cu for the measurements behind it): the tensor core adds into the
// fp32 accumulator with truncation, so hi*hi is accumulated in TMEM only over chunks of G2_CH
// k-blocks; two chunk accumulators ping-pong and the epilogue warps sum finished chunks in fp32
// registers.  The cross terms use their own accumulator, double-buffered per tile so the global
// stores of tile i overlap the MMAs of tile i+1.
// TMEM columns: [0,BN) hi chunk 0 | [BN,2BN) hi chunk 1 | [2BN,3BN) lo tile even | [3BN,4BN) lo odd.
#include <cstdlib>
#include "dv3_tc.cuh"

namespace dv3 {

constexpr int G2_BM = 128, G2_BK = 32;
constexpr int G2_THREADS = 512;
constexpr int G2_CH = 4;                 // k-blocks per hi*hi chunk accumulator
constexpr int G2_CONV_WARP0 = 12;        // first converter warp
constexpr int G2_EPI_WARP0 = 4;          // first epilogue warp (8 of them)

struct Gemm2Args {
  float* C;
  const float* bias;
  const float* addend;
  int ldc, ldadd, M, N, K1;
  int nk1, nk;                           // k-blocks of segment 1 / total
  int tiles_n, tiles;
  int accumulate;
};

template <int BN>
struct G2Cfg {
  static constexpr int STAGES = (BN == 128) ? 3 : 4;
  static constexpr uint32_t A_BYTES = G2_BM * G2_BK * 4;            // 16 KB
  static constexpr uint32_t B_BYTES = BN * G2_BK * 4;
  static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr uint32_t TMEM_COLS = (4 * BN <= 256) ? 256 : 512;
  static constexpr int NBAR = 3 * STAGES + 6;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + NBAR * 8 + 64;
};

template <int BN, bool MASK_HI>
__global__ void __launch_bounds__(G2_THREADS, 1)
umma2_gemm_kernel(const __grid_constant__ CUtensorMap mA1, const __grid_constant__ CUtensorMap mA2,
                  const __grid_constant__ CUtensorMap mW, Gemm2Args g) {
  using Cfg = G2Cfg<BN>;
  constexpr int ST = Cfg::STAGES;
  constexpr uint32_t A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);
  const uint32_t full0 = smem_u32(bars), conv0 = full0 + 8 * ST, empty0 = conv0 + 8 * ST,
                 tfull0 = empty0 + 8 * ST, tempty0 = tfull0 + 16, lempty0 = tempty0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = g.nk;
  const int nchunks = (nk + G2_CH - 1) / G2_CH;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(conv0 + 8 * s, 4);       // one arrive per converter warp
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 8);     // one arrive per epilogue warp
      mbar_init(lempty0 + 8 * b, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x) {
        const int m0 = (tile / g.tiles_n) * G2_BM, n0 = (tile % g.tiles_n) * BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % ST;
          const uint32_t ph = (it / ST) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
          mbar_expect_tx(full0 + 8 * s, A_BYTES + B_BYTES);
          int wk;
          if (kb < g.nk1) {
            tma_load_2d(base, &mA1, full0 + 8 * s, kb * G2_BK, m0);
            wk = kb * G2_BK;
          } else {
            tma_load_2d(base, &mA2, full0 + 8 * s, (kb - g.nk1) * G2_BK, m0);
            wk = g.K1 + (kb - g.nk1) * G2_BK;
          }
          tma_load_2d(base + 2 * A_BYTES, &mW, full0 + 8 * s, wk, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(G2_BM >> 4) << 24);
      int it = 0, cc = 0, tl = 0;
      for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++tl) {
        const int lb = tl & 1;
        mbar_wait(lempty0 + 8 * lb, ((tl >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc_lo = tmem_base + (2 + lb) * BN;
        int buf = 0;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % ST;
          const uint32_t ph = (it / ST) & 1;
          const int kin = kb % G2_CH;
          if (kin == 0) {
            buf = cc & 1;
            mbar_wait(tempty0 + 8 * buf, ((cc >> 1) & 1) ^ 1);
            tc_fence_after();
          }
          mbar_wait(conv0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t acc_hi = tmem_base + buf * BN;
#pragma unroll
          for (int k = 0; k < G2_BK / 8; ++k) {
            const uint64_t ah = umma_desc(base + k * 32);
            const uint64_t al = umma_desc(base + A_BYTES + k * 32);
            const uint64_t bh = umma_desc(base + 2 * A_BYTES + k * 32);
            const uint64_t bl = umma_desc(base + 2 * A_BYTES + B_BYTES + k * 32);
            umma_tf32(acc_lo, al, bh, idesc, (kb | k) != 0);
            umma_tf32(acc_lo, ah, bl, idesc, 1);
            umma_tf32(acc_hi, ah, bh, idesc, (kin | k) != 0);
          }
          umma_commit(empty0 + 8 * s);
          if (kin == G2_CH - 1 || kb == nk - 1) {
            umma_commit(tfull0 + 8 * buf);
            ++cc;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= G2_CONV_WARP0) {
    // ------------------------------ converters --------------------------------
    const int ct = threadIdx.x - G2_CONV_WARP0 * 32;     // 0..127
    int it = 0;
    for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x) {
      for (int kb = 0; kb < nk; ++kb, ++it) {
        const int s = it % ST;
        const uint32_t ph = (it / ST) & 1;
        mbar_wait(full0 + 8 * s, ph);
        uint8_t* base = smem + s * STAGE_BYTES;
        float4* ah = reinterpret_cast<float4*>(base);
        float4* al = reinterpret_cast<float4*>(base + A_BYTES);
        float4* bh = reinterpret_cast<float4*>(base + 2 * A_BYTES);
        float4* bl = reinterpret_cast<float4*>(base + 2 * A_BYTES + B_BYTES);
        auto conv = [](float4* hi, float4* lo, int i) {
          const float4 v = hi[i];
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          lo[i] = l;
          if (MASK_HI) hi[i] = h;
        };
#pragma unroll
        for (int u = 0; u < (int)(A_BYTES / 16 / 128); ++u) conv(ah, al, ct + u * 128);
#pragma unroll
        for (int u = 0; u < (int)(B_BYTES / 16 / 128); ++u) conv(bh, bl, ct + u * 128);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(conv0 + 8 * s);
      }
    }
  } else if (warp >= G2_EPI_WARP0) {
    // ------------------------------ epilogue ----------------------------------
    constexpr int HW = BN / 2;                            // columns per epilogue thread
    const int q = warp & 3, half = (warp - G2_EPI_WARP0) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HW);
    int cc = 0, tl = 0;
    for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++tl) {
      const int m0 = (tile / g.tiles_n) * G2_BM, n0 = (tile % g.tiles_n) * BN + half * HW;
      const int row = m0 + q * 32 + lane;
      const int lb = tl & 1;
      float sum[HW];
#pragma unroll
      for (int j = 0; j < HW; ++j) sum[j] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++cc) {
        const int buf = cc & 1;
        mbar_wait(tfull0 + 8 * buf, (cc >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < HW; c0 += 32) {
          uint32_t v[32];
          DV3_TMEM_LD32(v, lane_addr + (uint32_t)(buf * BN + c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      }
      // the last chunk commit also covers the cross-term accumulator of this tile
#pragma unroll
      for (int c0 = 0; c0 < HW; c0 += 32) {
        uint32_t v[32];
        DV3_TMEM_LD32(v, lane_addr + (uint32_t)((2 + lb) * BN + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(lempty0 + 8 * lb);

      if (row < g.M) {
        float* crow = g.C + (size_t)row * g.ldc;
        const float* arow = g.addend ? g.addend + (size_t)row * g.ldadd : nullptr;
        const bool vec = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                         (n0 + HW <= g.N) && !g.accumulate && !arow;
        if (vec) {
#pragma unroll
          for (int j = 0; j < HW; j += 4) {
            float4 o = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
            if (g.bias) {
              o.x += __ldg(g.bias + n0 + j); o.y += __ldg(g.bias + n0 + j + 1);
              o.z += __ldg(g.bias + n0 + j + 2); o.w += __ldg(g.bias + n0 + j + 3);
            }
            *reinterpret_cast<float4*>(crow + n0 + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < HW; ++j) {
            const int col = n0 + j;
            if (col < g.N) {
              float r = sum[j];
              if (g.bias) r += g.bias[col];
              if (arow) r += arow[col];
              if (g.accumulate) r += crow[col];
              crow[col] = r;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_base), "n"(Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ---- host ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 encode_fn2() {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(p);
  }
  return fn;
}

// [rows, K] fp32, row stride ld floats; box = 32 floats (one 128B swizzle row) x box_rows
static int make_map2(CUtensorMap* m, const float* base, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn2 fn = encode_fn2();
  DV3_REQUIRE(fn, DV3_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)G2_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DV3_REQUIRE(r == CUDA_SUCCESS, DV3_ERR_CUDA,
              "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld);
  return 0;
}

static int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static int g_mask_hi = -1;   // DV3_TC_MASK_HI=0: leave the raw fp32 word as the "hi" operand

template <int BN>
static int launch_umma2(const CUtensorMap& mA1, const CUtensorMap& mA2, const CUtensorMap& mW,
                        Gemm2Args g, double flops, cudaStream_t st) {
  using Cfg = G2Cfg<BN>;
  if (g_mask_hi < 0) {
    const char* e = getenv("DV3_TC_MASK_HI");
    g_mask_hi = (e && e[0] == '0') ? 0 : 1;
  }
  static bool attr = false;
  if (!attr) {
    DV3_CHECK_CUDA(cudaFuncSetAttribute(umma2_gemm_kernel<BN, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    DV3_CHECK_CUDA(cudaFuncSetAttribute(umma2_gemm_kernel<BN, false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr = true;
  }
  g.tiles_n = (g.N + BN - 1) / BN;
  g.tiles = g.tiles_n * ((g.M + G2_BM - 1) / G2_BM);
  const int grid = g.tiles < sm_count() ? g.tiles : sm_count();
  const bool prof = prof_on();
  if (prof) prof_begin(st);
  if (g_mask_hi)
    umma2_gemm_kernel<BN, true><<<grid, G2_THREADS, Cfg::SMEM, st>>>(mA1, mA2, mW, g);
  else
    umma2_gemm_kernel<BN, false><<<grid, G2_THREADS, Cfg::SMEM, st>>>(mA1, mA2, mW, g);
  if (prof) prof_end(st, 1, flops);
  DV3_CHECK_LAUNCH("umma2_gemm_kernel");
  return 0;
}

static bool tma_ok(const float* p, int ld) {
  return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0;
}

bool tc_gemm_raw_ok(const float* A1, int lda1, const float* A2, int lda2, const float* W, int ldw) {
  return tma_ok(A1, lda1) && (!A2 || tma_ok(A2, lda2)) && tma_ok(W, ldw);
}

// C = [A1|A2] W^T (+bias +addend) straight from fp32 operands.  A2 may be NULL (then K2 ignored).
int tc_gemm_raw(const float* A1, int lda1, int K1, const float* A2, int lda2, int K2,
                const float* W, int ldw, const float* bias, const float* addend, int ldadd,
                float* C, int ldc, int M, int N, int accumulate, cudaStream_t st) {
  if (!A2) K2 = 0;
  DV3_REQUIRE(M > 0 && N > 0 && K1 > 0 && K2 >= 0, DV3_ERR_BAD_SHAPE,
              "tc_gemm_raw: M=%d N=%d K1=%d K2=%d", M, N, K1, K2);
  DV3_REQUIRE(tc_gemm_raw_ok(A1, lda1, A2, lda2, W, ldw), DV3_ERR_BAD_SHAPE,
              "tc_gemm_raw: operands must be 16-byte aligned with row strides %% 4 == 0 "
              "(lda1=%d lda2=%d ldw=%d)", lda1, lda2, ldw);
  const int K = K1 + K2;
  // narrow N tiles when 128-wide ones would leave most of the machine idle
  const int tiles128 = ((N + 127) / 128) * ((M + G2_BM - 1) / G2_BM);
  const bool wide = N > 64 && tiles128 >= 96;
  const int BN = wide ? 128 : 64;
  CUtensorMap mA1, mA2, mW;
  DV3_TRY(make_map2(&mA1, A1, M, K1, lda1, G2_BM));
  if (A2) DV3_TRY(make_map2(&mA2, A2, M, K2, lda2, G2_BM));
  else mA2 = mA1;
  DV3_TRY(make_map2(&mW, W, N, K, ldw, BN));
  Gemm2Args g{};
  g.C = C; g.bias = bias; g.addend = addend; g.ldc = ldc; g.ldadd = ldadd; g.M = M; g.N = N;
  g.K1 = K1; g.nk1 = (K1 + G2_BK - 1) / G2_BK; g.nk = g.nk1 + (K2 + G2_BK - 1) / G2_BK;
  g.accumulate = accumulate;
  const double flops = 2.0 * M * N * K;
  if (wide) return launch_umma2<128>(mA1, mA2, mW, g, flops, st);
  return launch_umma2<64>(mA1, mA2, mW, g, flops, st);
}

}  // namespace dv3

// C ABI: tensor-core Linear straight from fp32 operands (no scratch).
extern "C" int dv3_linear_tc2_fwd(const float* A1, int32_t lda1, int32_t K1, const float* A2,
                                  int32_t lda2, int32_t K2, const float* W, int32_t ldw,
                                  const float* bias, const float* addend, int32_t ldadd, float* C,
                                  int32_t ldc, int32_t M, int32_t N, int32_t accumulate,
                                  void* stream) {
  using namespace dv3;
  DV3_REQUIRE(M >= 0 && N >= 0, DV3_ERR_BAD_SHAPE, "linear_tc2_fwd: M=%d N=%d", M, N);
  if (M == 0 || N == 0) return 0;
  DV3_REQUIRE(A1 && W && C, DV3_ERR_NULL, "linear_tc2_fwd: null pointer");
  return tc_gemm_raw(A1, lda1, K1, A2, lda2, K2, W, ldw, bias, addend, ldadd, C, ldc, M, N,
                     accumulate, static_cast<cudaStream_t>(stream));
}
