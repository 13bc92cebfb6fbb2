// 3xTF32 GEMM with the A operand split in the SM and fed to the tensor core from TENSOR MEMORY:
//     C[M,N] = [A1 | A2][M, K1+K2] B[N, K1+K2]^T  (+bias) (+addend)
// A is plain fp32 ([M, K] row-major); B comes as its tf32 hi / lo planes (weights: split once per
// optimizer step).  TMA brings the fp32 A tile (16 KB per 32-wide k-block instead of the 32 KB of
// hi / lo planes the pre-split kernel, dv3_umma2.cu, streams); eight converter warps read it once
// (thread = (row, 16 of the 32 k-columns), the next k-block's loads in flight under the stores),
// split it into hi = x & 0xFFFFE000, lo = x - hi and write both with tcgen05.st into TMEM (lane =
// row, column = k); tcgen05.mma then takes A from TMEM and only B from shared memory:
//     acc[0,BN)    += A_hi(tmem) [B_hi; B_lo]^T   (one MMA of width 2 BN: hi*hi | hi*lo)
//     acc[BN,2BN)  += A_lo(tmem) B_hi^T
// Accumulation, chunking and epilogue are those of dv3_umma2.cu (same chunk boundaries, same order:
// results are bit-identical to the pre-split kernel on the same tile).
//
// What bounds the skinny products (M = 1024, single wave), measured while building this kernel:
//  * scratch/probe/mma_dep.cu: a tf32 MMA of M = 128, K = 8 occupies the tensor pipe for
//    max(46, N / 2) clk whatever its accumulator (dependent chain or not) and wherever A lives, and a
//    tcgen05.commit behind 8 MMAs adds ~180 clk (pipe drain): 8 x 46 + 180 = 550 clk per k-block of
//    a 128 x 32 tile in the best case.
//  * a timeline of one CTA (globaltimer stamps per k-block): the single TMA-producing thread gets
//    its 3-4 boxes per k-block accepted only every ~300 ns even while every stage is free, and they
//    land ~900 ns later -- the per-SM TMA path at this box pattern (128-byte rows) is the pacer at
//    0.30-0.33 us per k-block, for the pre-split kernel (40 KB per k-block) and for this one (24 KB).
//  * variants that did not move that pace: two MMA issuer warps on alternate chunks; stage release
//    in pairs (one commit point per two k-blocks); four instead of eight converter warps.  Two
//    producer threads (A tiles | B tiles) are 8 % slower -- the TMA unit, not the issuing thread, is
//    the limit -- and loading A with per-thread LDG.128 instead of TMA is 2x slower (per-lane rows,
//    latency-bound).
// Net: 8-13 % faster than the pre-split kernel on the in-loop shapes (1024 x 512 x 512: 9.5 -> 8.6 us,
// the GRU product 1024 x 1536 x 1024: 22.8 -> 20.2 us with single-wave 128 x 96 tiles); used for
// single-wave launches only -- multi-wave products are tensor-pipe bound and stay on CTA pairs.
//
// Roles (512 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 4-7 epilogue,
// warps 8-15 converters (warp w owns TMEM lanes 32 (w % 4) .. and k-columns 16 ((w - 8) / 4) ..).
// TMEM columns: [0, 4 BN) two chunk accumulators of 2 BN | AT stages of 64 columns (hi 32 | lo 32).
#include <cstdlib>
#include "dv3_tc.cuh"

namespace dv3 {

int make_map2(CUtensorMap* m, const float* base, int rows, int K, int ld, int box_rows, bool mn);

constexpr int GT_BM = 128, GT_BK = 32, GT_THREADS = 512, GT_CH = 4;
constexpr int GT_EPI_WARP0 = 4, GT_CONV_WARP0 = 8;
constexpr uint32_t GT_A_BYTES = GT_BM * GT_BK * 4;     // 16 KB, fp32

template <int BN>
struct GtCfg {
  static constexpr uint32_t B_BYTES = BN * GT_BK * 4;
  static constexpr uint32_t STAGE_BYTES = GT_A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (int)(196608 / STAGE_BYTES) > 8 ? 8 : (int)(196608 / STAGE_BYTES);
  static constexpr int AT = (512 - 4 * BN) / 64 > 4 ? 4 : (512 - 4 * BN) / 64;   // TMEM A stages
  static constexpr uint32_t A_COL0 = 4 * BN;
  static constexpr int EPI_WARPS = 4;
  static constexpr int NBAR = 2 * STAGES + 2 * AT + 4;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + NBAR * 8 + 64;
  static_assert(AT >= 2, "TMEM: accumulators + at least two A stages");
};

struct GtMaps {
  CUtensorMap a1, a2, bh, bl;
};

struct GtArgs {
  float* C;
  const float* bias;
  const float* addend;
  int ldc, ldadd, M, N, K1;
  int nk1, nk;
  int tiles_n, tiles_m, tiles;
  int m_fast;
};

#define DV3_TMEM_ST16(taddr, v)                                                                   \
  asm volatile(                                                                                   \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                             \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"                  \
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),  \
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),          \
        "r"(v[14]), "r"(v[15])                                                                    \
      : "memory")

// D[tmem] (+)= A[tmem] B[smem]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t db,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_c), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BN, bool BMN>
__global__ void __launch_bounds__(GT_THREADS, 1)
umma2t_gemm_kernel(const __grid_constant__ GtMaps mp, GtArgs g) {
  using Cfg = GtCfg<BN>;
  constexpr int ST = Cfg::STAGES, AT = Cfg::AT;
  constexpr uint32_t B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t gt_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gt_smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * ST, aready0 = empty0 + 8 * ST,
                 afree0 = aready0 + 8 * AT, tfull0 = afree0 + 8 * AT, tempty0 = tfull0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = g.nk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < AT; ++a) {
      mbar_init(aready0 + 8 * a, 8);       // one arrive per converter warp
      mbar_init(afree0 + 8 * a, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, Cfg::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mp.a1); prefetch_tensormap(&mp.bh); prefetch_tensormap(&mp.bl);
    if (g.nk1 < g.nk) prefetch_tensormap(&mp.a2);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x) {
        const int tm = g.m_fast ? tile % g.tiles_m : tile / g.tiles_n;
        const int tn = g.m_fast ? tile / g.tiles_m : tile % g.tiles_n;
        const int m0 = tm * GT_BM, n0 = tn * BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % ST;
          mbar_wait(empty0 + 8 * s, ((it / ST) & 1) ^ 1);
          const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t bar = full0 + 8 * s;
          mbar_expect_tx(bar, STAGE_BYTES);
          const bool seg1 = kb < g.nk1;
          const int ak = seg1 ? kb * GT_BK : (kb - g.nk1) * GT_BK;
          const int wk = seg1 ? ak : g.K1 + ak;
          tma_load_2d(base, seg1 ? &mp.a1 : &mp.a2, bar, ak, m0);
          if (!BMN) {
            tma_load_2d(base + GT_A_BYTES, &mp.bh, bar, wk, n0);
            tma_load_2d(base + GT_A_BYTES + B_BYTES, &mp.bl, bar, wk, n0);
          } else {
            for (int jb = 0; jb < BN / 32; ++jb) {
              tma_load_2d(base + GT_A_BYTES + jb * 4096, &mp.bh, bar, n0 + 32 * jb, wk);
              tma_load_2d(base + GT_A_BYTES + B_BYTES + jb * 4096, &mp.bl, bar, n0 + 32 * jb, wk);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)BMN << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
    constexpr uint32_t idesc_w = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(2 * BN >> 3) << 17);
    constexpr uint32_t B_KU = (BMN ? 1024 : 32) >> 4, ST_U = STAGE_BYTES >> 4;
    const bool issuer = elect_one();
    const uint32_t unit0 = (smem_u32(smem) >> 4) & 0x3FFF;
    int it = 0, cc = 0;
    for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x) {
      int buf = 0;
      for (int kb = 0; kb < nk; ++kb, ++it) {
        const int s = it % ST, ta = it % AT;
        const int kin = kb % GT_CH;
        if (kin == 0) {
          buf = cc & 1;
          mbar_wait(tempty0 + 8 * buf, ((cc >> 1) & 1) ^ 1);
        }
        mbar_wait(full0 + 8 * s, (it / ST) & 1);           // B tiles (async-proxy writes) visible here
        mbar_wait(aready0 + 8 * ta, (it / AT) & 1);        // A hi / lo of this k-block are in TMEM
        tc_fence_after();
        const uint32_t bu = unit0 + s * ST_U + (GT_A_BYTES >> 4);
        const uint32_t acc = tmem_base + buf * (2 * BN);
        const uint32_t a_hi = tmem_base + Cfg::A_COL0 + ta * 64, a_lo = a_hi + 32;
        const bool last = kin == GT_CH - 1 || kb == nk - 1;
        if (issuer) {
#pragma unroll
          for (int k = 0; k < GT_BK / 8; ++k) {
            const uint64_t bh = umma_desc_units<BMN>(bu + k * B_KU);    // [B_hi; B_lo] when 2 BN wide
            umma_tf32_ts(acc, a_hi + 8 * k, bh, idesc_w, (kin | k) != 0);
            umma_tf32_ts(acc + BN, a_lo + 8 * k, bh, idesc, 1);
          }
          umma_commit(empty0 + 8 * s);
          umma_commit(afree0 + 8 * ta);
          if (last) umma_commit(tfull0 + 8 * buf);
        }
        if (last) ++cc;
        __syncwarp();
      }
    }
  } else if (warp >= GT_CONV_WARP0) {
    // ------------------------------ converters --------------------------------
    // thread = (row, 16 of the 32 k-columns): 4 x LDS.128 (the tile is 128B-swizzled: chunk c of row
    // r sits at c ^ (r & 7)), split, two tcgen05.st of 16 columns.  The loads of k-block i+1 are
    // issued before the stores of k-block i are awaited.
    const int q = warp & 3, ch = (warp - GT_CONV_WARP0) >> 2, r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::A_COL0 + 16 * ch;
    const int my_tiles = (g.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * nk;
    auto load = [&](int it, float4 (&v)[4]) {
      const int s = it % ST;
      mbar_wait(full0 + 8 * s, (it / ST) & 1);
      const uint8_t* arow = smem + s * STAGE_BYTES + r * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        v[c] = *reinterpret_cast<const float4*>(arow + (((4 * ch + c) ^ (r & 7)) << 4));
    };
    float4 cur[4], nxt[4];
    if (total > 0) load(0, cur);
    for (int it = 0; it < total; ++it) {
      const int ta = it % AT;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float e[4] = {cur[c].x, cur[c].y, cur[c].z, cur[c].w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t hb = __float_as_uint(e[u]) & 0xFFFFE000u;
          hi[4 * c + u] = hb;
          lo[4 * c + u] = __float_as_uint(e[u] - __uint_as_float(hb));
        }
      }
      mbar_wait(afree0 + 8 * ta, ((it / AT) & 1) ^ 1);
      tc_fence_after();
      DV3_TMEM_ST16(lane_addr + (uint32_t)(ta * 64), hi);
      DV3_TMEM_ST16(lane_addr + (uint32_t)(ta * 64 + 32), lo);
      if (it + 1 < total) load(it + 1, nxt);              // in flight under the stores
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(aready0 + 8 * ta);
#pragma unroll
      for (int c = 0; c < 4; ++c) cur[c] = nxt[c];
    }
  } else if (warp >= GT_EPI_WARP0 && warp < GT_EPI_WARP0 + Cfg::EPI_WARPS) {
    // ------------------------------ epilogue ----------------------------------
    constexpr int HW = BN;                                 // an epilogue thread drains its whole row
    const int q = warp & 3, half = 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int cc = 0;
    for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x) {
      const int nchunks = (nk + GT_CH - 1) / GT_CH;
      const int tm = g.m_fast ? tile % g.tiles_m : tile / g.tiles_n;
      const int tn = g.m_fast ? tile / g.tiles_m : tile % g.tiles_n;
      const int m0 = tm * GT_BM, n0 = tn * BN + half * HW;
      const int row = m0 + q * 32 + lane;
      float sum[HW];
#pragma unroll
      for (int j = 0; j < HW; ++j) sum[j] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++cc) {
        const int buf = cc & 1;
        mbar_wait(tfull0 + 8 * buf, (cc >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < HW; c0 += 16) {
          uint32_t v[16], w[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
                "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(lane_addr + (uint32_t)(buf * 2 * BN + c0)));
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]),
                "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]),
                "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
              : "r"(lane_addr + (uint32_t)(buf * 2 * BN + BN + c0)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) sum[c0 + j] += __uint_as_float(v[j]) + __uint_as_float(w[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      }
      if (row < g.M) {
        float* crow = g.C + (size_t)row * g.ldc;
        const float* arow = g.addend ? g.addend + (size_t)row * g.ldadd : nullptr;
        const bool vec = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                         (n0 + HW <= g.N) && !arow;
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
              float rr = sum[j];
              if (g.bias) rr += g.bias[col];
              if (arow) rr += arow[col];
              crow[col] = rr;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512)
                 : "memory");
  }
}

// ---- host ------------------------------------------------------------------------------------
template <int BN, bool BMN>
static int launch_umma2t(const GtMaps& mp, GtArgs g, double flops, cudaStream_t st) {
  using Cfg = GtCfg<BN>;
  auto kern = umma2t_gemm_kernel<BN, BMN>;
  static DeviceOnce attr;
  if (attr.need())
    DV3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)Cfg::SMEM));
  g.tiles_n = (g.N + BN - 1) / BN;
  g.tiles_m = (g.M + GT_BM - 1) / GT_BM;
  g.tiles = g.tiles_n * g.tiles_m;
  g.m_fast = g.N > g.M ? 1 : 0;
  const int grid = g.tiles < sm_count() ? g.tiles : sm_count();
  const bool prof = prof_on();
  if (prof) prof_begin(st);
  DV3_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(GT_THREADS), Cfg::SMEM, st, mp, g));
  if (prof) prof_end(st, 1, flops);
  DV3_CHECK_LAUNCH("umma2t_gemm_kernel");
  return 0;
}

// tile width by the L2 -> SM ingest model: waves x k-blocks x (16 KB + BN / 4 KB) at 64 B/clk
static int pick_bn_t(int M, int N) {
  const int tm = (M + GT_BM - 1) / GT_BM, sms = sm_count();
  const int bns[3] = {32, 64, 96};
  long long best = -1;
  int bn = 32;
  for (int i = 0; i < 3; ++i) {
    if (bns[i] > 32 && N <= bns[i] / 2) continue;
    const long long tiles = (long long)tm * ((N + bns[i] - 1) / bns[i]);
    const long long waves = (tiles + sms - 1) / sms;
    const long long t = waves * (16384 + 256 * bns[i]) + (waves - 1) * 4096;
    if (best < 0 || t < best) { best = t; bn = bns[i]; }
  }
  return bn;
}

// the launch would be a single wave of 128 x BN tiles: where the raw-A kernel beats the pre-split
// one (multi-wave products are tensor-pipe bound and faster on CTA pairs / 128 x 128 tiles)
bool tc_gemm_rawa_single_wave(int M, int N) {
  const int bn = pick_bn_t(M, N);
  return (long long)((M + GT_BM - 1) / GT_BM) * ((N + bn - 1) / bn) <= sm_count();
}

// true when the shapes / alignments allow the raw-A kernel (K-major fp32 A, pre-split B planes)
bool tc_gemm_rawa_ok(const float* A1, int lda1, int K1, const float* A2, int lda2, int K2,
                     const TcOperand& B, int M, int N) {
  auto ok = [](const float* p, int ld) {
    return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0;
  };
  if (!ok(A1, lda1) || (A2 && !ok(A2, lda2)) || !ok(B.hi, B.ld) || !ok(B.lo, B.ld)) return false;
  if (M < TC_MIN_ROWS || N < 8 || K1 <= 0 || K1 % 4 != 0 || (A2 && (K2 <= 0 || K2 % 4 != 0))) return false;
  if (A2 && K1 % GT_BK != 0) return false;              // the B k-offset of segment 2 must be block aligned
  return true;
}

int tc_gemm_rawa(const float* A1, int lda1, int K1, const float* A2, int lda2, int K2,
                 const TcOperand& B, const float* bias, const float* addend, int ldadd, float* C,
                 int ldc, int M, int N, cudaStream_t st) {
  if (!A2) K2 = 0;
  DV3_REQUIRE(tc_gemm_rawa_ok(A1, lda1, K1, A2, lda2, K2, B, M, N), DV3_ERR_BAD_SHAPE,
              "tc_gemm_rawa: M=%d N=%d K1=%d K2=%d lda=%d ldb=%d", M, N, K1, K2, lda1, B.ld);
  const int K = K1 + K2;
  int BN = pick_bn_t(M, N);
  if (const char* f = DV3_ENV("DV3_TCT_FORCE")) {          // experiment knob: tile width
    const int fb = atoi(f);
    if (fb == 32 || fb == 64 || fb == 96) BN = fb;
  }
  if (B.mn && BN == 96) BN = 64;                           // MN-major B tiles come in 32-row boxes of 32 k: keep powers of two
  GtMaps mp;
  DV3_TRY(make_map2(&mp.a1, A1, M, K1, lda1, GT_BM, false));
  if (A2) DV3_TRY(make_map2(&mp.a2, A2, M, K2, lda2, GT_BM, false));
  else mp.a2 = mp.a1;
  DV3_TRY(make_map2(&mp.bh, B.hi, N, K, B.ld, BN, B.mn));
  DV3_TRY(make_map2(&mp.bl, B.lo, N, K, B.ld, BN, B.mn));
  GtArgs g{};
  g.C = C; g.bias = bias; g.addend = addend; g.ldc = ldc; g.ldadd = ldadd; g.M = M; g.N = N;
  g.K1 = K1; g.nk1 = (K1 + GT_BK - 1) / GT_BK; g.nk = g.nk1 + (K2 + GT_BK - 1) / GT_BK;
  const double flops = 2.0 * M * N * K;
  if (B.mn) {
    if (BN == 64) return launch_umma2t<64, true>(mp, g, flops, st);
    return launch_umma2t<32, true>(mp, g, flops, st);
  }
  if (BN == 96) return launch_umma2t<96, false>(mp, g, flops, st);
  if (BN == 64) return launch_umma2t<64, false>(mp, g, flops, st);
  return launch_umma2t<32, false>(mp, g, flops, st);
}

}  // namespace dv3

static dv3::TcOperand to_op_t(const dv3_tc_operand* o) {
  dv3::TcOperand r{};
  if (o) { r.hi = o->hi; r.lo = o->lo; r.ld = o->ld; r.mn = o->mn_major != 0; }
  return r;
}

// C ABI (see include/dv3_b200.h)
extern "C" int dv3_gemm_tc_rawa(const float* A1, int32_t lda1, int32_t K1, const float* A2,
                                int32_t lda2, int32_t K2, const dv3_tc_operand* B, const float* bias,
                                const float* addend, int32_t ldadd, float* C, int32_t ldc, int32_t M,
                                int32_t N, void* stream) {
  using namespace dv3;
  DV3_REQUIRE(M >= 0 && N >= 0, DV3_ERR_BAD_SHAPE, "gemm_tc_rawa: M=%d N=%d", M, N);
  if (M == 0 || N == 0) return 0;
  DV3_REQUIRE(A1 && B && C && B->hi && B->lo, DV3_ERR_NULL, "gemm_tc_rawa: null pointer");
  return tc_gemm_rawa(A1, lda1, K1, A2, lda2, A2 ? K2 : 0, to_op_t(B), bias, addend, ldadd, C, ldc, M,
                      N, static_cast<cudaStream_t>(stream));
}
