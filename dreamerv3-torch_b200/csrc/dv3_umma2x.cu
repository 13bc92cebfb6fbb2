// The tcgen05 3xTF32 GEMM of dv3_umma2.cu on CTA pairs (cta_group::2):
//     C[M,N] = [A1 | A2] B^T  (+bias) (+addend) (+C),   pre-split hi/lo operand planes
//
// Two CTAs of a cluster (two SMs of one TPC) compute one 256 x BN tile.  Each CTA stages its own
// 128 rows of A and only HALF of the B tile (BN/2 rows); the MMA, issued by the leader CTA alone,
// reads A from both CTAs' shared memory for the two row halves and both B halves for all 256
// rows.  Per 128 x BN of output that cuts the shared-memory traffic that bounds the single-CTA
// kernel (TMA writes + UMMA operand reads at 128 B/clk): BN=128: 160 -> 120 KB per k-block,
// i.e. 1250 -> 940 clk against 768 clk of tensor time.
//
// Pair protocol (rank 0 = leader):
//   * TMA: both CTAs load into their own stage; every load signals the LEADER's full barrier
//     (.cta_group::2, barrier address with the peer bit cleared); the leader posts the expected
//     byte count of both CTAs.
//   * MMA: leader lane issues tcgen05.mma.cta_group::2 (M = 256); tcgen05.commit multicasts the
//     "stage free" / "chunk complete" arrivals to the same barrier offsets in both CTAs.
//   * epilogue: each CTA drains its own TMEM lanes (its 128 rows); "accumulator drained" arrivals
//     of both CTAs go to the leader's barriers (remote mbarrier.arrive).
//   * TMEM is allocated / freed with cta_group::2 by the same warp of both CTAs; cluster barriers
//     fence barrier initialisation and teardown.
// Accumulation scheme (chunk promotion of hi*hi, separate cross-term accumulator double-buffered
// per tile), operand orders (K-major / MN-major), K segments, split-K and tile order are those of
// the single-CTA kernel.
#include <cstdlib>
#include "dv3_tc.cuh"

namespace dv3 {

constexpr int X2_BM = 128, X2_BK = 32;     // per-CTA rows, floats per k-block
constexpr int X2_THREADS = 384;            // warp 0 TMA, warp 1 MMA/TMEM, warps 4-11 epilogue
constexpr int X2_CH = 4;
constexpr int X2_EPI_WARP0 = 4, X2_EPI_WARPS = 8;
constexpr uint32_t X2_PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address

struct GemmX2Args {
  float* C;
  const float* bias;
  const float* addend;
  int ldc, ldadd, M, N, K1;
  int nk1, nk;
  int tiles_n, tiles_m, tiles;   // tiles_m counts 256-row pair tiles
  int m_fast;
  int splitk, nkp, units;
  int accumulate;
};

struct GemmX2Maps {
  CUtensorMap a1h, a1l, a2h, a2l, bh, bl;
};

template <int BN>
struct X2Cfg {
  static constexpr int HB = BN / 2;                                  // B rows staged per CTA
  static constexpr int STAGES = (BN == 128) ? 4 : 5;
  static constexpr uint32_t A_BYTES = X2_BM * X2_BK * 4;             // 16 KB
  static constexpr uint32_t B_BYTES = HB * X2_BK * 4;                // 8 / 4 KB
  static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr uint32_t TMEM_COLS = (4 * BN <= 256) ? 256 : 512;
  static constexpr int NBAR = 2 * STAGES + 6;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + NBAR * 8 + 64;
};

// 2D TMA load into this CTA's smem, completion bytes credited to the pair leader's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map,
                                                 uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_c, uint64_t da, uint64_t db,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
// arrive on the leader CTA's copy of a barrier (local for rank 0, remote for rank 1)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & X2_PEER_MASK) : "memory");
}

template <bool MN>
__device__ __forceinline__ void load_tile_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                               int k0, int r0, int rows) {
  if (!MN) {
    tma_load_2d_pair(dst, map, bar, k0, r0);
  } else {
    for (int j = 0; j < rows / 32; ++j) tma_load_2d_pair(dst + j * 4096, map, bar, r0 + 32 * j, k0);
  }
}

template <int BN, bool AMN, bool BMN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(X2_THREADS, 1)
umma2x_gemm_kernel(const __grid_constant__ GemmX2Maps mp, GemmX2Args g) {
  using Cfg = X2Cfg<BN>;
  constexpr int ST = Cfg::STAGES, HB = Cfg::HB;
  constexpr uint32_t A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * ST, tfull0 = empty0 + 8 * ST,
                 tempty0 = tfull0 + 16, lempty0 = tempty0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nk = g.nk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(full0 + 8 * s, 1);                         // used in the leader only
      mbar_init(empty0 + 8 * s, 1);                        // one multicast commit per round
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 2 * X2_EPI_WARPS);        // leader: epilogue warps of both CTAs
      mbar_init(lempty0 + 8 * b, 2 * X2_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mp.a1h); prefetch_tensormap(&mp.a1l);
    prefetch_tensormap(&mp.bh); prefetch_tensormap(&mp.bl);
    if (g.nk1 < g.nk) { prefetch_tensormap(&mp.a2h); prefetch_tensormap(&mp.a2l); }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // peer barriers initialised
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int unit = pair; unit < g.units; unit += npairs) {
        const int tile = unit % g.tiles, kb0 = (unit / g.tiles) * g.nkp;
        const int kb1 = min(nk, kb0 + g.nkp);
        const int tm = g.m_fast ? tile % g.tiles_m : tile / g.tiles_n;
        const int tn = g.m_fast ? tile / g.tiles_m : tile % g.tiles_n;
        const int m0 = tm * (2 * X2_BM) + (int)rank * X2_BM, n0 = tn * BN + (int)rank * HB;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % ST;
          const uint32_t ph = (it / ST) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t bar = (full0 + 8 * s) & X2_PEER_MASK;      // the leader's barrier
          if (leader) mbar_expect_tx(full0 + 8 * s, 2 * STAGE_BYTES);
          const bool seg1 = kb < g.nk1;
          const int ak = seg1 ? kb * X2_BK : (kb - g.nk1) * X2_BK;
          const int wk = seg1 ? ak : g.K1 + ak;
          load_tile_pair<AMN>(base, seg1 ? &mp.a1h : &mp.a2h, bar, ak, m0, X2_BM);
          load_tile_pair<AMN>(base + A_BYTES, seg1 ? &mp.a1l : &mp.a2l, bar, ak, m0, X2_BM);
          load_tile_pair<BMN>(base + 2 * A_BYTES, &mp.bh, bar, wk, n0, HB);
          load_tile_pair<BMN>(base + 2 * A_BYTES + B_BYTES, &mp.bl, bar, wk, n0, HB);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) --------------------------
    // The warp stays converged; one elected lane issues the MMAs and commits.
    if (leader) {
      // c=F32, a=b=TF32; bit 15/16 = A/B MN-major; N>>3 at [17,23), M>>4 at [24,29): M = 256
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)AMN << 15) |
                                 ((uint32_t)BMN << 16) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)((2 * X2_BM) >> 4) << 24);
      constexpr uint32_t A_KU = (AMN ? 1024 : 32) >> 4, B_KU = (BMN ? 1024 : 32) >> 4;
      constexpr uint32_t A_PU = A_BYTES >> 4, B_PU = B_BYTES >> 4, ST_U = STAGE_BYTES >> 4;
      const bool issuer = elect_one();
      const uint32_t unit0 = (smem_u32(smem) >> 4) & 0x3FFF;
      int it = 0, cc = 0, tl = 0;
      for (int unit = pair; unit < g.units; unit += npairs, ++tl) {
        const int kb0 = (unit / g.tiles) * g.nkp, kb1 = min(nk, kb0 + g.nkp);
        const int lb = tl & 1;
        mbar_wait(lempty0 + 8 * lb, ((tl >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc_lo = tmem_base + (2 + lb) * BN;
        int buf = 0;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % ST;
          const uint32_t ph = (it / ST) & 1;
          const int kin = (kb - kb0) % X2_CH;
          if (kin == 0) {
            buf = cc & 1;
            mbar_wait(tempty0 + 8 * buf, ((cc >> 1) & 1) ^ 1);
          }
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t au = unit0 + s * ST_U, bu = au + 2 * A_PU;
          const uint32_t acc_hi = tmem_base + buf * BN;
          const bool last = kin == X2_CH - 1 || kb == kb1 - 1;
          if (issuer) {
#pragma unroll
            for (int k = 0; k < X2_BK / 8; ++k) {
              const uint64_t ah = umma_desc_units<AMN>(au + k * A_KU);
              const uint64_t al = umma_desc_units<AMN>(au + A_PU + k * A_KU);
              const uint64_t bh = umma_desc_units<BMN>(bu + k * B_KU);
              const uint64_t bl = umma_desc_units<BMN>(bu + B_PU + k * B_KU);
              umma_tf32_pair(acc_lo, al, bh, idesc, ((kb - kb0) | k) != 0);
              umma_tf32_pair(acc_lo, ah, bl, idesc, 1);
              umma_tf32_pair(acc_hi, ah, bh, idesc, (kin | k) != 0);
            }
            umma_commit_pair(empty0 + 8 * s);
            if (last) umma_commit_pair(tfull0 + 8 * buf);
          }
          if (last) ++cc;
          __syncwarp();
        }
      }
    }
  } else if (warp >= X2_EPI_WARP0) {
    // ------------------------------ epilogue (both CTAs, own 128 rows) --------------------
    constexpr int HW = BN / 2;
    const int q = warp & 3, half = (warp - X2_EPI_WARP0) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HW);
    int cc = 0, tl = 0;
    for (int unit = pair; unit < g.units; unit += npairs, ++tl) {
      const int tile = unit % g.tiles, ks = unit / g.tiles;
      const int nchunks = (min(nk, (ks + 1) * g.nkp) - ks * g.nkp + X2_CH - 1) / X2_CH;
      const int tm = g.m_fast ? tile % g.tiles_m : tile / g.tiles_n;
      const int tn = g.m_fast ? tile / g.tiles_m : tile % g.tiles_n;
      const int m0 = tm * (2 * X2_BM) + (int)rank * X2_BM, n0 = tn * BN + half * HW;
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
        if (lane == 0) mbar_arrive_leader(tempty0 + 8 * buf);
      }
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
      if (lane == 0) mbar_arrive_leader(lempty0 + 8 * lb);

      if (row < g.M && g.splitk > 1) {
        float* crow = g.C + (size_t)row * g.ldc;
        const float* arow = (g.addend && ks == 0) ? g.addend + (size_t)row * g.ldadd : nullptr;
#pragma unroll
        for (int j = 0; j < HW; ++j) {
          const int col = n0 + j;
          if (col < g.N) {
            float r = sum[j];
            if (g.bias && ks == 0) r += g.bias[col];
            if (arow) r += arow[col];
            atomicAdd(crow + col, r);
          }
        }
      } else if (row < g.M) {
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
  cluster_sync_all();                                       // nobody leaves while the pair is live
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_base), "n"(Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ---- host ------------------------------------------------------------------------------------
int make_map2(CUtensorMap* m, const float* base, int rows, int K, int ld, int box_rows, bool mn);

template <int BN, bool AMN, bool BMN>
static int launch_x2(const GemmX2Maps& mp, GemmX2Args g, double flops, cudaStream_t st) {
  using Cfg = X2Cfg<BN>;
  auto kern = umma2x_gemm_kernel<BN, AMN, BMN>;
  static DeviceOnce attr;
  if (attr.need())
    DV3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)Cfg::SMEM));
  g.tiles_n = (g.N + BN - 1) / BN;
  g.tiles_m = (g.M + 2 * X2_BM - 1) / (2 * X2_BM);
  g.tiles = g.tiles_n * g.tiles_m;
  g.m_fast = g.N > g.M ? 1 : 0;
  if (g.splitk < 1) g.splitk = 1;
  g.nkp = (g.nk + g.splitk - 1) / g.splitk;
  g.splitk = (g.nk + g.nkp - 1) / g.nkp;
  g.units = g.tiles * g.splitk;
  if (g.splitk > 1 && !g.accumulate)
    DV3_CHECK_CUDA(cudaMemset2DAsync(g.C, (size_t)g.ldc * 4, 0, (size_t)g.N * 4, g.M, st));
  const int max_pairs = sm_count() / 2;
  const int pairs = g.units < max_pairs ? g.units : max_pairs;
  const bool prof = prof_on();
  if (prof) prof_begin(st);
  kern<<<2 * pairs, X2_THREADS, Cfg::SMEM, st>>>(mp, g);
  if (prof) prof_end(st, 1, flops);
  DV3_CHECK_LAUNCH("umma2x_gemm_kernel");
  return 0;
}

template <int BN>
static int dispatch_x2(bool amn, bool bmn, const GemmX2Maps& mp, const GemmX2Args& g, double flops,
                       cudaStream_t st) {
  if (!amn && !bmn) return launch_x2<BN, false, false>(mp, g, flops, st);
  if (!amn && bmn) return launch_x2<BN, false, true>(mp, g, flops, st);
  if (amn && !bmn) return launch_x2<BN, true, false>(mp, g, flops, st);
  return launch_x2<BN, true, true>(mp, g, flops, st);
}

// Pair-tile GEMM from pre-split operands; BN = 128 or 64.  Same contract as tc_gemm_ops.
int tc_gemm_pair(const TcOperand& A1, int K1, const TcOperand* A2, int K2, const TcOperand& B,
                 const float* bias, const float* addend, int ldadd, float* C, int ldc, int M, int N,
                 int accumulate, int BN, int splitk, cudaStream_t st) {
  if (!A2) K2 = 0;
  const int K = K1 + K2;
  GemmX2Maps mp;
  DV3_TRY(make_map2(&mp.a1h, A1.hi, M, K1, A1.ld, X2_BM, A1.mn));
  DV3_TRY(make_map2(&mp.a1l, A1.lo, M, K1, A1.ld, X2_BM, A1.mn));
  if (A2) {
    DV3_TRY(make_map2(&mp.a2h, A2->hi, M, K2, A2->ld, X2_BM, A2->mn));
    DV3_TRY(make_map2(&mp.a2l, A2->lo, M, K2, A2->ld, X2_BM, A2->mn));
  } else {
    mp.a2h = mp.a1h; mp.a2l = mp.a1l;
  }
  DV3_TRY(make_map2(&mp.bh, B.hi, N, K, B.ld, BN / 2, B.mn));
  DV3_TRY(make_map2(&mp.bl, B.lo, N, K, B.ld, BN / 2, B.mn));
  GemmX2Args g{};
  g.C = C; g.bias = bias; g.addend = addend; g.ldc = ldc; g.ldadd = ldadd; g.M = M; g.N = N;
  g.K1 = K1; g.nk1 = (K1 + X2_BK - 1) / X2_BK; g.nk = g.nk1 + (K2 + X2_BK - 1) / X2_BK;
  g.accumulate = accumulate & 1;
  g.splitk = splitk;
  const double flops = 2.0 * M * N * K;
  if (BN == 128) return dispatch_x2<128>(A1.mn, B.mn, mp, g, flops, st);
  return dispatch_x2<64>(A1.mn, B.mn, mp, g, flops, st);
}

}  // namespace dv3
