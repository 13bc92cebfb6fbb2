// fp32-accurate GEMM on the tcgen05 tensor cores (persistent, warp-specialised):
//     C[M,N] = [A1 | A2][M, K1+K2] B[N, K1+K2]^T  (+bias) (+addend) (+C)
//
// 3xTF32: x = hi + lo with hi = x & 0xFFFFE000 (exact in tf32) and lo = x - hi (exact in fp32);
//     A B^T ~= A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T      (dropped lo*lo term ~2^-22 relative)
//
// Two operand modes:
//   * pre-split (the fast path): every operand is given as its (hi, lo) pair, in either storage
//     order -- "K-major" [rows, K] or "MN-major" [K, rows] -- selected through the UMMA
//     descriptors, so y = x W^T, dx = dy W and dW = dy^T x all read the same split arrays with
//     no transposes.
//   * raw (RAW=true): plain fp32 operands; TMA brings each tile once and four converter warps
//     split it in shared memory.  Halves L2 -> SM traffic, but measured 20-35 % slower: the
//     kernel is bound by shared-memory bandwidth (TMA writes + UMMA operand reads share
//     128 B/clk/SM: 160 KB per k-block pre-split vs 224 KB with the in-SM split), not by L2.
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
// WIDE (the default for BN <= 64 and single-wave BN = 128): two MMAs per k-step -- A_hi x [B_hi;B_lo]
// at width 2 BN, then A_lo x B_hi -- into two chunk accumulators of 2 BN columns each:
// [0,BN) hi*hi | [BN,2BN) cross terms | [2BN,4BN) the same for the other chunk.
// SK > 1: K partitioned over the CTAs of a cluster, partial tiles reduced through distributed
// shared memory (see the kernel's comment and DESIGN.md section 7 for when it pays).
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
  int tiles_n, tiles_m, tiles;
  int m_fast;                            // tile order: consecutive tiles walk M (1) or N (0)
  int splitk, nkp, units;                // K partitions, k-blocks per partition, tiles * splitk
  int accumulate;
  unsigned long long* stamps;            // debug: globaltimer stamps of CTA 0 (DV3_GEMM_TIMING=1)
};

#define G2_STAMP(slot)                                                      \
  do {                                                                      \
    if (g.stamps && blockIdx.x == 0) {                                      \
      unsigned long long t_;                                                \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                \
      g.stamps[slot] = t_;                                                  \
    }                                                                       \
  } while (0)

template <int BN>
struct G2Cfg {
  static constexpr int STAGES = (BN == 128) ? 3 : (BN == 64) ? 4 : 5;
  static constexpr uint32_t A_BYTES = G2_BM * G2_BK * 4;            // 16 KB
  static constexpr uint32_t B_BYTES = BN * G2_BK * 4;
  static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr uint32_t TMEM_COLS = (4 * BN <= 128) ? 128 : (4 * BN <= 256) ? 256 : 512;
  static constexpr int EPI_WARPS = (BN >= 64) ? 8 : 4;   // a thread drains >= 32 columns
  static constexpr int NBAR = 3 * STAGES + 6;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + NBAR * 8 + 64;
};

// MN-major tf32 operands only exist in the "128B swizzle, 32B atom" layout (layout type 1; TMA
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 32 MN-floats (128 B) per k, 32-byte chunks XORed
// with (k % 4), i.e. atoms of 4 k-rows = 512 B.  Canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in
// 16-byte units: SBO = 512 B between k-atoms, LBO = 4096 B between the 32-float-wide MN blocks
// (one TMA box [32 k x 32 mn] each).  The descriptor itself is umma_desc_units<true> (dv3_tc.cuh).

struct Gemm2Maps {
  CUtensorMap a1h, a1l, a2h, a2l, bh, bl;   // raw mode uses a1h, a2h, bh only
};

// Four consecutive columns of one result row: + bias + addend (`extras`: only K partition 0 adds
// them), then store, add into C, or -- split-K over CTAs -- fp32 atomics.
__device__ __forceinline__ void g2_emit4(const Gemm2Args& g, int row, int col, float4 v, bool extras,
                                         bool atomic, bool vec_ok) {
  if (row >= g.M || col >= g.N) return;
  float* cp = g.C + (size_t)row * g.ldc + col;
  const float* bp = (g.bias && extras) ? g.bias + col : nullptr;
  const float* ap = (g.addend && extras) ? g.addend + (size_t)row * g.ldadd + col : nullptr;
  if (vec_ok && col + 4 <= g.N) {
    if (bp) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bp));
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (ap) {
      const float4 a = *reinterpret_cast<const float4*>(ap);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    if (atomic) {
      atomicAdd(cp, v.x); atomicAdd(cp + 1, v.y); atomicAdd(cp + 2, v.z); atomicAdd(cp + 3, v.w);
    } else {
      if (g.accumulate) {
        const float4 c = *reinterpret_cast<const float4*>(cp);
        v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
      }
      *reinterpret_cast<float4*>(cp) = v;
    }
  } else {
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (col + j < g.N) {
        float r = e[j];
        if (bp) r += bp[j];
        if (ap) r += ap[j];
        if (atomic) {
          atomicAdd(cp + j, r);
        } else {
          if (g.accumulate) r += cp[j];
          cp[j] = r;
        }
      }
    }
  }
}

__device__ __forceinline__ bool g2_vec_ok(const Gemm2Args& g) {
  return ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
         (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) &&
         (!g.addend || ((g.ldadd & 3) == 0 && (reinterpret_cast<uintptr_t>(g.addend) & 15) == 0));
}

// one operand tile of `rows` rows for k-block starting at k0: K-major = one box [rows x 32 k],
// MN-major = rows/32 boxes [32 k x 32 rows], 4 KB apart
template <bool MN>
__device__ __forceinline__ void load_tile(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                          int k0, int r0, int rows) {
  if (!MN) {
    tma_load_2d(dst, map, bar, k0, r0);
  } else {
    for (int j = 0; j < rows / 32; ++j) tma_load_2d(dst + j * 4096, map, bar, r0 + 32 * j, k0);
  }
}

// WIDE: the B_hi and B_lo tiles sit back to back in a stage, so one MMA of width 2 BN computes
// A_hi [B_hi; B_lo]^T -- hi*hi into columns [0, BN) and hi*lo into [BN, 2 BN) -- and a second of
// width BN adds A_lo B_hi^T onto the latter: the A tile is read from shared memory twice per
// k-step instead of three times.  Both halves are then chunk accumulators (drained together).
// SK > 1: the launch is one wave of clusters of SK CTAs; the CTAs of a cluster take the K
// partitions of ONE tile and combine their partial sums through distributed shared memory in a
// fixed order (deterministic, no memset, no atomics): CTA r reduces columns [r BN/SK, (r+1) BN/SK).
// This lets the skinny M = 1024 products use 128 x 128 tiles (2.2x less operand traffic per
// flop than 128 x 32) and still fill the machine.
template <int BN, bool AMN, bool BMN, bool RAW, bool WIDE = false, int SK = 1>
__global__ void __launch_bounds__(G2_THREADS, 1)
umma2_gemm_kernel(const __grid_constant__ Gemm2Maps mp, Gemm2Args g) {
  constexpr bool MASK_HI = true;
  static_assert(SK == 1 || (!RAW && BN % (8 * SK) == 0), "cluster split-K: pre-split operands");
  // SK mode: exactly one work unit per CTA, unit = tile + partition * tiles like split-K
  const int unit_first = SK > 1 ? (int)(blockIdx.x / SK) + (int)(blockIdx.x % SK) * g.tiles
                                : (int)blockIdx.x;
  const int unit_step = SK > 1 ? g.units : (int)gridDim.x;
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
  if (threadIdx.x == 0) G2_STAMP(0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(conv0 + 8 * s, 4);       // one arrive per converter warp
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, Cfg::EPI_WARPS);     // one arrive per epilogue warp
      mbar_init(lempty0 + 8 * b, Cfg::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mp.a1h); prefetch_tensormap(&mp.bh);
    if (!RAW) { prefetch_tensormap(&mp.a1l); prefetch_tensormap(&mp.bl); }
    if (g.nk1 < g.nk) { prefetch_tensormap(&mp.a2h); if (!RAW) prefetch_tensormap(&mp.a2l); }
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
  // the prologue above overlapped the previous kernel's tail; its outputs (our operands, addend,
  // C when accumulating) may only be touched from here on
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) G2_STAMP(1);

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int unit = unit_first; unit < g.units; unit += unit_step) {
        const int tile = unit % g.tiles, kb0 = (unit / g.tiles) * g.nkp;
        const int kb1 = min(nk, kb0 + g.nkp);
        const int tm = g.m_fast ? tile % g.tiles_m : tile / g.tiles_n;
        const int tn = g.m_fast ? tile / g.tiles_m : tile % g.tiles_n;
        const int m0 = tm * G2_BM, n0 = tn * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % ST;
          const uint32_t ph = (it / ST) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t bar = full0 + 8 * s;
          mbar_expect_tx(bar, RAW ? (A_BYTES + B_BYTES) : STAGE_BYTES);
          const bool seg1 = kb < g.nk1;
          const int ak = seg1 ? kb * G2_BK : (kb - g.nk1) * G2_BK;
          const int wk = seg1 ? ak : g.K1 + ak;
          load_tile<AMN>(base, seg1 ? &mp.a1h : &mp.a2h, bar, ak, m0, G2_BM);
          load_tile<BMN>(base + 2 * A_BYTES, &mp.bh, bar, wk, n0, BN);
          if (!RAW) {
            load_tile<AMN>(base + A_BYTES, seg1 ? &mp.a1l : &mp.a2l, bar, ak, m0, G2_BM);
            load_tile<BMN>(base + 2 * A_BYTES + B_BYTES, &mp.bl, bar, wk, n0, BN);
          }
          if (it == 0) G2_STAMP(2);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    // The warp stays converged; one elected lane issues the MMAs and commits.
    {
      // c=F32, a=b=TF32; bit 15/16 = A/B MN-major; N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)AMN << 15) |
                                 ((uint32_t)BMN << 16) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(G2_BM >> 4) << 24);
      constexpr uint32_t A_KU = (AMN ? 1024 : 32) >> 4, B_KU = (BMN ? 1024 : 32) >> 4;
      constexpr uint32_t A_PU = A_BYTES >> 4, B_PU = B_BYTES >> 4, ST_U = STAGE_BYTES >> 4;
      const bool issuer = elect_one();
      const uint32_t unit0 = (smem_u32(smem) >> 4) & 0x3FFF;
      int it = 0, cc = 0, tl = 0;
      for (int unit = unit_first; unit < g.units; unit += unit_step, ++tl) {
        const int kb0 = (unit / g.tiles) * g.nkp, kb1 = min(nk, kb0 + g.nkp);
        const int lb = tl & 1;
        if (!WIDE) {
          mbar_wait(lempty0 + 8 * lb, ((tl >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t acc_lo = tmem_base + (2 + lb) * BN;
        int buf = 0;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % ST;
          const uint32_t ph = (it / ST) & 1;
          const int kin = (kb - kb0) % G2_CH;
          if (kin == 0) {
            buf = cc & 1;
            mbar_wait(tempty0 + 8 * buf, ((cc >> 1) & 1) ^ 1);
          }
          mbar_wait((RAW ? conv0 : full0) + 8 * s, ph);
          tc_fence_after();
          if (it == 0 && issuer) G2_STAMP(3);
          const uint32_t au = unit0 + s * ST_U, bu = au + 2 * A_PU;
          const uint32_t acc_hi = tmem_base + buf * BN;
          const bool last = kin == G2_CH - 1 || kb == kb1 - 1;
          if (issuer && WIDE) {
            constexpr uint32_t idesc_w = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(2 * BN >> 3) << 17);
            const uint32_t acc = tmem_base + buf * (2 * BN);
#pragma unroll
            for (int k = 0; k < G2_BK / 8; ++k) {
              const uint64_t ah = umma_desc_units<AMN>(au + k * A_KU);
              const uint64_t al = umma_desc_units<AMN>(au + A_PU + k * A_KU);
              const uint64_t bh = umma_desc_units<BMN>(bu + k * B_KU);    // [B_hi; B_lo] when 2 BN wide
              umma_tf32(acc, ah, bh, idesc_w, (kin | k) != 0);
              umma_tf32(acc + BN, al, bh, idesc, 1);
            }
          } else if (issuer) {
#pragma unroll
            for (int k = 0; k < G2_BK / 8; ++k) {
              const uint64_t ah = umma_desc_units<AMN>(au + k * A_KU);
              const uint64_t al = umma_desc_units<AMN>(au + A_PU + k * A_KU);
              const uint64_t bh = umma_desc_units<BMN>(bu + k * B_KU);
              const uint64_t bl = umma_desc_units<BMN>(bu + B_PU + k * B_KU);
              umma_tf32(acc_lo, al, bh, idesc, ((kb - kb0) | k) != 0);
              umma_tf32(acc_lo, ah, bl, idesc, 1);
              umma_tf32(acc_hi, ah, bh, idesc, (kin | k) != 0);
            }
          }
          if (issuer) {
            umma_commit(empty0 + 8 * s);
            if (last) umma_commit(tfull0 + 8 * buf);
            if (kb == kb1 - 1) G2_STAMP(4);
          }
          if (last) ++cc;
          __syncwarp();
        }
      }
    }
  } else if (RAW && warp >= G2_CONV_WARP0) {
    // ------------------------------ converters --------------------------------
    const int ct = threadIdx.x - G2_CONV_WARP0 * 32;     // 0..127
    int it = 0;
    for (int unit = unit_first; unit < g.units; unit += unit_step) {
      const int kb0 = (unit / g.tiles) * g.nkp, kb1 = min(nk, kb0 + g.nkp);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
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
  } else if (warp >= G2_EPI_WARP0 && warp < G2_EPI_WARP0 + Cfg::EPI_WARPS) {
    // ------------------------------ epilogue ----------------------------------
    constexpr int HW = BN / (Cfg::EPI_WARPS / 4);         // columns per epilogue thread
    const int q = warp & 3, half = (warp - G2_EPI_WARP0) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HW);
    int cc = 0, tl = 0;
    for (int unit = unit_first; unit < g.units; unit += unit_step, ++tl) {
      const int tile = unit % g.tiles, ks = unit / g.tiles;
      const int nchunks = (min(nk, (ks + 1) * g.nkp) - ks * g.nkp + G2_CH - 1) / G2_CH;
      const int tm = g.m_fast ? tile % g.tiles_m : tile / g.tiles_n;
      const int tn = g.m_fast ? tile / g.tiles_m : tile % g.tiles_n;
      const int m0 = tm * G2_BM, n0 = tn * BN + half * HW;
      const int row = m0 + q * 32 + lane;
      const int lb = tl & 1;
      float sum[HW];
#pragma unroll
      for (int j = 0; j < HW; ++j) sum[j] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++cc) {
        const int buf = cc & 1;
        mbar_wait(tfull0 + 8 * buf, (cc >> 1) & 1);
        tc_fence_after();
        if (c == nchunks - 1 && threadIdx.x == G2_EPI_WARP0 * 32) G2_STAMP(5);
#pragma unroll
        for (int c0 = 0; c0 < HW; c0 += 32) {
          if (WIDE) {
            uint32_t v[32], w[32];
            DV3_TMEM_LD32(v, lane_addr + (uint32_t)(buf * 2 * BN + c0));
            DV3_TMEM_LD32(w, lane_addr + (uint32_t)(buf * 2 * BN + BN + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(v[j]) + __uint_as_float(w[j]);
          } else {
            uint32_t v[32];
            DV3_TMEM_LD32(v, lane_addr + (uint32_t)(buf * BN + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      }
      if (!WIDE) {
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
      }

      if (threadIdx.x == G2_EPI_WARP0 * 32) G2_STAMP(6);
      if (SK > 1) {
        // cluster split-K: park the partial sums in shared memory (the pipeline stages are idle:
        // this CTA has a single unit and its last MMA has completed); reduced after the role code
        float* stg = reinterpret_cast<float*>(smem) + (size_t)(q * 32 + lane) * (BN + 4) + half * HW;
#pragma unroll
        for (int j = 0; j < HW; j += 4)
          *reinterpret_cast<float4*>(stg + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
      } else if (row < g.M && g.splitk > 1) {
        // split-K: partial sums meet in C through fp32 atomics (C zeroed / preloaded by the host
        // side of the launch); bias and addend ride on partition 0
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
  if (SK > 1) {
    __syncthreads();
    cluster_sync_all();                  // every partition's partial tile is parked
    if (threadIdx.x == G2_EPI_WARP0 * 32) G2_STAMP(8);
    if (warp >= G2_EPI_WARP0 && warp < G2_EPI_WARP0 + 8) {
      // this CTA's share: columns [rank CW, (rank+1) CW) of all 128 rows, one float4 per thread
      // and pass, consecutive lanes on consecutive 16 B of a row (coalesced both ways)
      constexpr int CW = BN / SK, C4 = CW / 4, PER = G2_BM * C4 / 256;
      const int rank = (int)cluster_ctarank();
      const int te = threadIdx.x - G2_EPI_WARP0 * 32;   // 0..255
      const int tile = (int)(blockIdx.x / SK);
      const int tm = g.m_fast ? tile % g.tiles_m : tile / g.tiles_n;
      const int tn = g.m_fast ? tile / g.tiles_m : tile % g.tiles_n;
      const bool vec_ok = g2_vec_ok(g);
      const uint32_t stage0 = smem_u32(smem);
      uint32_t peer[SK];
#pragma unroll
      for (int r = 0; r < SK; ++r)
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer[r]) : "r"(stage0), "r"(r));
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = i * 256 + te, rl = idx / C4, c4 = idx % C4;
        const uint32_t off = (uint32_t)(rl * (BN + 4) + rank * CW + 4 * c4) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < SK; ++r) {
          float4 v;
          asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "r"(peer[r] + off));
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        g2_emit4(g, tm * G2_BM + rl, tn * BN + rank * CW + 4 * c4, acc, true, false, vec_ok);
      }
    }
    if (threadIdx.x == G2_EPI_WARP0 * 32) G2_STAMP(9);
    // nobody's shared memory goes away under a peer's reads; nothing to publish, so the arrive
    // is relaxed (a release would first drain this CTA's result stores: ~1 us)
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    if (threadIdx.x == G2_EPI_WARP0 * 32) G2_STAMP(10);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) G2_STAMP(7);
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

// Tensor map of one operand plane.  K-major: memory is [rows, K] (row stride ld), box = 32 k x
// box_rows rows.  MN-major: memory is [K, rows] (row stride ld), box = 32 rows x 32 k.
int make_map2(CUtensorMap* m, const float* base, int rows, int K, int ld, int box_rows, bool mn) {
  EncodeTiledFn2 fn = encode_fn2();
  DV3_REQUIRE(fn, DV3_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)(mn ? rows : K), (cuuint64_t)(mn ? K : rows)};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)G2_BK, (cuuint32_t)(mn ? G2_BK : box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DV3_REQUIRE(r == CUDA_SUCCESS, DV3_ERR_CUDA,
              "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d mn=%d", (int)r, rows, K, ld,
              (int)mn);
  return 0;
}

// resident clusters of `sk` CTAs of the 128-wide kernel (0 = cluster launch unavailable)
template <int SK>
static int max_clusters_sk();

template <int BN, bool AMN, bool BMN, bool RAW, bool WIDE = false, int SK = 1>
static int launch_umma2(const Gemm2Maps& mp, Gemm2Args g, double flops, cudaStream_t st) {
  using Cfg = G2Cfg<BN>;
  auto kern = umma2_gemm_kernel<BN, AMN, BMN, RAW, WIDE, SK>;
  static DeviceOnce attr;
  if (attr.need())
    DV3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)Cfg::SMEM));
  g.tiles_n = (g.N + BN - 1) / BN;
  g.tiles_m = (g.M + G2_BM - 1) / G2_BM;
  g.tiles = g.tiles_n * g.tiles_m;
  // concurrently running CTAs should share the larger operand's tiles in L2: walk M fastest when
  // the B operand is the big one (measured on 1024 x 12288 x 5120: N-fastest re-read the 503 MB
  // weight planes 5x from DRAM)
  g.m_fast = g.N > g.M ? 1 : 0;
  if (g.splitk < 1) g.splitk = 1;
  g.nkp = (g.nk + g.splitk - 1) / g.splitk;
  if (SK == 1) g.splitk = (g.nk + g.nkp - 1) / g.nkp;   // no empty partition (a cluster tolerates them)
  g.units = g.tiles * g.splitk;
  if (SK == 1 && g.splitk > 1 && !g.accumulate)
    DV3_CHECK_CUDA(cudaMemset2DAsync(g.C, (size_t)g.ldc * 4, 0, (size_t)g.N * 4, g.M, st));
  const bool prof = prof_on();
  if (SK > 1) {
    DV3_REQUIRE(g.splitk == SK, DV3_ERR_BAD_SHAPE, "cluster split-K: %d partitions for SK=%d",
                g.splitk, SK);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.units);
    cfg.blockDim = dim3(G2_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = SK; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (prof) prof_begin(st);
    DV3_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, mp, g));
    if (prof) prof_end(st, 1, flops);
    DV3_CHECK_LAUNCH("umma2_gemm_kernel<cluster split-K>");
    return 0;
  }
  const int grid = g.units < sm_count() ? g.units : sm_count();
  if (prof) prof_begin(st);
  DV3_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(G2_THREADS), Cfg::SMEM, st, mp, g));
  if (prof) prof_end(st, 1, flops);
  DV3_CHECK_LAUNCH("umma2_gemm_kernel");
  return 0;
}

template <int BN, bool RAW, bool WIDE = false, int SK = 1>
static int dispatch_major(bool amn, bool bmn, const Gemm2Maps& mp, const Gemm2Args& g, double flops,
                          cudaStream_t st) {
  if (RAW || (!amn && !bmn)) return launch_umma2<BN, false, false, RAW, WIDE, SK>(mp, g, flops, st);
  if (!amn && bmn) return launch_umma2<BN, false, true, false, WIDE, SK>(mp, g, flops, st);
  if (amn && !bmn) return launch_umma2<BN, true, false, false, WIDE, SK>(mp, g, flops, st);
  return launch_umma2<BN, true, true, false, WIDE, SK>(mp, g, flops, st);
}

template <int SK>
static int max_clusters_sk() {
  static int per_dev[64];
  static DeviceOnce probed;
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = per_dev[dev & 63];
  if (probed.need()) {
    using Cfg = G2Cfg<128>;
    auto kern = umma2_gemm_kernel<128, false, false, false, true, SK>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sm_count() / SK * SK);
    cfg.blockDim = dim3(G2_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = SK; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int m = 0;
    if (cudaOccupancyMaxActiveClusters(&m, kern, &cfg) != cudaSuccess) m = 0;
    cudaGetLastError();
    n = m;
  }
  return n;
}

static bool tma_ok(const float* p, int ld) {
  return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0;
}

// N-tile width (and K partitions, when the caller allows them) by a small cost model:
// waves x k-blocks x (time per k-block).  The k-block time is set by shared-memory traffic (TMA
// writes + UMMA operand reads at 128 B/clk: 160 / 120 / 100 KB per k-block for BN = 128 / 64 /
// 32), measured 1250 / 940 / 780 clk.  Split-K pays a memset and an atomic epilogue.
static int max_clusters(int sk) {
  return sk == 2 ? max_clusters_sk<2>() : sk == 4 ? max_clusters_sk<4>() : max_clusters_sk<8>();
}

static void pick_shape(int M, int N, int nk, bool allow_splitk, bool allow_pair, bool allow_csk,
                       int* bn_out, int* splitk_out, int* pair_out, int* csk_out) {
  const int tm = (M + G2_BM - 1) / G2_BM, sms = sm_count();
  // single-CTA tiles 128 x {128,64,32}; pair tiles 256 x {128,64} (dv3_umma2x.cu): per CTA and
  // k-block 120 / 100 KB of shared-memory traffic -> 940 / 780 clk
  const int bns[5] = {128, 64, 32, 128, 64};
  // single-CTA tiles issue two MMAs per k-step (WIDE): 1130 (single wave) / 790 / 650 clk
  const long long cost[5] = {1250, 790, 650, 940, 780};
  long long best_t = -1;
  *bn_out = 32; *splitk_out = 1; *pair_out = 0; *csk_out = 0;
  for (int i = 0; i < 5; ++i) {
    const bool pair = i >= 3;
    if (pair && (!allow_pair || M <= G2_BM)) continue;
    if (bns[i] > 32 && N <= bns[i] / 2) continue;          // mostly padding
    const int slots = pair ? sms / 2 : sms;
    const int tiles = (pair ? (M + 2 * G2_BM - 1) / (2 * G2_BM) : tm) * ((N + bns[i] - 1) / bns[i]);
    int sk = 1;
    if (allow_splitk && tiles < slots) {
      sk = slots / tiles;
      const int cap = nk / 8;                              // >= 8 k-blocks per partition
      if (sk > cap) sk = cap;
      if (sk < 1) sk = 1;
    }
    const int nkp = (nk + sk - 1) / sk;
    const long long ck = (i == 0 && tiles * sk <= slots) ? 1130 : cost[i];
    long long t = (long long)((tiles * sk + slots - 1) / slots) * nkp * ck;
    if (sk > 1) t += 8000;
    if (pair) t += t / 50;                                 // ties go to the single-CTA kernel
    if (best_t < 0 || t < best_t) {
      best_t = t; *bn_out = bns[i]; *splitk_out = sk; *pair_out = pair ? 1 : 0;
    }
  }
  // 128 x 128 tiles with K partitioned over a cluster (one wave by construction).  Measured
  // (scratch/csk_test.py, 1024 x 512 x K): 7.9 us + 0.15 us per k-block against 3.4 + 0.32 for
  // BN = 32, i.e. the parked partial tile, two cluster barriers and the DSMEM reduction cost
  // about 8500 clk more per launch than the plain epilogue
  constexpr long long CSK_OVERHEAD = 8500;
  if (allow_csk) {
    const int tiles = tm * ((N + 127) / 128);
    for (int sk = 2; sk <= 8; sk *= 2) {
      if (nk / sk < G2_CH || tiles > max_clusters(sk)) continue;
      const long long t = (long long)((nk + sk - 1) / sk) * 1130 + CSK_OVERHEAD;
      if (t < best_t) {
        best_t = t; *bn_out = 128; *splitk_out = sk; *pair_out = 0; *csk_out = sk;
      }
    }
  }
}

// C = [A1|A2] B^T (+bias +addend) from pre-split operands.  A2 may be NULL (K2 ignored); it shares
// A1's storage order.  lo == NULL on all operands selects the raw mode (K-major only).
// accumulate: bit 0 = add into C; bit 1 = the kernel may partition K over CTAs and combine the
// partial sums with fp32 atomics (summation order then not fixed: used for parameter gradients).
int tc_gemm_ops(const TcOperand& A1, int K1, const TcOperand* A2, int K2, const TcOperand& B,
                const float* bias, const float* addend, int ldadd, float* C, int ldc, int M, int N,
                int accumulate, cudaStream_t st) {
  if (!A2) K2 = 0;
  DV3_REQUIRE(M > 0 && N > 0 && K1 > 0 && K2 >= 0, DV3_ERR_BAD_SHAPE,
              "tc_gemm: M=%d N=%d K1=%d K2=%d", M, N, K1, K2);
  const bool raw = !A1.lo;
  DV3_REQUIRE(tma_ok(A1.hi, A1.ld) && tma_ok(B.hi, B.ld) && (!A2 || tma_ok(A2->hi, A2->ld)),
              DV3_ERR_BAD_SHAPE,
              "tc_gemm: operands must be 16-byte aligned with row strides %% 4 == 0 (lda=%d ldb=%d)",
              A1.ld, B.ld);
  DV3_REQUIRE(raw ? (!B.lo && (!A2 || !A2->lo) && !A1.mn && !B.mn)
                  : (tma_ok(A1.lo, A1.ld) && tma_ok(B.lo, B.ld) && (!A2 || tma_ok(A2->lo, A2->ld))),
              DV3_ERR_BAD_SHAPE, "tc_gemm: mixed raw / pre-split operands");
  DV3_REQUIRE(!A2 || A2->mn == A1.mn, DV3_ERR_BAD_SHAPE, "tc_gemm: A segments differ in order");
  const int K = K1 + K2;
  const int nk_all = (K1 + G2_BK - 1) / G2_BK + (K2 + G2_BK - 1) / G2_BK;
  int BN = 32, splitk = 1, pair = 0, csk = 0;
  const char* e_pair = DV3_ENV("DV3_TC_PAIR");
  const char* e_csk = DV3_ENV("DV3_TC_CSK");
  const int allow_pair = (e_pair && e_pair[0] == '0') ? 0 : 1;
  const int allow_csk = (e_csk && e_csk[0] == '0') ? 0 : 1;
  pick_shape(M, N, nk_all, (accumulate & 2) != 0, allow_pair && !raw, allow_csk && !raw, &BN,
             &splitk, &pair, &csk);
  if (const char* f = DV3_ENV("DV3_TC_FORCE")) {       // experiment knob: "<BN>,<pair>[,<cluster K>]"
    int fb = 0, fp = 0, fc = 0;
    const int got = sscanf(f, "%d,%d,%d", &fb, &fp, &fc);
    if (got >= 2 && (fb == 32 || fb == 64 || fb == 128) &&
        !(fp && (fb == 32 || raw || M <= G2_BM))) {
      BN = fb; pair = fp; splitk = 1; csk = 0;
      const int tiles = ((M + G2_BM - 1) / G2_BM) * ((N + 127) / 128);
      if (got == 3 && (fc == 2 || fc == 4 || fc == 8) && fb == 128 && !fp && !raw &&
          nk_all / fc >= 1 && tiles <= max_clusters(fc)) {
        csk = fc; splitk = fc;
      }
    }
  }
  if (pair)
    return tc_gemm_pair(A1, K1, A2, K2, B, bias, addend, ldadd, C, ldc, M, N, accumulate, BN, splitk,
                        st);
  Gemm2Maps mp;
  DV3_TRY(make_map2(&mp.a1h, A1.hi, M, K1, A1.ld, G2_BM, A1.mn));
  if (!raw) DV3_TRY(make_map2(&mp.a1l, A1.lo, M, K1, A1.ld, G2_BM, A1.mn));
  else mp.a1l = mp.a1h;
  if (A2) {
    DV3_TRY(make_map2(&mp.a2h, A2->hi, M, K2, A2->ld, G2_BM, A2->mn));
    if (!raw) DV3_TRY(make_map2(&mp.a2l, A2->lo, M, K2, A2->ld, G2_BM, A2->mn));
    else mp.a2l = mp.a2h;
  } else {
    mp.a2h = mp.a1h; mp.a2l = mp.a1l;
  }
  DV3_TRY(make_map2(&mp.bh, B.hi, N, K, B.ld, BN, B.mn));
  if (!raw) DV3_TRY(make_map2(&mp.bl, B.lo, N, K, B.ld, BN, B.mn));
  else mp.bl = mp.bh;
  Gemm2Args g{};
  g.C = C; g.bias = bias; g.addend = addend; g.ldc = ldc; g.ldadd = ldadd; g.M = M; g.N = N;
  g.K1 = K1; g.nk1 = (K1 + G2_BK - 1) / G2_BK; g.nk = g.nk1 + (K2 + G2_BK - 1) / G2_BK;
  g.accumulate = accumulate & 1;
  g.splitk = splitk;
  if (const char* te = DV3_ENV("DV3_GEMM_TIMING")) g.stamps = te[0] == '1' ? po_timing_buffer() : nullptr;
  const double flops = 2.0 * M * N * K;
  if (raw) {
    if (BN == 128) return dispatch_major<128, true>(false, false, mp, g, flops, st);
    if (BN == 64) return dispatch_major<64, true>(false, false, mp, g, flops, st);
    return dispatch_major<32, true>(false, false, mp, g, flops, st);
  }
  // two-MMA issue: measured 12-20 % faster for BN = 32, 10-12 % for BN = 64, 5 % for BN = 128
  // while the launch is a single wave, 1-3 % slower for multi-wave BN = 128 (tensor-pipe bound)
  const char* e_wide = DV3_ENV("DV3_TC_WIDE");
  const int wide = (e_wide && e_wide[0] == '0') ? 0 : 1;
  if (csk == 2) return dispatch_major<128, false, true, 2>(A1.mn, B.mn, mp, g, flops, st);
  if (csk == 4) return dispatch_major<128, false, true, 4>(A1.mn, B.mn, mp, g, flops, st);
  if (csk == 8) return dispatch_major<128, false, true, 8>(A1.mn, B.mn, mp, g, flops, st);
  const int units128 = ((M + G2_BM - 1) / G2_BM) * ((N + 127) / 128) * splitk;
  if (wide) {
    if (BN == 128 && units128 <= sm_count())
      return dispatch_major<128, false, true>(A1.mn, B.mn, mp, g, flops, st);
    if (BN == 128) return dispatch_major<128, false>(A1.mn, B.mn, mp, g, flops, st);
    if (BN == 64) return dispatch_major<64, false, true>(A1.mn, B.mn, mp, g, flops, st);
    return dispatch_major<32, false, true>(A1.mn, B.mn, mp, g, flops, st);
  }
  if (BN == 128) return dispatch_major<128, false>(A1.mn, B.mn, mp, g, flops, st);
  if (BN == 64) return dispatch_major<64, false>(A1.mn, B.mn, mp, g, flops, st);
  return dispatch_major<32, false>(A1.mn, B.mn, mp, g, flops, st);
}

}  // namespace dv3

static dv3::TcOperand to_op(const dv3_tc_operand* o) {
  dv3::TcOperand r{};
  if (o) { r.hi = o->hi; r.lo = o->lo; r.ld = o->ld; r.mn = o->mn_major != 0; }
  return r;
}

// C ABI (see include/dv3_b200.h)
extern "C" int dv3_gemm_tc(const dv3_tc_operand* A1, int32_t K1, const dv3_tc_operand* A2,
                           int32_t K2, const dv3_tc_operand* B, const float* bias,
                           const float* addend, int32_t ldadd, float* C, int32_t ldc, int32_t M,
                           int32_t N, int32_t accumulate, void* stream) {
  using namespace dv3;
  DV3_REQUIRE(M >= 0 && N >= 0, DV3_ERR_BAD_SHAPE, "gemm_tc: M=%d N=%d", M, N);
  if (M == 0 || N == 0) return 0;
  DV3_REQUIRE(A1 && B && C && A1->hi && B->hi, DV3_ERR_NULL, "gemm_tc: null pointer");
  TcOperand a1 = to_op(A1), a2 = to_op(A2), b = to_op(B);
  return tc_gemm_ops(a1, K1, (A2 && A2->hi) ? &a2 : nullptr, K2, b, bias, addend, ldadd, C, ldc, M,
                     N, accumulate, static_cast<cudaStream_t>(stream));
}

extern "C" int dv3_split_tf32(const float* x, int32_t ld, int32_t rows, int32_t cols, float* hi,
                              float* lo, int32_t ld_out, void* stream) {
  using namespace dv3;
  DV3_REQUIRE(rows >= 0 && cols >= 0 && ld_out >= cols, DV3_ERR_BAD_SHAPE,
              "split_tf32: rows=%d cols=%d ld_out=%d", rows, cols, ld_out);
  if (rows == 0 || cols == 0) return 0;
  DV3_REQUIRE(x && hi && lo, DV3_ERR_NULL, "split_tf32: null pointer");
  return tc_split(x, ld, cols, nullptr, 0, 0, rows, hi, lo, static_cast<cudaStream_t>(stream),
                  ld_out);
}

// raw-operand convenience entry (no scratch, in-SM split)
extern "C" int dv3_linear_tc2_fwd(const float* A1, int32_t lda1, int32_t K1, const float* A2,
                                  int32_t lda2, int32_t K2, const float* W, int32_t ldw,
                                  const float* bias, const float* addend, int32_t ldadd, float* C,
                                  int32_t ldc, int32_t M, int32_t N, int32_t accumulate,
                                  void* stream) {
  using namespace dv3;
  DV3_REQUIRE(M >= 0 && N >= 0, DV3_ERR_BAD_SHAPE, "linear_tc2_fwd: M=%d N=%d", M, N);
  if (M == 0 || N == 0) return 0;
  DV3_REQUIRE(A1 && W && C, DV3_ERR_NULL, "linear_tc2_fwd: null pointer");
  TcOperand a1{A1, nullptr, lda1, false}, a2{A2, nullptr, lda2, false}, b{W, nullptr, ldw, false};
  return tc_gemm_ops(a1, K1, A2 ? &a2 : nullptr, K2, b, bias, addend, ldadd, C, ldc, M, N,
                     accumulate, static_cast<cudaStream_t>(stream));
}
