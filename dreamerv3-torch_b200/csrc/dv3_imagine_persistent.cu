// ImagBehavior._imagine forward as ONE persistent kernel (models.py:448-548 over networks.py:208-233,
// actor networks.py:657-700): all H steps of the actor-in-the-loop prior rollout, every GEMM on the
// tcgen05 tensor cores, LayerNorm / SiLU / GRU gates / categorical draw / actor head in the GEMM
// epilogues.  Replaces the ~12 launches per step of the stepwise path (dv3_imagine.cu).
//
// Decomposition.  Trajectories (rows) are independent, so the N rows are cut into row blocks of 128
// (one UMMA M tile) and a GROUP of 16 CTAs owns a row block for the whole horizon: no grid-wide
// barrier, groups drift freely.  Inside a group CTA j owns output-column slice j of every layer
// (32 of the 512 actor / hidden / deter columns, the matching 3 x 32 GRU gate columns, 2 of the 32
// categorical groups), i.e. each phase is a 128 x {32, 96, 64} x K product whose A operand (the
// previous layer's activations of the 128 rows, tf32 hi/lo planes in L2) is streamed by TMA and
// whose B operand is the CTA's rows of the weight planes.  The layer's epilogue runs on the
// accumulator in registers: + one-hot gather, LayerNorm statistics (partial sums exchanged inside the
// group), SiLU / gates / draw, fp32 outputs for backward and the hi/lo planes the next phase reads.
// Group synchronisation: one monotonically increasing counter per row block in global memory
// (red.release / ld.acquire at gpu scope), writer-side fence.proxy.async before TMA reads data written
// by generic stores.  (A 16-CTA thread-block cluster would give DSMEM for this exchange, but only 7
// such clusters are co-resident on a B200 with this kernel's shared memory -- measured,
// scratch/probe/cl16.cu -- and the rollout needs 8.)
//
// Bound: per-SM L2 -> SM ingest (64 B/clk): every CTA streams the whole 128 x K A operand of each
// phase (32 KB per 32-wide k-block) plus its weight slice; see DESIGN.md section 4.3.
//
// Shapes: S = C = 32, D = Hd = U = 512 (the reference's defaults for every suite), 1..3 actor layers,
// N a multiple of 128 with N / 128 * 16 <= SMs.  Anything else runs the stepwise path.
#include <cstdlib>
#include <type_traits>
#include "dv3_tc.cuh"

namespace dv3 {

int make_map2(CUtensorMap* m, const float* base, int rows, int K, int ld, int box_rows, bool mn);

constexpr int PI_BM = 128, PI_BK = 32;
constexpr int PI_GC = 16;                    // CTAs per row block
constexpr int PI_W = 512;                    // D = Hd = U
constexpr int PI_SL = PI_W / PI_GC;          // 32 columns per CTA
constexpr int PI_S = 32, PI_C = 32, PI_SC = PI_S * PI_C;
constexpr int PI_THREADS = 384;              // warp 0 TMA, warp 1 MMA, warps 4-11 epilogue
constexpr int PI_EPI_WARP0 = 4, PI_EPI_THREADS = 256;
constexpr int PI_CH = 4;                     // k-blocks per chunk accumulator (as dv3_umma2.cu)
constexpr uint32_t PI_A_BYTES = PI_BM * PI_BK * 4;                 // 16 KB per plane
constexpr uint32_t PI_PIPE_BYTES = 3 * (2 * PI_A_BYTES + 2 * 96 * 128);   // 168 KB
constexpr int PI_MAX_ST = 4;
constexpr int PI_MAX_L = 3;
constexpr int PI_NBAR = 2 * PI_MAX_ST + 4;
constexpr int PI_IDX_LD = 48;                // bytes per row of the staged class indices
constexpr int PI_T_LD = 36, PI_T2_LD = 68;   // row pitch (floats) of the 32- / 64-column staging tiles
// per-CTA constants in shared memory (float offsets): LayerNorm gamma | beta of the CTA's 32 columns
// per layer, b_ims of its 64 logits, head biases, head weight columns, action rows of W_in^T
constexpr int PI_C_LNA = 0, PI_C_LNIN = 64 * PI_MAX_L, PI_C_LNOUT = PI_C_LNIN + 64,
              PI_C_LNGRU = PI_C_LNOUT + 64, PI_C_BIMS = PI_C_LNGRU + 192, PI_C_BHEAD = PI_C_BIMS + 64,
              PI_C_WHEAD = PI_C_BHEAD + 32, PI_C_WACT = PI_C_WHEAD + 32 * 32, PI_CST = PI_C_WACT + 32 * 32;
constexpr size_t PI_SMEM = PI_PIPE_BYTES + 1024 + 256 + PI_BM * PI_IDX_LD + 2 * PI_BM * PI_T_LD * 4 +
                           PI_CST * 4;
static_assert(PI_SMEM <= 227 * 1024, "shared memory budget");
static_assert(PI_NBAR * 8 + 8 <= 256, "barrier block");
static_assert(PI_BM * PI_T2_LD <= 2 * PI_BM * PI_T_LD, "64-column tile spans both 32-column tiles");
constexpr int PI_STAMPS = 16;

template <int BN> struct PiGeo {
  static constexpr uint32_t STAGE = 2 * PI_A_BYTES + 2 * BN * 128;
  static constexpr int ST = (int)(PI_PIPE_BYTES / STAGE) > PI_MAX_ST ? PI_MAX_ST : (int)(PI_PIPE_BYTES / STAGE);
};

struct PiMaps {                              // [.][0] = hi plane, [.][1] = lo plane
  CUtensorMap dsp[2][2], asp[2][2], xsp[2], ysp[2];        // A operands, box 32 k x 128 rows
  CUtensorMap wa[PI_MAX_L][2], gru[2], out[2], ims[2];     // B operands, box 32 k x 32 rows
};

struct PiArgs {
  int N, H, L, A, dist;
  float eps, unimix, a_unimix, min_std, max_std;
  const float* Wa0T; const float* WinT;
  const float* a_ln_g[PI_MAX_L]; const float* a_ln_b[PI_MAX_L];
  const float *w_mean, *b_mean, *w_std, *b_std;
  const float *ln_in_g, *ln_in_b, *ln_gru_g, *ln_gru_b, *ln_out_g, *ln_out_b, *b_ims;
  const float* start_deter; const float* act_noise; const float* u_state;
  float *feat, *logit, *action; int32_t* idx;
  float *x_pre, *x, *g_pre, *y_pre, *y, *a_pre, *a_act, *a_mean_raw, *a_std_raw;
  float *dsp_hi[2], *dsp_lo[2], *asp_hi[2], *asp_lo[2], *xsp_hi, *xsp_lo, *ysp_hi, *ysp_lo;
  unsigned* cnt;                 // [row blocks][32] arrival counters (one 128-byte line each)
  float2* stats;                 // [row blocks][2][16][2][128] partial LayerNorm statistics
  float* hpart;                  // [row blocks][16][32][128] K-split partial sums of the actor heads
  unsigned* abort_flag;
  unsigned long long* stamps;    // debug: [H][PI_STAMPS] globaltimer stamps of CTA 0
};

#define DV3_TMEM_LD16(v, taddr)                                                                   \
  asm volatile(                                                                                   \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                   \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"           \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),      \
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),  \
        "=r"(v[14]), "=r"(v[15])                                                                  \
      : "r"(taddr))

// ---- waits with a watchdog: a protocol bug must end as wrong numbers + a flag, not as a hung GPU
__device__ __forceinline__ unsigned long long pi_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long PI_TIMEOUT_NS = 2000000000ull;

__device__ __noinline__ void pi_mbar_wait_slow(uint32_t bar, uint32_t parity, unsigned* abort_flag,
                                               int code) {
  const unsigned long long t0 = pi_now();
  for (;;) {
    for (int i = 0; i < 2048; ++i)
      if (mbar_try_wait(bar, parity)) return;
    if (*reinterpret_cast<volatile unsigned*>(abort_flag) != 0) return;
    if (pi_now() - t0 > PI_TIMEOUT_NS) {
      atomicCAS(abort_flag, 0u, (unsigned)(0x100 | code));
      return;
    }
  }
}
__device__ __forceinline__ void pi_mbar_wait(uint32_t bar, uint32_t parity, unsigned* abort_flag,
                                             int code) {
  for (int i = 0; i < 64; ++i)
    if (mbar_try_wait(bar, parity)) return;
  pi_mbar_wait_slow(bar, parity, abort_flag, code);
}

__device__ __forceinline__ unsigned pi_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __noinline__ void pi_event_wait(const unsigned* cnt, unsigned target, unsigned* abort_flag) {
  if (pi_ld_acquire(cnt) >= target) return;
  const unsigned long long t0 = pi_now();
  for (;;) {
    for (int i = 0; i < 1024; ++i)
      if (pi_ld_acquire(cnt) >= target) return;
    if (*reinterpret_cast<volatile unsigned*>(abort_flag) != 0) return;
    if (pi_now() - t0 > PI_TIMEOUT_NS) {
      atomicCAS(abort_flag, 0u, 0x200u);
      return;
    }
  }
}
__device__ __forceinline__ void pi_event_arrive(unsigned* cnt) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
}
__device__ __forceinline__ void pi_fence_proxy_async() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void pi_epi_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(PI_EPI_THREADS) : "memory");
}

// butterfly-ordered sum of 32 values held by one thread: the association order of warp_sum over
// lanes, so the categorical chain is bit-identical to the lane-per-class kernels (dv3_rowwise.cu)
__device__ __forceinline__ float pi_bfly_sum32(float (&a)[32]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < o; ++i) a[i] += a[i + o];
  return a[0];
}

__device__ __forceinline__ void pi_store16(float* p, const float* v) {
#pragma unroll
  for (int i = 0; i < 16; i += 4)
    *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

// The chain of unimix_probs + the supplied-uniform argmax (dv3_common.cuh, dv3_rowwise.cu) for one
// categorical held by ONE thread: classes [0, Cv) of l / u are valid.  Sums run in the butterfly's
// association order over 32 slots (invalid ones hold 0), so every value -- and the drawn index -- is
// bit-identical to the lane-per-class kernels.  unimix > 0.
__device__ __forceinline__ int pi_unimix_draw(float (&l)[32], const float (&u)[32], int Cv, float unimix) {
  float e[32];
  float m = -INFINITY;
#pragma unroll
  for (int c = 0; c < 32; ++c) if (c < Cv) m = fmaxf(m, l[c]);
#pragma unroll
  for (int c = 0; c < 32; ++c) { l[c] = (c < Cv) ? expf(l[c] - m) : 0.f; e[c] = l[c]; }
  const float s1 = pi_bfly_sum32(e);
#pragma unroll
  for (int c = 0; c < 32; ++c) l[c] = logf((l[c] / s1) * (1.f - unimix) + unimix / (float)Cv);
  float m2 = -INFINITY;
#pragma unroll
  for (int c = 0; c < 32; ++c) if (c < Cv) m2 = fmaxf(m2, l[c]);
#pragma unroll
  for (int c = 0; c < 32; ++c) e[c] = (c < Cv) ? expf(l[c] - m2) : 0.f;
  const float lse = m2 + logf(pi_bfly_sum32(e));
#pragma unroll
  for (int c = 0; c < 32; ++c) l[c] = l[c] - lse;
  float m3 = -INFINITY;
#pragma unroll
  for (int c = 0; c < 32; ++c) if (c < Cv) m3 = fmaxf(m3, l[c]);
#pragma unroll
  for (int c = 0; c < 32; ++c) { l[c] = (c < Cv) ? expf(l[c] - m3) : 0.f; e[c] = l[c]; }
  const float s3 = pi_bfly_sum32(e);
  float bv = -INFINITY;
  int kk = 0;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    if (c < Cv) {
      float sc = (l[c] / s3) / (-logf(u[c]));
      if (sc != sc) sc = -INFINITY;                       // NaN never wins
      if (c == 0 || sc > bv) { bv = sc; kk = c; }         // first index wins ties
    }
  }
  return kk;
}

__global__ void __launch_bounds__(PI_THREADS, 1)
imagine_persistent_fwd_kernel(const __grid_constant__ PiMaps mp, const PiArgs g) {
  extern __shared__ __align__(1024) uint8_t pi_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pi_smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PI_PIPE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + PI_NBAR);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * PI_MAX_ST,
                 tfull0 = empty0 + 8 * PI_MAX_ST, tempty0 = tfull0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = blockIdx.x / PI_GC, j = blockIdx.x % PI_GC;
  const int m0 = rb * PI_BM;
  unsigned* const cnt = g.cnt + rb * 32;
  unsigned* const abort_flag = g.abort_flag;
  const int L = g.L, H = g.H, N = g.N, A = g.A;
  const unsigned EPS = 2u * L + 7u;          // events per step

  if (threadIdx.x == 0) {
    for (int s = 0; s < PI_MAX_ST; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 8);         // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t pf = 0;                       // fill parity per stage
      int prev_st = 0, prev_bn = 0;
      auto load = [&](auto bn_tag, const CUtensorMap* a1, int nk1, const CUtensorMap* a2, int nk,
                      const CUtensorMap* b, int nbox, int brow0, int bstride) {
        constexpr int BN = decltype(bn_tag)::value;
        constexpr int ST = PiGeo<BN>::ST;
        constexpr uint32_t STAGE = PiGeo<BN>::STAGE;
        if (prev_bn != BN) {
          // the stage geometry changes: every stage of the old one must have been consumed
          for (int s = 0; s < prev_st; ++s)
            pi_mbar_wait(empty0 + 8 * s, ((pf >> s) & 1) ^ 1, abort_flag, 5);
          prev_bn = BN; prev_st = ST;
        }
        for (int kb = 0; kb < nk; ++kb) {
          const int s = kb % ST;
          pi_mbar_wait(empty0 + 8 * s, ((pf >> s) & 1) ^ 1, abort_flag, 1);
          pf ^= 1u << s;
          const uint32_t base = smem_u32(smem + s * STAGE);
          const uint32_t bar = full0 + 8 * s;
          mbar_expect_tx(bar, STAGE);
          const bool seg1 = kb < nk1;
          const CUtensorMap* am = seg1 ? a1 : a2;
          const int ak = (seg1 ? kb : kb - nk1) * PI_BK, wk = kb * PI_BK;
          tma_load_2d(base, am, bar, ak, m0);
          tma_load_2d(base + PI_A_BYTES, am + 1, bar, ak, m0);
          for (int bx = 0; bx < nbox; ++bx) {
            tma_load_2d(base + 2 * PI_A_BYTES + bx * 4096, b, bar, wk, brow0 + bx * bstride);
            tma_load_2d(base + 2 * PI_A_BYTES + BN * 128 + bx * 4096, b + 1, bar, wk,
                        brow0 + bx * bstride);
          }
        }
      };
      auto wait_ev = [&](unsigned e) {
        pi_event_wait(cnt, (e + 1) * PI_GC, abort_flag);
        pi_fence_proxy_async();
      };
      constexpr int NKW = PI_W / PI_BK;      // 16 k-blocks per 512-wide operand
      for (int k = 0; k < H; ++k) {
        const unsigned be = (unsigned)k * EPS;
        load(std::integral_constant<int, 32>{}, mp.dsp[k & 1], NKW, nullptr, NKW, mp.wa[0], 1,
             PI_SL * j, 0);
        for (int i = 1; i < L; ++i) {
          wait_ev(be + 2 * i - 1);
          load(std::integral_constant<int, 32>{}, mp.asp[(i - 1) & 1], NKW, nullptr, NKW, mp.wa[i], 1,
               PI_SL * j, 0);
        }
        if (k == H - 1) break;
        wait_ev(be + 2 * L + 1);
        load(std::integral_constant<int, 96>{}, mp.xsp, NKW, mp.dsp[k & 1], 2 * NKW, mp.gru, 3,
             PI_SL * j, PI_W);
        wait_ev(be + 2 * L + 3);
        load(std::integral_constant<int, 32>{}, mp.dsp[(k + 1) & 1], NKW, nullptr, NKW, mp.out, 1,
             PI_SL * j, 0);
        wait_ev(be + 2 * L + 5);
        load(std::integral_constant<int, 64>{}, mp.ysp, NKW, nullptr, NKW, mp.ims, 2, 2 * PI_SL * j,
             PI_SL);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    const bool issuer = elect_one();
    const uint32_t unit0 = (smem_u32(smem) >> 4) & 0x3FFF;
    uint32_t pc = 0;                         // consume parity per stage
    int cc = 0;
    auto mma = [&](auto bn_tag, int nk) {
      constexpr int BN = decltype(bn_tag)::value;
      constexpr int ST = PiGeo<BN>::ST;
      constexpr uint32_t ST_U = PiGeo<BN>::STAGE >> 4, A_PU = PI_A_BYTES >> 4;
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(PI_BM >> 4) << 24);
      constexpr uint32_t idesc_w = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(2 * BN >> 3) << 17);
      int buf = 0;
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % ST;
        const int kin = kb % PI_CH;
        if (kin == 0) {
          buf = cc & 1;
          pi_mbar_wait(tempty0 + 8 * buf, ((cc >> 1) & 1) ^ 1, abort_flag, 2);
        }
        pi_mbar_wait(full0 + 8 * s, (pc >> s) & 1, abort_flag, 3);
        pc ^= 1u << s;
        tc_fence_after();
        const uint32_t au = unit0 + s * ST_U, bu = au + 2 * A_PU;
        const uint32_t acc = tmem_base + buf * (2 * BN);
        const bool last = kin == PI_CH - 1 || kb == nk - 1;
        if (issuer) {
#pragma unroll
          for (int k8 = 0; k8 < PI_BK / 8; ++k8) {
            const uint64_t ah = umma_desc_units<false>(au + k8 * 2);
            const uint64_t al = umma_desc_units<false>(au + A_PU + k8 * 2);
            const uint64_t bh = umma_desc_units<false>(bu + k8 * 2);    // [B_hi; B_lo], 2 BN rows
            umma_tf32(acc, ah, bh, idesc_w, (kin | k8) != 0);
            umma_tf32(acc + BN, al, bh, idesc, 1);
          }
          umma_commit(empty0 + 8 * s);
          if (last) umma_commit(tfull0 + 8 * buf);
        }
        if (last) ++cc;
        __syncwarp();
      }
    };
    constexpr int NKW = PI_W / PI_BK;
    for (int k = 0; k < H; ++k) {
      for (int i = 0; i < L; ++i) mma(std::integral_constant<int, 32>{}, NKW);
      if (k == H - 1) break;
      mma(std::integral_constant<int, 96>{}, 2 * NKW);
      mma(std::integral_constant<int, 32>{}, NKW);
      mma(std::integral_constant<int, 64>{}, NKW);
    }
  } else if (warp >= PI_EPI_WARP0) {
    // ------------------------------ epilogue ----------------------------------
    // Thread = (row of the block, column half): the natural TMEM mapping (lane = row).  Rules that
    // came out of the phase stamps (DV3_IMAGINE_TIMING=1):
    //  * anything a peer waits for moves through shared-memory tiles so that warps write whole
    //    128-byte lines (thread = row accesses cost one L1 wavefront per lane: 17 us per gather);
    //  * outputs only the backward pass reads (pre-activations, fp32 activations, logits, one-hot
    //    rows) are stored AFTER the event arrive, straight from registers -- they drain under the
    //    next main loop instead of in front of a release;
    //  * per-layer constants live in shared memory: every acquire invalidates L1, so __ldg'd
    //    LayerNorm parameters cost an L2 round trip per phase;
    //  * the one-hot gathers run under a main loop (they only need the class indices).
    const int q = warp & 3, half = (warp - PI_EPI_WARP0) >> 2;
    const int et = threadIdx.x - PI_EPI_WARP0 * 32;                  // 0..255
    const int ew = warp - PI_EPI_WARP0;                              // 0..7
    const int rl = q * 32 + lane;                                    // row inside the block
    const size_t row = (size_t)m0 + rl;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int h16 = 16 * half;
    const int c16 = PI_SL * j + h16;                                 // first of this thread's 16 columns
    int cc = 0;
    unsigned ev = 0;                                                 // events passed so far
    unsigned spar = 0;                                               // stats buffer parity
    const int F = PI_SC + PI_W;
    uint8_t* const sidx8 = smem + PI_PIPE_BYTES + 256;               // [128][PI_IDX_LD] class indices
    float* const T0 = reinterpret_cast<float*>(sidx8 + PI_BM * PI_IDX_LD);   // [128][36] tile
    float* const T1 = T0 + PI_BM * PI_T_LD;                          // [128][36] tile (T0|T1 = [128][68])
    float* const cst = T1 + PI_BM * PI_T_LD;                         // per-CTA constants (PI_CST floats)
    const int nhA = (g.dist == 0 ? 2 : 1) * A;
    const bool writer = half == 0 && (rl & (PI_GC - 1)) == j;        // this thread publishes its row's action

    // ---- constants of this CTA's column slice -> shared memory
    {
      auto cp32 = [&](float* dst, const float* src) {                // 32 consecutive floats
        if (et < 32) dst[et] = __ldg(src + et);
      };
      for (int i = 0; i < L; ++i) {
        cp32(cst + PI_C_LNA + i * 64, g.a_ln_g[i] + PI_SL * j);
        cp32(cst + PI_C_LNA + i * 64 + 32, g.a_ln_b[i] + PI_SL * j);
      }
      cp32(cst + PI_C_LNIN, g.ln_in_g + PI_SL * j);
      cp32(cst + PI_C_LNIN + 32, g.ln_in_b + PI_SL * j);
      cp32(cst + PI_C_LNOUT, g.ln_out_g + PI_SL * j);
      cp32(cst + PI_C_LNOUT + 32, g.ln_out_b + PI_SL * j);
      for (int s3 = 0; s3 < 3; ++s3) {
        cp32(cst + PI_C_LNGRU + s3 * 32, g.ln_gru_g + s3 * PI_W + PI_SL * j);
        cp32(cst + PI_C_LNGRU + 96 + s3 * 32, g.ln_gru_b + s3 * PI_W + PI_SL * j);
      }
      cp32(cst + PI_C_BIMS, g.b_ims + 2 * PI_C * j);
      cp32(cst + PI_C_BIMS + 32, g.b_ims + 2 * PI_C * j + 32);
      if (et < 32) cst[PI_C_BHEAD + et] = et < A ? __ldg(g.b_mean + et)
                                                  : (et < nhA ? __ldg(g.b_std + et - A) : 0.f);
      for (int o = 0; o < nhA; ++o)
        cp32(cst + PI_C_WHEAD + o * 32,
             (o < A ? g.w_mean + (size_t)o * PI_W : g.w_std + (size_t)(o - A) * PI_W) + PI_SL * j);
      for (int a = 0; a < A; ++a) cp32(cst + PI_C_WACT + a * 32, g.WinT + (size_t)(PI_SC + a) * PI_W + PI_SL * j);
    }
    pi_epi_sync();

    auto stamp = [&](int k, int slot) {
      if (g.stamps && blockIdx.x == 0 && et == 0) g.stamps[(size_t)k * PI_STAMPS + slot] = pi_now();
    };
    // accumulator chunks -> registers; `between(c)` runs before chunk c is awaited
    auto drain = [&](auto nseg_tag, int BN, int nk, const int* segcol, float* sum, auto&& between) {
      constexpr int NSEG = decltype(nseg_tag)::value;
      const int nchunks = (nk + PI_CH - 1) / PI_CH;
#pragma unroll
      for (int i = 0; i < NSEG * 16; ++i) sum[i] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++cc) {
        between(c);
        const int buf = cc & 1;
        pi_mbar_wait(tfull0 + 8 * buf, (cc >> 1) & 1, abort_flag, 4);
        tc_fence_after();
#pragma unroll
        for (int sg = 0; sg < NSEG; ++sg) {
          uint32_t v[16], w[16];
          DV3_TMEM_LD16(v, tmem_row + (uint32_t)(buf * 2 * BN + segcol[sg]));
          DV3_TMEM_LD16(w, tmem_row + (uint32_t)(buf * 2 * BN + BN + segcol[sg]));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) sum[sg * 16 + i] += __uint_as_float(v[i]) + __uint_as_float(w[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      }
    };
    const auto no_between = [](int) {};
    auto arrive = [&]() {
      pi_epi_sync();
      if (et == 0) pi_event_arrive(cnt);
    };
    auto wait = [&](unsigned e) {
      if (et == 0) pi_event_wait(cnt, (e + 1) * PI_GC, abort_flag);
      pi_epi_sync();
    };
    // LayerNorm statistics of the row over all 32 partial vectors (16 CTAs x 2 halves) of n values
    // each: every thread publishes (sum, M2 about its own mean); after the group event every thread
    // combines the 32 partials in the same fixed order (Chan's formula) -> identical mean / rstd in
    // every CTA.  One event per call.
    auto ln_stats = [&](auto n_tag, const float* v, float& mean, float& rstd) {
      constexpr int n = decltype(n_tag)::value;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < n; ++i) s += v[i];
      const float mloc = s / (float)n;
      float m2 = 0.f;
#pragma unroll
      for (int i = 0; i < n; ++i) { const float dlt = v[i] - mloc; m2 = fmaf(dlt, dlt, m2); }
      float2* sb = g.stats + ((size_t)(rb * 2 + spar) * PI_GC) * 256;
      __stcg(sb + (size_t)j * 256 + half * 128 + rl, make_float2(s, m2));
      arrive();
      wait(ev);
      ++ev;
      float2 p[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) p[t] = __ldcg(sb + (size_t)(t >> 1) * 256 + (t & 1) * 128 + rl);
      float tot = 0.f;
#pragma unroll
      for (int t = 0; t < 32; ++t) tot += p[t].x;
      const float ntot = (float)(32 * n);
      mean = tot / ntot;
      float M2 = 0.f;
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const float dm = p[t].x / (float)n - mean;
        M2 += p[t].y + (float)n * dm * dm;
      }
      rstd = 1.f / sqrtf(M2 / ntot + g.eps);
      spar ^= 1;
    };
    // v <- SiLU(LN(v)) with this thread's 16 gamma / beta from the constants block
    auto ln_silu16 = [&](float* v, float mean, float rstd, const float* gb) {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = siluf_(fmaf((v[c] - mean) * rstd, gb[h16 + c], gb[32 + h16 + c]));
    };
    auto put16 = [&](float* T, int ldt, int col, const float* v) {
      float4* p = reinterpret_cast<float4*>(T + rl * ldt + col);
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    };
    auto get16 = [&](const float* T, int ldt, int col, float* v) {
      const float4* p = reinterpret_cast<const float4*>(T + rl * ldt + col);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = p[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    };
    // coalesced copy of a [128][NC] tile to global rows m0.. as fp32 (dst, row stride ld; may be
    // NULL) and / or as tf32 hi / lo planes (row stride PI_W)
    auto flush = [&](auto nc_tag, const float* T, int ldt, float* dst, size_t ld, float* hi, float* lo) {
      constexpr int NC4 = decltype(nc_tag)::value / 4;
#pragma unroll
      for (int i = 0; i < PI_BM * NC4 / PI_EPI_THREADS; ++i) {
        const int e = et + PI_EPI_THREADS * i, r = e / NC4, c4 = e % NC4;
        const float4 v = *reinterpret_cast<const float4*>(T + r * ldt + 4 * c4);
        if (dst) *reinterpret_cast<float4*>(dst + (size_t)r * ld + 4 * c4) = v;
        if (hi) {
          float4 hh;
          hh.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          hh.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          hh.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          hh.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          *reinterpret_cast<float4*>(hi + (size_t)r * PI_W + 4 * c4) = hh;
          *reinterpret_cast<float4*>(lo + (size_t)r * PI_W + 4 * c4) =
              make_float4(v.x - hh.x, v.y - hh.y, v.z - hh.z, v.w - hh.w);
        }
      }
    };
    // this thread's activations -> the tile, the tile -> hi / lo planes (coalesced), event "data"
    auto publish_planes = [&](const float* v, float* hi, float* lo) {
      put16(T1, PI_T_LD, h16, v);
      pi_epi_sync();
      flush(std::integral_constant<int, 32>{}, T1, PI_T_LD, nullptr, 0, hi + (size_t)m0 * PI_W + PI_SL * j,
            lo + (size_t)m0 * PI_W + PI_SL * j);
      pi_fence_proxy_async();
      arrive();
      ++ev;
    };
    // class indices of state k -> bytes in shared memory
    auto stage_idx = [&](size_t kb0) {
      const int r = et >> 1, s0 = (et & 1) * 16;
      const int4* ip = reinterpret_cast<const int4*>(g.idx + (kb0 + r) * PI_S + s0);
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int4 t = __ldcg(ip + i);
        pk[i] = (uint32_t)t.x | ((uint32_t)t.y << 8) | ((uint32_t)t.z << 16) | ((uint32_t)t.w << 24);
      }
      *reinterpret_cast<uint4*>(sidx8 + r * PI_IDX_LD + s0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    };
    // T[r][0..32) = sum over the S one-hot rows of WT (columns 32 j .. 32 j + 31), in group order:
    // warp ew takes rows 16 ew .. 16 ew + 15, a lane one float4 of four of them -- every load
    // instruction reads four whole 128-byte lines
    auto gather_tile = [&](const float* WT, float* T) {
      const int rs = lane >> 3, c4 = lane & 7;
      float4 acc[4];
#pragma unroll
      for (int rq = 0; rq < 4; ++rq) acc[rq] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* wbase = WT + PI_SL * j + 4 * c4;
#pragma unroll 2
      for (int s4 = 0; s4 < PI_S / 4; ++s4) {
        uint32_t pk[4];
#pragma unroll
        for (int rq = 0; rq < 4; ++rq)
          pk[rq] = *reinterpret_cast<const uint32_t*>(sidx8 + (ew * 16 + rq * 4 + rs) * PI_IDX_LD + 4 * s4);
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) {
#pragma unroll
          for (int rq = 0; rq < 4; ++rq) {
            const int id = (pk[rq] >> (8 * ss)) & 0xff;
            const float4 w = __ldg(reinterpret_cast<const float4*>(
                wbase + (size_t)((4 * s4 + ss) * PI_C + id) * PI_W));
            acc[rq].x += w.x; acc[rq].y += w.y; acc[rq].z += w.z; acc[rq].w += w.w;
          }
        }
      }
#pragma unroll
      for (int rq = 0; rq < 4; ++rq)
        *reinterpret_cast<float4*>(T + (ew * 16 + rq * 4 + rs) * PI_T_LD + 4 * c4) = acc[rq];
    };

    float h[16];                                                     // deter_k of this thread's columns
    {
      const float4* hp = reinterpret_cast<const float4*>(g.start_deter + row * PI_W + c16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = hp[i];
        h[4 * i] = t.x; h[4 * i + 1] = t.y; h[4 * i + 2] = t.z; h[4 * i + 3] = t.w;
      }
    }
    constexpr int NKW = PI_W / PI_BK;
    const size_t HN = (size_t)H * N;
    // head partials of the group: [16 CTAs][32 outputs][128 rows]
    float* const hpart = g.hpart + (size_t)rb * PI_GC * 32 * PI_BM;
    bool sampled_pending = false;                                    // event "sampled" arrived at, not yet awaited

    for (int k = 0; k < H; ++k) {
      const size_t kr = (size_t)k * N + row;                         // row of step k in [H,N,..] tensors
      const size_t kb0 = (size_t)k * N + m0;                         // first row of the block
      const bool more = k < H - 1;
      stamp(k, 0);
      // ---------------- actor trunk ----------------
      for (int i = 0; i < L; ++i) {
        float v[16], pre[16];
        const int seg[1] = {h16};
        if (i == 0) {
          // the first two chunks are drained at once (the MMA warp then runs on into chunks 2, 3);
          // the one-hot part of layer 0 is gathered meanwhile -- it needs the indices of state k,
          // i.e. the previous step's event "sampled"
          drain(std::integral_constant<int, 1>{}, 32, NKW, seg, v, [&](int c) {
            if (c != 2) return;
            if (sampled_pending) { wait(ev); ++ev; sampled_pending = false; }
            stage_idx(kb0);
            pi_epi_sync();
            gather_tile(g.Wa0T, T0);
            pi_epi_sync();
          });
          float gsum[16];
          get16(T0, PI_T_LD, h16, gsum);
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] += gsum[c];
        } else {
          drain(std::integral_constant<int, 1>{}, 32, NKW, seg, v, no_between);
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) pre[c] = v[c];
        float mean, rstd;
        ln_stats(std::integral_constant<int, 16>{}, v, mean, rstd);
        ln_silu16(v, mean, rstd, cst + PI_C_LNA + i * 64);
        if (i == L - 1) {
          // actor heads, K-split: this thread's 16 columns of every head output; the two halves of a
          // row meet in shared memory, the 16 CTAs' sums in global memory behind the data event
          auto head_part = [&](int o) {
            const float4* w4 = reinterpret_cast<const float4*>(cst + PI_C_WHEAD + o * 32 + h16);
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 w = w4[c];
              acc = fmaf(v[4 * c], w.x, acc); acc = fmaf(v[4 * c + 1], w.y, acc);
              acc = fmaf(v[4 * c + 2], w.z, acc); acc = fmaf(v[4 * c + 3], w.w, acc);
            }
            return acc;
          };
          if (half == 1)
            for (int o = 0; o < nhA; ++o) T1[o * PI_BM + rl] = head_part(o);
          pi_epi_sync();
          if (half == 0)
            for (int o = 0; o < nhA; ++o)
              __stcg(hpart + ((size_t)j * 32 + o) * PI_BM + rl, head_part(o) + T1[o * PI_BM + rl]);
          arrive();                                                  // head partials (nothing reads the
          ++ev;                                                      // top activations as an operand)
        } else {
          publish_planes(v, g.asp_hi[i & 1], g.asp_lo[i & 1]);
        }
        // backward-only outputs: after the arrive
        pi_store16(g.a_pre + (size_t)i * HN * PI_W + kr * PI_W + c16, pre);
        pi_store16(g.a_act + (size_t)i * HN * PI_W + kr * PI_W + c16, v);
        if (i == 0 && more) {
          // the one-hot part of img_in, under the next main loop (T0 is free: read before the
          // statistics event above)
          gather_tile(g.WinT, T0);
        }
      }
      stamp(k, 1);
      wait(ev - 1);
      stamp(k, 2);
      // ---------------- actor head: every thread finishes its own row's outputs ----------------
      // (the same fixed order in every CTA); thread `writer` publishes them.  img_in's action part
      // is accumulated on the fly: x_pre = one-hot rows (in T0) + sum_a action_a W_in^T[SC + a].
      float v[16];
      if (more) get16(T0, PI_T_LD, h16, v);
      {
        const size_t o0 = kr * A;
        float ho[32];                                                // head outputs (before the bias)
#pragma unroll 4
        for (int o = 0; o < nhA; ++o) {
          float acc = 0.f;
#pragma unroll
          for (int t = 0; t < PI_GC; ++t) acc += __ldcg(hpart + ((size_t)t * 32 + o) * PI_BM + rl);
          ho[o] = acc + cst[PI_C_BHEAD + o];
        }
        auto add_action = [&](int a, float sa) {
          const float4* wr = reinterpret_cast<const float4*>(cst + PI_C_WACT + a * 32 + h16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 w = wr[i];
            v[4 * i] = fmaf(sa, w.x, v[4 * i]); v[4 * i + 1] = fmaf(sa, w.y, v[4 * i + 1]);
            v[4 * i + 2] = fmaf(sa, w.z, v[4 * i + 2]); v[4 * i + 3] = fmaf(sa, w.w, v[4 * i + 3]);
          }
        };
        if (g.dist == 0) {
          for (int a = 0; a < A; ++a) {
            const float mraw = ho[a], sraw = ho[A + a];
            const float mean = tanhf(mraw);
            const float std = (g.max_std - g.min_std) * sigmoidf_(sraw + 2.f) + g.min_std;
            const float out = mean + std * g.act_noise[o0 + a];
            const float act = out * (1.f / fmaxf(fabsf(out), 1.f));
            if (writer) {
              g.a_mean_raw[o0 + a] = mraw;
              g.a_std_raw[o0 + a] = sraw;
              g.action[o0 + a] = act;
            }
            if (more) add_action(a, act);
          }
        } else {
          float l[32], u[32];
#pragma unroll
          for (int a = 0; a < 32; ++a) {
            l[a] = -INFINITY; u[a] = 1.f;
            if (a < A) {
              l[a] = ho[a];
              u[a] = g.act_noise[o0 + a];
              if (writer) g.a_mean_raw[o0 + a] = l[a];
            }
          }
          const int kk = pi_unimix_draw(l, u, A, g.a_unimix);
          if (writer) {
            for (int a = 0; a < A; ++a) g.action[o0 + a] = (a == kk) ? 1.f : 0.f;
          }
          if (more) add_action(kk, 1.f);
        }
      }
      if (!more) break;
      stamp(k, 3);

      // ---------------- img_in: x = SiLU(LN(W_in [onehot | action])) ----------------
      {
        float pre[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) pre[c] = v[c];
        float mean, rstd;
        ln_stats(std::integral_constant<int, 16>{}, v, mean, rstd);
        ln_silu16(v, mean, rstd, cst + PI_C_LNIN);
        publish_planes(v, g.xsp_hi, g.xsp_lo);
        pi_store16(g.x_pre + kr * PI_W + c16, pre);
        pi_store16(g.x + kr * PI_W + c16, v);
      }
      stamp(k, 4);
      // ---------------- GRU: gates on LN_{3D}([x | h] W_gru^T) ----------------
      {
        float gv[48];
        const int seg[3] = {h16, 32 + h16, 64 + h16};
        drain(std::integral_constant<int, 3>{}, 96, 2 * NKW, seg, gv, no_between);
        float mean, rstd;
        ln_stats(std::integral_constant<int, 48>{}, gv, mean, rstd);
        const float* gg = cst + PI_C_LNGRU + h16;
        const float* gb = cst + PI_C_LNGRU + 96 + h16;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float pr = fmaf((gv[c] - mean) * rstd, gg[c], gb[c]);
          const float pcn = fmaf((gv[16 + c] - mean) * rstd, gg[32 + c], gb[32 + c]);
          const float pu = fmaf((gv[32 + c] - mean) * rstd, gg[64 + c], gb[64 + c]);
          const float rg = sigmoidf_(pr);
          const float cn = tanhf(rg * pcn);
          const float u = sigmoidf_(pu - 1.f);
          h[c] = u * cn + (1.f - u) * h[c];
        }
        publish_planes(h, g.dsp_hi[(k + 1) & 1], g.dsp_lo[(k + 1) & 1]);
        float* gp = g.g_pre + kr * (3 * PI_W) + c16;
        pi_store16(gp, gv);
        pi_store16(gp + PI_W, gv + 16);
        pi_store16(gp + 2 * PI_W, gv + 32);
        pi_store16(g.feat + (kr + N) * F + PI_SC + c16, h);
      }
      stamp(k, 5);
      // ---------------- img_out: y = SiLU(LN(deter W_out^T)) ----------------
      {
        float pre[16];
        const int seg[1] = {h16};
        drain(std::integral_constant<int, 1>{}, 32, NKW, seg, v, no_between);
#pragma unroll
        for (int c = 0; c < 16; ++c) pre[c] = v[c];
        float mean, rstd;
        ln_stats(std::integral_constant<int, 16>{}, v, mean, rstd);
        ln_silu16(v, mean, rstd, cst + PI_C_LNOUT);
        publish_planes(v, g.ysp_hi, g.ysp_lo);
        pi_store16(g.y_pre + kr * PI_W + c16, pre);
        pi_store16(g.y + kr * PI_W + c16, v);
      }
      stamp(k, 6);
      // ---------------- imgs_stat + unimix categorical draw: thread = (row, group 2j + half) ----------
      {
        float l[32], u[32];
        const int grp = 2 * j + half;
        {
          // the uniforms do not depend on anything: fetched under the main loop
          const float4* up = reinterpret_cast<const float4*>(g.u_state + kr * PI_SC + grp * PI_C);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 t = __ldg(up + c4);
            u[4 * c4] = t.x; u[4 * c4 + 1] = t.y; u[4 * c4 + 2] = t.z; u[4 * c4 + 3] = t.w;
          }
        }
        const int seg[2] = {32 * half, 32 * half + 16};
        drain(std::integral_constant<int, 2>{}, 64, NKW, seg, l, no_between);
        const float* bi = cst + PI_C_BIMS + 32 * half;
#pragma unroll
        for (int c = 0; c < 32; ++c) l[c] += bi[c];
        put16(T0, PI_T2_LD, 32 * half, l);                           // logits: stored after the event
        put16(T0, PI_T2_LD, 32 * half + 16, l + 16);
        const int kk = pi_unimix_draw(l, u, PI_C, g.unimix);
        g.idx[(kr + N) * PI_S + grp] = kk;
        arrive();                                                    // event "sampled": awaited in the
        sampled_pending = true;                                      // next step, when the indices are needed
        flush(std::integral_constant<int, 64>{}, T0, PI_T2_LD, g.logit + (kb0 + N) * PI_SC + 2 * PI_C * j,
              PI_SC, nullptr, nullptr);
        float4* oh = reinterpret_cast<float4*>(g.feat + (kr + N) * F + grp * PI_C);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          oh[c4] = make_float4(kk == 4 * c4 ? 1.f : 0.f, kk == 4 * c4 + 1 ? 1.f : 0.f,
                               kk == 4 * c4 + 2 ? 1.f : 0.f, kk == 4 * c4 + 3 ? 1.f : 0.f);
        pi_epi_sync();                                               // T0 is free again
      }
      stamp(k, 7);
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
static unsigned long long* g_pi_stamps = nullptr;

int imagine_fwd_persistent(const dv3_rssm_dims* d, const dv3_rssm_params* p, const dv3_actor* a,
                           const dv3_imagine_io* io, const PiPlanes& pl, void* sync_ws,
                           cudaStream_t st, bool* used) {
  *used = false;
  // Opt-in (DV3_IMAGINE_PERSISTENT=1): measured on B200 at the reference's sizes (1024 x 15) the
  // rollout takes 2.16 ms as this one kernel against 1.88 ms as the stepwise launches replayed in a
  // CUDA graph -- the per-SM ingest of the A operand and the 11 dependent group events per step
  // bound it (profiles/README.md).
  const char* env = DV3_ENV("DV3_IMAGINE_PERSISTENT");
  if (!(env && env[0] == '1')) return 0;
  if (!a || !pl.ok) return 0;
  const int N = io->N, H = io->H, L = a->layers, A = d->actions;
  if (d->stoch != PI_S || d->classes != PI_C || d->deter != PI_W || d->hidden != PI_W ||
      a->units != PI_W || L < 1 || L > PI_MAX_L || A > 32 || (a->dist == 0 ? 2 * A : A) > 32)
    return 0;
  if (N % PI_BM != 0 || N <= 0 || H < 1) return 0;
  if (!(d->unimix > 0.f) || (a->dist != 0 && !(a->unimix > 0.f))) return 0;
  const int RB = N / PI_BM, G = RB * PI_GC;
  int dev = 0, coop = 0;
  DV3_CHECK_CUDA(cudaGetDevice(&dev));
  DV3_CHECK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop || G > sm_count()) return 0;
  auto kern = imagine_persistent_fwd_kernel;
  static DeviceOnce attr;
  if (attr.need())
    DV3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PI_SMEM));
  int per_sm = 0;
  DV3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PI_THREADS, PI_SMEM));
  if (per_sm * sm_count() < G) return 0;

  PiMaps mp;
  auto amap = [&](CUtensorMap* m, const float* hi, const float* lo) -> int {
    DV3_TRY(make_map2(m, hi, N, PI_W, PI_W, PI_BM, false));
    return make_map2(m + 1, lo, N, PI_W, PI_W, PI_BM, false);
  };
  auto bmap = [&](CUtensorMap* m, const PiPlane& w, int rows, int K) -> int {
    DV3_TRY(make_map2(m, w.hi, rows, K, w.ld, PI_SL, false));
    return make_map2(m + 1, w.lo, rows, K, w.ld, PI_SL, false);
  };
  for (int b = 0; b < 2; ++b) {
    DV3_TRY(amap(mp.dsp[b], pl.dsp_hi[b], pl.dsp_lo[b]));
    DV3_TRY(amap(mp.asp[b], pl.asp_hi[b], pl.asp_lo[b]));
  }
  DV3_TRY(amap(mp.xsp, pl.xsp_hi, pl.xsp_lo));
  DV3_TRY(amap(mp.ysp, pl.ysp_hi, pl.ysp_lo));
  for (int i = 0; i < PI_MAX_L; ++i) DV3_TRY(bmap(mp.wa[i], pl.wa[i < L ? i : 0], PI_W, PI_W));
  DV3_TRY(bmap(mp.gru, pl.gru, 3 * PI_W, 2 * PI_W));
  DV3_TRY(bmap(mp.out, pl.out, PI_W, PI_W));
  DV3_TRY(bmap(mp.ims, pl.ims, PI_SC, PI_W));

  PiArgs g{};
  g.N = N; g.H = H; g.L = L; g.A = A; g.dist = a->dist;
  g.eps = d->ln_eps; g.unimix = d->unimix; g.a_unimix = a->unimix;
  g.min_std = a->min_std; g.max_std = a->max_std;
  g.Wa0T = pl.Wa0T; g.WinT = pl.WinT;
  for (int i = 0; i < L; ++i) { g.a_ln_g[i] = a->ln_g[i]; g.a_ln_b[i] = a->ln_b[i]; }
  g.w_mean = a->w_mean; g.b_mean = a->b_mean; g.w_std = a->w_std; g.b_std = a->b_std;
  g.ln_in_g = p->ln_in_g; g.ln_in_b = p->ln_in_b; g.ln_gru_g = p->ln_gru_g; g.ln_gru_b = p->ln_gru_b;
  g.ln_out_g = p->ln_out_g; g.ln_out_b = p->ln_out_b; g.b_ims = p->b_ims;
  g.start_deter = io->start_deter; g.act_noise = io->act_noise; g.u_state = io->u_state;
  g.feat = io->feat; g.logit = io->logit; g.action = io->action; g.idx = io->idx;
  g.x_pre = io->x_pre; g.x = io->x; g.g_pre = io->g_pre; g.y_pre = io->y_pre; g.y = io->y;
  g.a_pre = io->a_pre; g.a_act = io->a_act; g.a_mean_raw = io->a_mean_raw; g.a_std_raw = io->a_std_raw;
  for (int b = 0; b < 2; ++b) {
    g.dsp_hi[b] = pl.dsp_hi[b]; g.dsp_lo[b] = pl.dsp_lo[b];
    g.asp_hi[b] = pl.asp_hi[b]; g.asp_lo[b] = pl.asp_lo[b];
  }
  g.xsp_hi = pl.xsp_hi; g.xsp_lo = pl.xsp_lo; g.ysp_hi = pl.ysp_hi; g.ysp_lo = pl.ysp_lo;
  // sync workspace: [RB][32] counters, one abort word, then the statistics buffers
  unsigned* sw = static_cast<unsigned*>(sync_ws);
  g.cnt = sw;
  g.abort_flag = sw + (size_t)RB * 32;
  g.stats = reinterpret_cast<float2*>(sw + (size_t)RB * 32 + 64);
  g.hpart = reinterpret_cast<float*>(g.stats + (size_t)RB * 2 * PI_GC * 256);
  DV3_CHECK_CUDA(cudaMemsetAsync(sw, 0, ((size_t)RB * 32 + 64) * sizeof(unsigned), st));
  g.stamps = nullptr;
  if (const char* te = DV3_ENV("DV3_IMAGINE_TIMING")) {
    if (te[0] == '1' && H <= 256) {
      if (!g_pi_stamps && cudaMalloc(&g_pi_stamps, 256 * PI_STAMPS * sizeof(unsigned long long)) != cudaSuccess)
        g_pi_stamps = nullptr;
      g.stamps = g_pi_stamps;
    }
  }
  void* args[] = {const_cast<PiMaps*>(&mp), &g};
  DV3_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(G), dim3(PI_THREADS),
                                             args, PI_SMEM, st));
  note_launch();
  *used = true;
  return 0;
}

size_t imagine_persistent_sync_bytes(int N) {
  const size_t RB = (size_t)(N + PI_BM - 1) / PI_BM;
  return (RB * 32 + 64) * sizeof(unsigned) + RB * 2 * PI_GC * 256 * sizeof(float2) +
         RB * PI_GC * 32 * PI_BM * sizeof(float) + 256;
}

}  // namespace dv3

// debug: [H][16] phase stamps (ns) of CTA 0 of the last persistent imagination launch run with
// DV3_IMAGINE_TIMING=1; also returns the watchdog word in host[H*16] when room is given
extern "C" int dv3_debug_imagine_timing(unsigned long long* host, int32_t H) {
  DV3_REQUIRE(dv3::g_pi_stamps && host && H > 0 && H <= 256, DV3_ERR_NULL,
              "debug_imagine_timing: no timing buffer (set DV3_IMAGINE_TIMING=1)");
  DV3_CHECK_CUDA(cudaMemcpy(host, dv3::g_pi_stamps, (size_t)H * dv3::PI_STAMPS * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost));
  return 0;
}
