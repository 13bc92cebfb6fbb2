// Persistent forward kernel of RSSM.observe for the latency-bound small-batch case (B <= 16).
//
// One cooperative launch runs all T steps.  The recurrent weights stay resident in shared memory
// for the whole sequence, column-sharded over the grid:
//     W_gru   [3D, Hd+D]  -> ceil(3D/G) output columns per CTA            (48 KB at D=Hd=512, G=128)
//     W_obs[:, :D] [Hd,D] -> ceil(Hd/G) columns per CTA                   ( 8 KB)
//     W_os    [S*C, Hd]   -> one categorical group (C columns) per CTA g < S   (64 KB)
// A step is four phases separated by a grid barrier (an atomic counter + generation flag in L2):
//   A  CTA b < B: previous-state select (is_first mix), one-hot gather of W_in^T, LN+SiLU -> x
//   B  all CTAs: g_pre[:, own columns] = W_gru [x, h]              (inputs read from L2)
//   C+D all CTAs: LN_3D + GRU gates for all rows (redundantly, rows kept in registers) -> h' in
//      smem; z_pre[:, own columns] = W_obs_d h' + (embed part, precomputed)
//   E  CTA g < S: LN+SiLU of z rows -> smem, logits of group g, unimix categorical draw
// Activations cross CTAs through global memory (L2) with .cg loads; every per-step tensor has its
// own address per t, so nothing stale can sit in L1.  All saved tensors of the stepwise path are
// written identically, so dv3_observe_bwd is unchanged.
//
// Reference: networks.py:174-206 (obs_step), 208-233 (img_step), 760-768 (GRUCell),
// tools.py:436-460 (OneHotDist), tools.py:806-850 (static_scan).
#include <cstdlib>
#include "dv3_persist.cuh"

namespace dv3 {

struct PoArgs {
  int B, T, S, C, D, Hd, A, E;
  float unimix, eps;
  const float *w_gru, *ln_gru_g, *ln_gru_b, *w_obs, *ln_obs_g, *ln_obs_b, *w_os, *b_os, *ln_in_g,
      *ln_in_b;
  const float *WinT, *pre_e;
  const float *action, *first_eff, *u_post;
  const int32_t* state_idx;
  const float* state_deter;
  const int32_t* init_idx;
  const float* init_deter;
  float *post_stoch, *post_logit, *deter, *hprev, *aprev, *x_pre, *x, *g_pre, *z_pre, *z;
  int32_t *post_idx, *sprev_idx;
  unsigned* bar;
  int ncg, bufw;
  int nrbC, rpbC, cpb;          // phase C/D: row blocks, rows per block, W_obs columns per CTA
  int nrbE, rpbE;               // phase E: row blocks per categorical group, rows per block
  unsigned long long* timing;   // debug: [T][8] %globaltimer stamps of CTA 0 (NULL = off)
};

__device__ __forceinline__ void po_stamp(const PoArgs& p, int t, int slot) {
  if (p.timing && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    p.timing[t * 8 + slot] = v;
  }
}

template <int DV>   // DV = D / 32
__global__ void __launch_bounds__(PO_THREADS, 1) observe_persistent_fwd_kernel(PoArgs p) {
  extern __shared__ __align__(16) float smf[];
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = p.B, T = p.T, S = p.S, C = p.C, D = p.D, Hd = p.Hd, A = p.A, E = p.E;
  const int SC = S * C, Kg = Hd + D, D3 = 3 * D;
  float* Wg = smf;
  float* Wo = Wg + (size_t)p.ncg * Kg;                    // [cpb][D + 4]
  float* Ws = Wo + (size_t)p.cpb * (D + 4);               // [C][Hd + 4]
  float* zs = Ws + (size_t)C * (Hd + 4);                  // [8 warps][Hd]  (phase A reuses row 0)
  float* hs = zs + (size_t)PO_WARPS * Hd;                 // [4 rows][D]    h' of this CTA's rows
  float* lng = hs + (size_t)4 * D;                        // GRU LN gamma [3D], beta [3D]
  float* lnb = lng + D3;
  float* lzg = lnb + D3;                                  // obs LN gamma [Hd], beta [Hd]
  float* lzb = lzg + Hd;
  float* part = lzb + Hd;
  float* red = part + PO_WARPS * PO_NV;
  int* sidx = reinterpret_cast<int*>(red + 128);
  float* sact = reinterpret_cast<float*>(sidx + ((S + 3) & ~3));
  float* buf = zs;
  float* xch = part;                                      // [4][32] warp-pair exchange

  const int g0 = min(cta * p.ncg, D3), gn = min(p.ncg, D3 - g0);
  // phase C/D: CTA = (row block rbC, column block cbC) -> rows rC0.., W_obs columns o0..o0+on
  const int rbC = cta % p.nrbC, cbC = cta / p.nrbC;
  const int rC0 = rbC * p.rpbC, rCn = max(0, min(p.rpbC, B - rC0));
  const int o0 = min(cbC * p.cpb, Hd), on = min(p.cpb, Hd - o0);
  // phase E: CTA = (group gE, row block rbE)
  const int gE = cta % S, rbE = cta / S;
  const int rE0 = rbE * p.rpbE, rEn = (rbE < p.nrbE) ? max(0, min(p.rpbE, B - rE0)) : 0;

  // resident weight slices
  for (int i = tid * 4; i < gn * Kg; i += PO_THREADS * 4)
    *reinterpret_cast<float4*>(Wg + i) =
        __ldg(reinterpret_cast<const float4*>(p.w_gru + (size_t)g0 * Kg + i));
  for (int i = tid * 4; i < on * D; i += PO_THREADS * 4) {
    const int r = i / D, k = i % D;
    *reinterpret_cast<float4*>(Wo + (size_t)r * (D + 4) + k) =
        __ldg(reinterpret_cast<const float4*>(p.w_obs + (size_t)(o0 + r) * (D + E) + k));
  }
  if (rEn > 0)
    for (int i = tid * 4; i < C * Hd; i += PO_THREADS * 4) {
      const int r = i / Hd, k = i % Hd;
      *reinterpret_cast<float4*>(Ws + (size_t)r * (Hd + 4) + k) =
          __ldg(reinterpret_cast<const float4*>(p.w_os + (size_t)gE * C * Hd + i));
    }
  for (int i = tid; i < D3; i += PO_THREADS) { lng[i] = p.ln_gru_g[i]; lnb[i] = p.ln_gru_b[i]; }
  for (int i = tid; i < Hd; i += PO_THREADS) { lzg[i] = p.ln_obs_g[i]; lzb[i] = p.ln_obs_b[i]; }
  __syncthreads();

  // constants the time loop would otherwise re-fetch from L2 after every barrier
  float lin_g[4], lin_b[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = tid + u * PO_THREADS;
    lin_g[u] = i < Hd ? p.ln_in_g[i] : 0.f;
    lin_b[u] = i < Hd ? p.ln_in_b[i] : 0.f;
  }
  const float bos = (lane < C) ? p.b_os[(size_t)gE * C + lane] : 0.f;
  const int init_i = tid < S ? p.init_idx[tid] : 0;
  float init_h[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = tid + u * PO_THREADS;
    init_h[u] = j < D ? p.init_deter[j] : 0.f;
  }

  unsigned gen = 0;
  for (int t = 0; t < T; ++t) {
    // ---------------- phase A: row b = cta ----------------
    po_stamp(p, t, 0);
    if (cta < B) {
      const int b = cta;
      const size_t bt = (size_t)b * T + t;
      // the reset flag, the previous state and the action are fetched TOGETHER (the previous state
      // through a pointer that is valid either way) and selected afterwards: one L2 round trip
      // instead of flag -> state
      const int32_t* pidx = t ? p.post_idx + (bt - 1) * S : (p.state_idx ? p.state_idx + (size_t)b * S : nullptr);
      const float* ph = t ? p.deter + (bt - 1) * D : (p.state_deter ? p.state_deter + (size_t)b * D : nullptr);
      const int32_t* pidx_l = pidx ? pidx : p.init_idx;
      const float* ph_l = ph ? ph : p.init_deter;
      const float fflag = __ldcg(p.first_eff + bt);
      int vi = 0;
      float vh[4], va = 0.f;
      if (tid < S) vi = __ldcg(pidx_l + tid);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = tid + u * PO_THREADS;
        vh[u] = j < D ? __ldcg(ph_l + j) : 0.f;
      }
      if (tid < A) va = __ldcg(p.action + bt * A + tid);
      const bool first = fflag != 0.f || pidx == nullptr;
      if (tid < S) {
        const int v = first ? init_i : vi;
        p.sprev_idx[bt * S + tid] = v;
        sidx[tid] = v + tid * C;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = tid + u * PO_THREADS;
        if (j < D) p.hprev[bt * D + j] = first ? init_h[u] : vh[u];
      }
      for (int j = tid + 4 * PO_THREADS; j < D; j += PO_THREADS)       // (D > 1024 never reaches here)
        p.hprev[bt * D + j] = first ? p.init_deter[j] : __ldcg(ph_l + j);
      if (tid < A) {
        const float v = first ? 0.f : va;
        p.aprev[bt * A + tid] = v;
        sact[tid] = v;
      }
      __syncthreads();
      // (the barrier's acquire has dropped L1: every load below is an L2 round trip, so all S + A of
      // a column are issued before the first add; same summation order as the stepwise kernel)
      for (int i = tid; i < Hd; i += PO_THREADS) {
        float acc = 0.f;
        if (S == 32) {
          float w[32];
#pragma unroll
          for (int s = 0; s < 32; ++s) w[s] = __ldg(p.WinT + (size_t)sidx[s] * Hd + i);
#pragma unroll
          for (int s = 0; s < 32; ++s) acc += w[s];
        } else {
          for (int s = 0; s < S; ++s) acc += p.WinT[(size_t)sidx[s] * Hd + i];
        }
        const float* wa = p.WinT + (size_t)SC * Hd + i;
        if (A <= 8) {
          float w[8];
#pragma unroll
          for (int a = 0; a < 8; ++a) w[a] = a < A ? __ldg(wa + (size_t)a * Hd) : 0.f;
#pragma unroll
          for (int a = 0; a < 8; ++a) if (a < A) acc = fmaf(sact[a], w[a], acc);
        } else {
          for (int a = 0; a < A; ++a) acc = fmaf(sact[a], wa[(size_t)a * Hd], acc);
        }
        buf[i] = acc;
        p.x_pre[bt * Hd + i] = acc;
      }
      __syncthreads();
      float s1[1] = {0.f};
      for (int i = tid; i < Hd; i += PO_THREADS) s1[0] += buf[i];
      block_sum<1>(s1, red);
      const float mean = s1[0] / (float)Hd;
      float s2[1] = {0.f};
      for (int i = tid; i < Hd; i += PO_THREADS) {
        const float dd = buf[i] - mean;
        s2[0] = fmaf(dd, dd, s2[0]);
      }
      block_sum<1>(s2, red);
      const float rstd = 1.f / sqrtf(s2[0] / (float)Hd + p.eps);
      for (int i = tid, u = 0; i < Hd; i += PO_THREADS, ++u) {
        const float gg = u < 4 ? lin_g[u < 4 ? u : 0] : p.ln_in_g[i];
        const float bb = u < 4 ? lin_b[u < 4 ? u : 0] : p.ln_in_b[i];
        p.x[bt * Hd + i] = siluf_(fmaf((buf[i] - mean) * rstd, gg, bb));
      }
    }
    po_stamp(p, t, 1);
    grid_barrier(p.bar, G, gen);
    po_stamp(p, t, 2);

    // ---------------- phase B: GRU pre-activations, own columns ----------------
    if (gn > 0) {
      float* gp = p.g_pre;
      gemv16<true, 6, 1>(Wg, gn, Kg, p.x + (size_t)t * Hd, T * Hd, Hd, p.hprev + (size_t)t * D, T * D, B,
                   part, [&](int m, int c, float r) {
                     gp[((size_t)m * T + t) * D3 + g0 + c] = r;
                   });
    }
    po_stamp(p, t, 3);
    grid_barrier(p.bar, G, gen);
    po_stamp(p, t, 4);

    // ------- phase C+D.  CTA (row block, column block): two warps per row.  Each warp takes the
    // LN_3D statistics of the whole row but evaluates the gates for its half of D only; h' goes to
    // smem, then lane = (W_obs column, quarter of D) does a serial dot over its quarter -------
    {
      const int rl = warp & 3, half = warp >> 2;
      const bool act = rl < rCn && on > 0;
      const int b = rC0 + rl;
      const size_t bt = (size_t)b * T + t;
      // the hoisted embed half of z_pre is an input: fetched with the phase's first loads, not after
      // the dot product (an exposed L2 round trip at the end of the phase)
      const float pre_e_v = (half == 0 && act && lane < on) ? __ldg(p.pre_e + bt * Hd + o0 + lane) : 0.f;
      if (act) {
        const float* row = p.g_pre + bt * D3;
        float v[3 * DV], hp[DV / 2];
        const int jb = half * (D >> 1);
#pragma unroll
        for (int i = 0; i < 3 * DV; ++i) v[i] = __ldcg(row + lane + 32 * i);
#pragma unroll
        for (int i = 0; i < DV / 2; ++i) hp[i] = __ldcg(p.hprev + bt * D + jb + lane + 32 * i);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 3 * DV; ++i) s += v[i];
        const float mean = warp_sum(s) / (float)D3;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 3 * DV; ++i) {
          const float dd = v[i] - mean;
          q = fmaf(dd, dd, q);
        }
        const float rstd = 1.f / sqrtf(warp_sum(q) / (float)D3 + p.eps);
#pragma unroll
        for (int i = 0; i < DV / 2; ++i) {
          // this warp's half of each part (selects, so that v[] keeps static register indices)
          const float vr = half ? v[DV / 2 + i] : v[i];
          const float vc = half ? v[DV + DV / 2 + i] : v[DV + i];
          const float vu = half ? v[2 * DV + DV / 2 + i] : v[2 * DV + i];
          const int j = jb + lane + 32 * i;
          const float pr = fmaf((vr - mean) * rstd, lng[j], lnb[j]);
          const float pc = fmaf((vc - mean) * rstd, lng[D + j], lnb[D + j]);
          const float pu = fmaf((vu - mean) * rstd, lng[2 * D + j], lnb[2 * D + j]);
          const float rg = sigmoidf_(pr);
          const float cc = tanhf(rg * pc);
          const float u = sigmoidf_(pu - 1.f);
          const float hn = u * cc + (1.f - u) * hp[i];
          hs[(size_t)rl * D + j] = hn;
          if (cbC == 0) p.deter[bt * D + j] = hn;
        }
      }
      __syncthreads();
      float acc = 0.f;
      const int cl = lane & 15, jq = half * 2 + (lane >> 4);
      if (act && cl < on) {
        const int j0 = jq * (D >> 2);
        const float* w = Wo + (size_t)cl * (D + 4) + j0;
        const float* h = hs + (size_t)rl * D + j0;
        float a0 = 0.f, a1 = 0.f;
        for (int k = 0; k < (D >> 2); k += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(w + k);
          const float4 hv = *reinterpret_cast<const float4*>(h + k);
          a0 = fmaf(hv.x, wv.x, a0); a1 = fmaf(hv.y, wv.y, a1);
          a0 = fmaf(hv.z, wv.z, a0); a1 = fmaf(hv.w, wv.w, a1);
        }
        acc = a0 + a1;
      }
      acc += __shfl_xor_sync(FULL, acc, 16);
      if (half == 1 && lane < 16) xch[rl * 32 + lane] = acc;
      __syncthreads();
      if (half == 0 && act && lane < on) {
        const size_t o = bt * Hd + o0 + lane;
        p.z_pre[o] = acc + xch[rl * 32 + lane] + pre_e_v;
      }
    }
    po_stamp(p, t, 5);
    grid_barrier(p.bar, G, gen);
    po_stamp(p, t, 6);

    // ------- phase E.  CTA (group, row block): two warps per row, lane = class.  Each warp
    // normalises and dots its own half of the z row (the serial expf / divide chain per lane is
    // what these phases cost); row sums and partial logits meet in smem, then the unimix draw ---
    {
      const int rl = warp & 3, half = warp >> 2;
      const bool act = rl < rEn;
      const int b = rE0 + rl;
      const size_t bt = (size_t)b * T + t;
      const bool valid = lane < C;
      float uu = 1.f, a0 = 0.f, a1 = 0.f;
      // LN + SiLU of the z row, split over the pair: each warp normalises the half of the row it
      // will dot against W_os (columns half*Hd/2 ..), the two row sums meet in shared memory
      if (act && half == 0 && valid) uu = __ldcs(p.u_post + (((size_t)t * B + b) * S + gE) * C + lane);
      const int nh = Hd >> 1, j0 = half * nh;
      float zv[16];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        zv[i] = (act && lane + 32 * i < nh) ? __ldcg(p.z_pre + bt * Hd + j0 + lane + 32 * i) : 0.f;
        s += zv[i];
      }
      s = warp_sum(s);
      if (act && lane == 0) red[rl * 8 + half] = s;
      __syncthreads();
      const float mean = (red[rl * 8] + red[rl * 8 + 1]) / (float)Hd;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float dd = (lane + 32 * i < nh) ? zv[i] - mean : 0.f;
        q = fmaf(dd, dd, q);
      }
      q = warp_sum(q);
      if (act && lane == 0) red[rl * 8 + 2 + half] = q;
      __syncthreads();
      if (act) {
        const float rstd = 1.f / sqrtf((red[rl * 8 + 2] + red[rl * 8 + 3]) / (float)Hd + p.eps);
        float* zr = zs + (size_t)rl * Hd;
        const bool wr = gE == 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = j0 + lane + 32 * i;
          if (lane + 32 * i < nh) {
            const float y = siluf_(fmaf((zv[i] - mean) * rstd, lzg[j], lzb[j]));
            zr[j] = y;
            if (wr) p.z[bt * Hd + j] = y;
          }
        }
        __syncwarp();
        const float* wrow = Ws + (size_t)min(lane, C - 1) * (Hd + 4) + j0;
        const float* x = zr + j0;
        for (int k = 0; k < nh; k += 4) {
          const float4 w = *reinterpret_cast<const float4*>(wrow + k);
          const float4 xv = *reinterpret_cast<const float4*>(x + k);
          a0 = fmaf(xv.x, w.x, a0); a1 = fmaf(xv.y, w.y, a1);
          a0 = fmaf(xv.z, w.z, a0); a1 = fmaf(xv.w, w.w, a1);
        }
        if (half == 1) xch[rl * 32 + lane] = a0 + a1;
      }
      __syncthreads();
      if (act && half == 0) {
        const float l = valid ? (a0 + a1) + xch[rl * 32 + lane] + bos : 0.f;
        const Unimix um = unimix_probs(l, valid, C, p.unimix);
        const int k = warp_argmax(um.probs / (-logf(uu)), valid, lane);
        if (valid) {
          const size_t o = (bt * S + gE) * C + lane;
          p.post_logit[o] = l;
          p.post_stoch[o] = (lane == k) ? 1.f : 0.f;
        }
        if (lane == 0) p.post_idx[bt * S + gE] = k;
      }
    }
    po_stamp(p, t, 7);
    grid_barrier(p.bar, G, gen);
  }
}

static unsigned long long* g_po_timing = nullptr;

// debug stamp buffer ([4096][8] u64), allocated on first use when DV3_OBSERVE_TIMING=1
unsigned long long* po_timing_buffer() {
  const char* te = DV3_ENV("DV3_OBSERVE_TIMING");
  if (!te || (te[0] != '1' && te[0] != '2')) return nullptr;
  if (!g_po_timing && cudaMalloc(&g_po_timing, 4096 * 8 * sizeof(unsigned long long)) != cudaSuccess)
    return nullptr;
  return g_po_timing;
}

static size_t po_smem_bytes(const PoArgs& a) {
  const size_t fl = (size_t)a.ncg * (a.Hd + a.D) + (size_t)a.cpb * (a.D + 4) +
                    (size_t)a.C * (a.Hd + 4) + (size_t)PO_WARPS * a.Hd + 4 * (size_t)a.D +
                    6 * (size_t)a.D + 2 * (size_t)a.Hd + PO_WARPS * PO_NV + 128 +
                    ((a.S + 3) & ~3) + ((a.A + 3) & ~3) + 16;
  return fl * 4;
}

template <int DV>
static int po_launch(PoArgs& a, int G, size_t smem, cudaStream_t st, bool* used) {
  auto kern = observe_persistent_fwd_kernel<DV>;
  DV3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  DV3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PO_THREADS, smem));
  int dev = 0, sms = 0;
  DV3_CHECK_CUDA(cudaGetDevice(&dev));
  DV3_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (per_sm * sms < G) return 0;   // cannot be co-resident: caller falls back to the stepwise path
  DV3_CHECK_CUDA(cudaMemsetAsync(a.bar, 0, 2 * sizeof(unsigned), st));
  void* args[] = {&a};
  DV3_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(G), dim3(PO_THREADS),
                                             args, smem, st));
  note_launch();
  *used = true;
  return 0;
}

// Runs the T-step recurrence of observe_fwd in one persistent kernel when the shapes allow it.
// *used == false on return means "not applicable": the caller runs the stepwise launches.
int observe_fwd_persistent(const dv3_rssm_dims* d, const dv3_rssm_params* p,
                           const dv3_observe_io* io, const float* WinT, const float* pre_e,
                           unsigned* bar, cudaStream_t st, bool* used) {
  *used = false;
  const char* env = DV3_ENV("DV3_OBSERVE_STEPWISE");
  if (env && env[0] == '1') return 0;
  const int D = d->deter, Hd = d->hidden, S = d->stoch, C = d->classes;
  if (io->B > PO_ROWS || D % 32 != 0 || (D / 32 != 2 && D / 32 != 4 && D / 32 != 8 && D / 32 != 16))
    return 0;
  if (Hd % 4 != 0 || Hd > 1024 || d->embed % 4 != 0 || C > 32) return 0;
  int dev = 0, sms = 0, coop = 0;
  DV3_CHECK_CUDA(cudaGetDevice(&dev));
  DV3_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DV3_CHECK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return 0;
  const int G = sms >= 128 ? 128 : sms;
  if (S > G || io->B > G || d->actions > PO_THREADS) return 0;
  PoArgs a{};
  a.B = io->B; a.T = io->T; a.S = S; a.C = C; a.D = D; a.Hd = Hd; a.A = d->actions; a.E = d->embed;
  a.unimix = d->unimix; a.eps = d->ln_eps;
  a.w_gru = p->w_gru; a.ln_gru_g = p->ln_gru_g; a.ln_gru_b = p->ln_gru_b; a.w_obs = p->w_obs;
  a.ln_obs_g = p->ln_obs_g; a.ln_obs_b = p->ln_obs_b; a.w_os = p->w_os; a.b_os = p->b_os;
  a.ln_in_g = p->ln_in_g; a.ln_in_b = p->ln_in_b;
  a.WinT = WinT; a.pre_e = pre_e;
  a.action = io->action; a.first_eff = io->first_eff; a.u_post = io->u_post;
  a.state_idx = io->state_idx; a.state_deter = io->state_deter;
  a.init_idx = io->init_idx; a.init_deter = io->init_deter;
  a.post_stoch = io->post_stoch; a.post_logit = io->post_logit; a.deter = io->deter;
  a.hprev = io->hprev; a.aprev = io->aprev; a.x_pre = io->x_pre; a.x = io->x; a.g_pre = io->g_pre;
  a.z_pre = io->z_pre; a.z = io->z; a.post_idx = io->post_idx; a.sprev_idx = io->sprev_idx;
  a.bar = bar;
  a.timing = nullptr;
  if (const char* te = DV3_ENV("DV3_OBSERVE_TIMING"))
    if (te[0] == '1' && io->T <= 4096) a.timing = po_timing_buffer();
  a.ncg = (3 * D + G - 1) / G;
  a.bufw = ((D > Hd ? D : Hd) + 3) & ~3;
  // phase C/D: up to 4 rows per CTA (two warps per row); the W_obs columns go over the rest
  a.nrbC = (io->B + 3) / 4;
  a.rpbC = (io->B + a.nrbC - 1) / a.nrbC;
  const int ncb = G / a.nrbC;
  a.cpb = (Hd + ncb - 1) / ncb;
  // phase E: G / S row blocks per categorical group
  a.nrbE = G / S;
  if (a.nrbE > io->B) a.nrbE = io->B;
  a.rpbE = (io->B + a.nrbE - 1) / a.nrbE;
  if (ncb < 1 || a.cpb > 16 || a.rpbC > 4 || a.rpbE > 4 || Hd % 8 != 0) return 0;
  const size_t smem = po_smem_bytes(a);
  if (smem > 220 * 1024) return 0;
  switch (D / 32) {
    case 2: return po_launch<2>(a, G, smem, st, used);
    case 4: return po_launch<4>(a, G, smem, st, used);
    case 8: return po_launch<8>(a, G, smem, st, used);
    default: return po_launch<16>(a, G, smem, st, used);
  }
}

}  // namespace dv3

// debug: copies the [T][8] phase stamps of the last timed persistent observe launch (ns)
extern "C" int dv3_debug_observe_timing(unsigned long long* host, int32_t T) {
  DV3_REQUIRE(dv3::g_po_timing && host && T > 0 && T <= 4096, DV3_ERR_NULL,
              "debug_observe_timing: no timing buffer (set DV3_OBSERVE_TIMING=1)");
  DV3_CHECK_CUDA(cudaMemcpy(host, dv3::g_po_timing, (size_t)T * 8 * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost));
  return 0;
}
