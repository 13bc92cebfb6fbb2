// Persistent backward kernel of RSSM.observe for the latency-bound small-batch case (B <= 16):
// the reverse-time recurrence of dv3_observe_bwd in one cooperative launch.
//
// The four transposed recurrent weights stay resident in shared memory for the whole sequence,
// sharded over a grid of S x ceil(B/4) CTAs; CTA = (row block rb of <= 4 sequences, column
// block cb = one categorical group):
//     W_os^T    [Hd, S*C]   -> ceil(Hd/G) rows per CTA            (all-row GEMV, phase 1b)
//     W_obs_d^T [D, Hd]     -> ceil(D/S) rows per CTA             (row-block GEMV, phase 2)
//     W_gru^T   [Hd+D, 3D]  -> ceil((Hd+D)/G) rows per CTA        (all-row GEMV, phase 3b)
//     W_in_s^T  [S*C, Hd]   -> the C rows of group cb             (row-block GEMV, phase 4)
// One step (t = T-1 .. 0) is six phases with five grid barriers:
//   1a  straight-through backward of the posterior sample, warp per (row, group cb), lane = class;
//       its recurrent input d stoch is what this same CTA produced in phase 4 of step t+1
//   1b  d_z = d_post_logit W_os                       (K = S*C, all rows, own columns)
//   2   LN+SiLU backward of z rows (two warps per row) -> d_z_pre; dh_z = d_z_pre W_obs_d
//   3a  GRU gate + LN_3D backward of this block's rows (CTAs with cb == 0) -> d_g_pre, direct dh
//   3b  [dx | dh_prev] = d_g_pre W_gru                (K = 3D, all rows, own columns)
//   4   LN+SiLU backward of x rows -> d_x_pre; d stoch_prev (group cb) = d_x_pre W_in_s; reset
//       routing (rows with is_first send their state gradient to RSSM.initial)
// Activations cross CTAs through global memory (L2) with .cg loads.  All per-step outputs are
// written exactly as the stepwise path writes them, so the bulk parameter-gradient contractions
// that follow are unchanged.
//
// Reference: the autograd backward of networks.py:174-233 (obs_step / img_step), 760-768 (GRUCell),
// tools.py:436-460 (straight-through OneHotDist), driven by tools.py:806-850 (static_scan).
#include <cstdlib>
#include "dv3_persist.cuh"

namespace dv3 {

struct PbArgs {
  int B, T, S, C, D, Hd;
  float unimix, eps;
  const float *first_eff, *post_logit, *hprev, *x_pre, *g_pre, *z_pre;
  const float *g_post_stoch, *g_post_logit, *g_deter, *dh_prior;
  const float *WosT, *WobsT, *WgruT, *WinT;
  const float *ln_obs_g, *ln_obs_b, *ln_gru_g, *ln_gru_b, *ln_in_g, *ln_in_b;
  float *d_post_logit, *d_z_pre, *d_z_ln, *d_g_pre, *d_g_ln, *d_x_pre, *d_x_ln;
  float *ds_rec, *dh_rec, *dinit_s, *dinit_h;
  float *d_z, *dh_z, *dhdir, *dxh;
  unsigned* bar;
  int nrb, ncz, cpd, ncx;
  unsigned long long* timing;
};

__device__ __forceinline__ void pb_stamp(const PbArgs& p, int t, int slot) {
  if (p.timing && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    p.timing[t * 8 + slot] = v;
  }
}

// LN+SiLU backward of one row shared by the two warps of a row pair (half 0 / 1): each warp owns
// the columns half*n/2 + lane + 32 i (i < HV2) and the three row-wide sums (mean, variance, the
// two LN-backward moments) meet in shared memory.  Halving the per-lane element count halves the
// serial chain of expf / divide latencies that dominates these phases (measured 4.6 of 7.6 us
// with one warp per row).  Must be called by ALL threads of the CTA (three block barriers);
// `act` gates the row.  xr: shared scratch, 8 floats per row.
//   in : pre (saved Linear output), dout (gradient w.r.t. the SiLU output) for the owned columns
//   emit(j, d_pre, d_ln) is called once per owned column
template <int HV2, typename Emit>
__device__ __forceinline__ void pair_ln_silu_bwd(bool act, const float (&pre)[HV2], float (&dout)[HV2],
                                                 int n, int lane, int half, const float* g,
                                                 const float* b, float eps, float* xr, Emit emit) {
  const int nh = n >> 1, j0 = half * nh;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < HV2; ++i) s += (lane + 32 * i < nh) ? pre[i] : 0.f;
  s = warp_sum(s);
  if (act && lane == 0) xr[half] = s;
  __syncthreads();
  const float mean = (xr[0] + xr[1]) / (float)n;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < HV2; ++i) {
    const float d = (lane + 32 * i < nh) ? pre[i] - mean : 0.f;
    q = fmaf(d, d, q);
  }
  q = warp_sum(q);
  if (act && lane == 0) xr[2 + half] = q;
  __syncthreads();
  const float rstd = 1.f / sqrtf((xr[2] + xr[3]) / (float)n + eps);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int i = 0; i < HV2; ++i) {
    const int j = j0 + lane + 32 * i;
    if (act && lane + 32 * i < nh) {
      const float xh = (pre[i] - mean) * rstd;
      const float v = fmaf(xh, g[j], b[j]);
      const float dv = dout[i] * silu_grad(v);
      const float dx = dv * g[j];
      dout[i] = dv;
      a0 += dx;
      a1 = fmaf(dx, xh, a1);
    }
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if (act && lane == 0) { xr[4 + half] = a0; xr[6 + half] = a1; }
  __syncthreads();
  const float m1 = (xr[4] + xr[5]) / (float)n, m2 = (xr[6] + xr[7]) / (float)n;
#pragma unroll
  for (int i = 0; i < HV2; ++i) {
    const int j = j0 + lane + 32 * i;
    if (act && lane + 32 * i < nh) {
      const float xh = (pre[i] - mean) * rstd;
      emit(j, rstd * (dout[i] * g[j] - m1 - xh * m2), dout[i]);
    }
  }
}

template <int DV, int HV>   // DV = D / 32, HV >= ceil(Hd / 32)
__global__ void __launch_bounds__(PO_THREADS, 1) observe_persistent_bwd_kernel(PbArgs p) {
  extern __shared__ __align__(16) float smf[];
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = p.B, T = p.T, S = p.S, C = p.C, D = p.D, Hd = p.Hd;
  const int SC = S * C, D3 = 3 * D, HD = Hd + D;
  float* Wz = smf;                                  // [ncz][SC]
  float* Wd = Wz + (size_t)p.ncz * SC;              // [cpd][Hd + 4]
  float* Wx = Wd + (size_t)p.cpd * (Hd + 4);        // [ncx][3D]
  float* Wi = Wx + (size_t)p.ncx * D3;              // [C][Hd + 4]
  float* lzg = Wi + (size_t)C * (Hd + 4);           // LN(obs) gamma/beta [Hd] each
  float* lzb = lzg + Hd;
  float* lxg = lzb + Hd;                            // LN(in) gamma/beta [Hd] each
  float* lxb = lxg + Hd;
  float* lgg = lxb + Hd;                            // LN(gru) gamma/beta [3D] each
  float* lgb = lgg + D3;
  float* rows = lgb + D3;                           // [4][Hd]  d_z_pre / d_x_pre of this block
  float* part = rows + (size_t)4 * Hd;              // gemv16 fold
  float* xch = part + PO_WARPS * PO_NV;             // [4][32] warp-pair exchange
  float* dsl = xch + 4 * 32;                        // [4][32] recurrent d stoch of (rows, group cb)
  float* pr2 = dsl + 4 * 32;                        // [4][8] warp-pair reductions (LN sums)

  const int rb = cta % p.nrb, cb = cta / p.nrb;
  const int r0 = rb * 4, rn = max(0, min(4, B - r0));
  const int z0 = min(cta * p.ncz, Hd), zn = min(p.ncz, Hd - z0);
  const int d0 = min(cb * p.cpd, D), dn = min(p.cpd, D - d0);
  const int x0 = min(cta * p.ncx, HD), xn = min(p.ncx, HD - x0);

  // resident weight slices + LayerNorm parameters
  for (int i = tid * 4; i < zn * SC; i += PO_THREADS * 4)
    *reinterpret_cast<float4*>(Wz + i) =
        __ldg(reinterpret_cast<const float4*>(p.WosT + (size_t)z0 * SC + i));
  for (int i = tid * 4; i < dn * Hd; i += PO_THREADS * 4) {
    const int r = i / Hd, k = i % Hd;
    *reinterpret_cast<float4*>(Wd + (size_t)r * (Hd + 4) + k) =
        __ldg(reinterpret_cast<const float4*>(p.WobsT + (size_t)(d0 + r) * Hd + k));
  }
  for (int i = tid * 4; i < xn * D3; i += PO_THREADS * 4)
    *reinterpret_cast<float4*>(Wx + i) =
        __ldg(reinterpret_cast<const float4*>(p.WgruT + (size_t)x0 * D3 + i));
  for (int i = tid * 4; i < C * Hd; i += PO_THREADS * 4) {
    const int r = i / Hd, k = i % Hd;
    *reinterpret_cast<float4*>(Wi + (size_t)r * (Hd + 4) + k) =
        __ldg(reinterpret_cast<const float4*>(p.WinT + (size_t)(cb * C + r) * Hd + k));
  }
  for (int i = tid; i < Hd; i += PO_THREADS) {
    lzg[i] = p.ln_obs_g[i]; lzb[i] = p.ln_obs_b[i];
    lxg[i] = p.ln_in_g[i];  lxb[i] = p.ln_in_b[i];
  }
  for (int i = tid; i < D3; i += PO_THREADS) { lgg[i] = p.ln_gru_g[i]; lgb[i] = p.ln_gru_b[i]; }
  for (int i = tid; i < 4 * 32; i += PO_THREADS) dsl[i] = 0.f;
  __syncthreads();

  const int rl = warp & 3, half = warp >> 2;
  const bool ract = rl < rn;
  const int b = r0 + rl;
  unsigned gen = 0;
  for (int t = T - 1; t >= 0; --t) {
    const size_t bt = (size_t)b * T + t;

    pb_stamp(p, t, 0);
    // The saved activations of step t-1 were written by the forward pass milliseconds ago and
    // have left L2; the warps that idle through phase 1a pull them in now so that the phases of
    // the next step start from L2 instead of DRAM latency.
    if (half == 1 && ract && t > 0) {
      const size_t bp = bt - 1;
      auto pf = [&](const float* base, int nfloats) {
        for (int i = lane * 32; i < nfloats; i += 32 * 32)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + i));
      };
      pf(p.z_pre + bp * Hd, Hd);
      pf(p.x_pre + bp * Hd, Hd);
      pf(p.post_logit + bp * SC + (size_t)cb * C, C);
      if (p.g_post_stoch) pf(p.g_post_stoch + bp * SC + (size_t)cb * C, C);
      if (p.g_post_logit) pf(p.g_post_logit + bp * SC + (size_t)cb * C, C);
      if (cb == 0) {
        pf(p.g_pre + bp * D3, D3);
        pf(p.hprev + bp * D, D);
        pf(p.dh_prior + bp * D, D);
        if (p.g_deter) pf(p.g_deter + bp * D, D);
      }
    }
    // ---------------- phase 1a: straight-through backward of the posterior sample ------------
    if (half == 0 && ract) {
      const bool valid = lane < C;
      const size_t o = bt * SC + (size_t)cb * C + lane;
      const float l = valid ? p.post_logit[o] : 0.f;
      float g = valid ? dsl[rl * 32 + lane] : 0.f;
      if (valid && p.g_post_stoch) g += p.g_post_stoch[o];
      const float m = warp_max(valid ? l : -INFINITY);
      const float e = valid ? expf(l - m) : 0.f;
      const float pp = e / warp_sum(e);
      const float q = valid ? (pp * (1.f - p.unimix) + p.unimix / (float)C) : 0.f;
      const float gq = g - warp_sum(g * q);
      const float dot = warp_sum(valid ? gq * pp : 0.f);
      float dl = (1.f - p.unimix) * pp * (gq - dot);
      if (valid) {
        if (p.g_post_logit) dl += p.g_post_logit[o];
        p.d_post_logit[o] = dl;
      }
    }
    grid_barrier(p.bar, G, gen);

    pb_stamp(p, t, 1);
    // ---------------- phase 1b: d_z = d_post_logit W_os, own columns, all rows ---------------
    if (zn > 0) {
      float* dz = p.d_z;
      gemv16<true, 4, 1>(Wz, zn, SC, p.d_post_logit + (size_t)t * SC, T * SC, SC, nullptr, 0, B, part,
                   [&](int m, int c, float r) { dz[(size_t)m * Hd + z0 + c] = r; });
    }
    grid_barrier(p.bar, G, gen);

    pb_stamp(p, t, 2);
    // ---------------- phase 2: LN+SiLU backward of z rows; dh_z = d_z_pre W_obs_d ------------
    {
      {
        constexpr int HV2 = (HV + 1) / 2;
        float pre[HV2], dout[HV2];
        const int nh = Hd >> 1;
#pragma unroll
        for (int i = 0; i < HV2; ++i) {
          const int j = half * nh + lane + 32 * i;
          const bool on = ract && lane + 32 * i < nh;
          pre[i] = on ? p.z_pre[bt * Hd + j] : 0.f;
          dout[i] = on ? __ldcg(p.d_z + (size_t)b * Hd + j) : 0.f;
        }
        float* rrow = rows + (size_t)rl * Hd;
        const bool wr = cb == 0;
        pair_ln_silu_bwd<HV2>(ract, pre, dout, Hd, lane, half, lzg, lzb, p.eps, pr2 + rl * 8,
                              [&](int j, float dp_, float dl_) {
                                rrow[j] = dp_;
                                if (wr) { p.d_z_pre[bt * Hd + j] = dp_; p.d_z_ln[bt * Hd + j] = dl_; }
                              });
      }
      __syncthreads();
      const int cl = lane & 15, kq = half * 2 + (lane >> 4);
      float acc = 0.f;
      if (ract && cl < dn) {
        const int k0 = kq * (Hd >> 2);
        const float* w = Wd + (size_t)cl * (Hd + 4) + k0;
        const float* x = rows + (size_t)rl * Hd + k0;
        float a0 = 0.f, a1 = 0.f;
        for (int k = 0; k < (Hd >> 2); k += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(w + k);
          const float4 xv = *reinterpret_cast<const float4*>(x + k);
          a0 = fmaf(xv.x, wv.x, a0); a1 = fmaf(xv.y, wv.y, a1);
          a0 = fmaf(xv.z, wv.z, a0); a1 = fmaf(xv.w, wv.w, a1);
        }
        acc = a0 + a1;
      }
      acc += __shfl_xor_sync(FULL, acc, 16);
      if (half == 1 && lane < 16) xch[rl * 32 + lane] = acc;
      __syncthreads();
      if (half == 0 && ract && lane < dn) p.dh_z[(size_t)b * D + d0 + lane] = acc + xch[rl * 32 + lane];
    }
    grid_barrier(p.bar, G, gen);

    pb_stamp(p, t, 3);
    // ---------------- phase 3a: GRU gates + LN_3D backward (CTAs with cb == 0) ---------------
    if (cb == 0) {
      float dp[3 * (DV / 2)], xh[3 * (DV / 2)];
      float mean = 0.f, rstd = 0.f;
      const int jb = half * (D >> 1);
      if (ract) {
        const float* row = p.g_pre + bt * D3;
        float v[3 * DV];
#pragma unroll
        for (int i = 0; i < 3 * DV; ++i) v[i] = row[lane + 32 * i];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 3 * DV; ++i) s += v[i];
        mean = warp_sum(s) / (float)D3;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 3 * DV; ++i) {
          const float dd = v[i] - mean;
          q = fmaf(dd, dd, q);
        }
        rstd = 1.f / sqrtf(warp_sum(q) / (float)D3 + p.eps);
        // every input of the gate backward up front (stores below would serialise later loads)
        float hpv[DV / 2], dv[DV / 2];
#pragma unroll
        for (int i = 0; i < DV / 2; ++i) {
          const int j = jb + lane + 32 * i;
          hpv[i] = p.hprev[bt * D + j];
          dv[i] = __ldcg(p.dh_z + (size_t)b * D + j) + p.dh_prior[bt * D + j] +
                  __ldcg(p.dh_rec + (size_t)b * D + j);
          if (p.g_deter) dv[i] += p.g_deter[bt * D + j];
        }
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < DV / 2; ++i) {
          const float vr = half ? v[DV / 2 + i] : v[i];
          const float vc = half ? v[DV + DV / 2 + i] : v[DV + i];
          const float vu = half ? v[2 * DV + DV / 2 + i] : v[2 * DV + i];
          const int j = jb + lane + 32 * i;
          const float xr = (vr - mean) * rstd, xc = (vc - mean) * rstd, xu = (vu - mean) * rstd;
          const float pr = fmaf(xr, lgg[j], lgb[j]);
          const float pc = fmaf(xc, lgg[D + j], lgb[D + j]);
          const float pu = fmaf(xu, lgg[2 * D + j], lgb[2 * D + j]);
          const float rg = sigmoidf_(pr);
          const float c = tanhf(rg * pc);
          const float u = sigmoidf_(pu - 1.f);
          const float hp = hpv[i];
          const float d = dv[i];
          const float du = d * (c - hp);
          const float dc = d * u;
          const float drc = dc * (1.f - c * c);
          const float dpr = drc * pc * rg * (1.f - rg);
          const float dpc = drc * rg;
          const float dpu = du * u * (1.f - u);
          p.dhdir[(size_t)b * D + j] = d * (1.f - u);
          dp[i] = dpr; dp[DV / 2 + i] = dpc; dp[DV + i] = dpu;
          xh[i] = xr;  xh[DV / 2 + i] = xc;  xh[DV + i] = xu;
          const float gx0 = dpr * lgg[j], gx1 = dpc * lgg[D + j], gx2 = dpu * lgg[2 * D + j];
          a0 += gx0 + gx1 + gx2;
          a1 = fmaf(gx0, xr, a1); a1 = fmaf(gx1, xc, a1); a1 = fmaf(gx2, xu, a1);
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        if (lane == 0) { pr2[(rl * 2 + half) * 2] = a0; pr2[(rl * 2 + half) * 2 + 1] = a1; }
      }
      __syncthreads();
      if (ract) {
        const float m1 = (pr2[(rl * 2) * 2] + pr2[(rl * 2 + 1) * 2]) / (float)D3;
        const float m2 = (pr2[(rl * 2) * 2 + 1] + pr2[(rl * 2 + 1) * 2 + 1]) / (float)D3;
#pragma unroll
        for (int i = 0; i < DV / 2; ++i) {
          const int j = jb + lane + 32 * i;
#pragma unroll
          for (int q3 = 0; q3 < 3; ++q3) {
            const int col = q3 * D + j;
            const float dpv = dp[q3 * (DV / 2) + i], xv = xh[q3 * (DV / 2) + i];
            p.d_g_ln[bt * D3 + col] = dpv;
            p.d_g_pre[bt * D3 + col] = rstd * (dpv * lgg[col] - m1 - xv * m2);
          }
        }
      }
    }
    grid_barrier(p.bar, G, gen);

    pb_stamp(p, t, 4);
    // ---------------- phase 3b: [dx | dh_prev] = d_g_pre W_gru (+ direct path), own columns ---
    if (xn > 0) {
      float* dxh = p.dxh;
      const float* dhd = p.dhdir;
      gemv16<true, 4, 1>(Wx, xn, D3, p.d_g_pre + (size_t)t * D3, T * D3, D3, nullptr, 0, B, part,
                   [&](int m, int c, float r) {
                     const int n = x0 + c;
                     if (n >= Hd) r += __ldcg(dhd + (size_t)m * D + (n - Hd));
                     dxh[(size_t)m * HD + n] = r;
                   });
    }
    grid_barrier(p.bar, G, gen);

    pb_stamp(p, t, 5);
    // ---------------- phase 4: LN+SiLU backward of x rows; d stoch_prev of group cb; routing --
    {
      const bool first = ract ? (p.first_eff[bt] != 0.f) : false;
      {
        constexpr int HV2 = (HV + 1) / 2;
        float pre[HV2], dout[HV2];
        const int nh = Hd >> 1;
#pragma unroll
        for (int i = 0; i < HV2; ++i) {
          const int j = half * nh + lane + 32 * i;
          const bool on = ract && lane + 32 * i < nh;
          pre[i] = on ? p.x_pre[bt * Hd + j] : 0.f;
          dout[i] = on ? __ldcg(p.dxh + (size_t)b * HD + j) : 0.f;
        }
        float* rrow = rows + (size_t)rl * Hd;
        const bool wr = cb == 0;
        pair_ln_silu_bwd<HV2>(ract, pre, dout, Hd, lane, half, lxg, lxb, p.eps, pr2 + rl * 8,
                              [&](int j, float dp_, float dl_) {
                                rrow[j] = dp_;
                                if (wr) { p.d_x_pre[bt * Hd + j] = dp_; p.d_x_ln[bt * Hd + j] = dl_; }
                              });
      }
      if (ract) {
        // reset routing of the deter gradient (this block's rows, once: CTAs with cb == 0)
        if (cb == 0) {
          float dv[DV / 2], di[DV / 2];
#pragma unroll
          for (int i = 0; i < DV / 2; ++i) {
            const int j = half * (D >> 1) + lane + 32 * i;
            dv[i] = __ldcg(p.dxh + (size_t)b * HD + Hd + j);
            di[i] = first ? p.dinit_h[(size_t)b * D + j] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < DV / 2; ++i) {
            const int j = half * (D >> 1) + lane + 32 * i;
            p.dh_rec[(size_t)b * D + j] = first ? 0.f : dv[i];
            if (first) p.dinit_h[(size_t)b * D + j] = di[i] + dv[i];
          }
        }
      }
      __syncthreads();
      float a0 = 0.f, a1 = 0.f;
      if (ract) {
        const int k0 = half * (Hd >> 1);
        const float* w = Wi + (size_t)min(lane, C - 1) * (Hd + 4) + k0;
        const float* x = rows + (size_t)rl * Hd + k0;
        for (int k = 0; k < (Hd >> 1); k += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(w + k);
          const float4 xv = *reinterpret_cast<const float4*>(x + k);
          a0 = fmaf(xv.x, wv.x, a0); a1 = fmaf(xv.y, wv.y, a1);
          a0 = fmaf(xv.z, wv.z, a0); a1 = fmaf(xv.w, wv.w, a1);
        }
        if (half == 1) xch[rl * 32 + lane] = a0 + a1;
      }
      __syncthreads();
      if (ract && half == 0 && lane < C) {
        const float ds = (a0 + a1) + xch[rl * 32 + lane];
        dsl[rl * 32 + lane] = first ? 0.f : ds;
        if (first) p.dinit_s[(size_t)b * SC + (size_t)cb * C + lane] += ds;
      }
      __syncwarp();
    }
    pb_stamp(p, t, 6);
  }
  // the gradient that reaches the caller's initial state
  if (ract && half == 0 && lane < C)
    p.ds_rec[(size_t)b * SC + (size_t)cb * C + lane] = dsl[rl * 32 + lane];
}

static size_t pb_smem_bytes(const PbArgs& a) {
  const size_t SC = (size_t)a.S * a.C;
  const size_t fl = (size_t)a.ncz * SC + (size_t)a.cpd * (a.Hd + 4) + (size_t)a.ncx * 3 * a.D +
                    (size_t)a.C * (a.Hd + 4) + 4 * (size_t)a.Hd + 6 * (size_t)a.D +
                    4 * (size_t)a.Hd + PO_WARPS * PO_NV + 4 * 32 + 4 * 32 + 32 + 16;
  return fl * 4;
}

template <int DV, int HV>
static int pb_launch(PbArgs& a, int G, size_t smem, cudaStream_t st, bool* used) {
  auto kern = observe_persistent_bwd_kernel<DV, HV>;
  DV3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  DV3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PO_THREADS, smem));
  int dev = 0, sms = 0;
  DV3_CHECK_CUDA(cudaGetDevice(&dev));
  DV3_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (per_sm * sms < G) return 0;   // cannot be co-resident: the caller runs the stepwise path
  DV3_CHECK_CUDA(cudaMemsetAsync(a.bar, 0, 2 * sizeof(unsigned), st));
  void* args[] = {&a};
  DV3_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(G), dim3(PO_THREADS),
                                             args, smem, st));
  note_launch();
  *used = true;
  return 0;
}

// Runs the reverse-time recurrence of observe_bwd in one persistent kernel when the shapes allow
// it; *used == false on return means "not applicable": the caller runs the stepwise launches.
// dh_rec / dinit_s / dinit_h must be zeroed by the caller; WosT, WobsT, WgruT, WinT are the
// transposed weights the stepwise path uses.
int observe_bwd_persistent(const dv3_rssm_dims* d, const dv3_rssm_params* p,
                           const dv3_observe_bwd_io* io, const ObsBwdShared& w, cudaStream_t st,
                           bool* used) {
  *used = false;
  const char* env = DV3_ENV("DV3_OBSERVE_STEPWISE");
  if (env && env[0] == '1') return 0;
  const int D = d->deter, Hd = d->hidden, S = d->stoch, C = d->classes, B = io->B;
  if (B > PO_ROWS || D % 32 != 0 || (D / 32 != 2 && D / 32 != 4 && D / 32 != 8 && D / 32 != 16))
    return 0;
  if (Hd % 16 != 0 || Hd > 1024 || C > 32 || (S * C) % 4 != 0) return 0;
  int dev = 0, sms = 0, coop = 0;
  DV3_CHECK_CUDA(cudaGetDevice(&dev));
  DV3_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DV3_CHECK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return 0;
  PbArgs a{};
  a.nrb = (B + 3) / 4;
  const int G = S * a.nrb;
  if (G > sms) return 0;
  a.B = B; a.T = io->T; a.S = S; a.C = C; a.D = D; a.Hd = Hd;
  a.unimix = d->unimix; a.eps = d->ln_eps;
  a.first_eff = io->first_eff; a.post_logit = io->post_logit; a.hprev = io->hprev;
  a.x_pre = io->x_pre; a.g_pre = io->g_pre; a.z_pre = io->z_pre;
  a.g_post_stoch = io->g_post_stoch; a.g_post_logit = io->g_post_logit; a.g_deter = io->g_deter;
  a.dh_prior = w.dh_prior;
  a.WosT = w.WosT; a.WobsT = w.WobsT; a.WgruT = w.WgruT; a.WinT = w.WinT;
  a.ln_obs_g = p->ln_obs_g; a.ln_obs_b = p->ln_obs_b; a.ln_gru_g = p->ln_gru_g;
  a.ln_gru_b = p->ln_gru_b; a.ln_in_g = p->ln_in_g; a.ln_in_b = p->ln_in_b;
  a.d_post_logit = io->d_post_logit; a.d_z_pre = io->d_z_pre; a.d_z_ln = io->d_z_ln;
  a.d_g_pre = io->d_g_pre; a.d_g_ln = io->d_g_ln; a.d_x_pre = io->d_x_pre; a.d_x_ln = io->d_x_ln;
  a.ds_rec = w.ds_rec; a.dh_rec = w.dh_rec; a.dinit_s = w.dinit_s; a.dinit_h = w.dinit_h;
  a.d_z = w.d_z; a.dh_z = w.dh_z; a.dhdir = w.dhdir; a.dxh = w.dxh;
  a.bar = w.bar;
  a.timing = nullptr;
  if (const char* te = DV3_ENV("DV3_OBSERVE_TIMING"))
    if (te[0] == '2' && io->T <= 4096) a.timing = po_timing_buffer();
  a.ncz = (Hd + G - 1) / G;
  a.cpd = (D + S - 1) / S;
  a.ncx = (Hd + D + G - 1) / G;
  if (a.cpd > 16) return 0;
  const size_t smem = pb_smem_bytes(a);
  if (smem > 220 * 1024) return 0;
  const int hv = (Hd + 31) / 32;
#define DV3_PB_DISPATCH(DVV)                                                   \
  do {                                                                         \
    if (hv <= 4) return pb_launch<DVV, 4>(a, G, smem, st, used);               \
    if (hv <= 16) return pb_launch<DVV, 16>(a, G, smem, st, used);             \
    return pb_launch<DVV, 32>(a, G, smem, st, used);                           \
  } while (0)
  switch (D / 32) {
    case 2: DV3_PB_DISPATCH(2);
    case 4: DV3_PB_DISPATCH(4);
    case 8: DV3_PB_DISPATCH(8);
    default: DV3_PB_DISPATCH(16);
  }
#undef DV3_PB_DISPATCH
}

}  // namespace dv3
