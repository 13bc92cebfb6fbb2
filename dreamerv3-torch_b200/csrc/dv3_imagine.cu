// ImagBehavior._imagine: H-step actor-in-the-loop prior rollout from N start states, and its
// backward ('dynamics' gradient: BPTT through img_step w.r.t. activations).
//
// Per step k:  feat_k = [one-hot stoch_k | deter_k] (detached for the actor, models.py:514)
//              a_k    = actor(feat_k).sample()       (networks.py:657-700, tools.py:594-598)
//              state_{k+1} = img_step(state_k, a_k)  (networks.py:208-233)
// The successor of the last step is never returned by the reference (models.py:546), so its
// img_step is not computed here.
// One-hot inputs: both Linear layers that consume `stoch` (W_in and the first actor layer) are
// evaluated as a gather-sum of S rows of the transposed weight plus a dense product over the
// remaining (action / deter) columns -- 2*S*C*width flops per row become S*width adds.
#include "dv3_common.cuh"

namespace dv3 {

__global__ void actor_normal_sample_kernel(const float* __restrict__ mean_raw,
                                           const float* __restrict__ std_raw,
                                           const float* __restrict__ eps, float min_std,
                                           float max_std, long long total,
                                           float* __restrict__ action) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float mean = tanhf(mean_raw[i]);
  const float std = (max_std - min_std) * sigmoidf_(std_raw[i] + 2.f) + min_std;
  const float out = mean + std * eps[i];
  action[i] = out * (1.f / fmaxf(fabsf(out), 1.f));
}

// g1 + g2 = gradient reaching the returned action; the clip factor is a constant (detached)
__global__ void actor_normal_sample_bwd_kernel(const float* __restrict__ mean_raw,
                                               const float* __restrict__ std_raw,
                                               const float* __restrict__ eps,
                                               const float* __restrict__ g1, int ld1, int c1,
                                               const float* __restrict__ g2, float min_std,
                                               float max_std, int N, int A,
                                               float* __restrict__ d_mean_raw,
                                               float* __restrict__ d_std_raw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * A) return;
  const int r = (int)(i / A), a = (int)(i % A);
  float g = 0.f;
  if (g1) g += g1[(size_t)r * ld1 + c1 + a];
  if (g2) g += g2[i];
  const float mean = tanhf(mean_raw[i]);
  const float sg = sigmoidf_(std_raw[i] + 2.f);
  const float std = (max_std - min_std) * sg + min_std;
  const float out = mean + std * eps[i];
  const float dout = g * (1.f / fmaxf(fabsf(out), 1.f));
  d_mean_raw[i] = dout * (1.f - mean * mean);
  d_std_raw[i] = dout * eps[i] * (max_std - min_std) * sg * (1.f - sg);
}

// Actor output heads fused with the draw: one warp per row computes the A (normal: 2A) dot
// products of the top activation with the head weights held in shared memory, then
//   normal: a = clip(tanh(mean) + std * eps)      (networks.py:693-700, tools.py:575-601)
//   onehot: only the logits are written; the unimix categorical draw follows in onehot_sample.
// Replaces two [N x U] x [U x A] SIMT GEMM launches (8 CTAs each) plus the sample kernel.
constexpr int AH_THREADS = 256;

__global__ void __launch_bounds__(AH_THREADS)
actor_head_kernel(const float* __restrict__ top, int ldt, const float* __restrict__ w_mean,
                  const float* __restrict__ b_mean, const float* __restrict__ w_std,
                  const float* __restrict__ b_std, const float* __restrict__ eps, float min_std,
                  float max_std, int N, int U, int A, float* __restrict__ mean_raw,
                  float* __restrict__ std_raw, float* __restrict__ action) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float ah_smem[];
  const int nh = w_std ? 2 : 1;
  float* ws = ah_smem;                         // [nh*A][U]
  for (int i = threadIdx.x * 4; i < A * U; i += AH_THREADS * 4) {
    *reinterpret_cast<float4*>(ws + i) = __ldg(reinterpret_cast<const float4*>(w_mean + i));
    if (w_std)
      *reinterpret_cast<float4*>(ws + A * U + i) = __ldg(reinterpret_cast<const float4*>(w_std + i));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = AH_THREADS / 32;
  for (int r = blockIdx.x * nw + warp; r < N; r += gridDim.x * nw) {
    const float* x = top + (size_t)r * ldt;
    float mine = 0.f;                          // lane j < nh*A ends up holding head value j
    for (int j = 0; j < nh * A; ++j) {
      const float* w = ws + (size_t)j * U;
      float acc = 0.f;
      for (int k = lane * 4; k < U; k += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(x + k);
        const float4 wv = *reinterpret_cast<const float4*>(w + k);
        acc = fmaf(xv.x, wv.x, acc); acc = fmaf(xv.y, wv.y, acc);
        acc = fmaf(xv.z, wv.z, acc); acc = fmaf(xv.w, wv.w, acc);
      }
      acc = warp_sum(acc);
      if (lane == j) mine = acc;
    }
    const size_t o = (size_t)r * A;
    if (lane < A) {
      mine += b_mean[lane];
      mean_raw[o + lane] = mine;
    } else if (lane < nh * A) {
      mine += b_std[lane - A];
      std_raw[o + lane - A] = mine;
    }
    if (w_std) {
      const float sraw = __shfl_sync(FULL, mine, (lane + A) & 31);
      if (lane < A) {
        const float mean = tanhf(mine);
        const float std = (max_std - min_std) * sigmoidf_(sraw + 2.f) + min_std;
        const float out = mean + std * eps[o + lane];
        action[o + lane] = out * (1.f / fmaxf(fabsf(out), 1.f));
      }
    }
  }
}

struct ImgFwdWs {
  float* WinT;   // [(SC+A), Hd]
  float* Wa0T;   // [F, U]
  float* a_add;  // [N, U]
  // tensor-core path: tf32 hi/lo planes of every GEMM input, written by the kernel that produces
  // the activation (no split passes inside the time loop)
  SplitOut dsp[2];   // deter of state k / k+1   [N, D]
  SplitOut xsp;      // x                        [N, Hd]
  SplitOut ysp;      // y                        [N, Hd]
  SplitOut asp[2];   // actor activations        [N, U]
  LinW gru, out, ims, a0d, al[16];
  void* pi_sync;     // counters + LayerNorm statistics of the persistent kernel
};

static SplitOut take_split(Arena& a, bool on, size_t rows, int cols) {
  SplitOut s;
  if (on) {
    s.hi = a.take<float>(rows * cols);
    s.lo = a.take<float>(rows * cols);
    s.ld = cols;
  }
  return s;
}

static void carve_img_fwd(Arena& a, const dv3_rssm_dims* d, const dv3_actor* act, int N,
                          ImgFwdWs& w) {
  const int SC = d->stoch * d->classes, D = d->deter, Hd = d->hidden;
  const bool tc = N >= TC_MIN_ROWS;
  w.WinT = a.take<float>((size_t)(SC + d->actions) * Hd);
  w.gru.reserve(a, tc, 3 * D, Hd + D);
  w.out.reserve(a, tc, Hd, D);
  w.ims.reserve(a, tc, SC, Hd);
  w.dsp[0] = take_split(a, tc, N, D);
  w.dsp[1] = take_split(a, tc, N, D);
  w.xsp = take_split(a, tc, N, Hd);
  w.ysp = take_split(a, tc, N, Hd);
  if (act) {
    const int U = act->units;
    w.Wa0T = a.take<float>((size_t)(SC + D) * U);
    w.a_add = a.take<float>((size_t)N * U);
    w.a0d.reserve(a, tc, U, D);
    for (int i = 1; i < act->layers; ++i) w.al[i].reserve(a, tc, U, U);
    w.asp[0] = take_split(a, tc, N, U);
    w.asp[1] = take_split(a, tc, N, U);
  } else {
    w.Wa0T = w.a_add = nullptr;
  }
  w.pi_sync = a.take<char>(imagine_persistent_sync_bytes(N));
}

static int check_actor(const dv3_rssm_dims* d, const dv3_actor* a, const char* who) {
  if (!a) return 0;
  DV3_REQUIRE(a->layers >= 1 && a->layers <= 16 && a->units >= 4 && a->units % 4 == 0,
              DV3_ERR_BAD_SHAPE, "%s: actor layers=%d units=%d", who, a->layers, a->units);
  DV3_REQUIRE(a->dist == 0 || a->dist == 1, DV3_ERR_BAD_SHAPE, "%s: actor dist=%d", who, a->dist);
  DV3_REQUIRE(a->dist == 0 || d->actions <= 32, DV3_ERR_BAD_SHAPE,
              "%s: onehot actor with %d > 32 actions", who, d->actions);
  DV3_REQUIRE(a->w && a->ln_g && a->ln_b && a->w_mean && a->b_mean, DV3_ERR_NULL,
              "%s: null actor parameter", who);
  DV3_REQUIRE(a->dist == 1 || (a->w_std && a->b_std), DV3_ERR_NULL, "%s: null actor std layer",
              who);
  return 0;
}

static int check_rssm_dims(const dv3_rssm_dims* d, const char* who) {
  DV3_REQUIRE(d, DV3_ERR_NULL, "%s: dims is NULL", who);
  DV3_REQUIRE(d->stoch >= 1 && d->classes >= 1 && d->classes <= 32 && d->deter % 4 == 0 &&
                  d->hidden % 4 == 0 && d->deter >= 4 && d->hidden >= 4 && d->actions >= 1,
              DV3_ERR_BAD_SHAPE, "%s: unsupported dims S=%d C=%d D=%d Hd=%d A=%d", who, d->stoch,
              d->classes, d->deter, d->hidden, d->actions);
  return 0;
}

}  // namespace dv3

using namespace dv3;

extern "C" size_t dv3_imagine_workspace_bytes(const dv3_rssm_dims* d, const dv3_actor* a, int32_t N,
                                              int32_t H) {
  if (!d || N <= 0 || H <= 0) return 0;
  Arena ar(nullptr, 0);
  ImgFwdWs w;
  carve_img_fwd(ar, d, a, N, w);
  return ar.used;
}

extern "C" int dv3_imagine_fwd(const dv3_rssm_dims* d, const dv3_rssm_params* p, const dv3_actor* a,
                               const dv3_imagine_io* io, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DV3_TRY(check_rssm_dims(d, "imagine_fwd"));
  DV3_TRY(check_actor(d, a, "imagine_fwd"));
  DV3_REQUIRE(p && io, DV3_ERR_NULL, "imagine_fwd: params/io is NULL");
  const int N = io->N, H = io->H;
  DV3_REQUIRE(N >= 0 && H >= 0, DV3_ERR_BAD_SHAPE, "imagine_fwd: N=%d H=%d", N, H);
  if (N == 0 || H == 0) return 0;
  const int S = d->stoch, C = d->classes, SC = S * C, D = d->deter, Hd = d->hidden, A = d->actions,
            F = SC + D;
  DV3_REQUIRE(io->start_idx && io->start_deter && io->u_state, DV3_ERR_NULL,
              "imagine_fwd: null input");
  DV3_REQUIRE(a ? io->act_noise != nullptr : io->given_action != nullptr, DV3_ERR_NULL,
              "imagine_fwd: need act_noise (actor) or given_action (no actor)");
  DV3_REQUIRE(io->feat && io->logit && io->action && io->idx && io->x_pre && io->x && io->g_pre &&
                  io->y_pre && io->y,
              DV3_ERR_NULL, "imagine_fwd: null output");
  DV3_REQUIRE(!a || (io->a_pre && io->a_act && io->a_mean_raw && (a->dist == 1 || io->a_std_raw)),
              DV3_ERR_NULL, "imagine_fwd: null actor activation buffer");
  Arena arena(io->workspace, io->workspace_bytes);
  ImgFwdWs w;
  carve_img_fwd(arena, d, a, N, w);
  DV3_REQUIRE(io->workspace && arena.ok(), DV3_ERR_WORKSPACE,
              "imagine_fwd: workspace %zu < %zu bytes", io->workspace_bytes, arena.used);
  const int U = a ? a->units : 0, L = a ? a->layers : 0;

  // weight forms: the caller's (made once per optimizer step, dv3_rssm_planes / dv3_actor_planes)
  // where supplied, else derived here per call
  const dv3_rssm_planes* rp = p->planes;
  const dv3_actor_planes* ap = a ? a->planes : nullptr;
  const float* WinT = w.WinT;
  if (rp && rp->w_in_t) WinT = rp->w_in_t;
  else DV3_TRY(launch_transpose(p->w_in, SC + A, Hd, SC + A, w.WinT, st));
  if (!w.gru.adopt(rp ? &rp->w_gru : nullptr, 0, false)) DV3_TRY(w.gru.prepare(p->w_gru, Hd + D, st));
  else w.gru.W = p->w_gru, w.gru.ldw = Hd + D;
  if (!w.out.adopt(rp ? &rp->w_out : nullptr, 0, false)) DV3_TRY(w.out.prepare(p->w_out, D, st));
  else w.out.W = p->w_out, w.out.ldw = D;
  if (!w.ims.adopt(rp ? &rp->w_ims : nullptr, 0, false)) DV3_TRY(w.ims.prepare(p->w_ims, Hd, st));
  else w.ims.W = p->w_ims, w.ims.ldw = Hd;
  const float* Wa0T = w.Wa0T;
  if (a) {
    if (ap && ap->w0_t) Wa0T = ap->w0_t;
    else DV3_TRY(launch_transpose(a->w[0], F, U, F, w.Wa0T, st));
    if (!w.a0d.adopt(ap ? &ap->w[0] : nullptr, SC, false)) DV3_TRY(w.a0d.prepare(a->w[0] + SC, F, st));
    else w.a0d.W = a->w[0] + SC, w.a0d.ldw = F;
    for (int i = 1; i < L; ++i) {
      if (!w.al[i].adopt(ap ? &ap->w[i] : nullptr, 0, false)) DV3_TRY(w.al[i].prepare(a->w[i], U, st));
      else w.al[i].W = a->w[i], w.al[i].ldw = U;
    }
  }
  // state 0
  DV3_TRY(copy_rows_i32(io->start_idx, S, N, S, io->idx, S, st));
  DV3_TRY(idx_to_onehot(io->start_idx, S, N, S, C, io->feat, F, st));
  DV3_TRY(copy_rows(io->start_deter, D, N, D, io->feat + SC, F, st));
  const bool tc = w.gru.tc;
  if (tc) DV3_TRY(tc_split(io->start_deter, D, D, nullptr, 0, 0, N, w.dsp[0].hi, w.dsp[0].lo, st));
  if (a && tc) {
    // the whole rollout as one persistent kernel when the shapes allow it (dv3_imagine_persistent.cu)
    auto plane = [](const LinW& W) { return PiPlane{W.hi, W.lo, W.ldp ? W.ldp : W.K}; };
    PiPlanes pl{};
    pl.ok = w.gru.tc && w.out.tc && w.ims.tc && w.a0d.tc && !w.gru.mn && !w.out.mn && !w.ims.mn &&
            !w.a0d.mn && L <= 3;
    for (int i = 1; i < L && i < 3; ++i) pl.ok = pl.ok && w.al[i].tc && !w.al[i].mn;
    if (pl.ok) {
      for (int b = 0; b < 2; ++b) {
        pl.dsp_hi[b] = w.dsp[b].hi; pl.dsp_lo[b] = w.dsp[b].lo;
        pl.asp_hi[b] = w.asp[b].hi; pl.asp_lo[b] = w.asp[b].lo;
      }
      pl.xsp_hi = w.xsp.hi; pl.xsp_lo = w.xsp.lo; pl.ysp_hi = w.ysp.hi; pl.ysp_lo = w.ysp.lo;
      pl.wa[0] = plane(w.a0d);
      for (int i = 1; i < L; ++i) pl.wa[i] = plane(w.al[i]);
      pl.gru = plane(w.gru); pl.out = plane(w.out); pl.ims = plane(w.ims);
      pl.Wa0T = Wa0T; pl.WinT = WinT;
      bool used = false;
      DV3_TRY(imagine_fwd_persistent(d, p, a, io, pl, w.pi_sync, st, &used));
      if (used) return 0;
    }
  }
  // C = [A1|A2] W^T: from the producers' hi/lo planes on the tensor-core path, from the fp32
  // activations on the CUDA-core path (N < 64 rows)
  const char* e_rawa = DV3_ENV("DV3_TC_RAWA");
  const bool rawa = !(e_rawa && e_rawa[0] == '0');
  auto lin = [&](const LinW& W, const float* A1, int lda1, int K1, const SplitOut& s1,
                 const float* A2, int lda2, int K2, const SplitOut* s2, const float* bias, float* Cc,
                 int ldc) -> int {
    // fp32 A split in the SM into tensor memory where that kernel covers the shape (bit-identical
    // to the pre-split product with the same tile; DV3_TC_RAWA=0 keeps the planes)
    if (rawa && W.rawa_ok(A1, lda1, K1, A2, lda2, K2, N))
      return W.apply_rawa(A1, lda1, K1, A2, lda2, K2, bias, nullptr, 0, Cc, ldc, N, st);
    if (W.tc) return W.apply_split(s1, K1, A2 ? s2 : nullptr, K2, bias, nullptr, 0, Cc, ldc, N, st);
    return W.apply(A1, lda1, K1, A2, lda2, K2, bias, nullptr, 0, Cc, ldc, N, nullptr, st);
  };

  for (int k = 0; k < H; ++k) {
    float* featk = io->feat + (size_t)k * N * F;
    const int32_t* idxk = io->idx + (size_t)k * N * S;
    float* actk = io->action + (size_t)k * N * A;
    const SplitOut& dcur = w.dsp[k & 1];
    const SplitOut& dnxt = w.dsp[(k + 1) & 1];
    if (a) {
      const size_t lstride = (size_t)H * N * U;
      float* pre0 = io->a_pre + (size_t)k * N * U;
      float* act0 = io->a_act + (size_t)k * N * U;
      // layer 0: deter columns dense, stoch columns gathered
      DV3_TRY(lin(w.a0d, featk + SC, F, D, dcur, nullptr, 0, 0, nullptr, nullptr, w.a_add, U));
      DV3_TRY(gather_ln_silu(idxk, S, S, C, nullptr, 0, 0, Wa0T, w.a_add, U, a->ln_g[0],
                             a->ln_b[0], d->ln_eps, N, U, pre0, U, act0, U, st, w.asp[0]));
      for (int i = 1; i < L; ++i) {
        float* prei = io->a_pre + i * lstride + (size_t)k * N * U;
        float* acti = io->a_act + i * lstride + (size_t)k * N * U;
        const float* prev = io->a_act + (i - 1) * lstride + (size_t)k * N * U;
        DV3_TRY(lin(w.al[i], prev, U, U, w.asp[(i - 1) & 1], nullptr, 0, 0, nullptr, nullptr, prei, U));
        DV3_TRY(ln_silu_fwd(prei, U, a->ln_g[i], a->ln_b[i], d->ln_eps, N, U, acti, U, st,
                            i + 1 < L ? w.asp[i & 1] : SplitOut()));
      }
      const float* top = io->a_act + (L - 1) * lstride + (size_t)k * N * U;
      float* mraw = io->a_mean_raw + (size_t)k * N * A;
      float* sraw = a->dist == 0 ? io->a_std_raw + (size_t)k * N * A : nullptr;
      const int nh = a->dist == 0 ? 2 : 1;
      const size_t ah_smem = (size_t)nh * A * U * sizeof(float);
      if (nh * A <= 32 && ah_smem <= 160 * 1024) {
        static DeviceOnce attr;
        if (attr.need())
          DV3_CHECK_CUDA(cudaFuncSetAttribute(actor_head_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              160 * 1024));
        const int nw = AH_THREADS / 32;
        int grid = (N + nw - 1) / nw;
        if (grid > 148) grid = 148;
        DV3_CHECK_CUDA(launch_pdl(actor_head_kernel, dim3(grid), dim3(AH_THREADS), ah_smem, st, top, U,
                                  a->w_mean, a->b_mean, a->dist == 0 ? a->w_std : nullptr, a->b_std,
                                  io->act_noise + (size_t)k * N * A, a->min_std, a->max_std, N, U, A,
                                  mraw, sraw, actk));
        DV3_CHECK_LAUNCH("actor_head_kernel");
      } else {
        DV3_TRY(linear1(top, U, a->w_mean, U, U, a->b_mean, mraw, A, N, A, 0, st));
        if (a->dist == 0) {
          DV3_TRY(linear1(top, U, a->w_std, U, U, a->b_std, sraw, A, N, A, 0, st));
          const long long tot = (long long)N * A;
          actor_normal_sample_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(
              mraw, sraw, io->act_noise + (size_t)k * N * A, a->min_std, a->max_std, tot, actk);
          DV3_CHECK_LAUNCH("actor_normal_sample_kernel");
        }
      }
      if (a->dist != 0)
        DV3_TRY(onehot_sample(mraw, A, io->act_noise + (size_t)k * N * A, A, 0, 0, a->unimix, N, 1,
                              A, nullptr, 0, actk, A, st));
    } else if (k < H - 1) {
      DV3_TRY(copy_rows(io->given_action + (size_t)k * N * A, A, N, A, actk, A, st));
    } else {
      DV3_TRY(fill_zero(actk, (size_t)N * A * 4, st));
    }
    if (k == H - 1) break;

    float* featn = io->feat + (size_t)(k + 1) * N * F;
    float* xpre = io->x_pre + (size_t)k * N * Hd;
    float* xk = io->x + (size_t)k * N * Hd;
    float* gpre = io->g_pre + (size_t)k * N * 3 * D;
    float* ypre = io->y_pre + (size_t)k * N * Hd;
    float* yk = io->y + (size_t)k * N * Hd;
    float* logn = io->logit + (size_t)(k + 1) * N * SC;
    DV3_TRY(gather_ln_silu(idxk, S, S, C, actk, A, A, WinT, nullptr, 0, p->ln_in_g, p->ln_in_b,
                           d->ln_eps, N, Hd, xpre, Hd, xk, Hd, st, w.xsp));
    DV3_TRY(lin(w.gru, xk, Hd, Hd, w.xsp, featk + SC, F, D, &dcur, nullptr, gpre, 3 * D));
    DV3_TRY(gru_gates_fwd(gpre, 3 * D, p->ln_gru_g, p->ln_gru_b, d->ln_eps, featk + SC, F, N, D,
                          featn + SC, F, st, dnxt));
    DV3_TRY(lin(w.out, featn + SC, F, D, dnxt, nullptr, 0, 0, nullptr, nullptr, ypre, Hd));
    DV3_TRY(ln_silu_fwd(ypre, Hd, p->ln_out_g, p->ln_out_b, d->ln_eps, N, Hd, yk, Hd, st, w.ysp));
    DV3_TRY(lin(w.ims, yk, Hd, Hd, w.ysp, nullptr, 0, 0, nullptr, p->b_ims, logn, SC));
    DV3_TRY(onehot_sample(logn, SC, io->u_state + (size_t)k * N * SC, SC, 0, 0, d->unimix, N, S, C,
                          io->idx + (size_t)(k + 1) * N * S, S, featn, F, st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
namespace dv3 {

struct ImgBwdWs {
  float *WimsT, *WoutT, *WgruT, *WinT;
  float *d_y, *dh_y, *dxh, *dxh_add, *dsa;
  SplitOut dlsp, dysp, dgsp, dxsp;   // hi/lo planes of d_logit, d_y_pre, d_g_pre, d_x_pre
  LinW ims, out, gru, in;   // over the transposed weights
};

static void carve_img_bwd(Arena& a, const dv3_rssm_dims* d, int N, ImgBwdWs& w) {
  const size_t SC = (size_t)d->stoch * d->classes, D = d->deter, Hd = d->hidden, A = d->actions;
  // W_in^T is padded to a multiple of 4 columns... the product d_x_pre @ W_in has N = SC+A
  w.WimsT = a.take<float>(Hd * SC);
  w.WoutT = a.take<float>(D * Hd);
  w.WgruT = a.take<float>((Hd + D) * 3 * D);
  w.WinT = a.take<float>((SC + A) * Hd);
  w.d_y = a.take<float>((size_t)N * Hd);
  w.dh_y = a.take<float>((size_t)N * D);
  w.dxh = a.take<float>((size_t)N * (Hd + D));
  w.dxh_add = a.take<float>((size_t)N * (Hd + D));
  w.dsa = a.take<float>((size_t)N * (SC + A));
  const bool tc = N >= TC_MIN_ROWS;
  w.ims.reserve(a, tc, (int)Hd, (int)SC);           // d_y   = d_logit @ W_ims      (K = SC)
  w.out.reserve(a, tc, (int)D, (int)Hd);            // dh_y  = d_y_pre @ W_out      (K = Hd)
  w.gru.reserve(a, tc, (int)(Hd + D), (int)(3 * D)); // dxh  = d_g_pre @ W_gru      (K = 3D)
  w.in.reserve(a, tc, (int)(SC + A), (int)Hd);      // dsa   = d_x_pre @ W_in       (K = Hd)
  w.dlsp = take_split(a, tc, N, (int)SC);
  w.dysp = take_split(a, tc, N, (int)Hd);
  w.dgsp = take_split(a, tc, N, (int)(3 * D));
  w.dxsp = take_split(a, tc, N, (int)Hd);
}

__global__ void add2_rows_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b,
                                 int ldb, int M, int n, float* __restrict__ out, int ldo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * n) return;
  const int r = (int)(i / n), c = (int)(i % n);
  float v = 0.f;
  if (a) v += a[(size_t)r * lda + c];
  if (b) v += b[(size_t)r * ldb + c];
  out[(size_t)r * ldo + c] = v;
}

}  // namespace dv3

extern "C" size_t dv3_imagine_bwd_workspace_bytes(const dv3_rssm_dims* d, const dv3_actor* a,
                                                  int32_t N, int32_t H) {
  (void)a;
  if (!d || N <= 0 || H <= 0) return 0;
  Arena ar(nullptr, 0);
  ImgBwdWs w;
  carve_img_bwd(ar, d, N, w);
  return ar.used;
}

extern "C" int dv3_imagine_bwd(const dv3_rssm_dims* d, const dv3_rssm_params* p, const dv3_actor* a,
                               const dv3_imagine_bwd_io* io, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DV3_TRY(check_rssm_dims(d, "imagine_bwd"));
  DV3_TRY(check_actor(d, a, "imagine_bwd"));
  DV3_REQUIRE(p && io, DV3_ERR_NULL, "imagine_bwd: params/io is NULL");
  const int N = io->N, H = io->H;
  DV3_REQUIRE(N >= 0 && H >= 0, DV3_ERR_BAD_SHAPE, "imagine_bwd: N=%d H=%d", N, H);
  if (N == 0 || H == 0) return 0;
  const int S = d->stoch, C = d->classes, SC = S * C, D = d->deter, Hd = d->hidden, A = d->actions,
            F = SC + D;
  DV3_REQUIRE(io->logit && io->feat && io->x_pre && io->g_pre && io->y_pre, DV3_ERR_NULL,
              "imagine_bwd: null saved tensor");
  DV3_REQUIRE(!a || (io->a_mean_raw && io->act_noise && io->d_mean_raw &&
                     (a->dist == 1 || (io->a_std_raw && io->d_std_raw))),
              DV3_ERR_NULL, "imagine_bwd: null actor tensor");
  DV3_REQUIRE(io->d_x_pre && io->d_x_ln && io->d_g_pre && io->d_g_ln && io->d_y_pre &&
                  io->d_y_ln && io->d_logit,
              DV3_ERR_NULL, "imagine_bwd: null output");
  Arena arena(io->workspace, io->workspace_bytes);
  ImgBwdWs w;
  carve_img_bwd(arena, d, N, w);
  DV3_REQUIRE(io->workspace && arena.ok(), DV3_ERR_WORKSPACE,
              "imagine_bwd: workspace %zu < %zu bytes", io->workspace_bytes, arena.used);

  // dx = dy W: with caller-supplied planes of W ([N_out, K_in] = [K, N] of this product) the weight
  // is read MN-major in place -- no transposed copy, no split; W_in^T comes K-major (w_in_t_sp)
  const dv3_rssm_planes* rp = p->planes;
  if (!w.ims.adopt(rp ? &rp->w_ims : nullptr, 0, true)) {
    DV3_TRY(launch_transpose(p->w_ims, Hd, SC, Hd, w.WimsT, st));
    DV3_TRY(w.ims.prepare(w.WimsT, SC, st));
  }
  if (!w.out.adopt(rp ? &rp->w_out : nullptr, 0, true)) {
    DV3_TRY(launch_transpose(p->w_out, D, Hd, D, w.WoutT, st));
    DV3_TRY(w.out.prepare(w.WoutT, Hd, st));
  }
  if (!w.gru.adopt(rp ? &rp->w_gru : nullptr, 0, true)) {
    DV3_TRY(launch_transpose(p->w_gru, Hd + D, 3 * D, Hd + D, w.WgruT, st));
    DV3_TRY(w.gru.prepare(w.WgruT, 3 * D, st));
  }
  if (!w.in.adopt(rp ? &rp->w_in_t_sp : nullptr, 0, false)) {
    DV3_TRY(launch_transpose(p->w_in, SC + A, Hd, SC + A, w.WinT, st));
    DV3_TRY(w.in.prepare(w.WinT, Hd, st));
  }
  DV3_TRY(fill_zero(w.dxh_add, (size_t)N * (Hd + D) * 4, st));

  // C = A W^T with A = a delta the producing kernel also wrote as hi/lo planes
  const char* e_rawa = DV3_ENV("DV3_TC_RAWA");
  const bool rawa = !(e_rawa && e_rawa[0] == '0');
  auto blin = [&](const LinW& W, const float* A1, int lda, int K, const SplitOut& sp,
                  const float* addend, int ldadd, float* Cc, int ldc) -> int {
    if (rawa && K <= 640 && W.rawa_ok(A1, lda, K, nullptr, 0, 0, N))     // long K: cluster split-K wins
      return W.apply_rawa(A1, lda, K, nullptr, 0, 0, nullptr, addend, ldadd, Cc, ldc, N, st);
    if (W.tc) return W.apply_split(sp, K, nullptr, 0, nullptr, addend, ldadd, Cc, ldc, N, st);
    return W.apply(A1, lda, K, nullptr, 0, 0, nullptr, addend, ldadd, Cc, ldc, N, nullptr, st);
  };

  auto actor_bwd = [&](int k, const float* d_a, int ld, int col) -> int {
    if (!a) return 0;
    const size_t o = (size_t)k * N * A;
    const float* gact = io->g_action ? io->g_action + o : nullptr;
    if (a->dist == 0) {
      const long long tot = (long long)N * A;
      actor_normal_sample_bwd_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(
          io->a_mean_raw + o, io->a_std_raw + o, io->act_noise + o, d_a, ld, col, gact, a->min_std,
          a->max_std, N, A, io->d_mean_raw + o, io->d_std_raw + o);
      DV3_CHECK_LAUNCH("actor_normal_sample_bwd_kernel");
      return 0;
    }
    return onehot_st_bwd(io->a_mean_raw + o, A, d_a ? d_a + col : nullptr, ld, gact, A, nullptr, 0,
                         a->unimix, N, 1, A, io->d_mean_raw + o, A, st);
  };

  // the last action feeds nothing downstream
  DV3_TRY(actor_bwd(H - 1, nullptr, 0, 0));

  // upstream state gradients: packed, or the column ranges of one [H,N,pitch] buffer
  DV3_REQUIRE(io->g_state_ld == 0 || (io->g_state_ld >= SC && io->g_state_ld % 4 == 0), DV3_ERR_BAD_SHAPE,
              "imagine_bwd: g_state_ld=%d", io->g_state_ld);
  const size_t lgs = io->g_state_ld ? (size_t)io->g_state_ld : (size_t)SC;
  const size_t lgd = io->g_state_ld ? (size_t)io->g_state_ld : (size_t)D;
  const float* ds_rec = nullptr;  // [N, SC+A] view into dsa once it is valid
  const float* dh_rec = nullptr;  // h columns of dxh
  for (int k = H - 2; k >= 0; --k) {
    const size_t on = (size_t)(k + 1) * N;  // row offset of state k+1
    const float* gs = io->g_stoch ? io->g_stoch + on * lgs : nullptr;
    const float* gl = io->g_logit ? io->g_logit + on * SC : nullptr;
    float* dlog = io->d_logit + on * SC;
    DV3_TRY(onehot_st_bwd(io->logit + on * SC, SC, gs, (int)lgs, ds_rec, SC + A, gl, SC, d->unimix, N, S,
                          C, dlog, SC, st, w.dlsp));
    DV3_TRY(blin(w.ims, dlog, SC, SC, w.dlsp, nullptr, 0, w.d_y, Hd));
    const size_t ok = (size_t)k * N;
    DV3_TRY(ln_silu_bwd(io->y_pre + ok * Hd, Hd, p->ln_out_g, p->ln_out_b, d->ln_eps, w.d_y, Hd, N,
                        Hd, io->d_y_pre + ok * Hd, Hd, io->d_y_ln + ok * Hd, Hd, st, w.dysp));
    DV3_TRY(blin(w.out, io->d_y_pre + ok * Hd, Hd, Hd, w.dysp, nullptr, 0, w.dh_y, D));
    const float* dh_in[4] = {w.dh_y, io->g_deter ? io->g_deter + on * lgd : nullptr, dh_rec, nullptr};
    const int ld_in[4] = {D, (int)lgd, Hd + D, 0};
    DV3_TRY(gru_gates_bwd(io->g_pre + ok * 3 * D, 3 * D, p->ln_gru_g, p->ln_gru_b, d->ln_eps,
                          io->feat + ok * F + SC, F, dh_in, ld_in, N, D, io->d_g_pre + ok * 3 * D,
                          3 * D, io->d_g_ln + ok * 3 * D, 3 * D, w.dxh_add + Hd, Hd + D, st, w.dgsp));
    DV3_TRY(blin(w.gru, io->d_g_pre + ok * 3 * D, 3 * D, 3 * D, w.dgsp, w.dxh_add, Hd + D, w.dxh,
                 Hd + D));
    dh_rec = w.dxh + Hd;
    DV3_TRY(ln_silu_bwd(io->x_pre + ok * Hd, Hd, p->ln_in_g, p->ln_in_b, d->ln_eps, w.dxh, Hd + D,
                        N, Hd, io->d_x_pre + ok * Hd, Hd, io->d_x_ln + ok * Hd, Hd, st, w.dxsp));
    // [d stoch_k | d action_k] = d_x_pre @ W_in
    DV3_TRY(blin(w.in, io->d_x_pre + ok * Hd, Hd, Hd, w.dxsp, nullptr, 0, w.dsa, SC + A));
    ds_rec = w.dsa;
    DV3_TRY(actor_bwd(k, w.dsa, SC + A, SC));
  }
  if (io->d_start_stoch) {
    add2_rows_kernel<<<(int)(((long long)N * SC + 255) / 256), 256, 0, st>>>(
        io->g_stoch, (int)lgs, ds_rec, SC + A, N, SC, io->d_start_stoch, SC);
    DV3_CHECK_LAUNCH("add2_rows_kernel");
  }
  if (io->d_start_deter) {
    add2_rows_kernel<<<(int)(((long long)N * D + 255) / 256), 256, 0, st>>>(
        io->g_deter, (int)lgd, dh_rec, Hd + D, N, D, io->d_start_deter, D);
    DV3_CHECK_LAUNCH("add2_rows_kernel");
  }
  return 0;
}
