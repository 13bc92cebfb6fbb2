// RSSM.observe: T-step posterior rollout and its BPTT, as a sequence of kernels enqueued on the
// caller's stream (no host sync, CUDA-graph capturable).
//
// What is hoisted out of the time loop (does not depend on the sampled state):
//   * embed @ W_obs[:, D:]^T for all B*T rows (one M=B*T contraction instead of T M=B products)
//   * the whole prior branch  deter -> _img_out_layers -> _imgs_stat_layer -> sample : the prior
//     sample is returned but never fed back (networks.py:198-206), so it runs once over B*T rows
//   * RSSM.initial (one row) and the is_first masks
// What stays in the loop is the true recurrence: one-hot gather of W_in -> LN/SiLU -> GRU GEMV ->
// LN/gates -> posterior GEMV -> LN/SiLU -> stat GEMV -> unimix sample.
//
// Reference: networks.py:127-143 (observe), 174-206 (obs_step), 208-233 (img_step),
// 99-125 + 235-239 (initial / get_stoch), tools.py:806-850 (static_scan).
#include "dv3_common.cuh"

namespace dv3 {

static int check_dims(const dv3_rssm_dims* d, const char* who) {
  DV3_REQUIRE(d, DV3_ERR_NULL, "%s: dims is NULL", who);
  DV3_REQUIRE(d->stoch >= 1 && d->classes >= 1 && d->classes <= 32, DV3_ERR_BAD_SHAPE,
              "%s: stoch=%d classes=%d (classes must be 1..32)", who, d->stoch, d->classes);
  DV3_REQUIRE(d->deter >= 4 && d->deter % 4 == 0 && d->hidden >= 4 && d->hidden % 4 == 0,
              DV3_ERR_BAD_SHAPE, "%s: deter=%d hidden=%d must be multiples of 4", who, d->deter,
              d->hidden);
  DV3_REQUIRE(d->actions >= 0 && d->embed >= 0 && d->embed % 4 == 0, DV3_ERR_BAD_SHAPE,
              "%s: actions=%d embed=%d (embed %% 4 == 0)", who, d->actions, d->embed);
  return 0;
}

// first_eff[b,t] = is_first[b,t] != 0, and every row at t == 0 when the caller passed no state
// (obs_step's `prev_state == None` branch, networks.py:176).
__global__ void first_eff_kernel(const float* __restrict__ is_first, int B, int T, int has_state,
                                 float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int t = i % T;
  out[i] = (is_first[i] != 0.f || (t == 0 && !has_state)) ? 1.f : 0.f;
}

// previous state / action seen by step t after the is_first mix (networks.py:176-193)
__global__ void obs_select_kernel(const float* __restrict__ first, int ldf,
                                  const int32_t* __restrict__ prev_idx, int ldpi,
                                  const float* __restrict__ prev_h, int ldph,
                                  const int32_t* __restrict__ init_idx,
                                  const float* __restrict__ init_deter,
                                  const float* __restrict__ action, int ldact, int S, int D, int A,
                                  int32_t* __restrict__ sprev, int lds, float* __restrict__ hprev,
                                  int ldh, float* __restrict__ aprev, int lda) {
  const int b = blockIdx.x;
  const bool f = first[(size_t)b * ldf] != 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x)
    sprev[(size_t)b * lds + i] = f ? init_idx[i] : prev_idx[(size_t)b * ldpi + i];
  for (int i = threadIdx.x; i < D; i += blockDim.x)
    hprev[(size_t)b * ldh + i] = f ? init_deter[i] : prev_h[(size_t)b * ldph + i];
  for (int i = threadIdx.x; i < A; i += blockDim.x)
    aprev[(size_t)b * lda + i] = f ? 0.f : action[(size_t)b * ldact + i];
}

struct ObsFwdWs {
  float* WinT;   // [(S*C+A), Hd]
  float* pre_e;  // [B*T, Hd]
  float* ascr;   // tensor-core split scratch for the bulk (B*T-row) products
  unsigned* bar; // grid-barrier words of the persistent kernel
  LinW obs_e, out, ims;
};

static void carve_fwd(Arena& a, const dv3_rssm_dims* d, int B, int T, ObsFwdWs& w) {
  const int SC = d->stoch * d->classes, D = d->deter, Hd = d->hidden, E = d->embed;
  const bool tc = B * T >= TC_MIN_ROWS;
  w.WinT = a.take<float>((size_t)(SC + d->actions) * Hd);
  w.pre_e = a.take<float>((size_t)B * T * Hd);
  w.bar = a.take<unsigned>(64);
  w.obs_e.reserve(a, tc && E > 0, Hd, E);
  w.out.reserve(a, tc, Hd, D);
  w.ims.reserve(a, tc, SC, Hd);
  int maxk = E > D ? E : D;
  if (Hd > maxk) maxk = Hd;
  w.ascr = tc ? a.take<float>((size_t)2 * B * T * maxk) : nullptr;
}

// RSSM.initial for one row: deter0 = tanh(W), stoch0 = mode(prior(deter0))
static int rssm_initial(const dv3_rssm_dims* d, const dv3_rssm_params* p, float* init_deter,
                        float* init_ypre, float* init_y, float* init_logit, int32_t* init_idx,
                        cudaStream_t st) {
  const int SC = d->stoch * d->classes;
  DV3_TRY(tanh_vec(p->w_init, d->deter, init_deter, st));
  DV3_TRY(linear1(init_deter, d->deter, p->w_out, d->deter, d->deter, nullptr, init_ypre,
                  d->hidden, 1, d->hidden, 0, st));
  DV3_TRY(ln_silu_fwd(init_ypre, d->hidden, p->ln_out_g, p->ln_out_b, d->ln_eps, 1, d->hidden,
                      init_y, d->hidden, st));
  DV3_TRY(linear1(init_y, d->hidden, p->w_ims, d->hidden, d->hidden, p->b_ims, init_logit, SC, 1,
                  SC, 0, st));
  DV3_TRY(onehot_sample(init_logit, SC, nullptr, 0, 0, 0, d->unimix, 1, d->stoch, d->classes,
                        init_idx, d->stoch, nullptr, 0, st));
  return 0;
}

}  // namespace dv3

using namespace dv3;

extern "C" size_t dv3_observe_workspace_bytes(const dv3_rssm_dims* d, int32_t B, int32_t T) {
  if (!d || B <= 0 || T <= 0) return 0;
  Arena a(nullptr, 0);
  ObsFwdWs w;
  carve_fwd(a, d, B, T, w);
  return a.used;
}

extern "C" int dv3_observe_fwd(const dv3_rssm_dims* d, const dv3_rssm_params* p,
                               const dv3_observe_io* io, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DV3_TRY(check_dims(d, "observe_fwd"));
  DV3_REQUIRE(p && io, DV3_ERR_NULL, "observe_fwd: params/io is NULL");
  const int B = io->B, T = io->T;
  DV3_REQUIRE(B >= 0 && T >= 0, DV3_ERR_BAD_SHAPE, "observe_fwd: B=%d T=%d", B, T);
  if (B == 0 || T == 0) return 0;  // empty batch / empty sequence: nothing to write
  const int S = d->stoch, C = d->classes, SC = S * C, D = d->deter, Hd = d->hidden, A = d->actions,
            E = d->embed;
  DV3_REQUIRE(io->embed && io->action && io->is_first && io->u_post && io->u_prior, DV3_ERR_NULL,
              "observe_fwd: null input");
  DV3_REQUIRE(io->post_stoch && io->post_logit && io->prior_stoch && io->prior_logit && io->deter,
              DV3_ERR_NULL, "observe_fwd: null output");
  DV3_REQUIRE(io->post_idx && io->prior_idx && io->sprev_idx && io->hprev && io->aprev &&
                  io->x_pre && io->x && io->g_pre && io->y_pre && io->y && io->z_pre && io->z &&
                  io->first_eff && io->init_deter && io->init_ypre && io->init_y &&
                  io->init_logit && io->init_idx,
              DV3_ERR_NULL, "observe_fwd: null saved-activation buffer");
  DV3_REQUIRE((io->state_idx == nullptr) == (io->state_deter == nullptr), DV3_ERR_NULL,
              "observe_fwd: state_idx and state_deter must be given together");
  Arena arena(io->workspace, io->workspace_bytes);
  ObsFwdWs w;
  carve_fwd(arena, d, B, T, w);
  DV3_REQUIRE(io->workspace && arena.ok(), DV3_ERR_WORKSPACE,
              "observe_fwd: workspace %zu < %zu bytes", io->workspace_bytes, arena.used);
  const int has_state = io->state_idx != nullptr;
  const int BT = B * T;

  // ---- hoisted work ----
  DV3_TRY(launch_transpose(p->w_in, SC + A, Hd, SC + A, w.WinT, st));
  DV3_TRY(w.obs_e.prepare(p->w_obs + D, D + E, st));
  DV3_TRY(w.out.prepare(p->w_out, D, st));
  DV3_TRY(w.ims.prepare(p->w_ims, Hd, st));
  DV3_TRY(w.obs_e.apply(io->embed, E, E, nullptr, 0, 0, nullptr, nullptr, 0, w.pre_e, Hd, BT, w.ascr,
                        st));
  first_eff_kernel<<<(BT + 255) / 256, 256, 0, st>>>(io->is_first, B, T, has_state, io->first_eff);
  DV3_CHECK_LAUNCH("first_eff_kernel");
  DV3_TRY(rssm_initial(d, p, io->init_deter, io->init_ypre, io->init_y, io->init_logit,
                       io->init_idx, st));

  // ---- the recurrence: one persistent kernel when the shapes allow, else a launch sequence ----
  bool persistent = false;
  DV3_TRY(observe_fwd_persistent(d, p, io, w.WinT, w.pre_e, w.bar, st, &persistent));
  for (int t = 0; t < T && !persistent; ++t) {
    const int32_t* prev_idx = t ? io->post_idx + (size_t)(t - 1) * S : io->state_idx;
    const int ldpi = t ? T * S : S;
    const float* prev_h = t ? io->deter + (size_t)(t - 1) * D : io->state_deter;
    const int ldph = t ? T * D : D;
    obs_select_kernel<<<B, 128, 0, st>>>(io->first_eff + t, T, prev_idx, ldpi, prev_h, ldph,
                                         io->init_idx, io->init_deter, io->action + (size_t)t * A,
                                         T * A, S, D, A, io->sprev_idx + (size_t)t * S, T * S,
                                         io->hprev + (size_t)t * D, T * D,
                                         io->aprev + (size_t)t * A, T * A);
    DV3_CHECK_LAUNCH("obs_select_kernel");
    // x = SiLU(LN(W_in [stoch, a]))
    DV3_TRY(gather_ln_silu(io->sprev_idx + (size_t)t * S, T * S, S, C, io->aprev + (size_t)t * A,
                           T * A, A, w.WinT, nullptr, 0, p->ln_in_g, p->ln_in_b, d->ln_eps, B, Hd,
                           io->x_pre + (size_t)t * Hd, T * Hd, io->x + (size_t)t * Hd, T * Hd, st));
    // g_pre = W_gru [x, h]
    LinearArgs g{};
    g.A[0] = io->x + (size_t)t * Hd; g.lda[0] = T * Hd; g.W[0] = p->w_gru; g.ldw[0] = Hd + D; g.K[0] = Hd;
    g.A[1] = io->hprev + (size_t)t * D; g.lda[1] = T * D; g.W[1] = p->w_gru + Hd; g.ldw[1] = Hd + D; g.K[1] = D;
    g.C = io->g_pre + (size_t)t * 3 * D; g.ldc = T * 3 * D; g.M = B; g.N = 3 * D;
    DV3_TRY(launch_linear(g, st));
    DV3_TRY(gru_gates_fwd(io->g_pre + (size_t)t * 3 * D, T * 3 * D, p->ln_gru_g, p->ln_gru_b,
                          d->ln_eps, io->hprev + (size_t)t * D, T * D, B, D,
                          io->deter + (size_t)t * D, T * D, st));
    // z = SiLU(LN(W_obs [deter, embed]))
    LinearArgs z{};
    z.A[0] = io->deter + (size_t)t * D; z.lda[0] = T * D; z.W[0] = p->w_obs; z.ldw[0] = D + E; z.K[0] = D;
    z.addend = w.pre_e + (size_t)t * Hd; z.ldadd = T * Hd;
    z.C = io->z_pre + (size_t)t * Hd; z.ldc = T * Hd; z.M = B; z.N = Hd;
    DV3_TRY(launch_linear(z, st));
    DV3_TRY(ln_silu_fwd(io->z_pre + (size_t)t * Hd, T * Hd, p->ln_obs_g, p->ln_obs_b, d->ln_eps, B,
                        Hd, io->z + (size_t)t * Hd, T * Hd, st));
    // posterior logits and draw
    DV3_TRY(linear1(io->z + (size_t)t * Hd, T * Hd, p->w_os, Hd, Hd, p->b_os,
                    io->post_logit + (size_t)t * SC, T * SC, B, SC, 0, st));
    DV3_TRY(onehot_sample(io->post_logit + (size_t)t * SC, T * SC, io->u_post + (size_t)t * B * SC,
                          SC, 0, 0, d->unimix, B, S, C, io->post_idx + (size_t)t * S, T * S,
                          io->post_stoch + (size_t)t * SC, T * SC, st));
  }

  // ---- prior branch for all B*T rows ----
  DV3_TRY(w.out.apply(io->deter, D, D, nullptr, 0, 0, nullptr, nullptr, 0, io->y_pre, Hd, BT, w.ascr,
                      st));
  DV3_TRY(ln_silu_fwd(io->y_pre, Hd, p->ln_out_g, p->ln_out_b, d->ln_eps, BT, Hd, io->y, Hd, st));
  DV3_TRY(w.ims.apply(io->y, Hd, Hd, nullptr, 0, 0, p->b_ims, nullptr, 0, io->prior_logit, SC, BT,
                      w.ascr, st));
  DV3_TRY(onehot_sample(io->prior_logit, SC, io->u_prior, SC, T, B, d->unimix, BT, S, C,
                        io->prior_idx, S, io->prior_stoch, SC, st));
  return 0;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
namespace dv3 {

struct ObsBwdWs {
  float *WosT, *WobsT, *WgruT, *WinT, *WimsT, *WoutT;
  float *d_y, *dh_prior;           // [B*T, Hd], [B*T, D]
  float *d_z, *dh_z;               // [B, Hd], [B, D]
  float *dxh, *dxh_add;            // [B, Hd+D]
  float *ds_tmp, *ds_rec, *dh_rec; // [B, SC], [B, SC], [B, D]
  float *dinit_s, *dinit_h;        // [B, SC], [B, D] per-row accumulators (deterministic)
  float* ascr;
  unsigned* bar;                   // grid-barrier word of the persistent kernel
  LinW ims, out, obs_e;            // bulk products over the transposed weights
};

static void carve_bwd(Arena& a, const dv3_rssm_dims* d, int B, int T, ObsBwdWs& w) {
  const size_t SC = (size_t)d->stoch * d->classes, D = d->deter, Hd = d->hidden, A = d->actions,
               E = d->embed;
  w.WosT = a.take<float>(Hd * SC);
  w.WobsT = a.take<float>((D + E) * Hd);
  w.WgruT = a.take<float>((Hd + D) * 3 * D);
  w.WinT = a.take<float>((SC + A) * Hd);
  w.WimsT = a.take<float>(Hd * SC);
  w.WoutT = a.take<float>(D * Hd);
  w.d_y = a.take<float>((size_t)B * T * Hd);
  w.dh_prior = a.take<float>((size_t)B * T * D);
  w.d_z = a.take<float>((size_t)B * Hd);
  w.dh_z = a.take<float>((size_t)B * D);
  w.dxh = a.take<float>((size_t)B * (Hd + D));
  w.dxh_add = a.take<float>((size_t)B * (Hd + D));
  w.ds_tmp = a.take<float>((size_t)B * SC);
  w.ds_rec = a.take<float>((size_t)B * SC);
  w.dh_rec = a.take<float>((size_t)B * D);
  w.dinit_s = a.take<float>((size_t)B * SC);
  w.dinit_h = a.take<float>((size_t)B * D);
  w.bar = a.take<unsigned>(64);
  const bool tc = B * T >= TC_MIN_ROWS;
  w.ims.reserve(a, tc, (int)Hd, (int)SC);     // d_y      = d_prior_logit @ W_ims   (K = SC)
  w.out.reserve(a, tc, (int)D, (int)Hd);      // dh_prior = d_y_pre @ W_out         (K = Hd)
  w.obs_e.reserve(a, tc && E > 0, (int)E, (int)Hd);   // d_embed = d_z_pre @ W_obs[:, D:]   (K = Hd)
  const size_t maxk = SC > Hd ? SC : Hd;
  w.ascr = tc ? a.take<float>((size_t)2 * B * T * maxk) : nullptr;
}

// Route the gradients that reached step t's *input* state: rows that were reset at t send them
// to RSSM.initial (accumulated per row, summed over rows at the end), the others hand them to
// step t-1 (straight-through into post_stoch[t-1], direct into deter[t-1]).
__global__ void obs_route_kernel(const float* __restrict__ first, int ldf,
                                 const float* __restrict__ ds, int SC, const float* __restrict__ dh,
                                 int lddh, int D, float* __restrict__ ds_rec,
                                 float* __restrict__ dh_rec, float* __restrict__ dinit_s,
                                 float* __restrict__ dinit_h) {
  const int b = blockIdx.x;
  const bool f = first[(size_t)b * ldf] != 0.f;
  for (int i = threadIdx.x; i < SC; i += blockDim.x) {
    const float v = ds[(size_t)b * SC + i];
    ds_rec[(size_t)b * SC + i] = f ? 0.f : v;
    if (f) dinit_s[(size_t)b * SC + i] += v;
  }
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const float v = dh[(size_t)b * lddh + i];
    dh_rec[(size_t)b * D + i] = f ? 0.f : v;
    if (f) dinit_h[(size_t)b * D + i] += v;
  }
}

// out[i] = sum_b in[b, i]  (fixed order)
__global__ void colsum_kernel(const float* __restrict__ in, int B, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += in[(size_t)b * n + i];
  out[i] = s;
}

}  // namespace dv3

extern "C" size_t dv3_observe_bwd_workspace_bytes(const dv3_rssm_dims* d, int32_t B, int32_t T) {
  if (!d || B <= 0 || T <= 0) return 0;
  Arena a(nullptr, 0);
  ObsBwdWs w;
  carve_bwd(a, d, B, T, w);
  return a.used;
}

extern "C" int dv3_observe_bwd(const dv3_rssm_dims* d, const dv3_rssm_params* p,
                               const dv3_observe_bwd_io* io, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DV3_TRY(check_dims(d, "observe_bwd"));
  DV3_REQUIRE(p && io, DV3_ERR_NULL, "observe_bwd: params/io is NULL");
  const int B = io->B, T = io->T;
  DV3_REQUIRE(B >= 0 && T >= 0, DV3_ERR_BAD_SHAPE, "observe_bwd: B=%d T=%d", B, T);
  if (B == 0 || T == 0) return 0;
  const int S = d->stoch, C = d->classes, SC = S * C, D = d->deter, Hd = d->hidden, A = d->actions,
            E = d->embed;
  DV3_REQUIRE(io->first_eff && io->post_logit && io->prior_logit && io->hprev && io->x_pre &&
                  io->g_pre && io->y_pre && io->z_pre,
              DV3_ERR_NULL, "observe_bwd: null saved tensor");
  DV3_REQUIRE(io->d_embed && io->d_x_pre && io->d_x_ln && io->d_g_pre && io->d_g_ln &&
                  io->d_y_pre && io->d_y_ln && io->d_z_pre && io->d_z_ln && io->d_post_logit &&
                  io->d_prior_logit && io->d_init_stoch && io->d_init_deter,
              DV3_ERR_NULL, "observe_bwd: null output");
  Arena arena(io->workspace, io->workspace_bytes);
  ObsBwdWs w;
  carve_bwd(arena, d, B, T, w);
  DV3_REQUIRE(io->workspace && arena.ok(), DV3_ERR_WORKSPACE,
              "observe_bwd: workspace %zu < %zu bytes", io->workspace_bytes, arena.used);
  const int BT = B * T;

  // transposed weights: every product below is then the K-contiguous "A W^T" form
  DV3_TRY(launch_transpose(p->w_os, Hd, SC, Hd, w.WosT, st));            // [Hd, SC]
  DV3_TRY(launch_transpose(p->w_obs, D + E, Hd, D + E, w.WobsT, st));    // [D+E, Hd]
  DV3_TRY(launch_transpose(p->w_gru, Hd + D, 3 * D, Hd + D, w.WgruT, st));  // [Hd+D, 3D]
  DV3_TRY(launch_transpose(p->w_in, SC + A, Hd, SC + A, w.WinT, st));    // [SC+A, Hd]
  DV3_TRY(launch_transpose(p->w_ims, Hd, SC, Hd, w.WimsT, st));          // [Hd, SC]
  DV3_TRY(launch_transpose(p->w_out, D, Hd, D, w.WoutT, st));            // [D, Hd]

  // ---- prior branch, all rows at once ----
  DV3_TRY(onehot_st_bwd(io->prior_logit, SC, io->g_prior_stoch, SC, nullptr, 0, io->g_prior_logit,
                        SC, d->unimix, BT, S, C, io->d_prior_logit, SC, st));
  DV3_TRY(w.ims.prepare(w.WimsT, SC, st));
  DV3_TRY(w.out.prepare(w.WoutT, Hd, st));
  DV3_TRY(w.obs_e.prepare(w.WobsT + (size_t)D * Hd, Hd, st));
  DV3_TRY(w.ims.apply(io->d_prior_logit, SC, SC, nullptr, 0, 0, nullptr, nullptr, 0, w.d_y, Hd, BT,
                      w.ascr, st));
  DV3_TRY(ln_silu_bwd(io->y_pre, Hd, p->ln_out_g, p->ln_out_b, d->ln_eps, w.d_y, Hd, BT, Hd,
                      io->d_y_pre, Hd, io->d_y_ln, Hd, st));
  DV3_TRY(w.out.apply(io->d_y_pre, Hd, Hd, nullptr, 0, 0, nullptr, nullptr, 0, w.dh_prior, D, BT,
                      w.ascr, st));

  DV3_TRY(fill_zero(w.ds_rec, (size_t)B * SC * 4, st));
  DV3_TRY(fill_zero(w.dh_rec, (size_t)B * D * 4, st));
  DV3_TRY(fill_zero(w.dinit_s, (size_t)B * SC * 4, st));
  DV3_TRY(fill_zero(w.dinit_h, (size_t)B * D * 4, st));
  DV3_TRY(fill_zero(w.dxh_add, (size_t)B * (Hd + D) * 4, st));

  // ---- reverse time: one persistent kernel when the shapes allow it ----
  ObsBwdShared sh{};
  sh.dh_prior = w.dh_prior; sh.WosT = w.WosT; sh.WobsT = w.WobsT; sh.WgruT = w.WgruT;
  sh.WinT = w.WinT; sh.ds_rec = w.ds_rec; sh.dh_rec = w.dh_rec; sh.dinit_s = w.dinit_s;
  sh.dinit_h = w.dinit_h; sh.d_z = w.d_z; sh.dh_z = w.dh_z; sh.dhdir = w.dxh_add; sh.dxh = w.dxh;
  sh.bar = w.bar;
  bool persistent = false;
  DV3_TRY(observe_bwd_persistent(d, p, io, sh, st, &persistent));
  for (int t = persistent ? -1 : T - 1; t >= 0; --t) {
    const float* gps = io->g_post_stoch ? io->g_post_stoch + (size_t)t * SC : nullptr;
    const float* gpl = io->g_post_logit ? io->g_post_logit + (size_t)t * SC : nullptr;
    DV3_TRY(onehot_st_bwd(io->post_logit + (size_t)t * SC, T * SC, gps, T * SC, w.ds_rec, SC, gpl,
                          T * SC, d->unimix, B, S, C, io->d_post_logit + (size_t)t * SC, T * SC,
                          st));
    DV3_TRY(linear1(io->d_post_logit + (size_t)t * SC, T * SC, w.WosT, SC, SC, nullptr, w.d_z, Hd,
                    B, Hd, 0, st));
    DV3_TRY(ln_silu_bwd(io->z_pre + (size_t)t * Hd, T * Hd, p->ln_obs_g, p->ln_obs_b, d->ln_eps,
                        w.d_z, Hd, B, Hd, io->d_z_pre + (size_t)t * Hd, T * Hd,
                        io->d_z_ln + (size_t)t * Hd, T * Hd, st));
    // d deter from the posterior branch: d_z_pre @ W_obs[:, :D]
    DV3_TRY(linear1(io->d_z_pre + (size_t)t * Hd, T * Hd, w.WobsT, Hd, Hd, nullptr, w.dh_z, D, B, D,
                    0, st));
    const float* dh_in[4] = {w.dh_z, w.dh_prior + (size_t)t * D, w.dh_rec,
                             io->g_deter ? io->g_deter + (size_t)t * D : nullptr};
    const int ld_in[4] = {D, T * D, D, T * D};
    DV3_TRY(gru_gates_bwd(io->g_pre + (size_t)t * 3 * D, T * 3 * D, p->ln_gru_g, p->ln_gru_b,
                          d->ln_eps, io->hprev + (size_t)t * D, T * D, dh_in, ld_in, B, D,
                          io->d_g_pre + (size_t)t * 3 * D, T * 3 * D,
                          io->d_g_ln + (size_t)t * 3 * D, T * 3 * D, w.dxh_add + Hd, Hd + D, st));
    // [dx | dh_prev] = d_g_pre @ W_gru  (+ the direct (1-u) path in the h columns)
    LinearArgs g{};
    g.A[0] = io->d_g_pre + (size_t)t * 3 * D; g.lda[0] = T * 3 * D; g.W[0] = w.WgruT; g.ldw[0] = 3 * D;
    g.K[0] = 3 * D; g.addend = w.dxh_add; g.ldadd = Hd + D; g.C = w.dxh; g.ldc = Hd + D; g.M = B;
    g.N = Hd + D;
    DV3_TRY(launch_linear(g, st));
    DV3_TRY(ln_silu_bwd(io->x_pre + (size_t)t * Hd, T * Hd, p->ln_in_g, p->ln_in_b, d->ln_eps,
                        w.dxh, Hd + D, B, Hd, io->d_x_pre + (size_t)t * Hd, T * Hd,
                        io->d_x_ln + (size_t)t * Hd, T * Hd, st));
    // d stoch_prev = d_x_pre @ W_in[:, :SC]
    DV3_TRY(linear1(io->d_x_pre + (size_t)t * Hd, T * Hd, w.WinT, Hd, Hd, nullptr, w.ds_tmp, SC, B,
                    SC, 0, st));
    obs_route_kernel<<<B, 256, 0, st>>>(io->first_eff + t, T, w.ds_tmp, SC, w.dxh + Hd, Hd + D, D,
                                        w.ds_rec, w.dh_rec, w.dinit_s, w.dinit_h);
    DV3_CHECK_LAUNCH("obs_route_kernel");
  }
  // ---- after the loop ----
  // d embed = d_z_pre @ W_obs[:, D:]
  if (E > 0)
    DV3_TRY(w.obs_e.apply(io->d_z_pre, Hd, Hd, nullptr, 0, 0, nullptr, nullptr, 0, io->d_embed, E, BT,
                          w.ascr, st));
  colsum_kernel<<<(SC + 255) / 256, 256, 0, st>>>(w.dinit_s, B, SC, io->d_init_stoch);
  DV3_CHECK_LAUNCH("colsum_kernel");
  colsum_kernel<<<(D + 255) / 256, 256, 0, st>>>(w.dinit_h, B, D, io->d_init_deter);
  DV3_CHECK_LAUNCH("colsum_kernel");
  if (io->d_state_deter) DV3_TRY(copy_rows(w.dh_rec, D, B, D, io->d_state_deter, D, st));
  if (io->d_state_stoch) DV3_TRY(copy_rows(w.ds_rec, SC, B, SC, io->d_state_stoch, SC, st));
  return 0;
}
