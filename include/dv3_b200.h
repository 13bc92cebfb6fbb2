/* dv3_b200.h -- C ABI of the B200-native DreamerV3 training hot path (libdv3_b200.so).
 *
 * The reference (ChenFengTsai/dreamerv3-torch) has no FFI: its boundary is the Python method
 * surface RSSM.observe / imagine_with_action, ImagBehavior._imagine, tools.lambda_return,
 * tools.DiscDist and RSSM.kl_loss.  Each entry point below names the reference code it
 * replaces (file:line relative to the reference tree).  INTEGRATION.md shows the ctypes stub a
 * reference maintainer would add.
 *
 * Conventions
 *  - plain C: pointers and sizes only, no torch types.  All tensors are contiguous fp32 in
 *    device memory (HBM) unless stated (int32 class indices), 16-byte aligned.
 *  - the caller allocates every input / output / workspace buffer; the library owns no memory
 *    and keeps no state between calls; parameters are read from the caller's storage each call.
 *  - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it (no host
 *    sync) and is re-entrant per stream.
 *  - return value: 0 on success, negative dv3_status otherwise; dv3_last_error() gives a
 *    thread-local message.
 *  - noise (uniforms / normals) is always an explicit input ("supplied uniforms" contract):
 *    categorical draw idx = argmax_k probs_k / (-log u_k).
 */
#ifndef DV3_B200_H
#define DV3_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DV3_ABI_VERSION 2

typedef enum {
  DV3_OK = 0,
  DV3_ERR_BAD_SHAPE = -1,    /* unsupported / inconsistent dims */
  DV3_ERR_NULL = -2,         /* required pointer is NULL */
  DV3_ERR_CUDA = -3,         /* a CUDA runtime call failed (message has the cudaError string) */
  DV3_ERR_WORKSPACE = -4     /* workspace too small */
} dv3_status;

int dv3_version(void);
const char* dv3_last_error(void);
/* The library reads its environment knobs (DV3_*) once and caches them;
 * call this after changing one inside a running process (tests, A/B measurements). */
void dv3_reload_env(void);
/* compute capability major*10+minor of the current device, or negative on error */
int dv3_device_arch(void);
/* number of kernels this library has launched in this process (all entry points) */
long long dv3_launch_count(void);
/* measurement aid for bench.py: when enabled, every GEMM launch is bracketed by CUDA events on
 * the caller's stream; dv3_prof_read waits for them and returns, per GEMM kind (index 0 = skinny
 * M<=32 kernel, 1 = tiled kernel), the summed device time in ms, the summed 2*M*N*K flops and
 * the launch count since the previous read.  Not capturable into a CUDA graph while enabled. */
void dv3_prof_enable(int on);
int dv3_prof_read(double* ms, double* flops, long long* launches);

/* ------------------------------------------------------------------------------------------
 * RSSM description (networks.py:13-97).  Weight layouts are PyTorch's [out, in] row-major,
 * exactly the tensors in the reference state_dict.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t stoch;    /* S: number of categorical groups (dyn_stoch, 32)          */
  int32_t classes;  /* C: classes per group (dyn_discrete, 32; must be <= 32)   */
  int32_t deter;    /* D: GRU state width (dyn_deter)                           */
  int32_t hidden;   /* Hd: dyn_hidden                                           */
  int32_t actions;  /* A                                                        */
  int32_t embed;    /* E: encoder output width                                  */
  float unimix;     /* 0.01                                                     */
  float ln_eps;     /* 1e-3                                                     */
} dv3_rssm_dims;

/* One operand of dv3_gemm_tc: tf32 hi / lo planes (hi = x with the 13 low mantissa bits cleared,
 * lo = x - hi) with a common row pitch ld (floats, % 4 == 0); mn_major == 0: stored [rows, K],
 * != 0: stored [K, rows]. */
typedef struct {
  const float* hi;
  const float* lo;
  int32_t ld;
  int32_t mn_major;
} dv3_tc_operand;

/* Optional derived forms of the RSSM weights, made by the CALLER once per optimizer step and
 * reused by every call until the weights change (dv3_split_tf32 / dv3_transpose, or the planes
 * dv3_adam_clip_step_planes writes).  With `planes == NULL` in dv3_rssm_params the library derives
 * what it needs inside each call (transposes + splits per call).  Used by dv3_imagine_fwd / _bwd
 * and dv3_img_step_fwd when the tensor-core path is taken (N >= 64 rows). */
typedef struct {
  dv3_tc_operand w_gru;      /* planes of _cell.layers.GRU_linear.weight [3D, Hd+D]          */
  dv3_tc_operand w_out;      /* planes of _img_out_layers.0.weight       [Hd, D]             */
  dv3_tc_operand w_ims;      /* planes of _imgs_stat_layer.weight        [S*C, Hd]           */
  const float* w_in_t;       /* (_img_in_layers.0.weight)^T              [S*C+A, Hd] fp32    */
  dv3_tc_operand w_in_t_sp;  /* planes of w_in_t                         [S*C+A, Hd]         */
} dv3_rssm_planes;

typedef struct {
  const float* w_in;      /* _img_in_layers.0.weight        [Hd, S*C+A]  */
  const float* ln_in_g;   /* _img_in_layers.1.weight        [Hd]         */
  const float* ln_in_b;   /* _img_in_layers.1.bias          [Hd]         */
  const float* w_gru;     /* _cell.layers.GRU_linear.weight [3D, Hd+D]   */
  const float* ln_gru_g;  /* _cell.layers.GRU_norm.weight   [3D]         */
  const float* ln_gru_b;  /* _cell.layers.GRU_norm.bias     [3D]         */
  const float* w_out;     /* _img_out_layers.0.weight       [Hd, D]      */
  const float* ln_out_g;  /* _img_out_layers.1.weight       [Hd]         */
  const float* ln_out_b;  /* _img_out_layers.1.bias         [Hd]         */
  const float* w_ims;     /* _imgs_stat_layer.weight        [S*C, Hd]    */
  const float* b_ims;     /* _imgs_stat_layer.bias          [S*C]        */
  const float* w_obs;     /* _obs_out_layers.0.weight       [Hd, D+E]    */
  const float* ln_obs_g;  /* _obs_out_layers.1.weight       [Hd]         */
  const float* ln_obs_b;  /* _obs_out_layers.1.bias         [Hd]         */
  const float* w_os;      /* _obs_stat_layer.weight         [S*C, Hd]    */
  const float* b_os;      /* _obs_stat_layer.bias           [S*C]        */
  const float* w_init;    /* W                              [1, D]       */
  const dv3_rssm_planes* planes;  /* optional caller-derived forms, or NULL                  */
} dv3_rssm_params;

/* ------------------------------------------------------------------------------------------
 * observe: T-step posterior rollout.  Replaces RSSM.observe (networks.py:127-143) =
 * tools.static_scan (tools.py:806-850) over RSSM.obs_step (networks.py:174-206), which calls
 * RSSM.img_step (208-233), GRUCell.forward (760-768), _suff_stats_layer (241-250) and
 * OneHotDist.sample (tools.py:452-460).
 *
 * Public tensors are batch-major [B,T,...] like the reference API; noise is time-major.
 * The `sv_*` tensors are activations saved for dv3_observe_bwd (all [B,T,...] batch-major).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, T;
  /* inputs */
  const float* embed;      /* [B,T,E]   */
  const float* action;     /* [B,T,A]   (never written; the reference zeroes is_first rows in place) */
  const float* is_first;   /* [B,T]     0/1 */
  const float* u_prior;    /* [T,B,S,C] uniforms in (0,1] for the prior draw  */
  const float* u_post;     /* [T,B,S,C] uniforms for the posterior draw       */
  const int32_t* state_idx;  /* optional [B,S]: class indices of a caller-supplied previous stoch, or NULL */
  const float* state_deter;  /* optional [B,D], or NULL (=> step 0 starts from RSSM.initial for every row) */
  /* outputs */
  float* post_stoch;   /* [B,T,S,C] one-hot */
  float* post_logit;   /* [B,T,S,C] */
  float* prior_stoch;  /* [B,T,S,C] one-hot */
  float* prior_logit;  /* [B,T,S,C] */
  float* deter;        /* [B,T,D]   */
  /* saved for backward */
  int32_t* post_idx;   /* [B,T,S] */
  int32_t* prior_idx;  /* [B,T,S] */
  float* first_eff;    /* [B,T]    effective reset mask: is_first, and every row at t=0 if no state */
  int32_t* sprev_idx;  /* [B,T,S]  previous-step stoch after the is_first mix */
  float* hprev;        /* [B,T,D]  previous deter after the mix */
  float* aprev;        /* [B,T,A]  action after the is_first zeroing */
  float* x_pre;        /* [B,T,Hd] W_in [s,a] (pre LayerNorm) */
  float* x;            /* [B,T,Hd] SiLU(LN(x_pre)) */
  float* g_pre;        /* [B,T,3D] W_gru [x,h] (pre LayerNorm) */
  float* y_pre;        /* [B,T,Hd] */
  float* y;            /* [B,T,Hd] */
  float* z_pre;        /* [B,T,Hd] */
  float* z;            /* [B,T,Hd] */
  /* RSSM.initial (networks.py:99-125) intermediates, one row */
  float* init_deter;   /* [D]   tanh(W) */
  float* init_ypre;    /* [Hd]  */
  float* init_y;       /* [Hd]  */
  float* init_logit;   /* [S*C] */
  int32_t* init_idx;   /* [S]   argmax = mode */
  /* scratch */
  void* workspace;
  size_t workspace_bytes;
} dv3_observe_io;

size_t dv3_observe_workspace_bytes(const dv3_rssm_dims* d, int32_t B, int32_t T);
int dv3_observe_fwd(const dv3_rssm_dims* d, const dv3_rssm_params* p, const dv3_observe_io* io,
                    void* stream);

/* Backward of observe: BPTT through the T steps.  Consumes upstream gradients of the five
 * public outputs (any may be NULL = zero) and produces d embed plus, for every Linear /
 * LayerNorm on the path, the gradient w.r.t. its OUTPUT for all B*T rows ("deltas").  The
 * caller turns deltas into parameter gradients with plain GEMMs (dW = delta^T @ input) --
 * contractions over B*T = 1024 rows that do not sit inside the time loop.
 * Straight-through estimator, unimix and the renormalisation inside torch's Categorical are
 * differentiated exactly as autograd does for tools.py:436-460. */
typedef struct {
  int32_t B, T;
  /* saved forward state: same tensors as dv3_observe_io */
  const float* first_eff;
  const float* post_logit;
  const float* prior_logit;
  const float* hprev;
  const float* x_pre;
  const float* g_pre;
  const float* y_pre;
  const float* z_pre;
  /* upstream gradients, batch-major, NULL = zeros */
  const float* g_post_stoch;   /* [B,T,S,C] */
  const float* g_post_logit;   /* [B,T,S,C] */
  const float* g_prior_stoch;  /* [B,T,S,C] */
  const float* g_prior_logit;  /* [B,T,S,C] */
  const float* g_deter;        /* [B,T,D]   */
  /* outputs */
  float* d_embed;      /* [B,T,E] */
  float* d_x_pre;      /* [B,T,Hd]  delta of _img_in_layers.0 output   */
  float* d_x_ln;       /* [B,T,Hd]  grad wrt LayerNorm affine output   */
  float* d_g_pre;      /* [B,T,3D]  delta of GRU_linear output         */
  float* d_g_ln;       /* [B,T,3D]  */
  float* d_y_pre;      /* [B,T,Hd]  */
  float* d_y_ln;       /* [B,T,Hd]  */
  float* d_z_pre;      /* [B,T,Hd]  */
  float* d_z_ln;       /* [B,T,Hd]  */
  float* d_post_logit; /* [B,T,S*C] delta of _obs_stat_layer output    */
  float* d_prior_logit;/* [B,T,S*C] delta of _imgs_stat_layer output   */
  float* d_init_stoch; /* [S*C]  summed grad reaching RSSM.initial's stoch */
  float* d_init_deter; /* [D]    summed grad reaching RSSM.initial's deter */
  float* d_state_deter;/* optional [B,D]: grad wrt caller-supplied state deter, or NULL */
  float* d_state_stoch;/* optional [B,S*C]: grad wrt caller-supplied state stoch, or NULL */
  void* workspace;
  size_t workspace_bytes;
} dv3_observe_bwd_io;

size_t dv3_observe_bwd_workspace_bytes(const dv3_rssm_dims* d, int32_t B, int32_t T);
int dv3_observe_bwd(const dv3_rssm_dims* d, const dv3_rssm_params* p,
                    const dv3_observe_bwd_io* io, void* stream);

/* ------------------------------------------------------------------------------------------
 * imagine: H-step actor-in-the-loop prior rollout from N start states.  Replaces
 * ImagBehavior._imagine (models.py:448-548) with policy = actor MLP (networks.py:657-700);
 * with actor == NULL and given actions it is RSSM.imagine_with_action (networks.py:145-152).
 * Time-major [H,N,...] like the reference.
 * ---------------------------------------------------------------------------------------- */
/* Optional derived forms of the actor trunk weights (see dv3_rssm_planes). */
typedef struct {
  dv3_tc_operand w[16];  /* planes of Actor_linear{i}.weight: [U, F] for i == 0, then [U, U]  */
  const float* w0_t;     /* (Actor_linear0.weight)^T [F, U] fp32 (one-hot gather form)        */
} dv3_actor_planes;

typedef struct {
  int32_t layers;        /* trunk layers (2 or 5) */
  int32_t units;         /* U */
  int32_t dist;          /* 0 = 'normal' (tanh mean, learned std, clip 1.0), 1 = 'onehot' */
  float min_std, max_std;/* 0.1, 1.0 */
  float unimix;          /* 0.01 (onehot) */
  const float* const* w;     /* [layers] Actor_linear{i}.weight: [U, F] then [U, U]  */
  const float* const* ln_g;  /* [layers] Actor_norm{i}.weight [U] */
  const float* const* ln_b;  /* [layers] */
  const float* w_mean; const float* b_mean;  /* mean_layer [A, U], [A] */
  const float* w_std;  const float* b_std;   /* std_layer  [A, U], [A] (normal only) */
  const dv3_actor_planes* planes;            /* optional caller-derived forms, or NULL */
} dv3_actor;

typedef struct {
  int32_t N, H;
  /* inputs */
  const int32_t* start_idx;   /* [N,S]  class indices of the (one-hot) start stoch */
  const float* start_deter;   /* [N,D]  */
  const float* act_noise;     /* [H,N,A] N(0,1) draws (normal) or uniforms (onehot); unused if actor==NULL */
  const float* u_state;       /* [H,N,S,C] uniforms for the prior draws (last step's are not consumed) */
  const float* given_action;  /* [H-1,N,A] when actor == NULL (action k drives state k -> k+1) */
  /* outputs: state k for k=0..H-1 where state 0 = start, action k, feat k */
  float* feat;        /* [H,N,S*C+D] = [one-hot stoch | deter]  (row k is also states.stoch/deter k) */
  float* logit;       /* [H,N,S,C]  rows 1..H-1 (row 0 is the caller's start logit, not written) */
  float* action;      /* [H,N,A] */
  int32_t* idx;       /* [H,N,S]  class indices of state k (row 0 copies start_idx) */
  /* saved for backward (rows k = step k, i.e. the transition state k -> k+1; row H-1 unused) */
  float* x_pre; float* x;        /* [H,N,Hd] */
  float* g_pre;                  /* [H,N,3D] */
  float* y_pre; float* y;        /* [H,N,Hd] */
  float* a_pre;       /* [layers,H,N,U] actor trunk pre-LN */
  float* a_act;       /* [layers,H,N,U] actor trunk post-SiLU */
  float* a_mean_raw;  /* [H,N,A] mean_layer output (logits for onehot) */
  float* a_std_raw;   /* [H,N,A] (normal only) */
  void* workspace; size_t workspace_bytes;
} dv3_imagine_io;

/* single steps, exported for teacher-forced parity checks and for acting:
 *   dv3_obs_step_fwd = RSSM.obs_step (networks.py:174-206): observe io with T == 1 + state
 *   dv3_img_step_fwd = RSSM.img_step (networks.py:208-233): imagine io with H == 2, given_action */
int dv3_obs_step_fwd(const dv3_rssm_dims* d, const dv3_rssm_params* p, const dv3_observe_io* io,
                     void* stream);
size_t dv3_imagine_workspace_bytes(const dv3_rssm_dims* d, const dv3_actor* a, int32_t N, int32_t H);
int dv3_imagine_fwd(const dv3_rssm_dims* d, const dv3_rssm_params* p, const dv3_actor* a,
                    const dv3_imagine_io* io, void* stream);
int dv3_img_step_fwd(const dv3_rssm_dims* d, const dv3_rssm_params* p, const dv3_imagine_io* io,
                     void* stream);

/* Backward of imagine ('dynamics' gradient, models.py:513-517 with feat detached): BPTT through
 * img_step w.r.t. activations, producing the gradient reaching each sampled action and from it
 * the deltas of the actor heads.  RSSM deltas are produced too (for callers that train the
 * world model through imagination); the actor trunk backward is not recurrent (its input is
 * detached) and is done by the caller in bulk over H*N rows. */
typedef struct {
  int32_t N, H;
  const float* logit;          /* [H,N,S,C] as written by imagine_fwd (row 0 unused) */
  const float* feat;           /* [H,N,F] */
  const float* x_pre; const float* g_pre; const float* y_pre;
  const float* a_mean_raw; const float* a_std_raw; const float* act_noise;
  /* upstream grads (time-major, NULL = zeros) */
  const float* g_stoch;   /* [H,N,S,C] grad wrt states.stoch */
  const float* g_deter;   /* [H,N,D]   grad wrt states.deter */
  const float* g_logit;   /* [H,N,S,C] grad wrt states.logit (rows >= 1) */
  const float* g_action;  /* [H,N,A]   grad wrt returned actions */
  /* outputs */
  float* d_mean_raw;      /* [H,N,A] delta of actor mean_layer output */
  float* d_std_raw;       /* [H,N,A] */
  float* d_x_pre; float* d_x_ln;   /* [H,N,Hd] (rows 0..H-2) */
  float* d_g_pre; float* d_g_ln;   /* [H,N,3D] */
  float* d_y_pre; float* d_y_ln;   /* [H,N,Hd] */
  float* d_logit;                  /* [H,N,S*C] delta of _imgs_stat_layer output (rows 1..H-1) */
  float* d_start_stoch;            /* [N,S*C] grad wrt start stoch */
  float* d_start_deter;            /* [N,D]   */
  void* workspace; size_t workspace_bytes;
  int32_t g_state_ld;              /* row pitch (floats) of g_stoch / g_deter; 0 = packed (S*C / D).
                                    * With S*C+D the two are the column ranges of ONE [H,N,S*C+D]
                                    * gradient of the feature buffer (g_deter = g_stoch + S*C): no
                                    * split copies of the upstream gradient */
} dv3_imagine_bwd_io;

size_t dv3_imagine_bwd_workspace_bytes(const dv3_rssm_dims* d, const dv3_actor* a, int32_t N, int32_t H);
int dv3_imagine_bwd(const dv3_rssm_dims* d, const dv3_rssm_params* p, const dv3_actor* a,
                    const dv3_imagine_bwd_io* io, void* stream);

/* ------------------------------------------------------------------------------------------
 * lambda return.  Replaces tools.lambda_return + static_scan_for_lambda_return
 * (tools.py:682-728).  Time-major [H,N]; bootstrap [N].
 *   R_t = r_t + c_t * ((1-lambda) * v_{t+1} + lambda * R_{t+1}),  v_H = R_H = bootstrap.
 * lambda_ is a double so that (1-lambda) is rounded to fp32 once, as the reference's Python
 * scalar is (bit-identical fp32 results).
 * ---------------------------------------------------------------------------------------- */
int dv3_lambda_return_fwd(const float* reward, const float* value, const float* pcont,
                          const float* bootstrap, double lambda_, int32_t H, int32_t N,
                          float* ret, void* stream);
int dv3_lambda_return_bwd(const float* value, const float* pcont, const float* bootstrap,
                          const float* ret, const float* g_ret, double lambda_, int32_t H,
                          int32_t N, float* d_reward, float* d_value, float* d_pcont,
                          float* d_bootstrap, void* stream);

/* ------------------------------------------------------------------------------------------
 * symlog two-hot head.  Replaces tools.DiscDist.log_prob / mean / mode (tools.py:463-513) with
 * tools.symlog / symexp (22-27).  logits [R,K] (K = 255), x [R], buckets [K] (pass the tensor
 * torch.linspace(-20,20,255) built so bucket values are bit-identical to the reference's).
 * ---------------------------------------------------------------------------------------- */
int dv3_twohot_logprob_fwd(const float* logits, const float* x, const float* buckets, int32_t R,
                           int32_t K, float* logprob, void* stream);
int dv3_twohot_logprob_bwd(const float* logits, const float* x, const float* buckets,
                           const float* g_logprob, int32_t R, int32_t K, float* d_logits,
                           void* stream);
int dv3_twohot_mean_fwd(const float* logits, const float* buckets, int32_t R, int32_t K,
                        float* mean, void* stream);
int dv3_twohot_mean_bwd(const float* logits, const float* buckets, const float* g_mean, int32_t R,
                        int32_t K, float* d_logits, void* stream);

/* ------------------------------------------------------------------------------------------
 * KL balance.  Replaces RSSM.kl_loss (networks.py:272-290) over
 * Independent(OneHotDist(unimix),1) with torch's categorical KL.  logits [R,S,C].
 * outputs [R]: loss = dyn_scale*max(dyn,free) + rep_scale*max(rep,free); value = raw KL;
 * post_ent / prior_ent = entropies (the metrics of models.py:156-168).  Backward takes the
 * gradient of `loss` only (value/dyn/rep/entropies are metrics).
 * ---------------------------------------------------------------------------------------- */
int dv3_kl_balance_fwd(const float* post_logit, const float* prior_logit, int32_t R, int32_t S,
                       int32_t C, float unimix, float free_nats, float dyn_scale, float rep_scale,
                       float* loss, float* value, float* dyn, float* rep, float* post_ent,
                       float* prior_ent, void* stream);
int dv3_kl_balance_bwd(const float* post_logit, const float* prior_logit, const float* g_loss,
                       int32_t R, int32_t S, int32_t C, float unimix, float free_nats,
                       float dyn_scale, float rep_scale, float* d_post_logit,
                       float* d_prior_logit, void* stream);

/* ------------------------------------------------------------------------------------------
 * Building blocks (the kernels the ops above are made of), exported so that the host side can
 * run the non-recurrent bulk parts (actor trunk backward over H*N rows, parameter-gradient
 * reductions) and so that each kernel is testable against the oracle in isolation.
 * ---------------------------------------------------------------------------------------- */
/* C[M,N] (ldc) = [A1|A2] W^T + bias + addend;  W = [W1|W2] given as two K-segments of one
 * [N, *] row-major weight (A2/W2 may be NULL).  accumulate != 0 adds into C.  K1,K2 % 4 == 0. */
int dv3_linear_fwd(const float* A1, int32_t lda1, const float* W1, int32_t ldw1, int32_t K1,
                   const float* A2, int32_t lda2, const float* W2, int32_t ldw2, int32_t K2,
                   const float* bias, const float* addend, int32_t ldadd, float* C, int32_t ldc,
                   int32_t M, int32_t N, int32_t accumulate, void* stream);
/* C[M,N] = op(A) op(W)^T + bias + addend on the tcgen05 tensor cores, fp32-accurate through a
 * 3xTF32 split (hi*hi + hi*lo + lo*hi; fp32 accumulation in TMEM, promoted every 128 k to fp32
 * registers because the tensor core truncates its accumulator adds; TMA-staged 128B-swizzled
 * operand tiles).  op(A) is [M,K]: transA == 0 -> A is stored [M,K] (row stride lda), transA != 0
 * -> A is stored [K,M].  op(W) is [N,K]: transW == 0 -> stored [N,K] (the nn.Linear layout),
 * transW != 0 -> stored [K,N].  This covers y = x W^T, dx = dy W (transW) and dW = dy^T x (both).
 * scratch: dv3_linear_tc_scratch_bytes(M,N,K) bytes for the split (zero-padded) operands. */
size_t dv3_linear_tc_scratch_bytes(int32_t M, int32_t N, int32_t K);
int dv3_linear_tc_fwd(const float* A, int32_t lda, int32_t transA, const float* W, int32_t ldw,
                      int32_t transW, const float* bias, const float* addend, int32_t ldadd,
                      float* C, int32_t ldc, int32_t M, int32_t N, int32_t K, void* scratch,
                      size_t scratch_bytes, void* stream);
/* The tcgen05 GEMM proper: C[M,N] = [A1|A2] B^T (+bias +addend, +C when accumulate != 0) from
 * operands already split into tf32 hi/lo planes (dv3_split_tf32, or emitted by the row kernels).
 * mn_major == 0: the operand is stored [rows, K] (row stride ld); mn_major != 0: stored
 * [K, rows] -- so y = x W^T, dx = dy W and dW = dy^T x read the same planes, no transposes.
 * Planes are 16-byte aligned, ld % 4 == 0.  lo == NULL on all operands: raw fp32, split in the SM.
 * accumulate: bit 0 = add into C; bit 1 = allow split-K (deep contractions with few output tiles
 * -- the dW products -- are partitioned along K and combined with fp32 atomics; the summation
 * order is then not fixed, so callers that need run-to-run bit-identical results leave it 0).
 * Persistent kernels: single-CTA 128 x {128,64,32} tiles (dv3_umma2.cu) or cta_group::2 CTA pairs
 * with 256 x {128,64} tiles (dv3_umma2x.cu), chosen per shape by a cost model; TMA 128B-swizzled
 * stages, fp32 accumulation in TMEM promoted to registers every 128 k.
 * Replaces: every nn.Linear product of the path (networks.py:48-78 RSSM layers, 623-655 MLP,
 * 742-768 GRUCell) and the dx / dW matmuls autograd derives from them. */
int dv3_gemm_tc(const dv3_tc_operand* A1, int32_t K1, const dv3_tc_operand* A2, int32_t K2,
                const dv3_tc_operand* B, const float* bias, const float* addend, int32_t ldadd,
                float* C, int32_t ldc, int32_t M, int32_t N, int32_t accumulate, void* stream);
/* The same product with the A operand as plain fp32 ([M, K] row-major, 16-byte aligned, row strides
 * % 4 == 0) and B as tf32 planes: the fp32 tile is split in the SM and fed to the tensor core from
 * tensor memory (dv3_umma2t.cu) -- half the A bytes per k-block of dv3_gemm_tc, bit-identical
 * results.  For the skinny products of the time loops (M >= 64; K1 % 32 == 0 when A2 is given). */
int dv3_gemm_tc_rawa(const float* A1, int32_t lda1, int32_t K1, const float* A2, int32_t lda2,
                     int32_t K2, const dv3_tc_operand* B, const float* bias, const float* addend,
                     int32_t ldadd, float* C, int32_t ldc, int32_t M, int32_t N, void* stream);
/* hi = x with the 13 low mantissa bits cleared, lo = x - hi; x is [rows, cols] with row stride
 * ld, the planes have row stride ld_out >= cols (pad columns are zeroed). */
int dv3_split_tf32(const float* x, int32_t ld, int32_t rows, int32_t cols, float* hi, float* lo,
                   int32_t ld_out, void* stream);
/* Same product straight from fp32 operands, no scratch and no pre-pass: C = [A1|A2] W^T (+bias
 * +addend, +C when accumulate != 0).  TMA loads each fp32 tile once; the hi/lo split is done
 * inside the SM by converter warps of the persistent GEMM kernel (dv3_umma2.cu).  A1/A2/W must be
 * 16-byte aligned with row strides (in floats) that are multiples of 4; A2 may be NULL. */
int dv3_linear_tc2_fwd(const float* A1, int32_t lda1, int32_t K1, const float* A2, int32_t lda2,
                       int32_t K2, const float* W, int32_t ldw, const float* bias,
                       const float* addend, int32_t ldadd, float* C, int32_t ldc, int32_t M,
                       int32_t N, int32_t accumulate, void* stream);
/* out[C,R] = in[R,C]^T  (in has row stride ld) */
int dv3_transpose(const float* in, int32_t ld, int32_t R, int32_t C, float* out, void* stream);
/* out = SiLU(LayerNorm(pre)) row-wise; networks.py:48-58 (Linear->LN->SiLU blocks) */
int dv3_ln_silu_fwd(const float* pre, int32_t ld, const float* g, const float* b, float eps,
                    int32_t M, int32_t n, float* out, int32_t ldo, void* stream);
int dv3_ln_silu_bwd(const float* pre, int32_t ld, const float* g, const float* b, float eps,
                    const float* d_out, int32_t ldd, int32_t M, int32_t n, float* d_pre,
                    float* d_ln, int32_t ldp, void* stream);
/* The same two kernels also writing the tf32 hi/lo planes (row pitch lds >= n, n % 4 == 0 keeps
 * pad columns out of play) of their output -- the A operand of the dv3_gemm_tc that consumes it,
 * so no separate dv3_split_tf32 pass is needed. */
int dv3_ln_silu_fwd_split(const float* pre, int32_t ld, const float* g, const float* b, float eps,
                          int32_t M, int32_t n, float* out, int32_t ldo, float* hi, float* lo,
                          int32_t lds, void* stream);
int dv3_ln_silu_bwd_split(const float* pre, int32_t ld, const float* g, const float* b, float eps,
                          const float* d_out, int32_t ldd, int32_t M, int32_t n, float* d_pre,
                          float* d_ln, int32_t ldp, float* hi, float* lo, int32_t lds,
                          void* stream);
/* LayerNorm affine-parameter gradients over all M rows: dg[j] = sum_r d_ln[r,j]*xhat[r,j],
 * db[j] = sum_r d_ln[r,j]; xhat is recomputed from the saved pre-LN rows.  n <= 2048.
 * (autograd of nn.LayerNorm's weight / bias, networks.py:56, 628, 758) */
int dv3_ln_param_grads(const float* pre, int32_t ld, const float* d_ln, int32_t ldl, float eps,
                       int32_t M, int32_t n, float* dg, float* db, void* stream);
/* the same sums ADDED onto dg / db (no zeroing): accumulation into a pre-zeroed flat gradient */
int dv3_ln_param_grads_acc(const float* pre, int32_t ld, const float* d_ln, int32_t ldl, float eps,
                           int32_t M, int32_t n, float* dg, float* db, void* stream);
/* LayerNorm-GRU gate block (networks.py:760-768): parts = LN_3D(g_pre); r = sig(p0);
 * c = tanh(r*p1); u = sig(p2 - 1); h_new = u*c + (1-u)*h.  bwd also returns d_g_ln (gradient
 * w.r.t. the LayerNorm affine output) and d_h = the direct (1-u) path only. */
int dv3_gru_gates_fwd(const float* g_pre, int32_t ldg, const float* g, const float* b, float eps,
                      const float* h, int32_t ldh, int32_t M, int32_t D, float* h_new, int32_t ldn,
                      void* stream);
int dv3_gru_gates_bwd(const float* g_pre, int32_t ldg, const float* g, const float* b, float eps,
                      const float* h, int32_t ldh, const float* d_h_new, int32_t ldd, int32_t M,
                      int32_t D, float* d_g_pre, float* d_g_ln, int32_t ldp, float* d_h,
                      int32_t ldo, void* stream);
/* Linear over a one-hot input fused with LN+SiLU: pre = addend + sum_s WT[s*C+idx[s]] +
 * act @ WT[S*C:], out = SiLU(LN(pre)).  WT is the TRANSPOSED weight [S*C+A, n]; idx [M,S];
 * act [M,A] or NULL; addend [M,n] or NULL.  (networks.py:48-58 applied to cat(one-hot, action)) */
int dv3_onehot_linear_ln_silu(const int32_t* idx, int32_t S, int32_t C, const float* act,
                              int32_t A, const float* WT, const float* addend, const float* g,
                              const float* b, float eps, int32_t M, int32_t n, float* pre,
                              float* out, void* stream);
/* unimix categorical draw / mode for logits [M,S,C] (C<=32): idx [M,S] and optional one-hot.
 * u == NULL -> mode (first argmax).  tools.py:436-460 */
int dv3_onehot_sample(const float* logits, const float* u, float unimix, int32_t M, int32_t S,
                      int32_t C, int32_t* idx, float* onehot, int32_t ld_onehot, void* stream);
/* d_logits = ext + straight-through backward of a sample whose value-grad is g_sample */
int dv3_onehot_st_bwd(const float* logits, const float* g_sample, const float* ext, float unimix,
                      int32_t M, int32_t S, int32_t C, float* d_logits, void* stream);
/* one-hot fp32 [M,S*C] (row stride ld) from indices [M,S] */
int dv3_idx_to_onehot(const int32_t* idx, int32_t M, int32_t S, int32_t C, float* out, int32_t ld,
                      void* stream);

/* tools.Optimizer.__call__ (tools.py:760-776) after backward, on one flat fp32 buffer of n
 * (multiple of 4) parameters: global-norm clip (coef = min(1, clip / (norm + 1e-6)); clip <= 0
 * disables it) + torch.optim.Adam update; p *= decay_mul first (decoupled weight decay of the
 * reference, 1.0 = off).  step: device float, incremented here.  ctl: 4 floats out (grad norm
 * before clipping, clip coefficient, lr / bias_correction1, sqrt(bias_correction2)).
 * scratch: >= 512 floats. */
int dv3_adam_clip_step(float* p, const float* g, float* m, float* v, long long n, float lr,
                       float beta1, float beta2, float eps, float clip, float decay_mul,
                       float* step, float* ctl, float* scratch, void* stream);
/* The same update, also writing the tf32 hi / lo planes of the updated parameters into flat
 * buffers of the same layout (hi + lo == p exactly): the operand planes dv3_gemm_tc reads for
 * every weight of the next step, produced by the pass that writes the weights. */
int dv3_adam_clip_step_planes(float* p, const float* g, float* m, float* v, long long n, float lr,
                              float beta1, float beta2, float eps, float clip, float decay_mul,
                              float* step, float* ctl, float* scratch, float* hi, float* lo,
                              void* stream);

/* ------------------------------------------------------------------------------------------
 * Step tail: the element-wise / reduction chains around the rollouts, one launch each
 * (dv3_tail.cu).  Every array is contiguous fp32; "rows" R are the B*T or H*N samples.
 * ---------------------------------------------------------------------------------------- */
/* tools.symlog (tools.py:22-23), element-wise: the MLP encoder's input transform (networks.py:333) */
int dv3_symlog(const float* x, long long n, float* out, void* stream);
/* Squared-error heads: logprob[r] = -sum_j (mode[r,j] - t(value[r,j]))^2.
 * use_symlog != 0: t = symlog and terms below tol are dropped -- tools.SymlogDist.log_prob
 * (tools.py:546-572, dist 'mse', agg 'sum', tol 1e-8); use_symlog == 0: t = identity --
 * tools.MSEDist.log_prob (tools.py:520-543).  bwd: d_mode = g_logprob[r] * d logprob / d mode. */
int dv3_sqerr_logprob_fwd(const float* mode, const float* value, int32_t R, int32_t n,
                          int32_t use_symlog, float tol, float* logprob, void* stream);
int dv3_sqerr_logprob_bwd(const float* mode, const float* value, const float* g_logprob, int32_t R,
                          int32_t n, int32_t use_symlog, float tol, float* d_mode, void* stream);
/* tools.Bernoulli.log_prob (tools.py:604-628): -softplus(l)(1-x) - softplus(-l)x, element-wise */
int dv3_bernoulli_logprob_fwd(const float* logit, const float* x, long long n, float* logprob,
                              void* stream);
int dv3_bernoulli_logprob_bwd(const float* logit, const float* x, const float* g, long long n,
                              float* d_logit, void* stream);
/* model_loss = mean_r sum_i scales[i] * terms[i][r]  (models.py:140-152; a term is a head's
 * log_prob with scale -loss_scale, or the KL loss with scale 1): out is one float; neg_out
 * (optional, [n_terms, R]) receives -terms[i][r], the per-head losses the metrics report.
 * bwd fills grads [n_terms, R] with g[0] * scales[i] / R.  terms / scales are HOST arrays. */
int dv3_loss_mean_fwd(const float* const* losses, const float* scales, int32_t n_losses, int32_t R,
                      float* out, float* neg_out, void* stream);
int dv3_loss_mean_bwd(const float* g, const float* scales, int32_t n_losses, int32_t R, float* grads,
                      void* stream);
/* ImagBehavior._compute_target (models.py:620-638), time-major [H,N]:
 * discount = gamma * sigmoid(cont_logit); weights[t] = prod_{i<t} discount[i] (weights[0] = 1) */
int dv3_discount_weights_fwd(const float* cont_logit, float gamma, int32_t H, int32_t N,
                             float* discount, float* weights, void* stream);
int dv3_discount_bwd(const float* cont_logit, const float* g_discount, float gamma, long long n,
                     float* d_logit, void* stream);
/* RewardEMA.__call__ (models.py:11-26): 5 % / 95 % torch.quantile ('linear') of x[n] (n <= 2^20;
 * the four order statistics by exact radix select in one CTA), ema_vals = alpha q + (1-alpha) ema_vals in place,
 * offset_scale = {ema_vals[0], max(ema_vals[1] - ema_vals[0], 1)}. */
int dv3_reward_ema(const float* x, int32_t n, double alpha, float* ema_vals, float* offset_scale,
                   void* stream);
/* ImagBehavior._compute_actor_loss + the entropy term (models.py:393-397, 640-681) over
 * count = (H-1)*N leading elements of time-major [H,N] arrays.  mode 0 'dynamics':
 * -w ((target-off)/scale - (base-off)/scale) - c_ent ent; mode 1 'reinforce':
 * -w logp (target - base) - c_ent ent.  loss = mean; normed (optional) = (target-off)/scale.
 * offset_scale NULL = reward_EMA off.  bwd writes d_target [count] (mode 0), d_logp [total]
 * (mode 1) and d_entropy [total], total = H*N. */
int dv3_actor_loss_fwd(const float* target, const float* base, const float* weights,
                       const float* entropy, const float* logp, const float* offset_scale,
                       float entropy_coef, int32_t mode, int32_t count, float* normed, float* loss,
                       void* stream);
int dv3_actor_loss_bwd(const float* g_loss, const float* target, const float* base,
                       const float* weights, const float* offset_scale, float entropy_coef,
                       int32_t mode, int32_t count, int32_t total, float* d_target, float* d_entropy,
                       float* d_logp, void* stream);
/* value loss (models.py:419-429): mean_i weights[i] * (-lp_target[i] - lp_slow[i]) (lp_slow may be
 * NULL); bwd: d_lp[i] = -weights[i] g / count for both log-prob inputs */
int dv3_value_loss_fwd(const float* lp_target, const float* lp_slow, const float* weights,
                       int32_t count, float* loss, void* stream);
int dv3_value_loss_bwd(const float* g_loss, const float* weights, int32_t count, float* d_lp,
                       void* stream);
/* 'normal' actor distribution (networks.py:693-700, tools.py:575-601) from the raw head outputs
 * [R,A]: mean = tanh(mean_raw), std = (max-min) sigmoid(std_raw + 2) + min; entropy [R] and (logp
 * != NULL) log_prob(action) [R].  bwd takes g_entropy / g_logp (either may be NULL). */
int dv3_normal_policy_fwd(const float* mean_raw, const float* std_raw, const float* action,
                          float min_std, float max_std, int32_t R, int32_t A, float* entropy,
                          float* logp, void* stream);
int dv3_normal_policy_bwd(const float* mean_raw, const float* std_raw, const float* action,
                          const float* g_entropy, const float* g_logp, float min_std, float max_std,
                          int32_t R, int32_t A, float* d_mean_raw, float* d_std_raw, float* d_action,
                          void* stream);
/* tools.tensorstats (tools.py:949-958): out4 = {mean, unbiased std, min, max} of x[n] */
int dv3_tensorstats(const float* x, long long n, float* out4, void* stream);
/* Backward of RSSM.initial (networks.py:99-125): deter0 = tanh(W), stoch0 = mode(prior head(deter0))
 * with straight-through log-probs -- one row, evaluated once per observe call.  init_* are the
 * intermediates dv3_observe_fwd returned; g_norm [S*C] / g_deter0 [D] are dv3_observe_bwd's
 * d_init_stoch / d_init_deter.  The six parameter gradients are ADDED onto the given buffers (the
 * bulk dW sums of the same parameters, or zeros).  scratch: dv3_rssm_initial_bwd_scratch_floats. */
size_t dv3_rssm_initial_bwd_scratch_floats(const dv3_rssm_dims* d);
int dv3_rssm_initial_bwd(const dv3_rssm_dims* d, const dv3_rssm_params* p, const float* init_deter,
                         const float* init_ypre, const float* init_y, const float* init_logit,
                         const float* g_norm, const float* g_deter0, float* d_w_init, float* d_w_out,
                         float* d_ln_out_g, float* d_ln_out_b, float* d_w_ims, float* d_b_ims,
                         float* scratch, void* stream);
/* out[j] = sum_r x[r,j] (+ out[j] when accumulate != 0): the gradient of a Linear bias from the
 * deltas of its output rows (autograd of networks.py:640-655); row chunks meet through fp32 atomics */
int dv3_col_sum(const float* x, int32_t ld, int32_t M, int32_t n, float* out, int32_t accumulate,
                void* stream);
/* slow-critic update (models.py:683-689) over flat buffers: dst = mix * src + (1 - mix) * dst */
int dv3_ema_mix(float* dst, const float* src, long long n, double mix, void* stream);

/* ------------------------------------------------------------------------------------------
 * 4x4 stride-2 convolution / transposed convolution as GEMMs (dv3_conv.cu): the image encoder /
 * decoder of reference networks.py:448-585 on channels-last [pixels, C] matrices.
 * im2col: x [n,H,W,C] -> cols [n*H/2*W/2, 16*C], column (ky*4+kx)*C + c = x[n, 2oy-1+ky, 2ox-1+kx, c]
 * (zero outside): the patch matrix of Conv2dSamePad(k=4, s=2) (networks.py:771-799), written as fp32
 * (cols, optional) and / or as the tf32 hi / lo planes dv3_gemm_tc reads.
 * col2im: the adjoint (ConvTranspose2d(k=4, s=2, padding=1), networks.py:533-556): out [n,H,W,C] from
 * cols [n*H/2*W/2, 16*C], gathered (deterministic), plus bias[c] (optional) and a constant shift.
 * ---------------------------------------------------------------------------------------- */
int dv3_im2col_s2k4(const float* x, int32_t n, int32_t H, int32_t W, int32_t C, float* cols, float* hi,
                    float* lo, void* stream);
int dv3_col2im_s2k4(const float* cols, int32_t n, int32_t H, int32_t W, int32_t C, const float* bias,
                    float shift, float* out, void* stream);

/* Debug aid: with DV3_OBSERVE_TIMING=1 in the environment the persistent observe kernel stamps
 * %globaltimer (ns) at its 8 phase boundaries per step on CTA 0; this copies [T][8] stamps out. */
int dv3_debug_observe_timing(unsigned long long* host, int32_t T);
/* The same for the persistent imagination forward (DV3_IMAGINE_TIMING=1): [H][16] stamps of the
 * first CTA's epilogue at the phase boundaries of each step (start, actor trunk, head, img_in, GRU,
 * img_out, imgs_stat + draw). */
int dv3_debug_imagine_timing(unsigned long long* host, int32_t H);

#ifdef __cplusplus
}
#endif
#endif /* DV3_B200_H */
