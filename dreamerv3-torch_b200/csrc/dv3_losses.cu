// Loss-side kernels: lambda-return recurrence, symlog two-hot head, KL balance.
// All are HBM streams: each input element is read once, with coalesced accesses.
//
// Reference semantics (paths relative to the reference tree):
//   tools.lambda_return / static_scan_for_lambda_return      tools.py:682-728
//   tools.DiscDist.log_prob / mean / mode, symlog, symexp     tools.py:22-27, 463-513
//   RSSM.kl_loss over Independent(OneHotDist(unimix),1)       networks.py:272-290
#include "dv3_common.cuh"

namespace dv3 {

// ------------------------------------------------------------------------------------------
// lambda return.  Time-major [H,N]; thread n walks its column backwards.  A warp reads 32
// consecutive floats per time step -> fully coalesced.
//   inputs_t = r_t + c_t * v_{t+1} * (1-lambda);  R_t = inputs_t + c_t * lambda * R_{t+1}
// (the reference's own operation order, so fp32 results are bit-identical to it)
// ------------------------------------------------------------------------------------------
__global__ void lambda_return_fwd_kernel(const float* __restrict__ reward,
                                         const float* __restrict__ value,
                                         const float* __restrict__ pcont,
                                         const float* __restrict__ bootstrap, float lambda_,
                                         float oml, int H, int N, float* __restrict__ ret) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float boot = bootstrap[n];
  float last = boot, vnext = boot;
  for (int t = H - 1; t >= 0; --t) {
    const size_t o = (size_t)t * N + n;
    const float c = pcont[o];
    const float inp = __fadd_rn(reward[o], __fmul_rn(__fmul_rn(c, vnext), oml));
    last = __fadd_rn(inp, __fmul_rn(__fmul_rn(c, lambda_), last));
    ret[o] = last;
    vnext = value[o];
  }
}

// A_t = dL/dR_t (total) = g_t + c_{t-1} * lambda * A_{t-1}
__global__ void lambda_return_bwd_kernel(const float* __restrict__ value,
                                         const float* __restrict__ pcont,
                                         const float* __restrict__ bootstrap,
                                         const float* __restrict__ ret,
                                         const float* __restrict__ g_ret, float lambda_,
                                         float oml, int H, int N, float* __restrict__ d_reward,
                                         float* __restrict__ d_value, float* __restrict__ d_pcont,
                                         float* __restrict__ d_bootstrap) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float boot = bootstrap[n];
  float A = 0.f, cprev = 0.f;
  if (d_value) d_value[n] = 0.f;  // value[0] never enters the return
  float dboot = 0.f;
  for (int t = 0; t < H; ++t) {
    const size_t o = (size_t)t * N + n;
    const float c = pcont[o];
    A = g_ret[o] + cprev * lambda_ * A;
    const bool lastt = (t == H - 1);
    const float vnext = lastt ? boot : value[o + N];
    const float rnext = lastt ? boot : ret[o + N];
    if (d_reward) d_reward[o] = A;
    if (d_pcont) d_pcont[o] = A * (oml * vnext + lambda_ * rnext);
    const float dv = A * c * oml;
    if (lastt) {
      dboot = dv + A * c * lambda_;
    } else if (d_value) {
      d_value[o + N] = dv;
    }
    cprev = c;
  }
  if (d_bootstrap) d_bootstrap[n] = dboot;
}

// ------------------------------------------------------------------------------------------
// two-hot.  One warp per row of K (=255) logits: lane j owns elements j, j+32, ...
// ------------------------------------------------------------------------------------------
constexpr int TH_MAXPER = 8;  // K <= 256

__device__ __forceinline__ float symlogf_(float x) {
  const float s = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
  return s * logf(fabsf(x) + 1.f);
}
__device__ __forceinline__ float symexpf_(float x) {
  const float s = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
  return s * (expf(fabsf(x)) - 1.f);
}

struct TwoHot {
  int below, above;
  float w_below, w_above;
};

// x is already symlog-transformed
__device__ __forceinline__ TwoHot twohot_target(float x, const float* __restrict__ buckets, int K,
                                                int lane) {
  int le = 0, gt = 0;
  for (int k = lane; k < K; k += 32) {
    const float bk = buckets[k];
    le += (bk <= x);
    gt += (bk > x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    le += __shfl_xor_sync(FULL, le, o);
    gt += __shfl_xor_sync(FULL, gt, o);
  }
  TwoHot t;
  t.below = min(max(le - 1, 0), K - 1);
  t.above = min(max(K - gt, 0), K - 1);
  const bool equal = t.below == t.above;
  const float db = equal ? 1.f : fabsf(buckets[t.below] - x);
  const float da = equal ? 1.f : fabsf(buckets[t.above] - x);
  const float tot = db + da;
  t.w_below = da / tot;
  t.w_above = db / tot;
  return t;
}

__global__ void __launch_bounds__(256)
twohot_logprob_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ x,
                          const float* __restrict__ buckets, int R, int K,
                          float* __restrict__ logprob) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* row = logits + (size_t)r * K;
  float v[TH_MAXPER];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    const int k = lane + 32 * i;
    v[i] = (k < K) ? row[k] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) s += (lane + 32 * i < K) ? expf(v[i] - m) : 0.f;
  const float lse = m + logf(warp_sum(s));
  const TwoHot t = twohot_target(symlogf_(x[r]), buckets, K, lane);
  if (lane == 0) {
    // (target * log_pred).sum(): two (or one doubled) non-zero terms
    const float lb = row[t.below] - lse, la = row[t.above] - lse;
    logprob[r] = (t.below == t.above) ? (t.w_below + t.w_above) * lb
                                      : (t.w_below * lb + t.w_above * la);
  }
}

// d logits_k = g * (target_k - softmax_k * sum(target))
__global__ void __launch_bounds__(256)
twohot_logprob_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ x,
                          const float* __restrict__ buckets, const float* __restrict__ g, int R,
                          int K, float* __restrict__ d_logits) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* row = logits + (size_t)r * K;
  float v[TH_MAXPER];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    const int k = lane + 32 * i;
    v[i] = (k < K) ? row[k] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    v[i] = (lane + 32 * i < K) ? expf(v[i] - m) : 0.f;
    s += v[i];
  }
  const float inv = 1.f / warp_sum(s);
  const TwoHot t = twohot_target(symlogf_(x[r]), buckets, K, lane);
  const float gr = g[r];
  const float tsum = t.w_below + t.w_above;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    const int k = lane + 32 * i;
    if (k < K) {
      float tgt = 0.f;
      if (k == t.below) tgt += t.w_below;
      if (k == t.above) tgt += t.w_above;
      d_logits[(size_t)r * K + k] = gr * (tgt - v[i] * inv * tsum);
    }
  }
}

__global__ void __launch_bounds__(256)
twohot_mean_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ buckets, int R,
                       int K, float* __restrict__ mean) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* row = logits + (size_t)r * K;
  float v[TH_MAXPER];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    const int k = lane + 32 * i;
    v[i] = (k < K) ? row[k] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    v[i] = (lane + 32 * i < K) ? expf(v[i] - m) : 0.f;
    s += v[i];
  }
  const float tot = warp_sum(s);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    const int k = lane + 32 * i;
    if (k < K) acc += (v[i] / tot) * buckets[k];
  }
  acc = warp_sum(acc);
  if (lane == 0) mean[r] = symexpf_(acc);
}

// y = symexp(m), m = sum p_k b_k:  d l_k = g * exp|m| * p_k (b_k - m)   (0 at m == 0, as autograd
// gives for sign(m)*(exp|m|-1))
__global__ void __launch_bounds__(256)
twohot_mean_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ buckets,
                       const float* __restrict__ g, int R, int K, float* __restrict__ d_logits) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* row = logits + (size_t)r * K;
  float v[TH_MAXPER];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    const int k = lane + 32 * i;
    v[i] = (k < K) ? row[k] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    v[i] = (lane + 32 * i < K) ? expf(v[i] - m) : 0.f;
    s += v[i];
  }
  const float tot = warp_sum(s);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    const int k = lane + 32 * i;
    v[i] = v[i] / tot;
    if (k < K) acc += v[i] * buckets[k];
  }
  const float mu = warp_sum(acc);
  const float dydm = (mu == 0.f) ? 0.f : expf(fabsf(mu));
  const float gr = g[r] * dydm;
#pragma unroll
  for (int i = 0; i < TH_MAXPER; ++i) {
    const int k = lane + 32 * i;
    if (k < K) d_logits[(size_t)r * K + k] = gr * v[i] * (buckets[k] - mu);
  }
}

// ------------------------------------------------------------------------------------------
// KL balance.  One CTA per row (b,t), one warp per group (S <= 32 warps), lane = class.
// ------------------------------------------------------------------------------------------
struct Cat {
  float p, q, norm;  // softmax(l), unimixed probs, normalised log q
};

__device__ __forceinline__ Cat cat_of(float l, bool valid, int C, float unimix) {
  Cat o;
  const float m = warp_max(valid ? l : -INFINITY);
  const float e = valid ? expf(l - m) : 0.f;
  o.p = e / warp_sum(e);
  float lp = l;
  if (unimix > 0.f) lp = logf(o.p * (1.f - unimix) + unimix / (float)C);
  const float m2 = warp_max(valid ? lp : -INFINITY);
  const float s2 = warp_sum(valid ? expf(lp - m2) : 0.f);
  o.norm = lp - (m2 + logf(s2));
  const float m3 = warp_max(valid ? o.norm : -INFINITY);
  const float e3 = valid ? expf(o.norm - m3) : 0.f;
  o.q = e3 / warp_sum(e3);
  return o;
}

__global__ void __launch_bounds__(1024)
kl_balance_fwd_kernel(const float* __restrict__ post, const float* __restrict__ prior, int S, int C,
                      float unimix, float free_nats, float dyn_scale, float rep_scale,
                      float* __restrict__ loss, float* __restrict__ value,
                      float* __restrict__ dyn, float* __restrict__ rep,
                      float* __restrict__ post_ent, float* __restrict__ prior_ent) {
  __shared__ float part[3][32];
  const int lane = threadIdx.x & 31, s = threadIdx.x >> 5;
  const int r = blockIdx.x;
  const bool valid = lane < C;
  const size_t o = ((size_t)r * S + s) * C + lane;
  const Cat P = cat_of(valid ? post[o] : 0.f, valid, C, unimix);
  const Cat Q = cat_of(valid ? prior[o] : 0.f, valid, C, unimix);
  // torch _kl_categorical_categorical: t = p*(logp-logq); q==0 -> inf; p==0 -> 0
  float t = 0.f, ep = 0.f, eq = 0.f;
  if (valid) {
    t = P.q * (P.norm - Q.norm);
    if (Q.q == 0.f) t = INFINITY;
    if (P.q == 0.f) t = 0.f;
    ep = fmaxf(P.norm, -3.402823466e38f) * P.q;
    eq = fmaxf(Q.norm, -3.402823466e38f) * Q.q;
  }
  t = warp_sum(t); ep = warp_sum(ep); eq = warp_sum(eq);
  if (lane == 0) { part[0][s] = t; part[1][s] = ep; part[2][s] = eq; }
  __syncthreads();
  if (s == 0) {
    float kl = warp_sum(lane < S ? part[0][lane] : 0.f);
    float hp = warp_sum(lane < S ? part[1][lane] : 0.f);
    float hq = warp_sum(lane < S ? part[2][lane] : 0.f);
    if (lane == 0) {
      const float d = fmaxf(kl, free_nats);
      if (value) value[r] = kl;
      if (dyn) dyn[r] = d;
      if (rep) rep[r] = d;
      if (loss) loss[r] = dyn_scale * d + rep_scale * d;
      if (post_ent) post_ent[r] = -hp;
      if (prior_ent) prior_ent[r] = -hq;
    }
  }
}

// rep: d/d post logits of KL(P||sg Q);  dyn: d/d prior logits of KL(sg P||Q); each passes the
// clip only where KL >= free (torch.clip's subgradient).
//   dKL/dP_k = logP_k - logQ_k (+1, cancels);  dKL/dQ_k = -P_k/Q_k
//   chain through q = (1-r) softmax(l) + r/C:  d l_m = (1-r) p_m (gq_m - <gq,p>)
__global__ void __launch_bounds__(1024)
kl_balance_bwd_kernel(const float* __restrict__ post, const float* __restrict__ prior,
                      const float* __restrict__ g_loss, int S, int C, float unimix,
                      float free_nats, float dyn_scale, float rep_scale,
                      float* __restrict__ d_post, float* __restrict__ d_prior) {
  __shared__ float part[32];
  __shared__ float kl_sh;
  const int lane = threadIdx.x & 31, s = threadIdx.x >> 5;
  const int r = blockIdx.x;
  const bool valid = lane < C;
  const size_t o = ((size_t)r * S + s) * C + lane;
  const Cat P = cat_of(valid ? post[o] : 0.f, valid, C, unimix);
  const Cat Q = cat_of(valid ? prior[o] : 0.f, valid, C, unimix);
  float t = 0.f;
  if (valid) {
    t = P.q * (P.norm - Q.norm);
    if (Q.q == 0.f) t = INFINITY;
    if (P.q == 0.f) t = 0.f;
  }
  t = warp_sum(t);
  if (lane == 0) part[s] = t;
  __syncthreads();
  if (s == 0) {
    const float kl = warp_sum(lane < S ? part[lane] : 0.f);
    if (lane == 0) kl_sh = kl;
  }
  __syncthreads();
  const float pass = (kl_sh >= free_nats) ? g_loss[r] : 0.f;
  // rep -> post logits
  {
    const float gq = valid ? (P.norm - Q.norm) : 0.f;
    const float dot = warp_sum(valid ? gq * P.p : 0.f);
    if (valid && d_post) d_post[o] = pass * rep_scale * (1.f - unimix) * P.p * (gq - dot);
  }
  // dyn -> prior logits
  {
    const float gq = valid ? -(P.q / Q.q) : 0.f;
    const float dot = warp_sum(valid ? gq * Q.p : 0.f);
    if (valid && d_prior) d_prior[o] = pass * dyn_scale * (1.f - unimix) * Q.p * (gq - dot);
  }
}

}  // namespace dv3

using namespace dv3;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int dv3_lambda_return_fwd(const float* reward, const float* value, const float* pcont,
                                     const float* bootstrap, double lambda_, int32_t H, int32_t N,
                                     float* ret, void* stream) {
  DV3_REQUIRE(H >= 0 && N >= 0, DV3_ERR_BAD_SHAPE, "lambda_return_fwd: H=%d N=%d", H, N);
  if (H == 0 || N == 0) return 0;  // empty horizon / batch: nothing to write
  DV3_REQUIRE(reward && value && pcont && bootstrap && ret, DV3_ERR_NULL,
              "lambda_return_fwd: null pointer");
  lambda_return_fwd_kernel<<<(N + 127) / 128, 128, 0, ST(stream)>>>(reward, value, pcont,
                                                                    bootstrap, (float)lambda_,
                                                                    (float)(1.0 - lambda_), H, N,
                                                                    ret);
  DV3_CHECK_LAUNCH("lambda_return_fwd_kernel");
  return 0;
}

extern "C" int dv3_lambda_return_bwd(const float* value, const float* pcont,
                                     const float* bootstrap, const float* ret, const float* g_ret,
                                     double lambda_, int32_t H, int32_t N, float* d_reward,
                                     float* d_value, float* d_pcont, float* d_bootstrap,
                                     void* stream) {
  DV3_REQUIRE(H >= 0 && N >= 0, DV3_ERR_BAD_SHAPE, "lambda_return_bwd: H=%d N=%d", H, N);
  if (H == 0 || N == 0) return 0;
  DV3_REQUIRE(value && pcont && bootstrap && ret && g_ret, DV3_ERR_NULL,
              "lambda_return_bwd: null pointer");
  lambda_return_bwd_kernel<<<(N + 127) / 128, 128, 0, ST(stream)>>>(
      value, pcont, bootstrap, ret, g_ret, (float)lambda_, (float)(1.0 - lambda_), H, N, d_reward,
      d_value, d_pcont, d_bootstrap);
  DV3_CHECK_LAUNCH("lambda_return_bwd_kernel");
  return 0;
}

static int twohot_check(const void* a, const void* b, const void* c, int R, int K,
                        const char* who) {
  DV3_REQUIRE(R >= 0 && K >= 2 && K <= 32 * TH_MAXPER, DV3_ERR_BAD_SHAPE, "%s: R=%d K=%d (K<=256)",
              who, R, K);
  DV3_REQUIRE(R == 0 || (a && b && c), DV3_ERR_NULL, "%s: null pointer", who);
  return 0;
}

extern "C" int dv3_twohot_logprob_fwd(const float* logits, const float* x, const float* buckets,
                                      int32_t R, int32_t K, float* logprob, void* stream) {
  DV3_TRY(twohot_check(logits, x, buckets, R, K, "twohot_logprob_fwd"));
  if (R == 0) return 0;
  DV3_REQUIRE(logprob, DV3_ERR_NULL, "twohot_logprob_fwd: null output");
  twohot_logprob_fwd_kernel<<<(R + 7) / 8, 256, 0, ST(stream)>>>(logits, x, buckets, R, K, logprob);
  DV3_CHECK_LAUNCH("twohot_logprob_fwd_kernel");
  return 0;
}

extern "C" int dv3_twohot_logprob_bwd(const float* logits, const float* x, const float* buckets,
                                      const float* g_logprob, int32_t R, int32_t K,
                                      float* d_logits, void* stream) {
  DV3_TRY(twohot_check(logits, x, buckets, R, K, "twohot_logprob_bwd"));
  if (R == 0) return 0;
  DV3_REQUIRE(g_logprob && d_logits, DV3_ERR_NULL, "twohot_logprob_bwd: null pointer");
  twohot_logprob_bwd_kernel<<<(R + 7) / 8, 256, 0, ST(stream)>>>(logits, x, buckets, g_logprob, R,
                                                                 K, d_logits);
  DV3_CHECK_LAUNCH("twohot_logprob_bwd_kernel");
  return 0;
}

extern "C" int dv3_twohot_mean_fwd(const float* logits, const float* buckets, int32_t R, int32_t K,
                                   float* mean, void* stream) {
  DV3_TRY(twohot_check(logits, buckets, mean, R, K, "twohot_mean_fwd"));
  if (R == 0) return 0;
  twohot_mean_fwd_kernel<<<(R + 7) / 8, 256, 0, ST(stream)>>>(logits, buckets, R, K, mean);
  DV3_CHECK_LAUNCH("twohot_mean_fwd_kernel");
  return 0;
}

extern "C" int dv3_twohot_mean_bwd(const float* logits, const float* buckets, const float* g_mean,
                                   int32_t R, int32_t K, float* d_logits, void* stream) {
  DV3_TRY(twohot_check(logits, buckets, g_mean, R, K, "twohot_mean_bwd"));
  if (R == 0) return 0;
  DV3_REQUIRE(d_logits, DV3_ERR_NULL, "twohot_mean_bwd: null output");
  twohot_mean_bwd_kernel<<<(R + 7) / 8, 256, 0, ST(stream)>>>(logits, buckets, g_mean, R, K,
                                                              d_logits);
  DV3_CHECK_LAUNCH("twohot_mean_bwd_kernel");
  return 0;
}

extern "C" int dv3_kl_balance_fwd(const float* post_logit, const float* prior_logit, int32_t R,
                                  int32_t S, int32_t C, float unimix, float free_nats,
                                  float dyn_scale, float rep_scale, float* loss, float* value,
                                  float* dyn, float* rep, float* post_ent, float* prior_ent,
                                  void* stream) {
  DV3_REQUIRE(R >= 0 && S >= 1 && S <= 32 && C >= 1 && C <= 32, DV3_ERR_BAD_SHAPE,
              "kl_balance_fwd: R=%d S=%d C=%d (S,C <= 32)", R, S, C);
  if (R == 0) return 0;
  DV3_REQUIRE(post_logit && prior_logit, DV3_ERR_NULL, "kl_balance_fwd: null pointer");
  kl_balance_fwd_kernel<<<R, S * 32, 0, ST(stream)>>>(post_logit, prior_logit, S, C, unimix,
                                                      free_nats, dyn_scale, rep_scale, loss, value,
                                                      dyn, rep, post_ent, prior_ent);
  DV3_CHECK_LAUNCH("kl_balance_fwd_kernel");
  return 0;
}

extern "C" int dv3_kl_balance_bwd(const float* post_logit, const float* prior_logit,
                                  const float* g_loss, int32_t R, int32_t S, int32_t C,
                                  float unimix, float free_nats, float dyn_scale, float rep_scale,
                                  float* d_post_logit, float* d_prior_logit, void* stream) {
  DV3_REQUIRE(R >= 0 && S >= 1 && S <= 32 && C >= 1 && C <= 32, DV3_ERR_BAD_SHAPE,
              "kl_balance_bwd: R=%d S=%d C=%d (S,C <= 32)", R, S, C);
  if (R == 0) return 0;
  DV3_REQUIRE(post_logit && prior_logit && g_loss, DV3_ERR_NULL, "kl_balance_bwd: null pointer");
  kl_balance_bwd_kernel<<<R, S * 32, 0, ST(stream)>>>(post_logit, prior_logit, g_loss, S, C,
                                                      unimix, free_nats, dyn_scale, rep_scale,
                                                      d_post_logit, d_prior_logit);
  DV3_CHECK_LAUNCH("kl_balance_bwd_kernel");
  return 0;
}
