// Step-tail kernels: the element-wise / reduction chains around the rollouts that the reference
// spells as dozens of tiny torch ops per call (and autograd doubles in backward).  Each is one
// launch here:
//   symlog                                   tools.py:22-23        (encoder input, networks.py:333)
//   squared-error log-probs                  tools.MSEDist 520-543, tools.SymlogDist 546-572
//   Bernoulli log-prob                       tools.py:604-628      (cont head, models.py:137-147)
//   weighted mean of per-row losses          models.py:140-152     (model_loss)
//   discount / cumulative weights            models.py:620-638     (_compute_target)
//   RewardEMA: 5/95 % quantiles + EMA        models.py:11-26
//   actor loss, value loss                   models.py:640-681, 419-429
//   normal-policy entropy / log-prob         networks.py:693-700, tools.py:575-601
//   tensorstats                              tools.py:949-958
// All are HBM streams over a few hundred KB; what they buy is launch count (the whole step is one
// CUDA graph whose length is the sum of its kernels' latencies).
#include "dv3_common.cuh"

namespace dv3 {

constexpr int TL_THREADS = 256;

__device__ __forceinline__ float symlog_(float x) {
  const float s = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
  return s * logf(fabsf(x) + 1.f);
}
// torch.nn.functional.softplus (beta 1, threshold 20)
__device__ __forceinline__ float softplus_(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_grad_(float x) { return x > 20.f ? 1.f : sigmoidf_(x); }

__global__ void symlog_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = symlog_(x[i]);
}

// logprob[r] = -sum_j d_j,  d_j = (mode - t(value))^2,  t = symlog or identity; symlog form drops
// d_j < tol (tools.py:558-563).  One warp per row.
__global__ void __launch_bounds__(TL_THREADS)
sqerr_logprob_fwd_kernel(const float* __restrict__ mode, const float* __restrict__ value, int R, int n,
                         int use_symlog, float tol, float* __restrict__ logprob) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (TL_THREADS >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* m = mode + (size_t)r * n;
  const float* v = value + (size_t)r * n;
  float acc = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float t = use_symlog ? symlog_(v[j]) : v[j];
    const float e = m[j] - t;
    float d = e * e;
    if (use_symlog && d < tol) d = 0.f;
    acc += d;
  }
  acc = warp_sum(acc);
  if (lane == 0) logprob[r] = -acc;
}

__global__ void __launch_bounds__(TL_THREADS)
sqerr_logprob_bwd_kernel(const float* __restrict__ mode, const float* __restrict__ value,
                         const float* __restrict__ g, int R, int n, int use_symlog, float tol,
                         float* __restrict__ d_mode) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (TL_THREADS >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* m = mode + (size_t)r * n;
  const float* v = value + (size_t)r * n;
  const float gr = g[r];
  for (int j = lane; j < n; j += 32) {
    const float t = use_symlog ? symlog_(v[j]) : v[j];
    const float e = m[j] - t;
    const bool drop = use_symlog && (e * e < tol);
    d_mode[(size_t)r * n + j] = drop ? 0.f : -2.f * e * gr;
  }
}

__global__ void bernoulli_logprob_fwd_kernel(const float* __restrict__ logit,
                                             const float* __restrict__ x, long long n,
                                             float* __restrict__ logprob) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float l = logit[i], xv = x[i];
  logprob[i] = -softplus_(l) * (1.f - xv) - softplus_(-l) * xv;
}

__global__ void bernoulli_logprob_bwd_kernel(const float* __restrict__ logit,
                                             const float* __restrict__ x,
                                             const float* __restrict__ g, long long n,
                                             float* __restrict__ d_logit) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float l = logit[i], xv = x[i];
  d_logit[i] = g[i] * (-softplus_grad_(l) * (1.f - xv) + softplus_grad_(-l) * xv);
}

// out[0] = (1/R) sum_r sum_i scale_i * loss_i[r]; one block, fixed order.
struct LossPtrs {
  const float* p[8];
  float scale[8];
  int n;
};

__global__ void __launch_bounds__(1024)
loss_mean_kernel(LossPtrs lp, int R, float* __restrict__ out, float* __restrict__ neg_out) {
  __shared__ float red[4 * 32];
  float acc[1] = {0.f};
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < lp.n; ++i) {
      const float v = lp.p[i][r];
      s += lp.scale[i] * v;
      if (neg_out) neg_out[(size_t)i * R + r] = -v;
    }
    acc[0] += s;
  }
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) out[0] = acc[0] / (float)R;
}

// grads[i][r] = g[0] * scale_i / R   (grads is [n, R])
__global__ void loss_mean_bwd_kernel(const float* __restrict__ g, LossPtrs lp, int R,
                                     float* __restrict__ grads) {
  const int i = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) grads[(size_t)i * R + r] = g[0] * lp.scale[i] / (float)R;
}

// discount[t,n] = gamma * sigmoid(logit[t,n]); weights[0] = 1, weights[t] = prod_{i<t} discount[i]
__global__ void discount_weights_kernel(const float* __restrict__ logit, float gamma, int H, int N,
                                        float* __restrict__ discount, float* __restrict__ weights) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float w = 1.f;
  for (int t = 0; t < H; ++t) {
    const size_t o = (size_t)t * N + n;
    const float d = gamma * sigmoidf_(logit[o]);
    discount[o] = d;
    weights[o] = w;
    w = w * d;
  }
}

__global__ void discount_bwd_kernel(const float* __restrict__ logit, const float* __restrict__ g,
                                    float gamma, long long n, float* __restrict__ d_logit) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = sigmoidf_(logit[i]);
  d_logit[i] = g[i] * gamma * s * (1.f - s);
}

// ---- RewardEMA (models.py:11-26): torch.quantile('linear') at 5 % / 95 % + EMA -------------------
// torch sorts all n values (several radix-sort launches) and gathers two neighbours per quantile.
// Only four order statistics are needed, so one CTA selects them by three-level radix select on
// the monotonic integer image of the floats (11 + 11 + 10 bits): a histogram pass over the values
// per level (they sit in L2), a warp-parallel prefix scan per wanted rank.  Exact, like a sort.
constexpr int EMA_MAX = 1 << 20;
constexpr int EMA_BINS = 2048;

__device__ __forceinline__ unsigned ema_key(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ema_unkey(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// warp w finds, in hist[w * stride .. + nbins), the bin holding the element of (0-based) rank
// rk[w] and the rank left inside that bin
__device__ __forceinline__ void ema_pick(const unsigned* hist, int nbins, unsigned rank, int lane,
                                         unsigned* bin_out, unsigned* rank_out) {
  const int per = nbins / 32;
  unsigned mine = 0;
  for (int i = 0; i < per; ++i) mine += hist[lane * per + i];
  unsigned incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned excl = incl - mine;
  const bool here = rank >= excl && rank < incl;
  const unsigned who = __ballot_sync(FULL, here);
  const int src = __ffs(who) - 1;                    // exactly one lane (rank < n)
  unsigned bin = 0, left = 0;
  if (lane == src) {
    unsigned acc = excl;
    for (int i = 0; i < per; ++i) {
      const unsigned c = hist[lane * per + i];
      if (rank < acc + c) { bin = lane * per + i; left = rank - acc; break; }
      acc += c;
    }
  }
  *bin_out = __shfl_sync(FULL, bin, src);
  *rank_out = __shfl_sync(FULL, left, src);
}

__global__ void __launch_bounds__(1024)
reward_ema_kernel(const float* __restrict__ x, int n, float alpha, float one_minus_alpha,
                  float* __restrict__ ema, float* __restrict__ out) {
  __shared__ unsigned hist[4 * EMA_BINS];
  __shared__ unsigned pref[4], left[4];
  __shared__ float val[4], res[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // wanted ranks: floor / ceil of q (n-1) for q = 0.05, 0.95 (rank computed in fp32 like torch)
  unsigned want[4];
  float wgt[2];
#pragma unroll
  for (int qi = 0; qi < 2; ++qi) {
    const float q = qi == 0 ? 0.05f : 0.95f;
    const float rank = __fmul_rn(q, (float)(n - 1));
    const float lo = floorf(rank);
    want[2 * qi] = (unsigned)lo;
    want[2 * qi + 1] = (unsigned)ceilf(rank);
    wgt[qi] = __fsub_rn(rank, lo);
  }
  // level 0: one histogram of the top 11 bits serves all four ranks
  for (int i = tid; i < EMA_BINS; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) atomicAdd(&hist[ema_key(x[i]) >> 21], 1u);
  __syncthreads();
  if (warp < 4) {
    unsigned b, l;
    ema_pick(hist, EMA_BINS, want[warp], lane, &b, &l);
    if (lane == 0) { pref[warp] = b; left[warp] = l; }
  }
  __syncthreads();
  // level 1: next 11 bits among the values sharing each rank's top bits
  for (int i = tid; i < 4 * EMA_BINS; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  {
    const unsigned p0 = pref[0], p1 = pref[1], p2 = pref[2], p3 = pref[3];
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned k = ema_key(x[i]);
      const unsigned top = k >> 21, mid = (k >> 10) & 2047u;
      if (top == p0) atomicAdd(&hist[mid], 1u);
      if (top == p1) atomicAdd(&hist[EMA_BINS + mid], 1u);
      if (top == p2) atomicAdd(&hist[2 * EMA_BINS + mid], 1u);
      if (top == p3) atomicAdd(&hist[3 * EMA_BINS + mid], 1u);
    }
  }
  __syncthreads();
  if (warp < 4) {
    unsigned b, l;
    ema_pick(hist + warp * EMA_BINS, EMA_BINS, left[warp], lane, &b, &l);
    __syncwarp();
    if (lane == 0) { pref[warp] = (pref[warp] << 11) | b; left[warp] = l; }
  }
  __syncthreads();
  // level 2: the last 10 bits
  for (int i = tid; i < 4 * EMA_BINS; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  {
    const unsigned p0 = pref[0], p1 = pref[1], p2 = pref[2], p3 = pref[3];
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned k = ema_key(x[i]);
      const unsigned top = k >> 10, low = k & 1023u;
      if (top == p0) atomicAdd(&hist[low], 1u);
      if (top == p1) atomicAdd(&hist[EMA_BINS + low], 1u);
      if (top == p2) atomicAdd(&hist[2 * EMA_BINS + low], 1u);
      if (top == p3) atomicAdd(&hist[3 * EMA_BINS + low], 1u);
    }
  }
  __syncthreads();
  if (warp < 4) {
    unsigned b, l;
    ema_pick(hist + warp * EMA_BINS, 1024, left[warp], lane, &b, &l);
    if (lane == 0) val[warp] = ema_unkey((pref[warp] << 10) | b);
  }
  __syncthreads();
  if (tid < 2) {
    // ATen's lerp: a + w (b - a) for w < 0.5, else b - (b - a)(1 - w)
    const float a = val[2 * tid], b = val[2 * tid + 1], w = wgt[tid];
    const float diff = __fsub_rn(b, a);
    const float qv = (w < 0.5f) ? __fadd_rn(a, __fmul_rn(w, diff))
                                : __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.f, w)));
    const float e = __fadd_rn(__fmul_rn(alpha, qv), __fmul_rn(one_minus_alpha, ema[tid]));
    ema[tid] = e;
    res[tid] = e;
  }
  __syncthreads();
  if (tid == 0) {
    out[0] = res[0];                                  // offset
    out[1] = fmaxf(__fsub_rn(res[1], res[0]), 1.f);   // scale = clip(hi - lo, min=1)
  }
}

// ---- actor loss (models.py:640-681 + 393-397) ---------------------------------------------------
// mode 0 ('dynamics'): term = -w * ((tgt-off)/scale - (base-off)/scale) - c_ent * ent
// mode 1 ('reinforce'): term = -w * logp * (tgt - base)            - c_ent * ent
// loss = mean over the Hm*N terms; normed[t,n] = (tgt-off)/scale is written for the metrics.
// os = {offset, scale} on the device (NULL: offset 0, scale 1 = reward_EMA off).
__global__ void __launch_bounds__(1024)
actor_loss_fwd_kernel(const float* __restrict__ tgt, const float* __restrict__ base,
                      const float* __restrict__ w, const float* __restrict__ ent,
                      const float* __restrict__ logp, const float* __restrict__ os, float c_ent,
                      int mode, int cnt, float* __restrict__ normed, float* __restrict__ loss) {
  __shared__ float red[4 * 32];
  const float off = os ? os[0] : 0.f, sc = os ? os[1] : 1.f;
  float acc[1] = {0.f};
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const float nt = (tgt[i] - off) / sc;
    if (normed) normed[i] = nt;
    float at;
    if (mode == 0) at = nt - (base[i] - off) / sc;
    else at = logp[i] * (tgt[i] - base[i]);
    acc[0] += -w[i] * at - c_ent * ent[i];
  }
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) loss[0] = acc[0] / (float)cnt;
}

__global__ void actor_loss_bwd_kernel(const float* __restrict__ g, const float* __restrict__ tgt,
                                      const float* __restrict__ base, const float* __restrict__ w,
                                      const float* __restrict__ os, float c_ent, int mode, int cnt,
                                      int total, float* __restrict__ d_tgt, float* __restrict__ d_ent,
                                      float* __restrict__ d_logp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;                    // total = H*N >= cnt = (H-1)*N
  const float gs = g[0] / (float)cnt;
  const bool in = i < cnt;
  d_ent[i] = in ? -c_ent * gs : 0.f;
  if (mode == 0) {
    const float sc = os ? os[1] : 1.f;
    if (in) d_tgt[i] = -w[i] * gs / sc;
  } else {
    d_logp[i] = in ? -w[i] * (tgt[i] - base[i]) * gs : 0.f;
  }
}

// value loss (models.py:419-429): mean_i w_i * (-(lp_target_i) - lp_slow_i)
__global__ void __launch_bounds__(1024)
value_loss_fwd_kernel(const float* __restrict__ lp1, const float* __restrict__ lp2,
                      const float* __restrict__ w, int cnt, float* __restrict__ loss) {
  __shared__ float red[4 * 32];
  float acc[1] = {0.f};
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    float v = -lp1[i];
    if (lp2) v -= lp2[i];
    acc[0] += w[i] * v;
  }
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) loss[0] = acc[0] / (float)cnt;
}

__global__ void value_loss_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                      int cnt, float* __restrict__ d_lp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) d_lp[i] = -w[i] * g[0] / (float)cnt;
}

// ---- 'normal' actor distribution (networks.py:693-700 + tools.ContDist) -------------------------
// mean = tanh(mr); std = (max-min) sigmoid(sr + 2) + min
// ent[r]  = sum_a 0.5 + 0.5 log(2 pi) + log std
// logp[r] = sum_a -(x-mean)^2 / (2 std^2) - log std - log sqrt(2 pi)
__global__ void normal_policy_fwd_kernel(const float* __restrict__ mr, const float* __restrict__ sr,
                                         const float* __restrict__ x, float min_std, float max_std,
                                         int R, int A, float* __restrict__ ent,
                                         float* __restrict__ logp) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float e = 0.f, lp = 0.f;
  for (int a = 0; a < A; ++a) {
    const size_t o = (size_t)r * A + a;
    const float mean = tanhf(mr[o]);
    const float std = (max_std - min_std) * sigmoidf_(sr[o] + 2.f) + min_std;
    const float ls = logf(std);
    e += 0.5f + 0.9189385332046727f + ls;             // 0.5 log(2 pi)
    const float dx = x[o] - mean;
    lp += -(dx * dx) / (2.f * std * std) - ls - 0.9189385332046727f;
  }
  ent[r] = e;
  if (logp) logp[r] = lp;
}

__global__ void normal_policy_bwd_kernel(const float* __restrict__ mr, const float* __restrict__ sr,
                                         const float* __restrict__ x, const float* __restrict__ g_ent,
                                         const float* __restrict__ g_logp, float min_std,
                                         float max_std, int R, int A, float* __restrict__ d_mr,
                                         float* __restrict__ d_sr, float* __restrict__ d_x) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * A) return;
  const int r = (int)(i / A);
  const float mean = tanhf(mr[i]);
  const float sg = sigmoidf_(sr[i] + 2.f);
  const float std = (max_std - min_std) * sg + min_std;
  const float ge = g_ent ? g_ent[r] : 0.f, gl = g_logp ? g_logp[r] : 0.f;
  const float dx = x[i] - mean;
  const float var = std * std;
  const float d_mean = gl * dx / var;
  const float d_std = ge / std + gl * (dx * dx / (var * std) - 1.f / std);
  d_mr[i] = d_mean * (1.f - mean * mean);
  d_sr[i] = d_std * (max_std - min_std) * sg * (1.f - sg);
  if (d_x) d_x[i] = -gl * dx / var;
}

// ---- tensorstats (tools.py:949-958): mean, std (unbiased), min, max; one block ------------------
__global__ void __launch_bounds__(1024)
tensorstats_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float red[4 * 32];
  __shared__ float smin[32], smax[32];
  float acc[1] = {0.f};
  float mn = INFINITY, mx = -INFINITY;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = x[i];
    acc[0] += v;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  block_sum<1>(acc, red);
  const float mean = acc[0] / (float)n;
  float q[1] = {0.f};
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = x[i] - mean;
    q[0] = fmaf(d, d, q[0]);
  }
  block_sum<1>(q, red);
  mn = -warp_max(-mn);
  mx = warp_max(mx);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { smin[wid] = mn; smax[wid] = mx; }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    mn = lane < nw ? smin[lane] : INFINITY;
    mx = lane < nw ? smax[lane] : -INFINITY;
    mn = -warp_max(-mn);
    mx = warp_max(mx);
    if (lane == 0) {
      out[0] = mean;
      out[1] = sqrtf(q[0] / (float)(n > 1 ? n - 1 : 1));
      out[2] = mn;
      out[3] = mx;
    }
  }
}

// param = (1 - mix) * param + mix * src  over a flat buffer (slow critic, models.py:683-689)
__global__ void ema_mix_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n,
                               float mix, float one_minus_mix) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __fadd_rn(__fmul_rn(mix, src[i]), __fmul_rn(one_minus_mix, dst[i]));
}

// out[j] += sum_r x[r, j]: a Linear bias gradient.  Block = 32 columns x 8 row-strided warps over
// one chunk of rows; the chunks of a column meet through fp32 atomics (out is pre-zeroed by the
// caller or by the launcher).
__global__ void __launch_bounds__(256)
col_sum_kernel(const float* __restrict__ x, int ld, int M, int n, int rows_per_block,
               float* __restrict__ out) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float acc = 0.f;
  if (j < n)
    for (int r = r0 + warp; r < r1; r += 8) acc += x[(size_t)r * ld + j];
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && j < n) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w][lane];
    atomicAdd(out + j, s);
  }
}

// narrow matrices (n <= 8, e.g. the 3 image channels): one thread strides the flat array, so that
// the loads stay coalesced; lanes holding the same column meet in shared memory
__global__ void __launch_bounds__(256)
col_sum_narrow_kernel(const float* __restrict__ x, long long total, int n, float* __restrict__ out) {
  __shared__ float part[8];
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % n);
    const float v = x[e];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += (k == c) ? v : 0.f;
  }
  if (threadIdx.x < 8) part[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float s = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0 && k < n) atomicAdd(&part[k], s);
  }
  __syncthreads();
  if (threadIdx.x < n) atomicAdd(out + threadIdx.x, part[threadIdx.x]);
}

static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

}  // namespace dv3

using namespace dv3;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int dv3_symlog(const float* x, long long n, float* out, void* stream) {
  if (n <= 0) return 0;
  DV3_REQUIRE(x && out, DV3_ERR_NULL, "symlog: null pointer");
  symlog_kernel<<<nblk(n, 256), 256, 0, ST>>>(x, n, out);
  DV3_CHECK_LAUNCH("symlog_kernel");
  return 0;
}

extern "C" int dv3_sqerr_logprob_fwd(const float* mode, const float* value, int32_t R, int32_t n,
                                     int32_t use_symlog, float tol, float* logprob, void* stream) {
  if (R <= 0) return 0;
  DV3_REQUIRE(mode && value && logprob && n > 0, DV3_ERR_NULL, "sqerr_logprob_fwd: null / n=%d", n);
  sqerr_logprob_fwd_kernel<<<nblk(R, TL_THREADS / 32), TL_THREADS, 0, ST>>>(mode, value, R, n,
                                                                             use_symlog, tol, logprob);
  DV3_CHECK_LAUNCH("sqerr_logprob_fwd_kernel");
  return 0;
}

extern "C" int dv3_sqerr_logprob_bwd(const float* mode, const float* value, const float* g_logprob,
                                     int32_t R, int32_t n, int32_t use_symlog, float tol,
                                     float* d_mode, void* stream) {
  if (R <= 0) return 0;
  DV3_REQUIRE(mode && value && g_logprob && d_mode && n > 0, DV3_ERR_NULL, "sqerr_logprob_bwd: null");
  sqerr_logprob_bwd_kernel<<<nblk(R, TL_THREADS / 32), TL_THREADS, 0, ST>>>(
      mode, value, g_logprob, R, n, use_symlog, tol, d_mode);
  DV3_CHECK_LAUNCH("sqerr_logprob_bwd_kernel");
  return 0;
}

extern "C" int dv3_bernoulli_logprob_fwd(const float* logit, const float* x, long long n,
                                         float* logprob, void* stream) {
  if (n <= 0) return 0;
  DV3_REQUIRE(logit && x && logprob, DV3_ERR_NULL, "bernoulli_logprob_fwd: null pointer");
  bernoulli_logprob_fwd_kernel<<<nblk(n, 256), 256, 0, ST>>>(logit, x, n, logprob);
  DV3_CHECK_LAUNCH("bernoulli_logprob_fwd_kernel");
  return 0;
}

extern "C" int dv3_bernoulli_logprob_bwd(const float* logit, const float* x, const float* g,
                                         long long n, float* d_logit, void* stream) {
  if (n <= 0) return 0;
  DV3_REQUIRE(logit && x && g && d_logit, DV3_ERR_NULL, "bernoulli_logprob_bwd: null pointer");
  bernoulli_logprob_bwd_kernel<<<nblk(n, 256), 256, 0, ST>>>(logit, x, g, n, d_logit);
  DV3_CHECK_LAUNCH("bernoulli_logprob_bwd_kernel");
  return 0;
}

extern "C" int dv3_loss_mean_fwd(const float* const* losses, const float* scales, int32_t n_losses,
                                 int32_t R, float* out, float* neg_out, void* stream) {
  DV3_REQUIRE(losses && scales && out && n_losses >= 1 && n_losses <= 8 && R > 0, DV3_ERR_BAD_SHAPE,
              "loss_mean_fwd: n_losses=%d R=%d", n_losses, R);
  LossPtrs lp{};
  lp.n = n_losses;
  for (int i = 0; i < n_losses; ++i) {
    DV3_REQUIRE(losses[i], DV3_ERR_NULL, "loss_mean_fwd: loss %d is NULL", i);
    lp.p[i] = losses[i];
    lp.scale[i] = scales[i];
  }
  loss_mean_kernel<<<1, 1024, 0, ST>>>(lp, R, out, neg_out);
  DV3_CHECK_LAUNCH("loss_mean_kernel");
  return 0;
}

extern "C" int dv3_loss_mean_bwd(const float* g, const float* scales, int32_t n_losses, int32_t R,
                                 float* grads, void* stream) {
  DV3_REQUIRE(g && scales && grads && n_losses >= 1 && n_losses <= 8 && R > 0, DV3_ERR_BAD_SHAPE,
              "loss_mean_bwd: n_losses=%d R=%d", n_losses, R);
  LossPtrs lp{};
  lp.n = n_losses;
  for (int i = 0; i < n_losses; ++i) lp.scale[i] = scales[i];
  loss_mean_bwd_kernel<<<dim3(nblk(R, 256), n_losses), 256, 0, ST>>>(g, lp, R, grads);
  DV3_CHECK_LAUNCH("loss_mean_bwd_kernel");
  return 0;
}

extern "C" int dv3_discount_weights_fwd(const float* cont_logit, float gamma, int32_t H, int32_t N,
                                        float* discount, float* weights, void* stream) {
  if (H <= 0 || N <= 0) return 0;
  DV3_REQUIRE(cont_logit && discount && weights, DV3_ERR_NULL, "discount_weights_fwd: null pointer");
  discount_weights_kernel<<<nblk(N, 128), 128, 0, ST>>>(cont_logit, gamma, H, N, discount, weights);
  DV3_CHECK_LAUNCH("discount_weights_kernel");
  return 0;
}

extern "C" int dv3_discount_bwd(const float* cont_logit, const float* g_discount, float gamma,
                                long long n, float* d_logit, void* stream) {
  if (n <= 0) return 0;
  DV3_REQUIRE(cont_logit && g_discount && d_logit, DV3_ERR_NULL, "discount_bwd: null pointer");
  discount_bwd_kernel<<<nblk(n, 256), 256, 0, ST>>>(cont_logit, g_discount, gamma, n, d_logit);
  DV3_CHECK_LAUNCH("discount_bwd_kernel");
  return 0;
}

extern "C" int dv3_reward_ema(const float* x, int32_t n, double alpha, float* ema_vals,
                              float* offset_scale, void* stream) {
  DV3_REQUIRE(x && ema_vals && offset_scale, DV3_ERR_NULL, "reward_ema: null pointer");
  DV3_REQUIRE(n >= 1 && n <= EMA_MAX, DV3_ERR_BAD_SHAPE, "reward_ema: n=%d outside [1, %d]", n,
              EMA_MAX);
  // alpha * q and (1 - alpha) * ema with the python scalars rounded to fp32 once, like torch
  reward_ema_kernel<<<1, 1024, 0, ST>>>(x, n, (float)alpha, (float)(1.0 - alpha), ema_vals,
                                        offset_scale);
  DV3_CHECK_LAUNCH("reward_ema_kernel");
  return 0;
}

extern "C" int dv3_actor_loss_fwd(const float* target, const float* base, const float* weights,
                                  const float* entropy, const float* logp, const float* offset_scale,
                                  float entropy_coef, int32_t mode, int32_t count, float* normed,
                                  float* loss, void* stream) {
  DV3_REQUIRE(target && base && weights && entropy && loss && count > 0 && (mode == 0 || mode == 1),
              DV3_ERR_NULL, "actor_loss_fwd: null pointer / count=%d mode=%d", count, mode);
  DV3_REQUIRE(mode == 0 || logp, DV3_ERR_NULL, "actor_loss_fwd: reinforce needs logp");
  actor_loss_fwd_kernel<<<1, 1024, 0, ST>>>(target, base, weights, entropy, logp, offset_scale,
                                             entropy_coef, mode, count, normed, loss);
  DV3_CHECK_LAUNCH("actor_loss_fwd_kernel");
  return 0;
}

extern "C" int dv3_actor_loss_bwd(const float* g_loss, const float* target, const float* base,
                                  const float* weights, const float* offset_scale, float entropy_coef,
                                  int32_t mode, int32_t count, int32_t total, float* d_target,
                                  float* d_entropy, float* d_logp, void* stream) {
  DV3_REQUIRE(g_loss && target && base && weights && d_entropy && count > 0 && total >= count,
              DV3_ERR_NULL, "actor_loss_bwd: null pointer / count=%d total=%d", count, total);
  DV3_REQUIRE(mode == 0 ? d_target != nullptr : d_logp != nullptr, DV3_ERR_NULL,
              "actor_loss_bwd: missing output for mode %d", mode);
  actor_loss_bwd_kernel<<<nblk(total, 256), 256, 0, ST>>>(g_loss, target, base, weights, offset_scale,
                                                          entropy_coef, mode, count, total, d_target,
                                                          d_entropy, d_logp);
  DV3_CHECK_LAUNCH("actor_loss_bwd_kernel");
  return 0;
}

extern "C" int dv3_value_loss_fwd(const float* lp_target, const float* lp_slow, const float* weights,
                                  int32_t count, float* loss, void* stream) {
  DV3_REQUIRE(lp_target && weights && loss && count > 0, DV3_ERR_NULL, "value_loss_fwd: null pointer");
  value_loss_fwd_kernel<<<1, 1024, 0, ST>>>(lp_target, lp_slow, weights, count, loss);
  DV3_CHECK_LAUNCH("value_loss_fwd_kernel");
  return 0;
}

extern "C" int dv3_value_loss_bwd(const float* g_loss, const float* weights, int32_t count,
                                  float* d_lp, void* stream) {
  DV3_REQUIRE(g_loss && weights && d_lp && count > 0, DV3_ERR_NULL, "value_loss_bwd: null pointer");
  value_loss_bwd_kernel<<<nblk(count, 256), 256, 0, ST>>>(g_loss, weights, count, d_lp);
  DV3_CHECK_LAUNCH("value_loss_bwd_kernel");
  return 0;
}

extern "C" int dv3_normal_policy_fwd(const float* mean_raw, const float* std_raw, const float* action,
                                     float min_std, float max_std, int32_t R, int32_t A,
                                     float* entropy, float* logp, void* stream) {
  if (R <= 0) return 0;
  DV3_REQUIRE(mean_raw && std_raw && action && entropy && A > 0, DV3_ERR_NULL,
              "normal_policy_fwd: null pointer");
  normal_policy_fwd_kernel<<<nblk(R, 128), 128, 0, ST>>>(mean_raw, std_raw, action, min_std, max_std,
                                                         R, A, entropy, logp);
  DV3_CHECK_LAUNCH("normal_policy_fwd_kernel");
  return 0;
}

extern "C" int dv3_normal_policy_bwd(const float* mean_raw, const float* std_raw, const float* action,
                                     const float* g_entropy, const float* g_logp, float min_std,
                                     float max_std, int32_t R, int32_t A, float* d_mean_raw,
                                     float* d_std_raw, float* d_action, void* stream) {
  if (R <= 0) return 0;
  DV3_REQUIRE(mean_raw && std_raw && action && d_mean_raw && d_std_raw && A > 0, DV3_ERR_NULL,
              "normal_policy_bwd: null pointer");
  normal_policy_bwd_kernel<<<nblk((long long)R * A, 256), 256, 0, ST>>>(
      mean_raw, std_raw, action, g_entropy, g_logp, min_std, max_std, R, A, d_mean_raw, d_std_raw,
      d_action);
  DV3_CHECK_LAUNCH("normal_policy_bwd_kernel");
  return 0;
}

extern "C" int dv3_tensorstats(const float* x, long long n, float* out4, void* stream) {
  DV3_REQUIRE(x && out4 && n > 0, DV3_ERR_NULL, "tensorstats: null pointer / n=%lld", n);
  tensorstats_kernel<<<1, 1024, 0, ST>>>(x, n, out4);
  DV3_CHECK_LAUNCH("tensorstats_kernel");
  return 0;
}

extern "C" int dv3_ema_mix(float* dst, const float* src, long long n, double mix, void* stream) {
  if (n <= 0) return 0;
  DV3_REQUIRE(dst && src, DV3_ERR_NULL, "ema_mix: null pointer");
  ema_mix_kernel<<<nblk(n, 256), 256, 0, ST>>>(dst, src, n, (float)mix, (float)(1.0 - mix));
  DV3_CHECK_LAUNCH("ema_mix_kernel");
  return 0;
}

extern "C" int dv3_col_sum(const float* x, int32_t ld, int32_t M, int32_t n, float* out,
                           int32_t accumulate, void* stream) {
  DV3_REQUIRE(x && out && n > 0 && M >= 0 && ld >= n, DV3_ERR_NULL, "col_sum: null pointer / M=%d n=%d ld=%d",
              M, n, ld);
  if (!accumulate) DV3_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)n * 4, ST));
  if (M == 0) return 0;
  if (n <= 8 && ld == n) {
    const long long total = (long long)M * n;
    int blocks = nblk(total, 256 * 8);
    if (blocks > 592) blocks = 592;
    col_sum_narrow_kernel<<<blocks, 256, 0, ST>>>(x, total, n, out);
    DV3_CHECK_LAUNCH("col_sum_narrow_kernel");
    return 0;
  }
  const int cb = nblk(n, 32);
  int rb = (2 * 148 + cb - 1) / cb;                 // about two blocks per SM in total
  if (rb > (M + 63) / 64) rb = (M + 63) / 64;       // at least 64 rows per block
  if (rb < 1) rb = 1;
  const int rpb = (M + rb - 1) / rb;
  col_sum_kernel<<<dim3(cb, rb), 256, 0, ST>>>(x, ld, M, n, rpb, out);
  DV3_CHECK_LAUNCH("col_sum_kernel");
  return 0;
}

// ---- backward of RSSM.initial (networks.py:99-125), one row ----------------------------------------
//   deter0 = tanh(W);  y0 = SiLU(LN(W_out deter0));  lg = W_ims y0 + b_ims;  norm = unimix log-probs
// upstream: g_norm [S*C] (gradient reaching the straight-through mode's log-probs), g_deter0 [D].
// One CTA: lane = class for the categorical chain, then the two GEMVs and the LayerNorm backward.
// Outputs: d_lg [S*C] (= d b_ims), d_ypre [Hd], d_ln [Hd] (LN affine-output gradient), xhat [Hd],
// d_w_init [D]; the two rank-1 weight gradients follow in rank1_add_kernel.
namespace dv3 {

// out[0..N) (shared, zero-initialised by the caller) += sum_r v[r] * W[r, 0..N)   (W row-major
// [R, N], N % 4 == 0).  Warp w takes rows w, w+32, ...; a lane keeps float4 column slices in
// registers (8 loads in flight per row), the 32 warps meet through shared-memory atomics.
__device__ __forceinline__ void vec_mat_accum(const float* __restrict__ v, const float* __restrict__ W,
                                              int R, int N, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int cb = 0; cb < N; cb += 1024) {
    float4 acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = warp; r < R; r += nw) {
      const float g = v[r];
      const float* row = W + (size_t)r * N + cb;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = lane * 4 + q * 128;
        if (cb + c < N) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(row + c));
          acc[q].x = fmaf(g, w.x, acc[q].x); acc[q].y = fmaf(g, w.y, acc[q].y);
          acc[q].z = fmaf(g, w.z, acc[q].z); acc[q].w = fmaf(g, w.w, acc[q].w);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = cb + lane * 4 + q * 128;
      if (c < N) {
        atomicAdd(out + c, acc[q].x); atomicAdd(out + c + 1, acc[q].y);
        atomicAdd(out + c + 2, acc[q].z); atomicAdd(out + c + 3, acc[q].w);
      }
    }
  }
}

__global__ void __launch_bounds__(1024)
rssm_initial_bwd_kernel(const float* __restrict__ init_deter, const float* __restrict__ init_ypre,
                        const float* __restrict__ init_logit, const float* __restrict__ w_out,
                        const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                        const float* __restrict__ w_ims, const float* __restrict__ g_norm,
                        const float* __restrict__ g_deter0, int S, int C, int D, int Hd, float unimix,
                        float eps, float* __restrict__ d_lg, float* __restrict__ d_ypre,
                        float* __restrict__ d_ln, float* __restrict__ xhat_out,
                        float* __restrict__ d_w_init) {
  extern __shared__ float sm[];
  float* s_dlg = sm;                 // [S*C]
  float* s_v = s_dlg + S * C;        // [Hd] d_y0, then d_ypre
  float* s_d = s_v + Hd;             // [D]  d_deter0
  __shared__ float red[4 * 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  // 1. categorical chain, one warp per group
  for (int s = warp; s < S; s += nw) {
    const bool valid = lane < C;
    const float l = valid ? init_logit[s * C + lane] : 0.f;
    const Unimix um = unimix_probs(l, valid, C, unimix);
    const float g = valid ? g_norm[s * C + lane] : 0.f;
    // norm = l' - lse(l'):  d l'_k = g_k - softmax(l')_k sum(g)
    const float gs = warp_sum(g);
    float dl = g - um.probs * gs;
    if (unimix > 0.f) {
      // l' = log(q), q = (1-u) p + u/C, p = softmax(l)
      const float q = um.p * (1.f - unimix) + unimix / (float)C;
      const float dp = valid ? (1.f - unimix) * dl / q : 0.f;
      const float dot = warp_sum(valid ? um.p * dp : 0.f);
      dl = um.p * (dp - dot);
    }
    if (valid) { s_dlg[s * C + lane] = dl; d_lg[s * C + lane] = dl; }
  }
  const int SC = S * C;
  for (int j = tid; j < Hd; j += blockDim.x) s_v[j] = 0.f;
  for (int k = tid; k < D; k += blockDim.x) s_d[k] = 0.f;
  __syncthreads();
  // 2. d_y0[j] = sum_sc d_lg[sc] W_ims[sc, j]
  vec_mat_accum(s_dlg, w_ims, SC, Hd, s_v);
  __syncthreads();
  // 3. LayerNorm + SiLU backward of the row
  float acc[1] = {0.f};
  for (int j = tid; j < Hd; j += blockDim.x) acc[0] += init_ypre[j];
  block_sum<1>(acc, red);
  const float mean = acc[0] / (float)Hd;
  float q2[1] = {0.f};
  for (int j = tid; j < Hd; j += blockDim.x) { const float d = init_ypre[j] - mean; q2[0] = fmaf(d, d, q2[0]); }
  block_sum<1>(q2, red);
  const float rstd = 1.f / sqrtf(q2[0] / (float)Hd + eps);
  float m[2] = {0.f, 0.f};
  for (int j = tid; j < Hd; j += blockDim.x) {
    const float xh = (init_ypre[j] - mean) * rstd;
    const float v = fmaf(xh, ln_g[j], ln_b[j]);
    const float dv = s_v[j] * silu_grad(v);
    d_ln[j] = dv;
    xhat_out[j] = xh;
    const float dx = dv * ln_g[j];
    m[0] += dx;
    m[1] = fmaf(dx, xh, m[1]);
  }
  block_sum<2>(m, red);
  const float m1 = m[0] / (float)Hd, m2 = m[1] / (float)Hd;
  for (int j = tid; j < Hd; j += blockDim.x) {
    const float xh = (init_ypre[j] - mean) * rstd;
    const float dp = rstd * (d_ln[j] * ln_g[j] - m1 - xh * m2);
    s_v[j] = dp;
    d_ypre[j] = dp;
  }
  __syncthreads();
  // 4. d_deter0[k] = sum_j d_ypre[j] W_out[j, k] + g_deter0[k];  d W = d_deter0 (1 - deter0^2)
  vec_mat_accum(s_v, w_out, Hd, D, s_d);
  __syncthreads();
  for (int k = tid; k < D; k += blockDim.x) {
    const float a = s_d[k] + (g_deter0 ? g_deter0[k] : 0.f);
    const float t = init_deter[k];
    d_w_init[k] = a * (1.f - t * t);
  }
}

// out[i, j] += a[i] * b[j]
__global__ void rank1_add_kernel(const float* __restrict__ a, const float* __restrict__ b, int M, int N,
                                 float* __restrict__ out, int ld) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)M * N) return;
  const int i = (int)(e / N), j = (int)(e % N);
  out[(size_t)i * ld + j] += a[i] * b[j];
}

// the vector gradients of the initial path, added onto the caller's buffers in one launch
__global__ void initial_vec_add_kernel(const float* __restrict__ d_lg, const float* __restrict__ d_ln,
                                       const float* __restrict__ xhat, const float* __restrict__ dw,
                                       int SC, int Hd, int D, float* __restrict__ d_b_ims,
                                       float* __restrict__ d_ln_g, float* __restrict__ d_ln_b,
                                       float* __restrict__ d_w_init) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < SC) { d_b_ims[i] += d_lg[i]; return; }
  i -= SC;
  if (i < Hd) { d_ln_g[i] += d_ln[i] * xhat[i]; return; }
  i -= Hd;
  if (i < Hd) { d_ln_b[i] += d_ln[i]; return; }
  i -= Hd;
  if (i < D) d_w_init[i] += dw[i];
}

}  // namespace dv3

extern "C" size_t dv3_rssm_initial_bwd_scratch_floats(const dv3_rssm_dims* d) {
  if (!d) return 0;
  return (size_t)d->stoch * d->classes + 3 * (size_t)d->hidden + d->deter;
}

extern "C" int dv3_rssm_initial_bwd(const dv3_rssm_dims* d, const dv3_rssm_params* p,
                                    const float* init_deter, const float* init_ypre,
                                    const float* init_y, const float* init_logit, const float* g_norm,
                                    const float* g_deter0, float* d_w_init, float* d_w_out,
                                    float* d_ln_out_g, float* d_ln_out_b, float* d_w_ims,
                                    float* d_b_ims, float* scratch, void* stream) {
  DV3_REQUIRE(d && p && init_deter && init_ypre && init_y && init_logit && g_norm && d_w_init &&
                  d_w_out && d_ln_out_g && d_ln_out_b && d_w_ims && d_b_ims && scratch,
              DV3_ERR_NULL, "rssm_initial_bwd: null pointer");
  const int S = d->stoch, C = d->classes, SC = S * C, D = d->deter, Hd = d->hidden;
  DV3_REQUIRE(C <= 32, DV3_ERR_BAD_SHAPE, "rssm_initial_bwd: classes=%d > 32", C);
  // scratch: d_lg [SC] | d_ypre [Hd] | d_ln [Hd] | xhat [Hd] | d_w_init [D]
  float* s_dlg = scratch;
  float* s_dyp = s_dlg + SC;
  float* s_dln = s_dyp + Hd;
  float* s_xh = s_dln + Hd;
  float* s_dw = s_xh + Hd;
  const size_t smem = (size_t)(SC + Hd + D) * 4;
  DV3_REQUIRE(smem <= 48 * 1024 && Hd % 4 == 0 && D % 4 == 0, DV3_ERR_BAD_SHAPE,
              "rssm_initial_bwd: S*C + Hd + D = %d too large or widths not multiples of 4", SC + Hd + D);
  rssm_initial_bwd_kernel<<<1, 1024, smem, ST>>>(init_deter, init_ypre, init_logit, p->w_out, p->ln_out_g,
                                                 p->ln_out_b, p->w_ims, g_norm, g_deter0, S, C, D, Hd,
                                                 d->unimix, d->ln_eps, s_dlg, s_dyp, s_dln, s_xh, s_dw);
  DV3_CHECK_LAUNCH("rssm_initial_bwd_kernel");
  // all outputs are ADDED onto the caller's buffers (which hold the bulk contributions, or zeros)
  rank1_add_kernel<<<nblk((long long)SC * Hd, 256), 256, 0, ST>>>(s_dlg, init_y, SC, Hd, d_w_ims, Hd);
  DV3_CHECK_LAUNCH("rank1_add_kernel");
  rank1_add_kernel<<<nblk((long long)Hd * D, 256), 256, 0, ST>>>(s_dyp, init_deter, Hd, D, d_w_out, D);
  DV3_CHECK_LAUNCH("rank1_add_kernel");
  initial_vec_add_kernel<<<nblk(SC + 2 * Hd + D, 256), 256, 0, ST>>>(s_dlg, s_dln, s_xh, s_dw, SC, Hd, D,
                                                                    d_b_ims, d_ln_out_g, d_ln_out_b, d_w_init);
  DV3_CHECK_LAUNCH("initial_vec_add_kernel");
  return 0;
}
