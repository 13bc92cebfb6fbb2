// Fused global-norm clip + Adam over one flat parameter buffer: the optimizer call sequence of the
// reference's tools.Optimizer (tools.py:760-776: clip_grad_norm_ -> Adam.step) as three launches
// instead of ~10 multi-tensor launches per optimizer.
//   1. partial sums of squares of the flat gradient (fixed grid, fixed order -> deterministic)
//   2. one block: total norm, clip coefficient min(1, clip / (norm + 1e-6)), step += 1, bias
//      corrections (double precision, like torch's host/capturable path)
//   3. update: g *= coef; m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2;
//              p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)          (torch.optim.Adam)
// The step counter and the control block live on the device, so the sequence is graph-capturable.
#include "dv3_common.cuh"

namespace dv3 {

constexpr int OP_THREADS = 256;
constexpr int OP_BLOCKS = 296;     // 2 per SM

__global__ void __launch_bounds__(OP_THREADS)
sqsum_partial_kernel(const float* __restrict__ g, long long n4, float* __restrict__ partials) {
  __shared__ float red[4 * 32];
  float acc[1] = {0.f};
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * OP_THREADS + threadIdx.x; i < n4;
       i += (long long)gridDim.x * OP_THREADS) {
    const float4 v = g4[i];
    acc[0] = fmaf(v.x, v.x, acc[0]);
    acc[0] = fmaf(v.y, v.y, acc[0]);
    acc[0] = fmaf(v.z, v.z, acc[0]);
    acc[0] = fmaf(v.w, v.w, acc[0]);
  }
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc[0];
}

// ctl: [0] total norm, [1] clip coefficient, [2] step size lr / bc1, [3] sqrt(bc2)
__global__ void adam_prepare_kernel(const float* __restrict__ partials, int np, float clip, float lr,
                                    float beta1, float beta2, float* __restrict__ step,
                                    float* __restrict__ ctl) {
  __shared__ float red[4 * 32];
  float acc[1] = {0.f};
  for (int i = threadIdx.x; i < np; i += blockDim.x) acc[0] += partials[i];
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(acc[0]);
    float coef = 1.f;
    if (clip > 0.f) coef = fminf(clip / (norm + 1e-6f), 1.f);
    const float s = step[0] + 1.f;
    step[0] = s;
    const double bc1 = 1.0 - pow((double)beta1, (double)s);
    const double bc2 = 1.0 - pow((double)beta2, (double)s);
    ctl[0] = norm;
    ctl[1] = coef;
    ctl[2] = (float)((double)lr / bc1);
    ctl[3] = (float)sqrt(bc2);
  }
}

__global__ void __launch_bounds__(OP_THREADS)
adam_update_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                   float* __restrict__ v, long long n4, float beta1, float beta2, float eps,
                   float decay_mul, const float* __restrict__ ctl, float* __restrict__ hi,
                   float* __restrict__ lo) {
  const float coef = ctl[1], step_size = ctl[2], bc2s = ctl[3];
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = (long long)blockIdx.x * OP_THREADS + threadIdx.x; i < n4;
       i += (long long)gridDim.x * OP_THREADS) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    auto upd = [&](float& pe, float& me, float& ve, float ge) {
      ge *= coef;
      pe *= decay_mul;
      me = me + omb1 * (ge - me);
      ve = beta2 * ve + omb2 * ge * ge;
      const float denom = sqrtf(ve) / bc2s + eps;
      pe -= step_size * me / denom;
    };
    upd(pp.x, mm.x, vv.x, gg.x);
    upd(pp.y, mm.y, vv.y, gg.y);
    upd(pp.z, mm.z, vv.z, gg.z);
    upd(pp.w, mm.w, vv.w, gg.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
    if (hi) {
      // tf32 operand planes of the updated parameters (hi + lo == p exactly): the GEMMs of the
      // next step read them, no per-weight split pass
      float4 h, l;
      h.x = __uint_as_float(__float_as_uint(pp.x) & 0xFFFFE000u);
      h.y = __uint_as_float(__float_as_uint(pp.y) & 0xFFFFE000u);
      h.z = __uint_as_float(__float_as_uint(pp.z) & 0xFFFFE000u);
      h.w = __uint_as_float(__float_as_uint(pp.w) & 0xFFFFE000u);
      l.x = pp.x - h.x; l.y = pp.y - h.y; l.z = pp.z - h.z; l.w = pp.w - h.w;
      reinterpret_cast<float4*>(hi)[i] = h;
      reinterpret_cast<float4*>(lo)[i] = l;
    }
  }
}

}  // namespace dv3

static int adam_clip_step_any(float* p, const float* g, float* m, float* v, long long n, float lr,
                              float beta1, float beta2, float eps, float clip, float decay_mul,
                              float* step, float* ctl, float* scratch, float* hi, float* lo,
                              void* stream) {
  using namespace dv3;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DV3_REQUIRE(n >= 0 && n % 4 == 0, DV3_ERR_BAD_SHAPE, "adam_clip_step: n=%lld must be a multiple of 4", n);
  if (n == 0) return 0;
  DV3_REQUIRE(p && g && m && v && step && ctl && scratch, DV3_ERR_NULL, "adam_clip_step: null pointer");
  const long long n4 = n / 4;
  int blocks = (int)((n4 + OP_THREADS - 1) / OP_THREADS);
  if (blocks > OP_BLOCKS) blocks = OP_BLOCKS;
  sqsum_partial_kernel<<<blocks, OP_THREADS, 0, st>>>(g, n4, scratch);
  DV3_CHECK_LAUNCH("sqsum_partial_kernel");
  adam_prepare_kernel<<<1, 256, 0, st>>>(scratch, blocks, clip, lr, beta1, beta2, step, ctl);
  DV3_CHECK_LAUNCH("adam_prepare_kernel");
  adam_update_kernel<<<blocks, OP_THREADS, 0, st>>>(p, g, m, v, n4, beta1, beta2, eps, decay_mul,
                                                    ctl, hi, lo);
  DV3_CHECK_LAUNCH("adam_update_kernel");
  return 0;
}

extern "C" int dv3_adam_clip_step(float* p, const float* g, float* m, float* v, long long n,
                                  float lr, float beta1, float beta2, float eps, float clip,
                                  float decay_mul, float* step, float* ctl, float* scratch,
                                  void* stream) {
  return adam_clip_step_any(p, g, m, v, n, lr, beta1, beta2, eps, clip, decay_mul, step, ctl,
                            scratch, nullptr, nullptr, stream);
}

extern "C" int dv3_adam_clip_step_planes(float* p, const float* g, float* m, float* v, long long n,
                                         float lr, float beta1, float beta2, float eps, float clip,
                                         float decay_mul, float* step, float* ctl, float* scratch,
                                         float* hi, float* lo, void* stream) {
  DV3_REQUIRE(hi && lo, DV3_ERR_NULL, "adam_clip_step_planes: null plane buffer");
  return adam_clip_step_any(p, g, m, v, n, lr, beta1, beta2, eps, clip, decay_mul, step, ctl,
                            scratch, hi, lo, stream);
}
