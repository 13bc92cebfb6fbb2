// 4x4 stride-2 convolution / transposed convolution of the image encoder and decoder
// (reference networks.py:448-585: ConvEncoder = [Conv2dSamePad(k=4, s=2) -> ImgChLayerNorm -> SiLU] x 4,
// ConvDecoder = Linear -> [ConvTranspose2d(k=4, s=2, p=1) -> ImgChLayerNorm -> SiLU] x 3 -> ConvTranspose2d)
// as GEMMs on the tcgen05 3xTF32 kernel.  Activations are channels-last matrices [pixels, C], so the
// channel LayerNorm + SiLU is the row kernel the MLPs use.  Two data-movement kernels make every
// product of both layers' forward and backward a plain GEMM:
//
//   im2col : big grid [n, H, W, C]  ->  cols [n * H/2 * W/2, 16 C],  cols[(n,oy,ox), (ky,kx,c)] =
//            x[n, 2 oy - 1 + ky, 2 ox - 1 + kx, c]   (zero outside; pad 1 = the 'same' padding of
//            k = 4, s = 2 on even sizes).  Writes the tf32 hi / lo planes the GEMM reads.
//   col2im : its adjoint, as a gather: x[n, y, x, c] = sum of the (at most) 2 x 2 taps (ky, kx) of the
//            parity of (y + 1, x + 1) -- deterministic, no atomics; optional bias / constant shift.
//
//   conv    fwd  y = im2col(x) W2^T            bwd  dx = col2im(dy W2),  dW2 = dy^T im2col(x)
//   deconv  fwd  y = col2im(x W2d)             bwd  dx = im2col(dy) W2d^T,  dW2d = x^T im2col(dy)
// with W2 [Cout, (ky,kx,ci)] / W2d [Cin, (ky,kx,co)] the weights permuted to the column order.
#include "dv3_common.cuh"

namespace dv3 {

// one thread per 4 consecutive channels of one (row, tap); C % 4 == 0
__global__ void __launch_bounds__(256)
im2col_s2k4_vec_kernel(const float* __restrict__ x, int n, int H, int W, int C, long long total4,
                       float* __restrict__ cols, float* __restrict__ hi, float* __restrict__ lo) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total4) return;
  const int c4n = C >> 2;
  const int c4 = (int)(e % c4n);
  long long t = e / c4n;
  const int tap = (int)(t % 16);
  t /= 16;
  const int Wo = W >> 1, Ho = H >> 1;
  const int ox = (int)(t % Wo);
  t /= Wo;
  const int oy = (int)(t % Ho);
  const int img = (int)(t / Ho);
  const int iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (iy >= 0 && iy < H && ix >= 0 && ix < W)
    v = *reinterpret_cast<const float4*>(x + (((size_t)img * H + iy) * W + ix) * C + 4 * c4);
  const size_t o = (size_t)e * 4;      // == row * 16 C + tap * C + 4 c4
  if (cols) *reinterpret_cast<float4*>(cols + o) = v;
  if (hi) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi + o) = h;
    *reinterpret_cast<float4*>(lo + o) = l;
  }
}

// scalar form (C = 3: the image itself)
__global__ void __launch_bounds__(256)
im2col_s2k4_kernel(const float* __restrict__ x, int n, int H, int W, int C, long long total,
                   float* __restrict__ cols, float* __restrict__ hi, float* __restrict__ lo) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c = (int)(e % C);
  long long t = e / C;
  const int tap = (int)(t % 16);
  t /= 16;
  const int Wo = W >> 1, Ho = H >> 1;
  const int ox = (int)(t % Wo);
  t /= Wo;
  const int oy = (int)(t % Ho);
  const int img = (int)(t / Ho);
  const int iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
  float v = 0.f;
  if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((size_t)img * H + iy) * W + ix) * C + c];
  if (cols) cols[e] = v;
  if (hi) {
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[e] = h;
    lo[e] = v - h;
  }
}

// x[img, y, xx, c] = bias[c] + shift + sum over the taps that reach (y, xx)
__global__ void __launch_bounds__(256)
col2im_s2k4_kernel(const float* __restrict__ cols, int n, int H, int W, int C,
                   const float* __restrict__ bias, float shift, long long total,
                   float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c = (int)(e % C);
  long long t = e / C;
  const int xx = (int)(t % W);
  t /= W;
  const int y = (int)(t % H);
  const int img = (int)(t / H);
  const int Wo = W >> 1, Ho = H >> 1;
  float acc = shift + (bias ? bias[c] : 0.f);
  const int ky0 = (y + 1) & 1, kx0 = (xx + 1) & 1;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int ky = ky0 + 2 * a;
    const int oy = (y + 1 - ky) >> 1;
    if (oy < 0 || oy >= Ho) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int kx = kx0 + 2 * b;
      const int ox = (xx + 1 - kx) >> 1;
      if (ox < 0 || ox >= Wo) continue;
      acc += cols[((((size_t)img * Ho + oy) * Wo + ox) * 16 + (ky * 4 + kx)) * C + c];
    }
  }
  out[e] = acc;
}

}  // namespace dv3

using namespace dv3;

extern "C" int dv3_im2col_s2k4(const float* x, int32_t n, int32_t H, int32_t W, int32_t C, float* cols,
                               float* hi, float* lo, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DV3_REQUIRE(n >= 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C >= 1, DV3_ERR_BAD_SHAPE,
              "im2col_s2k4: n=%d H=%d W=%d C=%d (even H, W)", n, H, W, C);
  DV3_REQUIRE(x && (cols || (hi && lo)) && (!hi == !lo), DV3_ERR_NULL, "im2col_s2k4: null pointer");
  const long long total = (long long)n * (H / 2) * (W / 2) * 16 * C;
  if (total == 0) return 0;
  if (C % 4 == 0) {
    const long long t4 = total / 4;
    im2col_s2k4_vec_kernel<<<(unsigned)((t4 + 255) / 256), 256, 0, st>>>(x, n, H, W, C, t4, cols, hi, lo);
    DV3_CHECK_LAUNCH("im2col_s2k4_vec_kernel");
  } else {
    im2col_s2k4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, n, H, W, C, total, cols, hi, lo);
    DV3_CHECK_LAUNCH("im2col_s2k4_kernel");
  }
  return 0;
}

extern "C" int dv3_col2im_s2k4(const float* cols, int32_t n, int32_t H, int32_t W, int32_t C,
                               const float* bias, float shift, float* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DV3_REQUIRE(n >= 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C >= 1, DV3_ERR_BAD_SHAPE,
              "col2im_s2k4: n=%d H=%d W=%d C=%d (even H, W)", n, H, W, C);
  DV3_REQUIRE(cols && out, DV3_ERR_NULL, "col2im_s2k4: null pointer");
  const long long total = (long long)n * H * W * C;
  if (total == 0) return 0;
  col2im_s2k4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cols, n, H, W, C, bias, shift, total, out);
  DV3_CHECK_LAUNCH("col2im_s2k4_kernel");
  return 0;
}
