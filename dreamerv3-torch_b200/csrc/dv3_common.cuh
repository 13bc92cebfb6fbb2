// Shared device helpers + host-side error plumbing for libdv3_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/dv3_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdv3_b200 targets sm_100a (B200) only"
#endif

namespace dv3 {

// ---- host error state -------------------------------------------------------------------
void set_error(const char* fmt, ...);
void note_launch();
// optional CUDA-event timing of the GEMM launches (dv3_prof_*): kind 0 = skinny, 1 = tiled
bool prof_on();
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st, int kind, double flops);
int cuda_fail(cudaError_t e, const char* what);

#define DV3_CHECK_CUDA(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::dv3::cuda_fail(_e, #expr); \
  } while (0)

// every kernel launch of the library goes through this macro: error check + launch counter
#define DV3_CHECK_LAUNCH(name)                                   \
  do {                                                           \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return ::dv3::cuda_fail(_e, name);    \
    ::dv3::note_launch();                                        \
  } while (0)

#define DV3_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::dv3::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define DV3_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// ---- programmatic dependent launch ---------------------------------------------------------
// Kernels of the long dependent chains (imagination steps, MLP layers) are launched with the
// programmatic-stream-serialization attribute: the next kernel's CTAs become resident and run
// their prologue (barrier init, TMEM allocation, tensor-map fetch) while the previous kernel
// drains; pdl_wait() then blocks until the previous grid has completed and its memory is
// visible.  Every kernel launched through launch_pdl() must call pdl_wait() before its first
// global-memory access.  Off by default (DV3_PDL=1 turns it on): on the dmc_proprio step the early
// residency of 200 KB GEMM CTAs cost more than the hidden prologues saved (16.76 vs 16.33 ms).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();
int pdl_mode();

// ---- process-level caches ------------------------------------------------------------------
// Environment knobs are read once and cached; dv3_reload_env() (tests, A/B runs) drops the cache.
struct EnvSlot {
  unsigned epoch = 0;
  bool has = false;
  char val[64] = {0};
};
const char* env_cached(const char* name, EnvSlot& slot);
#define DV3_ENV(name) ([]() -> const char* { static ::dv3::EnvSlot slot_; return ::dv3::env_cached(name, slot_); }())

// "once per device" latch for per-context settings (cudaFuncSetAttribute): a process may drive
// several GPUs through the library.
struct DeviceOnce {
  unsigned long long done = 0;
  bool need() {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (done & bit) return false;
    done |= bit;
    return true;
  }
};
int sm_count();     // multiprocessors of the current device (cached per device)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  const int pm = pdl_mode();
  at[0].val.programmaticStreamSerializationAllowed = (pm == 1 || (pm == 2 && smem <= 64 * 1024)) ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- device helpers ----------------------------------------------------------------------
constexpr int WARP = 32;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// Block-wide sum of up to 4 values at once; every thread gets the totals.  `red` is a
// shared array of at least 4*32 floats.  Ends with a barrier, so `red` is reusable.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * 32 + wid] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float t = (lane < nw) ? red[i * 32 + lane] : 0.f;
    v[i] = warp_sum(t);
  }
  __syncthreads();
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// d/dx [x*sigmoid(x)]
__device__ __forceinline__ float silu_grad(float x) {
  const float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}

// ---- unimix categorical helpers (one class per lane) ---------------------------------------
// probs are built exactly along the reference's chain (tools.py:439-441 + torch Categorical):
//   p = softmax(l); p = (1-r)p + r/C; l' = log p; norm = l' - logsumexp(l'); probs = softmax(norm)
struct Unimix {
  float p;      // softmax(l)
  float norm;   // normalised log-prob after unimix
  float probs;  // softmax(norm)
};

__device__ __forceinline__ Unimix unimix_probs(float l, bool valid, int C, float unimix) {
  Unimix o;
  const float NEG = -INFINITY;
  float m = warp_max(valid ? l : NEG);
  float e = valid ? expf(l - m) : 0.f;
  float s = warp_sum(e);
  o.p = e / s;
  float lp = l;
  if (unimix > 0.f) {
    const float pm = o.p * (1.f - unimix) + unimix / (float)C;
    lp = logf(pm);
  }
  float m2 = warp_max(valid ? lp : NEG);
  float s2 = warp_sum(valid ? expf(lp - m2) : 0.f);
  o.norm = lp - (m2 + logf(s2));
  float m3 = warp_max(valid ? o.norm : NEG);
  float e3 = valid ? expf(o.norm - m3) : 0.f;
  float s3 = warp_sum(e3);
  o.probs = e3 / s3;
  return o;
}

// first-index argmax across the warp
__device__ __forceinline__ int warp_argmax(float v, bool valid, int lane) {
  float bv = valid ? v : -INFINITY;
  int bi = valid ? lane : 0x7fffffff;
  if (valid && v != v) bv = -INFINITY;  // NaN never wins
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(FULL, bv, o);
    const int oi = __shfl_xor_sync(FULL, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  return bi;
}

// ---- internal host launchers (definitions spread over the .cu files) ----------------------
struct LinearArgs {
  const float* A[2];
  const float* W[2];
  int lda[2], ldw[2], K[2];
  const float* bias;
  const float* addend;
  int ldadd;
  float* C;
  int ldc, M, N, accumulate;
};
int launch_linear(const LinearArgs& g, cudaStream_t st);
int linear1(const float* A, int lda, const float* W, int ldw, int K, const float* bias, float* C,
            int ldc, int M, int N, int accumulate, cudaStream_t st);
int launch_transpose(const float* in, int ld, int R, int C, float* out, cudaStream_t st);

// optional tf32 hi/lo planes [M, ld] written by a producer kernel beside its fp32 output
struct SplitOut {
  float* hi = nullptr;
  float* lo = nullptr;
  int ld = 0;
};

// Row-wise kernels.  Every matrix argument is (pointer, row stride) so that batch-major
// [B,T,w] tensors can be addressed per time step as (base + t*w, T*w).
int ln_silu_fwd(const float* pre, int ld, const float* g, const float* b, float eps, int M, int n,
                float* out, int ldo, cudaStream_t st, SplitOut so = SplitOut());
int ln_silu_bwd(const float* pre, int ld, const float* g, const float* b, float eps,
                const float* d_out, int ldd, int M, int n, float* d_pre, int ldp, float* d_ln,
                int ldl, cudaStream_t st, SplitOut so = SplitOut());
// x_pre = addend + sum_g WT[g*C+idx[g]] + sum_a act[a]*WT[S*C+a];  x = SiLU(LN(x_pre))
int gather_ln_silu(const int32_t* idx, int ldi, int S, int C, const float* act, int lda, int A,
                   const float* WT, const float* addend, int ldadd, const float* g, const float* b,
                   float eps, int M, int n, float* pre, int ldp, float* out, int ldo,
                   cudaStream_t st, SplitOut so = SplitOut());
int gru_gates_fwd(const float* g_pre, int ldg, const float* g, const float* b, float eps,
                  const float* h, int ldh, int M, int D, float* h_new, int ldn, cudaStream_t st,
                  SplitOut so = SplitOut());
// dh_in: up to 4 addends (NULL = skip), each with its own row stride
int gru_gates_bwd(const float* g_pre, int ldg, const float* g, const float* b, float eps,
                  const float* h, int ldh, const float* const dh_in[4], const int ld_in[4], int M,
                  int D, float* d_g_pre, int ldp, float* d_g_ln, int ldl, float* dh_direct,
                  int ldd, cudaStream_t st, SplitOut so = SplitOut());
// u row r is read at row perm(r) = (r % permT) * permB + r / permT when permT > 0
int onehot_sample(const float* logits, int ldl, const float* u, int ldu, int permT, int permB,
                  float unimix, int M, int S, int C, int32_t* idx, int ldi, float* onehot, int ldo,
                  cudaStream_t st);
int onehot_st_bwd(const float* logits, int ldl, const float* g1, int ldg1, const float* g2,
                  int ldg2, const float* ext, int lde, float unimix, int M, int S, int C,
                  float* d_logits, int ldd, cudaStream_t st, SplitOut so = SplitOut());
int idx_to_onehot(const int32_t* idx, int ldi, int M, int S, int C, float* out, int ld,
                  cudaStream_t st);
int copy_rows(const float* in, int ldi, int M, int n, float* out, int ldo, cudaStream_t st);
int copy_rows_i32(const int32_t* in, int ldi, int M, int n, int32_t* out, int ldo, cudaStream_t st);
int tanh_vec(const float* in, int n, float* out, cudaStream_t st);
int fill_zero(void* p, size_t bytes, cudaStream_t st);

// ---- tensor-core Linear (dv3_umma.cu) -----------------------------------------------------
int tc_split(const float* a1, int ld1, int K1, const float* a2, int ld2, int K2, int M, float* hi,
             float* lo, cudaStream_t st, int Kp = 0);
int tc_split_t(const float* in, int ld, int R, int C, float* hi, float* lo, cudaStream_t st,
               int Cp = 0);
int tc_gemm(const float* Ah, const float* Al, const float* Wh, const float* Wl, const float* bias,
            const float* addend, int ldadd, float* C, int ldc, int M, int N, int K, int accumulate,
            cudaStream_t st);
// one operand of the persistent tcgen05 GEMM (dv3_umma2.cu): hi/lo planes with a common row
// stride; mn == false: stored [rows, K] ("K-major"), mn == true: stored [K, rows].
// lo == nullptr on every operand = raw fp32 operands, split inside the SM (slower, no pre-pass).
struct TcOperand {
  const float* hi;
  const float* lo;
  int ld;
  bool mn;
};
int tc_gemm_ops(const TcOperand& A1, int K1, const TcOperand* A2, int K2, const TcOperand& B,
                const float* bias, const float* addend, int ldadd, float* C, int ldc, int M, int N,
                int accumulate, cudaStream_t st);
// the same product from a plain fp32 A operand, split in the SM into tensor memory (dv3_umma2t.cu);
// tc_gemm_rawa_ok: shapes / alignments the kernel covers
bool tc_gemm_rawa_ok(const float* A1, int lda1, int K1, const float* A2, int lda2, int K2,
                     const TcOperand& B, int M, int N);
bool tc_gemm_rawa_single_wave(int M, int N);
int tc_gemm_rawa(const float* A1, int lda1, int K1, const float* A2, int lda2, int K2,
                 const TcOperand& B, const float* bias, const float* addend, int ldadd, float* C,
                 int ldc, int M, int N, cudaStream_t st);
// the same product on CTA pairs (cta_group::2, 256 x BN tiles; dv3_umma2x.cu)
int tc_gemm_pair(const TcOperand& A1, int K1, const TcOperand* A2, int K2, const TcOperand& B,
                 const float* bias, const float* addend, int ldadd, float* C, int ldc, int M, int N,
                 int accumulate, int BN, int splitk, cudaStream_t st);
// persistent (single cooperative launch) forward recurrence of observe; *used == false -> not
// applicable for these shapes, run the stepwise launches instead (dv3_observe_persistent.cu)
int observe_fwd_persistent(const dv3_rssm_dims* d, const dv3_rssm_params* p,
                           const dv3_observe_io* io, const float* WinT, const float* pre_e,
                           unsigned* bar, cudaStream_t st, bool* used);
// buffers dv3_observe_bwd shares with its persistent recurrence kernel
// (dv3_observe_persistent_bwd.cu); the transposed weights are the stepwise path's
struct ObsBwdShared {
  const float *dh_prior, *WosT, *WobsT, *WgruT, *WinT;
  float *ds_rec, *dh_rec, *dinit_s, *dinit_h;      // [B,SC], [B,D], [B,SC], [B,D]; zeroed by caller
  float *d_z, *dh_z, *dhdir, *dxh;                 // scratch [B,Hd], [B,D], [B,D], [B,Hd+D]
  unsigned* bar;
};
int observe_bwd_persistent(const dv3_rssm_dims* d, const dv3_rssm_params* p,
                           const dv3_observe_bwd_io* io, const ObsBwdShared& w, cudaStream_t st,
                           bool* used);
// persistent imagination forward (dv3_imagine_persistent.cu): the operand planes dv3_imagine_fwd
// has prepared; *used == false on return -> shapes not covered, run the stepwise launches
struct PiPlane { const float* hi; const float* lo; int ld; };
struct PiPlanes {
  bool ok;
  float *dsp_hi[2], *dsp_lo[2], *asp_hi[2], *asp_lo[2], *xsp_hi, *xsp_lo, *ysp_hi, *ysp_lo;
  PiPlane wa[3], gru, out, ims;
  const float* Wa0T;
  const float* WinT;
};
int imagine_fwd_persistent(const dv3_rssm_dims* d, const dv3_rssm_params* p, const dv3_actor* a,
                           const dv3_imagine_io* io, const PiPlanes& pl, void* sync_ws,
                           cudaStream_t st, bool* used);
size_t imagine_persistent_sync_bytes(int N);
unsigned long long* po_timing_buffer();   // debug stamps (DV3_OBSERVE_TIMING=1 fwd, =2 bwd)
// rows below this go to the CUDA-core kernels (a 128-row MMA tile would be mostly padding)
constexpr int TC_MIN_ROWS = 64;

// bump allocator over the caller's workspace
struct Arena {
  char* base; size_t size, used;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
  template <typename T> T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* r = base ? reinterpret_cast<T*>(base + used) : nullptr;
    used += bytes;
    return r;
  }
  bool ok() const { return used <= size; }
};

// A Linear weight [N,K] (row stride ldw) with, optionally, its tf32 hi/lo split for the tensor
// core path.  apply(): C = [A1|A2] W^T + bias + addend.
struct LinW {
  const float* W = nullptr;
  int ldw = 0, N = 0, K = 0;
  float* hi = nullptr;
  float* lo = nullptr;
  bool tc = false;
  int ldp = 0;        // plane pitch (0: K)
  bool mn = false;    // planes stored [K, N] (a weight used through its transpose)

  // use caller-supplied planes of (a column block of) the weight instead of splitting per call
  bool adopt(const dv3_tc_operand* o, int col0, bool stored_kn) {
    if (!tc || !o || !o->hi || !o->lo || o->mn_major != 0 || (o->ld & 3) || (col0 & 3)) return false;
    hi = const_cast<float*>(o->hi) + col0;
    lo = const_cast<float*>(o->lo) + col0;
    ldp = o->ld;
    mn = stored_kn;
    return true;
  }

  void reserve(Arena& a, bool use_tc, int n, int k) {
    N = n; K = k; tc = use_tc && (k % 4 == 0);
    if (tc) { hi = a.take<float>((size_t)n * k); lo = a.take<float>((size_t)n * k); }
  }
  int prepare(const float* w, int ld, cudaStream_t st) {
    W = w; ldw = ld;
    if (tc) return tc_split(w, ld, K, nullptr, 0, 0, N, hi, lo, st);
    return 0;
  }
  // C = [A1|A2] W^T from A planes a producer kernel already wrote (tensor-core path only)
  int apply_split(const SplitOut& a1, int K1, const SplitOut* a2, int K2, const float* bias,
                  const float* addend, int ldadd, float* C, int ldc, int M, cudaStream_t st) const {
    TcOperand o1{a1.hi, a1.lo, a1.ld, false}, o2{nullptr, nullptr, 0, false},
        b{hi, lo, ldp ? ldp : K, mn};
    if (a2) { o2.hi = a2->hi; o2.lo = a2->lo; o2.ld = a2->ld; }
    return tc_gemm_ops(o1, K1, a2 ? &o2 : nullptr, K2, b, bias, addend, ldadd, C, ldc, M, N, 0, st);
  }
  // C = [A1|A2] W^T from the fp32 activations (A split in the SM into tensor memory); false when
  // the raw-A kernel does not cover the shapes -> the caller uses apply_split
  bool rawa_ok(const float* A1, int lda1, int K1, const float* A2, int lda2, int K2, int M) const {
    if (!tc || !tc_gemm_rawa_single_wave(M, N)) return false;
    TcOperand b{hi, lo, ldp ? ldp : K, mn};
    return tc_gemm_rawa_ok(A1, lda1, K1, A2, lda2, A2 ? K2 : 0, b, M, N);
  }
  int apply_rawa(const float* A1, int lda1, int K1, const float* A2, int lda2, int K2,
                 const float* bias, const float* addend, int ldadd, float* C, int ldc, int M,
                 cudaStream_t st) const {
    TcOperand b{hi, lo, ldp ? ldp : K, mn};
    return tc_gemm_rawa(A1, lda1, K1, A2, lda2, A2 ? K2 : 0, b, bias, addend, ldadd, C, ldc, M, N, st);
  }
  // ascr: 2*M*K floats of scratch (only used on the tensor-core path)
  int apply(const float* A1, int lda1, int K1, const float* A2, int lda2, int K2, const float* bias,
            const float* addend, int ldadd, float* C, int ldc, int M, float* ascr,
            cudaStream_t st) const {
    if (tc) {
      float* ah = ascr;
      float* al = ascr + (size_t)M * K;
      DV3_TRY(tc_split(A1, lda1, K1, A2, lda2, A2 ? K2 : 0, M, ah, al, st));
      return tc_gemm(ah, al, hi, lo, bias, addend, ldadd, C, ldc, M, N, K, 0, st);
    }
    LinearArgs g{};
    g.A[0] = A1; g.lda[0] = lda1; g.W[0] = W; g.ldw[0] = ldw; g.K[0] = K1;
    if (A2) { g.A[1] = A2; g.lda[1] = lda2; g.W[1] = W + K1; g.ldw[1] = ldw; g.K[1] = K2; }
    g.bias = bias; g.addend = addend; g.ldadd = ldadd; g.C = C; g.ldc = ldc; g.M = M; g.N = N;
    return launch_linear(g, st);
  }
};

}  // namespace dv3
