"""torch.autograd bindings of the C-ABI kernels (libdv3_b200.so).

Each Function allocates its outputs / saved activations / workspace as torch tensors (the
library owns no memory), enqueues the kernels on torch's current stream and, in backward,
turns the per-row deltas the recurrent kernels return into parameter gradients with split-K
tensor-core contractions over all rows (dW = delta^T @ input) -- those do not sit inside a time
loop.

Every contraction goes through ``gemm_tc`` on ``Split`` operands (tf32 hi/lo planes): planes are
made once per tensor -- by the kernel that produces the activation, or ``split`` -- and reused by
the forward product and both backward products in either storage order.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib as L

LN_EPS = 1e-3


def _empty(*shape, like=None, dtype=torch.float32, device=None):
    return torch.empty(*shape, dtype=dtype, device=device if device is not None else like.device)


def _c(t):
    return t if t is None or t.is_contiguous() else t.contiguous()


def _f32(t):
    return None if t is None else _c(t.to(torch.float32))


# --------------------------------------------------------------------------------------
# lambda return                                              (reference tools.py:682-728)
# --------------------------------------------------------------------------------------
class _LambdaReturn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, reward, value, pcont, bootstrap, lambda_):
        H, N = reward.shape
        reward, value, pcont, bootstrap = map(_f32, (reward, value, pcont, bootstrap))
        ret = torch.empty_like(reward)
        L.check(L.lib().dv3_lambda_return_fwd(L.fptr(reward), L.fptr(value), L.fptr(pcont),
                                              L.fptr(bootstrap), float(lambda_), H, N,
                                              L.fptr(ret), L.stream_ptr()), "lambda_return_fwd")
        ctx.save_for_backward(value, pcont, bootstrap, ret)
        ctx.lambda_ = float(lambda_)
        return ret

    @staticmethod
    def backward(ctx, g):
        value, pcont, bootstrap, ret = ctx.saved_tensors
        H, N = ret.shape
        g = _f32(g)
        d_r, d_v, d_c = torch.empty_like(ret), torch.empty_like(ret), torch.empty_like(ret)
        d_b = torch.empty_like(bootstrap)
        L.check(L.lib().dv3_lambda_return_bwd(L.fptr(value), L.fptr(pcont), L.fptr(bootstrap),
                                              L.fptr(ret), L.fptr(g), ctx.lambda_, H, N,
                                              L.fptr(d_r), L.fptr(d_v), L.fptr(d_c), L.fptr(d_b),
                                              L.stream_ptr()), "lambda_return_bwd")
        return d_r, d_v, d_c, d_b, None


def lambda_return_hn(reward, value, pcont, bootstrap, lambda_):
    """[H,N] time-major in, [H,N] out."""
    return _LambdaReturn.apply(reward, value, pcont, bootstrap, lambda_)


# --------------------------------------------------------------------------------------
# symlog two-hot                                             (reference tools.py:463-513)
# --------------------------------------------------------------------------------------
class _TwohotLogprob(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, x, buckets):
        K = logits.shape[-1]
        lg = _f32(logits).reshape(-1, K)
        xx = _f32(x).reshape(-1)
        R = lg.shape[0]
        out = _empty(R, like=lg)
        L.check(L.lib().dv3_twohot_logprob_fwd(L.fptr(lg), L.fptr(xx), L.fptr(buckets), R, K,
                                               L.fptr(out), L.stream_ptr()), "twohot_logprob_fwd")
        ctx.save_for_backward(lg, xx, buckets)
        ctx.shape = logits.shape
        return out.reshape(logits.shape[:-1])

    @staticmethod
    def backward(ctx, g):
        lg, xx, buckets = ctx.saved_tensors
        R, K = lg.shape
        g = _f32(g).reshape(-1)
        d = torch.empty_like(lg)
        L.check(L.lib().dv3_twohot_logprob_bwd(L.fptr(lg), L.fptr(xx), L.fptr(buckets), L.fptr(g),
                                               R, K, L.fptr(d), L.stream_ptr()),
                "twohot_logprob_bwd")
        return d.reshape(ctx.shape), None, None


class _TwohotMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, buckets):
        K = logits.shape[-1]
        lg = _f32(logits).reshape(-1, K)
        R = lg.shape[0]
        out = _empty(R, like=lg)
        L.check(L.lib().dv3_twohot_mean_fwd(L.fptr(lg), L.fptr(buckets), R, K, L.fptr(out),
                                            L.stream_ptr()), "twohot_mean_fwd")
        ctx.save_for_backward(lg, buckets)
        ctx.shape = logits.shape
        return out.reshape(tuple(logits.shape[:-1]) + (1,))

    @staticmethod
    def backward(ctx, g):
        lg, buckets = ctx.saved_tensors
        R, K = lg.shape
        g = _f32(g).reshape(-1)
        d = torch.empty_like(lg)
        L.check(L.lib().dv3_twohot_mean_bwd(L.fptr(lg), L.fptr(buckets), L.fptr(g), R, K,
                                            L.fptr(d), L.stream_ptr()), "twohot_mean_bwd")
        return d.reshape(ctx.shape), None


def twohot_logprob(logits, x, buckets):
    return _TwohotLogprob.apply(logits, x, buckets)


def twohot_mean(logits, buckets):
    return _TwohotMean.apply(logits, buckets)


# --------------------------------------------------------------------------------------
# KL balance                                              (reference networks.py:272-290)
# --------------------------------------------------------------------------------------
class _KLBalance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, post_logit, prior_logit, free, dyn_scale, rep_scale, unimix):
        S, Cc = post_logit.shape[-2:]
        lead = post_logit.shape[:-2]
        po = _f32(post_logit).reshape(-1, S, Cc)
        pr = _f32(prior_logit).reshape(-1, S, Cc)
        R = po.shape[0]
        outs = [_empty(R, like=po) for _ in range(6)]
        L.check(L.lib().dv3_kl_balance_fwd(L.fptr(po), L.fptr(pr), R, S, Cc, unimix, free,
                                           dyn_scale, rep_scale, *[L.fptr(o) for o in outs],
                                           L.stream_ptr()), "kl_balance_fwd")
        ctx.save_for_backward(po, pr)
        ctx.cfg = (free, dyn_scale, rep_scale, unimix, post_logit.shape)
        outs = [o.reshape(lead) for o in outs]
        ctx.mark_non_differentiable(*outs[1:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_loss, *_):
        po, pr = ctx.saved_tensors
        free, dyn_scale, rep_scale, unimix, shape = ctx.cfg
        R, S, Cc = po.shape
        g = _f32(g_loss).reshape(-1)
        d_po, d_pr = torch.empty_like(po), torch.empty_like(pr)
        L.check(L.lib().dv3_kl_balance_bwd(L.fptr(po), L.fptr(pr), L.fptr(g), R, S, Cc, unimix,
                                           free, dyn_scale, rep_scale, L.fptr(d_po), L.fptr(d_pr),
                                           L.stream_ptr()), "kl_balance_bwd")
        return d_po.reshape(shape), d_pr.reshape(shape), None, None, None, None


def kl_balance(post_logit, prior_logit, free, dyn_scale, rep_scale, unimix):
    """-> loss, value, dyn, rep, post_entropy, prior_entropy (each lead-shaped)."""
    return _KLBalance.apply(post_logit, prior_logit, float(free), float(dyn_scale),
                            float(rep_scale), float(unimix))


# --------------------------------------------------------------------------------------
# building blocks (used by tests and by the bulk actor backward)
# --------------------------------------------------------------------------------------
def ln_silu_fwd(pre, g, b, eps=LN_EPS, with_split=False):
    """SiLU(LayerNorm(pre)); with_split also returns the tf32 hi/lo planes of the result, written
    by the same kernel (the A operand of the next layer's GEMM)."""
    M, n = pre.shape
    out = torch.empty_like(pre)
    if with_split and n % 4 == 0:
        hi, lo = torch.empty_like(pre), torch.empty_like(pre)
        L.check(L.lib().dv3_ln_silu_fwd_split(L.fptr(pre), n, L.fptr(g), L.fptr(b), eps, M, n,
                                              L.fptr(out), n, L.fptr(hi), L.fptr(lo), n,
                                              L.stream_ptr()), "ln_silu_fwd_split")
        return out, Split(hi, lo, M, n)
    L.check(L.lib().dv3_ln_silu_fwd(L.fptr(pre), n, L.fptr(g), L.fptr(b), eps, M, n, L.fptr(out),
                                    n, L.stream_ptr()), "ln_silu_fwd")
    return (out, split(out)) if with_split else out


def ln_silu_bwd(pre, g, b, d_out, eps=LN_EPS, with_split=False):
    """-> (d_pre, d_ln): gradient w.r.t. the Linear output and w.r.t. the LN affine output
    (+ the Split of d_pre when with_split)."""
    M, n = pre.shape
    d_pre, d_ln = torch.empty_like(pre), torch.empty_like(pre)
    if with_split and n % 4 == 0:
        hi, lo = torch.empty_like(pre), torch.empty_like(pre)
        L.check(L.lib().dv3_ln_silu_bwd_split(L.fptr(pre), n, L.fptr(g), L.fptr(b), eps,
                                              L.fptr(d_out), n, M, n, L.fptr(d_pre), L.fptr(d_ln),
                                              n, L.fptr(hi), L.fptr(lo), n, L.stream_ptr()),
                "ln_silu_bwd_split")
        return d_pre, d_ln, Split(hi, lo, M, n)
    L.check(L.lib().dv3_ln_silu_bwd(L.fptr(pre), n, L.fptr(g), L.fptr(b), eps, L.fptr(d_out), n,
                                    M, n, L.fptr(d_pre), L.fptr(d_ln), n, L.stream_ptr()),
            "ln_silu_bwd")
    return (d_pre, d_ln, split(d_pre)) if with_split else (d_pre, d_ln)


def linear_fwd(a1, w1, a2=None, w2=None, bias=None, addend=None):
    """C = [a1|a2] [w1|w2]^T + bias + addend.  w* are [N,K*] (may be column views of one weight)."""
    M, K1 = a1.shape
    N = w1.shape[0]
    out = _empty(M, N, like=a1)
    K2 = a2.shape[1] if a2 is not None else 0
    for t in (a1, w1, a2, w2):
        if t is not None and t.stride(-1) != 1:
            raise L.Dv3Error("linear_fwd: innermost stride must be 1")
    as_f = lambda t: None if t is None else C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_float))
    L.check(L.lib().dv3_linear_fwd(as_f(a1), a1.stride(0), as_f(w1), w1.stride(0), K1,
                                   as_f(a2), a2.stride(0) if a2 is not None else 0,
                                   as_f(w2), w2.stride(0) if w2 is not None else 0, K2,
                                   L.fptr(bias), L.fptr(addend), N, L.fptr(out), N, M, N, 0,
                                   L.stream_ptr()), "linear_fwd")
    return out


class Split:
    """tf32 hi/lo planes of a 2-D fp32 tensor [rows, cols]: hi = x with the 13 low mantissa bits
    cleared, lo = x - hi (so hi + lo == x exactly).  Both planes are [rows, ld] with ld = cols
    rounded up to 4 (zero pad) so that every row is 16-byte aligned for TMA.  One Split serves
    every product the tensor takes part in: as [rows, K] ("K-major") or, read transposed, as
    [K, rows] ("MN-major") -- selected by descriptor bits in the kernel, never by copying."""
    __slots__ = ("hi", "lo", "rows", "cols")

    def __init__(self, hi, lo, rows, cols):
        self.hi, self.lo, self.rows, self.cols = hi, lo, rows, cols

    @property
    def ld(self):
        return self.hi.shape[1]

    def prefix(self, rows):
        """The first ``rows`` rows (a contiguous view of the same planes)."""
        return Split(self.hi[:rows], self.lo[:rows], rows, self.cols)

    def operand(self, transposed):
        o = L.TcOperand()
        o.hi, o.lo, o.ld, o.mn_major = _raw(self.hi), _raw(self.lo), self.ld, int(transposed)
        return o


def _raw(t):
    return None if t is None else C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_float))


def split(x):
    """x: 2-D fp32 CUDA tensor (row-strided views allowed) -> Split."""
    if isinstance(x, Split):
        return x
    if x.dim() != 2 or x.dtype != torch.float32 or x.stride(1) != 1:
        x = x.reshape(-1, x.shape[-1]).to(torch.float32).contiguous()
    rows, cols = x.shape
    ld = (cols + 3) & ~3
    hi = torch.empty(rows, ld, dtype=torch.float32, device=x.device)
    lo = torch.empty(rows, ld, dtype=torch.float32, device=x.device)
    if rows and cols:
        L.check(L.lib().dv3_split_tf32(_raw(x), x.stride(0), rows, cols, _raw(hi), _raw(lo), ld,
                                       L.stream_ptr()), "split_tf32")
    return Split(hi, lo, rows, cols)


_WEIGHT_EPOCH = [0]


def invalidate_weight_splits():
    """Called by whatever rewrites parameters in place without going through autograd's version
    counter (the fused Adam step, the slow-critic EMA)."""
    _WEIGHT_EPOCH[0] += 1


def split_param(w):
    """Split of a weight.  For an nn.Parameter the planes are cached on the parameter object
    itself and reused until the parameter changes: tools.Optimizer bumps a global epoch after
    every step (torch's fused Adam does not move the autograd version counter), and the version
    counter / storage pointer catch load_state_dict and re-assignment."""
    if not isinstance(w, torch.nn.Parameter):
        return split(w.detach())
    tag = (_WEIGHT_EPOCH[0], w._version, w.data_ptr(), tuple(w.shape))
    hit = getattr(w, "_dv3_split", None)
    if hit is not None and hit[0] == tag:
        return hit[1]
    sp = split(w.detach())
    w._dv3_split = (tag, sp)
    return sp


def gemm_tc(A, B, a_t=False, b_t=False, A2=None, bias=None, addend=None, out=None,
            accumulate=False, split_k=False):
    """C[M,N] = [op(A) | op(A2)] op(B)^T (+bias +addend, +out when accumulate) on the persistent
    tcgen05 3xTF32 kernel.  A, A2, B are Splits (tensors are split on the fly).  op(X) = X when
    the flag is False (X stored [rows, K]) and X^T when True (X stored [K, rows]).  So:
    y = x W^T -> gemm_tc(x, W);  dx = dy W -> gemm_tc(dy, W, b_t=True);
    dW = dy^T x -> gemm_tc(dy, x, a_t=True, b_t=True)."""
    A, B = split(A), split(B)
    M, K1 = (A.cols, A.rows) if a_t else (A.rows, A.cols)
    N, K = (B.cols, B.rows) if b_t else (B.rows, B.cols)
    K2 = 0
    a2 = None
    if A2 is not None:
        A2 = split(A2)
        K2 = A2.rows if a_t else A2.cols
        a2 = A2.operand(a_t)
    if K1 + K2 != K:
        raise L.Dv3Error(f"gemm_tc: contraction mismatch {K1}+{K2} vs {K}")
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.hi.device)
    if M == 0 or N == 0:
        return out
    if K == 0:
        if not accumulate:
            out.zero_()
        return out
    a1, b = A.operand(a_t), B.operand(b_t)
    L.check(L.lib().dv3_gemm_tc(C.byref(a1), K1, C.byref(a2) if a2 is not None else None, K2,
                                C.byref(b), L.fptr(bias), L.fptr(addend), N, _raw(out),
                                out.stride(0), M, N, int(accumulate) | (2 if split_k else 0),
                                L.stream_ptr()), "gemm_tc")
    return out


def linear_tc(a, w, bias=None, addend=None, trans_a=False, trans_w=False):
    """C[M,N] = op(a) op(w)^T (+bias +addend), fp32-accurate on the tensor cores.  op(a) is
    [M,K] (a stored [K,M] when trans_a); op(w) is [N,K] (w stored [K,N] when trans_w)."""
    return gemm_tc(a, w, a_t=trans_a, b_t=trans_w, bias=bias, addend=addend)


def linear_tc_fwd(a, w, bias=None, addend=None):
    return linear_tc(a, w, bias, addend)


def linear_tc2(a1, w, a2=None, bias=None, addend=None, out=None, accumulate=False):
    """C[M,N] = [a1|a2] w^T (+bias +addend) on the persistent raw-operand tcgen05 kernel
    (dv3_umma2.cu): a1 [M,K1], a2 [M,K2] | None, w [N,K1+K2]; row-strided 2-D views allowed
    (innermost stride 1, 16-byte aligned rows)."""
    M, K1 = a1.shape
    N = w.shape[0]
    K2 = a2.shape[1] if a2 is not None else 0
    if out is None:
        out = _empty(M, N, like=a1)
    if M == 0 or N == 0:
        return out
    for t in (a1, a2, w):
        if t is not None and (t.stride(-1) != 1 or t.dtype != torch.float32):
            raise L.Dv3Error("linear_tc2: fp32 operands with innermost stride 1 required")
    raw = lambda t: None if t is None else C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_float))
    L.check(L.lib().dv3_linear_tc2_fwd(raw(a1), a1.stride(0), K1, raw(a2),
                                       a2.stride(0) if a2 is not None else 0, K2, raw(w),
                                       w.stride(0), L.fptr(bias), L.fptr(addend), N, raw(out),
                                       out.stride(0), M, N, int(accumulate), L.stream_ptr()),
            "linear_tc2_fwd")
    return out


def _tag(t):
    return (t._version, t.data_ptr(), tuple(t.shape))


def attach_split(t, sp):
    """Remember the hi/lo planes of activation tensor ``t`` on the tensor object (valid while the
    tensor is not modified in place: version counter + storage pointer are checked)."""
    t._dv3_split = (_tag(t), sp)
    return t


def split_of_attached(t):
    hit = getattr(t, "_dv3_split", None)
    return hit[1] if hit is not None and hit[0] == _tag(t) else None


def split_of(x2d, src=None):
    """Split of a 2-D activation; reuses planes attached to ``src`` (the tensor the caller was
    handed, possibly a higher-rank view of the same memory) or to ``x2d`` itself."""
    for t in (src, x2d):
        hit = getattr(t, "_dv3_split", None) if t is not None else None
        if hit is not None and hit[0] == _tag(t) and hit[1].rows == x2d.shape[0] \
                and hit[1].cols == x2d.shape[1]:
            return hit[1]
    sp = split(x2d)
    if src is not None and src.is_contiguous():
        attach_split(src, sp)
    return sp


_LAST_OUT_SPLIT = [None]


class _DenseLnSilu(torch.autograd.Function):
    """SiLU(LayerNorm(x W^T)) for x [M,K], W [U,K]: the Linear(no bias)+LN(eps 1e-3)+SiLU block of
    the reference MLPs (networks.py:623-632).  All three contractions (y, dx, dW) run on the
    tensor-core GEMM from one split each of x, W and dy; LN/SiLU forward and backward on the row
    kernels."""

    @staticmethod
    def forward(ctx, x, W, g, b, xs):
        Ws = split_param(W)
        pre = gemm_tc(xs, Ws)
        out, osp = ln_silu_fwd(pre, _c(g.detach()), _c(b.detach()), with_split=True)
        _LAST_OUT_SPLIT[0] = osp
        ctx.save_for_backward(g.detach(), b.detach(), pre, xs.hi, xs.lo, Ws.hi, Ws.lo)
        ctx.shapes = (xs.rows, xs.cols, Ws.rows, Ws.cols)
        return out

    @staticmethod
    def backward(ctx, d_out):
        g, b, pre, xh, xl, Wh, Wl = ctx.saved_tensors
        M, Kd, U, _ = ctx.shapes
        xs, Ws = Split(xh, xl, M, Kd), Split(Wh, Wl, U, Kd)
        d_pre, d_ln, ds = ln_silu_bwd(pre, _c(g), _c(b), _f32(d_out), with_split=True)
        dx = dW = dg = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm_tc(ds, Ws, b_t=True)                    # dy W
        if ctx.needs_input_grad[1]:
            dW = gemm_tc(ds, xs, a_t=True, b_t=True, split_k=True)   # dy^T x
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            dg, db = _ln_grads(pre, d_ln)
        return dx, dW, dg, db, None


class _LinearBias(torch.autograd.Function):
    """x W^T + bias (the MLP output heads, networks.py:640-655) on the tensor-core GEMM."""

    @staticmethod
    def forward(ctx, x, W, bias, xs):
        Ws = split_param(W)
        ctx.save_for_backward(xs.hi, xs.lo, Ws.hi, Ws.lo)
        ctx.shapes = (xs.rows, xs.cols, Ws.rows, Ws.cols)
        return gemm_tc(xs, Ws, bias=None if bias is None else _c(bias.detach()))

    @staticmethod
    def backward(ctx, d_out):
        xh, xl, Wh, Wl = ctx.saved_tensors
        M, Kd, N, _ = ctx.shapes
        xs, Ws = Split(xh, xl, M, Kd), Split(Wh, Wl, N, Kd)
        d_out = _f32(d_out)
        ds = split(d_out)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm_tc(ds, Ws, b_t=True)
        if ctx.needs_input_grad[1]:
            dW = gemm_tc(ds, xs, a_t=True, b_t=True, split_k=True)
        if ctx.needs_input_grad[2]:
            db = d_out.sum(0)
        return dx, dW, db, None


def dense_ln_silu(x, W, g, b):
    """Linear(no bias) + LayerNorm + SiLU over the last axis.  The hi/lo planes of the result are
    attached to the returned tensor, so the next layer's GEMM starts without a split pass; the
    planes of ``x`` are looked up on ``x`` the same way (several heads reading one feature tensor
    split it once)."""
    lead = x.shape[:-1]
    x2 = _f32(x).reshape(-1, x.shape[-1])
    out = _DenseLnSilu.apply(x2, W, g, b, split_of(x2, x))
    osp, _LAST_OUT_SPLIT[0] = _LAST_OUT_SPLIT[0], None
    out = out.reshape(tuple(lead) + (W.shape[0],))
    return attach_split(out, osp) if osp is not None else out


def linear_bias(x, W, bias):
    lead = x.shape[:-1]
    x2 = _f32(x).reshape(-1, x.shape[-1])
    out = _LinearBias.apply(x2, W, bias, split_of(x2, x))
    return out.reshape(tuple(lead) + (W.shape[0],))


def onehot_sample(logits, u, unimix):
    """logits [M,S,C], u [M,S,C] or None (mode) -> (idx int32 [M,S], onehot fp32 [M,S,C])."""
    M, S, Cc = logits.shape
    idx = _empty(M, S, like=logits, dtype=torch.int32)
    hot = torch.empty_like(logits)
    L.check(L.lib().dv3_onehot_sample(L.fptr(logits), L.fptr(u), unimix, M, S, Cc, L.iptr(idx),
                                      L.fptr(hot), S * Cc, L.stream_ptr()), "onehot_sample")
    return idx, hot


def onehot_st_bwd(logits, g_sample, ext, unimix):
    M, S, Cc = logits.shape
    d = torch.empty_like(logits)
    L.check(L.lib().dv3_onehot_st_bwd(L.fptr(logits), L.fptr(g_sample), L.fptr(ext), unimix, M, S,
                                      Cc, L.fptr(d), L.stream_ptr()), "onehot_st_bwd")
    return d


def _xhat(pre):
    return F.layer_norm(pre, (pre.shape[-1],), None, None, LN_EPS)


def _ln_grads(pre2d, d_ln2d):
    """LayerNorm weight / bias gradients from the saved pre-LN rows and d(LN output)."""
    M, n = pre2d.shape
    if n > 2048 or pre2d.stride(1) != 1 or d_ln2d.stride(1) != 1:
        return (d_ln2d * _xhat(pre2d)).sum(0), d_ln2d.sum(0)
    dg = torch.empty(n, dtype=torch.float32, device=pre2d.device)
    db = torch.empty(n, dtype=torch.float32, device=pre2d.device)
    L.check(L.lib().dv3_ln_param_grads(_raw(pre2d), pre2d.stride(0), _raw(d_ln2d), d_ln2d.stride(0),
                                       LN_EPS, M, n, L.fptr(dg), L.fptr(db), L.stream_ptr()),
            "ln_param_grads")
    return dg, db


# --------------------------------------------------------------------------------------
# RSSM parameter pack
# --------------------------------------------------------------------------------------
def make_dims(stoch, classes, deter, hidden, actions, embed, unimix):
    return L.RssmDims(stoch, classes, deter, hidden, actions, embed, unimix, LN_EPS)


def pack_rssm(params):
    """params: list of 17 tensors ordered as L.RSSM_PARAM_FIELDS -> (struct, keepalive)."""
    keep = [_c(p.detach()) for p in params]
    st = L.RssmParams()
    for name, t in zip(L.RSSM_PARAM_FIELDS, keep):
        setattr(st, name, L.fptr(t))
    return st, keep


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------
# observe                                                 (reference networks.py:127-143)
# --------------------------------------------------------------------------------------
class _Observe(torch.autograd.Function):
    """inputs: embed [B,T,E], action [B,T,A], is_first [B,T], u_prior/u_post [T,B,S,C],
    state_idx int32 [B,S] | None, state_deter [B,D] | None, dims tuple, *17 RSSM params.
    outputs: post_stoch, post_logit, prior_stoch, prior_logit [B,T,S,C], deter [B,T,D],
    aprev [B,T,A] (the action after the is_first zeroing; non-differentiable)."""

    @staticmethod
    def forward(ctx, embed, action, is_first, u_prior, u_post, state_idx, state_deter, dims,
                *params):
        S, Cc, D, Hd, A, E, unimix = dims
        B, T = embed.shape[:2]
        dev = embed.device
        embed, action, is_first = _f32(embed), _f32(action), _f32(is_first)
        u_prior, u_post = _f32(u_prior), _f32(u_post)
        d = make_dims(S, Cc, D, Hd, A, E, unimix)
        pst, keep = pack_rssm(params)
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        i = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
        o = dict(post_stoch=f(B, T, S, Cc), post_logit=f(B, T, S, Cc), prior_stoch=f(B, T, S, Cc),
                 prior_logit=f(B, T, S, Cc), deter=f(B, T, D), post_idx=i(B, T, S),
                 prior_idx=i(B, T, S), first_eff=f(B, T), sprev_idx=i(B, T, S), hprev=f(B, T, D),
                 aprev=f(B, T, A), x_pre=f(B, T, Hd), x=f(B, T, Hd), g_pre=f(B, T, 3 * D),
                 y_pre=f(B, T, Hd), y=f(B, T, Hd), z_pre=f(B, T, Hd), z=f(B, T, Hd),
                 init_deter=f(D), init_ypre=f(Hd), init_y=f(Hd), init_logit=f(S * Cc),
                 init_idx=i(S))
        ws = _ws(L.lib().dv3_observe_workspace_bytes(C.byref(d), B, T), dev)
        io = L.fill(L.ObserveIO(), B=B, T=T, embed=embed, action=action, is_first=is_first,
                    u_prior=u_prior, u_post=u_post, state_idx=state_idx,
                    state_deter=_f32(state_deter), workspace=ws, workspace_bytes=ws.numel(), **o)
        L.check(L.lib().dv3_observe_fwd(C.byref(d), C.byref(pst), C.byref(io), L.stream_ptr()),
                "observe_fwd")
        ctx.dims = dims
        ctx.BT = (B, T)
        ctx.has_state = state_idx is not None
        ctx.save_for_backward(embed, o["first_eff"], o["post_logit"], o["prior_logit"],
                              o["hprev"], o["x_pre"], o["g_pre"], o["y_pre"], o["z_pre"],
                              o["sprev_idx"], o["aprev"], o["x"], o["y"], o["z"], o["deter"],
                              *keep)
        ctx.mark_non_differentiable(o["aprev"], o["post_idx"], o["prior_idx"])
        return (o["post_stoch"], o["post_logit"], o["prior_stoch"], o["prior_logit"], o["deter"],
                o["aprev"], o["post_idx"], o["prior_idx"])

    @staticmethod
    def backward(ctx, g_post_stoch, g_post_logit, g_prior_stoch, g_prior_logit, g_deter, *_):
        (embed, first_eff, post_logit, prior_logit, hprev, x_pre, g_pre, y_pre, z_pre, sprev_idx,
         aprev, x, y, z, deter, *params) = ctx.saved_tensors
        S, Cc, D, Hd, A, E, unimix = ctx.dims
        B, T = ctx.BT
        dev = embed.device
        SC = S * Cc
        d = make_dims(S, Cc, D, Hd, A, E, unimix)
        pst, keep = pack_rssm(params)
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        o = dict(d_embed=f(B, T, E), d_x_pre=f(B, T, Hd), d_x_ln=f(B, T, Hd),
                 d_g_pre=f(B, T, 3 * D), d_g_ln=f(B, T, 3 * D), d_y_pre=f(B, T, Hd),
                 d_y_ln=f(B, T, Hd), d_z_pre=f(B, T, Hd), d_z_ln=f(B, T, Hd),
                 d_post_logit=f(B, T, SC), d_prior_logit=f(B, T, SC), d_init_stoch=f(SC),
                 d_init_deter=f(D))
        d_state_deter = f(B, D) if ctx.has_state and ctx.needs_input_grad[6] else None
        ws = _ws(L.lib().dv3_observe_bwd_workspace_bytes(C.byref(d), B, T), dev)
        io = L.fill(L.ObserveBwdIO(), B=B, T=T, first_eff=first_eff, post_logit=post_logit,
                    prior_logit=prior_logit, hprev=hprev, x_pre=x_pre, g_pre=g_pre, y_pre=y_pre,
                    z_pre=z_pre, g_post_stoch=_f32(g_post_stoch), g_post_logit=_f32(g_post_logit),
                    g_prior_stoch=_f32(g_prior_stoch), g_prior_logit=_f32(g_prior_logit),
                    g_deter=_f32(g_deter), d_state_deter=d_state_deter, d_state_stoch=None,
                    workspace=ws, workspace_bytes=ws.numel(), **o)
        L.check(L.lib().dv3_observe_bwd(C.byref(d), C.byref(pst), C.byref(io), L.stream_ptr()),
                "observe_bwd")
        P = dict(zip(L.RSSM_PARAM_FIELDS, params))
        need = dict(zip(L.RSSM_PARAM_FIELDS, ctx.needs_input_grad[8:]))
        G = {k: None for k in L.RSSM_PARAM_FIELDS}
        if any(need.values()):
            r2 = lambda t: t.reshape(B * T, -1)
            hot = torch.empty(B * T, SC, dtype=torch.float32, device=dev)
            L.check(L.lib().dv3_idx_to_onehot(L.iptr(sprev_idx.reshape(B * T, S)), B * T, S, Cc,
                                              L.fptr(hot), SC, L.stream_ptr()), "idx_to_onehot")
            dx, dg, dy, dz = (split(r2(o[k])) for k in ("d_x_pre", "d_g_pre", "d_y_pre", "d_z_pre"))
            dpo, dpr = r2(o["d_post_logit"]), r2(o["d_prior_logit"])
            dpos, dprs = split(dpo), split(dpr)
            deters = split(r2(deter))
            # dW = delta^T @ input over all B*T rows: transposed-operand tcgen05 products with the
            # rows (K = B*T) partitioned over the SMs; column blocks of one weight are written
            # through strided output views
            dw = lambda d, inp, out=None: gemm_tc(d, inp, a_t=True, b_t=True, out=out, split_k=True)
            G["w_in"] = torch.empty(Hd, SC + A, dtype=torch.float32, device=dev)
            dw(dx, hot, G["w_in"][:, :SC])
            dw(dx, r2(aprev), G["w_in"][:, SC:])
            G["ln_in_g"], G["ln_in_b"] = _ln_grads(r2(x_pre), r2(o["d_x_ln"]))
            G["w_gru"] = torch.empty(3 * D, Hd + D, dtype=torch.float32, device=dev)
            dw(dg, r2(x), G["w_gru"][:, :Hd])
            dw(dg, r2(hprev), G["w_gru"][:, Hd:])
            G["ln_gru_g"], G["ln_gru_b"] = _ln_grads(r2(g_pre), r2(o["d_g_ln"]))
            G["w_out"] = dw(dy, deters)
            G["ln_out_g"], G["ln_out_b"] = _ln_grads(r2(y_pre), r2(o["d_y_ln"]))
            G["w_ims"] = dw(dprs, r2(y))
            G["b_ims"] = dpr.sum(0)
            G["w_obs"] = torch.empty(Hd, D + E, dtype=torch.float32, device=dev)
            dw(dz, deters, G["w_obs"][:, :D])
            dw(dz, r2(embed), G["w_obs"][:, D:])
            G["ln_obs_g"], G["ln_obs_b"] = _ln_grads(r2(z_pre), r2(o["d_z_ln"]))
            G["w_os"] = dw(dpos, r2(z))
            G["b_os"] = dpo.sum(0)
            # RSSM.initial (networks.py:99-125): tanh(W) -> prior head -> mode (straight-through
            # on the normalised log-probs).  One row; differentiated with autograd.
            names = ["w_init", "w_out", "ln_out_g", "ln_out_b", "w_ims", "b_ims"]
            with torch.enable_grad():
                leaf = {k: P[k].detach().requires_grad_(True) for k in names}
                deter0 = torch.tanh(leaf["w_init"])
                y0 = F.silu(F.layer_norm(deter0 @ leaf["w_out"].t(), (Hd,), leaf["ln_out_g"],
                                         leaf["ln_out_b"], LN_EPS))
                lg = (y0 @ leaf["w_ims"].t() + leaf["b_ims"]).reshape(S, Cc)
                if unimix > 0:
                    lg = torch.log(F.softmax(lg, -1) * (1.0 - unimix) + unimix / Cc)
                norm = lg - torch.logsumexp(lg, -1, keepdim=True)
                gi = torch.autograd.grad([norm.reshape(-1), deter0.reshape(-1)],
                                         [leaf[k] for k in names],
                                         [o["d_init_stoch"], o["d_init_deter"]])
            for k, g in zip(names, gi):
                G[k] = g if G[k] is None else G[k] + g
        grads = [G[k] if need[k] else None for k in L.RSSM_PARAM_FIELDS]
        d_embed = o["d_embed"] if ctx.needs_input_grad[0] else None
        return (d_embed, None, None, None, None, None, d_state_deter, None, *grads)


def observe(embed, action, is_first, u_prior, u_post, state_idx, state_deter, dims, params):
    return _Observe.apply(embed, action, is_first, u_prior, u_post, state_idx, state_deter, dims,
                          *params)


# --------------------------------------------------------------------------------------
# imagine                                                   (reference models.py:448-548)
# --------------------------------------------------------------------------------------
class ActorSpec:
    """Flat view of the actor MLP (networks.py:588-700) for the kernels."""

    def __init__(self, layers, units, dist, min_std, max_std, unimix):
        self.layers, self.units, self.dist = layers, units, dist
        self.min_std, self.max_std, self.unimix = min_std, max_std, unimix

    @property
    def n_params(self):
        return 3 * self.layers + (4 if self.dist == "normal" else 2)

    def pack(self, params):
        """params: [w_0, g_0, b_0, w_1, ...] + [w_mean, b_mean (, w_std, b_std)]."""
        keep = [_c(p.detach()) for p in params]
        Lr = self.layers
        w = L.float_ptr_array(keep[0:3 * Lr:3])
        g = L.float_ptr_array(keep[1:3 * Lr:3])
        b = L.float_ptr_array(keep[2:3 * Lr:3])
        a = L.Actor()
        a.layers, a.units = Lr, self.units
        a.dist = 0 if self.dist == "normal" else 1
        a.min_std, a.max_std, a.unimix = self.min_std, self.max_std, self.unimix
        a.w, a.ln_g, a.ln_b = w, g, b
        a.w_mean, a.b_mean = L.fptr(keep[3 * Lr]), L.fptr(keep[3 * Lr + 1])
        if self.dist == "normal":
            a.w_std, a.b_std = L.fptr(keep[3 * Lr + 2]), L.fptr(keep[3 * Lr + 3])
        return a, (keep, w, g, b)


class _Imagine(torch.autograd.Function):
    """inputs: start_idx int32 [N,S], start_deter [N,D], act_noise [H,N,A], u_state [H,N,S,C],
    given_action [H-1,N,A] | None, start_logit [N,S,C] | None, H, dims, actor spec | None, the 17
    RSSM params, then the actor params.  outputs: feat [H,N,F] (= [one-hot stoch | deter] of
    state k), logit [H,N,S,C] (row 0 = start_logit or zeros), action [H,N,A], idx int32 [H,N,S]."""

    @staticmethod
    def forward(ctx, start_idx, start_deter, act_noise, u_state, given_action, start_logit, H,
                dims, spec, *params):
        S, Cc, D, Hd, A, E, unimix = dims
        N = start_idx.shape[0]
        dev = start_deter.device
        SC, Fw = S * Cc, S * Cc + D
        rssm_params, actor_params = params[:17], params[17:]
        d = make_dims(S, Cc, D, Hd, A, E, unimix)
        pst, keep_r = pack_rssm(rssm_params)
        act_struct, keep_a = (spec.pack(actor_params) if spec is not None else (None, None))
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        U, Lr = (spec.units, spec.layers) if spec is not None else (0, 0)
        o = dict(feat=f(H, N, Fw), logit=f(H, N, S, Cc), action=f(H, N, A),
                 idx=torch.empty(H, N, S, dtype=torch.int32, device=dev),
                 x_pre=f(H, N, Hd), x=f(H, N, Hd), g_pre=f(H, N, 3 * D), y_pre=f(H, N, Hd),
                 y=f(H, N, Hd))
        if start_logit is not None:
            o["logit"][0].copy_(start_logit.detach().reshape(N, S, Cc))
        else:
            o["logit"][0].zero_()
        if spec is not None:
            o.update(a_pre=f(Lr, H, N, U), a_act=f(Lr, H, N, U), a_mean_raw=f(H, N, A))
            if spec.dist == "normal":
                o["a_std_raw"] = f(H, N, A)
        aptr = C.byref(act_struct) if act_struct is not None else None
        ws = _ws(L.lib().dv3_imagine_workspace_bytes(C.byref(d), aptr, N, H), dev)
        act_noise, u_state = _f32(act_noise), _f32(u_state)
        io = L.fill(L.ImagineIO(), N=N, H=H, start_idx=_c(start_idx),
                    start_deter=_f32(start_deter), act_noise=act_noise, u_state=u_state,
                    given_action=_f32(given_action), workspace=ws, workspace_bytes=ws.numel(),
                    **o)
        L.check(L.lib().dv3_imagine_fwd(C.byref(d), C.byref(pst), aptr, C.byref(io),
                                        L.stream_ptr()), "imagine_fwd")
        ctx.dims, ctx.spec, ctx.NH = dims, spec, (N, H)
        ctx.n_actor = len(actor_params)
        saved = [o["logit"], o["feat"], o["x_pre"], o["g_pre"], o["y_pre"], act_noise]
        if spec is not None:
            saved += [o["a_pre"], o["a_act"], o["a_mean_raw"]]
            if spec.dist == "normal":
                saved.append(o["a_std_raw"])
        ctx.n_saved = len(saved)
        ctx.save_for_backward(*saved, *[p.detach() for p in params])
        ctx.mark_non_differentiable(o["idx"])
        # the actor heads' raw outputs at every step are returned as differentiable outputs: the
        # caller builds the policy distribution (entropy, log-prob) from them instead of running
        # the actor MLP a second time over all H*N rows (reference models.py:349 re-evaluates)
        empty = o["action"].new_zeros(0)
        mean_raw = o["a_mean_raw"] if spec is not None else empty
        std_raw = o["a_std_raw"] if spec is not None and spec.dist == "normal" else empty
        return o["feat"], o["logit"], o["action"], o["idx"], mean_raw, std_raw

    @staticmethod
    def backward(ctx, g_feat, g_logit, g_action, _g_idx, g_mean_raw=None, g_std_raw=None):
        S, Cc, D, Hd, A, E, unimix = ctx.dims
        spec = ctx.spec
        N, H = ctx.NH
        SC = S * Cc
        saved = ctx.saved_tensors
        logit, feat, x_pre, g_pre, y_pre, act_noise = saved[:6]
        params = saved[ctx.n_saved:]
        rssm_params, actor_params = params[:17], params[17:]
        if spec is None:
            raise L.Dv3Error("imagine backward without an actor (imagine_with_action) is not "
                             "differentiated: the reference only uses it under no_grad-style "
                             "video prediction (models.py:196-204)")
        a_pre, a_act, a_mean_raw = saved[6:9]
        a_std_raw = saved[9] if spec.dist == "normal" else None
        dev = feat.device
        d = make_dims(S, Cc, D, Hd, A, E, unimix)
        pst, keep_r = pack_rssm(rssm_params)
        act_struct, keep_a = spec.pack(actor_params)
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        g_stoch = g_deter = None
        if g_feat is not None:
            g_feat = _f32(g_feat)
            g_stoch = g_feat[..., :SC].contiguous()
            g_deter = g_feat[..., SC:].contiguous()
        o = dict(d_mean_raw=f(H, N, A), d_x_pre=f(H, N, Hd), d_x_ln=f(H, N, Hd),
                 d_g_pre=f(H, N, 3 * D), d_g_ln=f(H, N, 3 * D), d_y_pre=f(H, N, Hd),
                 d_y_ln=f(H, N, Hd), d_logit=f(H, N, SC))
        if spec.dist == "normal":
            o["d_std_raw"] = f(H, N, A)
        ws = _ws(L.lib().dv3_imagine_bwd_workspace_bytes(C.byref(d), C.byref(act_struct), N, H),
                 dev)
        io = L.fill(L.ImagineBwdIO(), N=N, H=H, logit=logit, feat=feat, x_pre=x_pre, g_pre=g_pre,
                    y_pre=y_pre, a_mean_raw=a_mean_raw, a_std_raw=a_std_raw, act_noise=act_noise,
                    g_stoch=g_stoch, g_deter=g_deter, g_logit=_f32(g_logit),
                    g_action=_f32(g_action), d_start_stoch=None, d_start_deter=None,
                    workspace=ws, workspace_bytes=ws.numel(), **o)
        L.check(L.lib().dv3_imagine_bwd(C.byref(d), C.byref(pst), C.byref(act_struct),
                                        C.byref(io), L.stream_ptr()), "imagine_bwd")
        if any(ctx.needs_input_grad[9:9 + 17]):
            raise L.Dv3Error("imagine backward w.r.t. RSSM parameters is not implemented: the "
                             "reference freezes the world model while imagining "
                             "(models.py:335 RequiresGrad(self.actor))")
        # actor trunk backward over all H*N rows at once (its input feat is detached)
        Lr, U = spec.layers, spec.units
        HN = H * N
        dm = o["d_mean_raw"].reshape(HN, A)
        if g_mean_raw is not None and g_mean_raw.numel():
            dm = dm + _f32(g_mean_raw).reshape(HN, A)
        top = a_act[Lr - 1].reshape(HN, U)
        ga = [None] * len(actor_params)
        dw = lambda d, inp: gemm_tc(d, inp, a_t=True, b_t=True, split_k=True)
        tops = split(top)
        ga[3 * Lr], ga[3 * Lr + 1] = dw(dm, tops), dm.sum(0)
        d_act = dm @ actor_params[3 * Lr]
        if spec.dist == "normal":
            ds = o["d_std_raw"].reshape(HN, A)
            if g_std_raw is not None and g_std_raw.numel():
                ds = ds + _f32(g_std_raw).reshape(HN, A)
            ga[3 * Lr + 2], ga[3 * Lr + 3] = dw(ds, tops), ds.sum(0)
            d_act = d_act + ds @ actor_params[3 * Lr + 2]
        for i in range(Lr - 1, -1, -1):
            pre = a_pre[i].reshape(HN, U)
            d_pre, d_ln, dps = ln_silu_bwd(pre, actor_params[3 * i + 1], actor_params[3 * i + 2],
                                           _c(d_act), with_split=True)
            inp = feat.reshape(HN, -1) if i == 0 else a_act[i - 1].reshape(HN, U)
            ga[3 * i] = dw(dps, inp)
            ga[3 * i + 1], ga[3 * i + 2] = _ln_grads(pre, d_ln)
            if i > 0:
                d_act = gemm_tc(dps, split(actor_params[3 * i]), b_t=True)
        need = ctx.needs_input_grad[9 + 17:]
        ga = [g if n else None for g, n in zip(ga, need)]
        return (None, None, None, None, None, None, None, None, None, *([None] * 17), *ga)


def imagine_full(start_idx, start_deter, act_noise, u_state, given_action, H, dims, spec,
                 rssm_params, actor_params, start_logit=None):
    """-> feat, logit, action, idx, actor mean_raw [H,N,A], actor std_raw [H,N,A] (empty for a
    one-hot actor / no actor); the last two are differentiable w.r.t. the actor parameters."""
    return _Imagine.apply(start_idx, start_deter, act_noise, u_state, given_action, start_logit, H,
                          dims, spec, *rssm_params, *actor_params)


def imagine(start_idx, start_deter, act_noise, u_state, given_action, H, dims, spec, rssm_params,
            actor_params, start_logit=None):
    """-> feat [H,N,F], logit [H,N,S,C], action [H,N,A], idx int32 [H,N,S]."""
    return imagine_full(start_idx, start_deter, act_noise, u_state, given_action, H, dims, spec,
                        rssm_params, actor_params, start_logit)[:4]
