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

import contextlib
import ctypes as C

import os

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


# --------------------------------------------------------------------------------------
# gradient sink: while tools.Optimizer runs a backward pass, the parameter gradients computed by
# the Functions below (dW = delta^T x split-K products, LayerNorm affine sums, bias column sums)
# are ACCUMULATED STRAIGHT INTO views of the optimizer's flat, pre-zeroed gradient buffer instead
# of being returned to autograd -- no per-product memset, no gather of ~100 gradient tensors
# afterwards (reference tools.py:760-776 has clip_grad_norm_ / Adam walk the per-parameter grads).
# --------------------------------------------------------------------------------------
_SINK = {}          # id(parameter) -> view of the flat gradient buffer (armed optimizers only)
_SINK_DIRTY = set() # ids whose view has been written in this backward pass


def arm_grad_sink(params, views):
    for p, v in zip(params, views):
        _SINK[id(p)] = v


def disarm_grad_sink(params):
    written = []
    for p in params:
        written.append(id(p) in _SINK_DIRTY)
        _SINK.pop(id(p), None)
        _SINK_DIRTY.discard(id(p))
    return written


_SUNK = object()     # marker: this gradient went into the sink (autograd gets None)

# Parameter-gradient work that nothing downstream in the backward pass depends on (the bulk dW
# products of the RSSM after the observe recurrence) runs on a side stream, i.e. as a parallel branch
# of the captured step graph under the encoder's backward; in data-parallel runs the all-reduce of
# that slice of the flat gradient follows on the same stream.  The optimizer joins the stream before
# it reads the gradient (join_grad_streams).  DV3_OBS_DW_SIDE=0 keeps everything on one stream.
_GRAD_STREAMS = {}
_GRAD_PENDING = []   # (stream, tensors the branch reads that were allocated on another stream)


def grad_side_enabled():
    return os.environ.get("DV3_OBS_DW_SIDE", "1") != "0"


def grad_side_stream(device):
    dev = torch.device(device)
    st = _GRAD_STREAMS.get(dev)
    if st is None:
        st = _GRAD_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return st


def join_grad_streams():
    cur = torch.cuda.current_stream()
    for st, _keep in _GRAD_PENDING:
        cur.wait_stream(st)
    _GRAD_PENDING.clear()


def _sink_of(pid):
    return _SINK.get(pid) if pid is not None else None


def _sink_gemm(pid, d, inp, cols=None):
    """dW = d^T inp accumulated into the sink view of parameter ``pid`` (optionally into a column
    block of it); returns None when there is no armed sink (caller computes a fresh tensor)."""
    view = _sink_of(pid)
    if view is None:
        return None
    out = view if cols is None else view[:, cols[0]:cols[1]]
    gemm_tc(d, inp, a_t=True, b_t=True, out=out, accumulate=True, split_k=True)
    _SINK_DIRTY.add(pid)
    return view


def _sink_add(pid, fn_out, fresh):
    """Vector gradients (bias / LayerNorm sums).  ``fn_out(out)`` writes the gradient into ``out``
    (overwriting); ``fresh()`` returns it as a new tensor.  -> None if sunk, else the tensor."""
    view = _sink_of(pid)
    if view is None:
        return fresh()
    if pid in _SINK_DIRTY:
        view.add_(fresh().reshape(view.shape))
    else:
        fn_out(view)
        _SINK_DIRTY.add(pid)
    return None


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
    flat = getattr(w, "_dv3_flat_split", None)
    if flat is not None and flat[0] == (w._version, w.data_ptr()):
        return flat[1]          # planes written by the fused Adam step that wrote the weight
    tag = (_WEIGHT_EPOCH[0], w._version, w.data_ptr(), tuple(w.shape))
    hit = getattr(w, "_dv3_split", None)
    if hit is not None and hit[0] == tag:
        return hit[1]
    sp = split(w.detach())
    w._dv3_split = (tag, sp)
    return sp


def transpose_param(w):
    """(W^T as a contiguous fp32 [K, N] tensor, its Split), cached on the parameter per weight
    epoch like split_param: the one-hot gather form of a Linear over [one-hot | dense] inputs."""
    tag = (_WEIGHT_EPOCH[0], w._version, w.data_ptr(), tuple(w.shape))
    hit = getattr(w, "_dv3_wt", None) if isinstance(w, torch.nn.Parameter) else None
    if hit is not None and hit[0] == tag:
        return hit[1]
    wd = w.detach()
    N, Kd = wd.shape
    wt = torch.empty(Kd, N, dtype=torch.float32, device=wd.device)
    L.check(L.lib().dv3_transpose(_raw(wd), wd.stride(0), N, Kd, L.fptr(wt), L.stream_ptr()),
            "transpose")
    out = (wt, split(wt))
    if isinstance(w, torch.nn.Parameter):
        w._dv3_wt = (tag, out)
    return out


def _operand(o, sp):
    o.hi, o.lo, o.ld, o.mn_major = _raw(sp.hi), _raw(sp.lo), sp.ld, 0


def rssm_planes(params):
    """dv3_rssm_planes for the 17 RSSM parameters (L.RSSM_PARAM_FIELDS order): the weight forms
    the imagination calls read, made once per optimizer step (planes written by the fused Adam
    step where available) instead of inside every call.  -> (struct, keepalive)."""
    P = dict(zip(L.RSSM_PARAM_FIELDS, params))
    pl = L.RssmPlanes()
    keep = []
    for name in ("w_gru", "w_out", "w_ims"):
        if P[name].shape[1] % 4:
            return None, None
        sp = split_param(P[name])
        _operand(getattr(pl, name), sp)
        keep.append(sp)
    wt, wtsp = transpose_param(P["w_in"])
    pl.w_in_t = L.fptr(wt)
    _operand(pl.w_in_t_sp, wtsp)
    keep += [wt, wtsp]
    return pl, keep


def actor_planes(spec, params):
    """dv3_actor_planes for the flat actor parameter list [w_0, g_0, b_0, w_1, ...]."""
    pl = L.ActorPlanes()
    keep = []
    for i in range(spec.layers):
        w = params[3 * i]
        if w.shape[1] % 4:
            return None, None
        sp = split_param(w)
        _operand(pl.w[i], sp)
        keep.append(sp)
    wt, _ = transpose_param(params[0])
    pl.w0_t = L.fptr(wt)
    keep.append(wt)
    return pl, keep


def split_param_like(w):
    """Split of a detached weight tensor (saved for backward): no parameter object to cache on."""
    return split(w)


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


def gemm_tc_rawa(a, B, b_t=False, a2=None, bias=None, addend=None, out=None):
    """C[M,N] = [a | a2] op(B)^T (+bias +addend) with the A operand as plain fp32 rows: the tile is
    split inside the SM and fed to the tensor core from tensor memory (dv3_umma2t.cu); B is a Split
    (tensors are split on the fly).  Bit-identical to gemm_tc on the same tile; M >= 64."""
    B = split(B)
    M, K1 = a.shape
    K2 = a2.shape[1] if a2 is not None else 0
    N, K = (B.cols, B.rows) if b_t else (B.rows, B.cols)
    if K1 + K2 != K:
        raise L.Dv3Error(f"gemm_tc_rawa: contraction mismatch {K1}+{K2} vs {K}")
    if a.stride(1) != 1 or (a2 is not None and a2.stride(1) != 1):
        raise L.Dv3Error("gemm_tc_rawa: A rows must be contiguous")
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    b = B.operand(b_t)
    L.check(L.lib().dv3_gemm_tc_rawa(_raw(a), a.stride(0), K1, _raw(a2),
                                     a2.stride(0) if a2 is not None else 0, K2, C.byref(b),
                                     L.fptr(bias), _raw(addend),
                                     addend.stride(0) if addend is not None else 0, _raw(out),
                                     out.stride(0), M, N, L.stream_ptr()), "gemm_tc_rawa")
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
        ctx.pids = (id(W), id(g), id(b))
        return out

    @staticmethod
    def backward(ctx, d_out):
        g, b, pre, xh, xl, Wh, Wl = ctx.saved_tensors
        M, Kd, U, _ = ctx.shapes
        wid, gid, bid = ctx.pids
        xs, Ws = Split(xh, xl, M, Kd), Split(Wh, Wl, U, Kd)
        d_pre, d_ln, ds = ln_silu_bwd(pre, _c(g), _c(b), _f32(d_out), with_split=True)
        dx = dW = dg = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm_tc(ds, Ws, b_t=True)                    # dy W
        if ctx.needs_input_grad[1] and _sink_gemm(wid, ds, xs) is None:
            dW = gemm_tc(ds, xs, a_t=True, b_t=True, split_k=True)   # dy^T x
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            dg, db = _ln_grads(pre, d_ln, gid, bid)
        return dx, dW, dg, db, None


class _LinearBias(torch.autograd.Function):
    """x W^T + bias (the MLP output heads, networks.py:640-655) on the tensor-core GEMM."""

    @staticmethod
    def forward(ctx, x, W, bias, xs):
        Ws = split_param(W)
        ctx.save_for_backward(xs.hi, xs.lo, Ws.hi, Ws.lo)
        ctx.shapes = (xs.rows, xs.cols, Ws.rows, Ws.cols)
        ctx.pids = (id(W), id(bias) if bias is not None else None)
        return gemm_tc(xs, Ws, bias=None if bias is None else _c(bias.detach()))

    @staticmethod
    def backward(ctx, d_out):
        xh, xl, Wh, Wl = ctx.saved_tensors
        M, Kd, N, _ = ctx.shapes
        wid, bid = ctx.pids
        xs, Ws = Split(xh, xl, M, Kd), Split(Wh, Wl, N, Kd)
        d_out = _f32(d_out)
        ds = split(d_out)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm_tc(ds, Ws, b_t=True)
        if ctx.needs_input_grad[1] and _sink_gemm(wid, ds, xs) is None:
            dW = gemm_tc(ds, xs, a_t=True, b_t=True, split_k=True)
        if ctx.needs_input_grad[2]:
            db = _bias_grad(d_out, bid)
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


def _ln_grads_into(pre2d, d_ln2d, dg, db, accumulate=False):
    M, n = pre2d.shape
    fn = L.lib().dv3_ln_param_grads_acc if accumulate else L.lib().dv3_ln_param_grads
    L.check(fn(_raw(pre2d), pre2d.stride(0), _raw(d_ln2d), d_ln2d.stride(0), LN_EPS, M, n,
               L.fptr(dg), L.fptr(db), L.stream_ptr()), "ln_param_grads")


def _ln_grads(pre2d, d_ln2d, gid=None, bid=None):
    """LayerNorm weight / bias gradients from the saved pre-LN rows and d(LN output).  With armed
    sinks for both parameters (ids ``gid`` / ``bid``) the kernel writes into the flat gradient
    buffer and (None, None) is returned."""
    M, n = pre2d.shape
    kernel_ok = n <= 2048 and pre2d.stride(1) == 1 and d_ln2d.stride(1) == 1
    vg, vb = _sink_of(gid), _sink_of(bid)
    if vg is not None and vb is not None and kernel_ok:
        _ln_grads_into(pre2d, d_ln2d, vg, vb, accumulate=True)   # the sink is pre-zeroed
        _SINK_DIRTY.update((gid, bid))
        return None, None
    if not kernel_ok:
        dg, db = (d_ln2d * _xhat(pre2d)).sum(0), d_ln2d.sum(0)
    else:
        dg = torch.empty(n, dtype=torch.float32, device=pre2d.device)
        db = torch.empty(n, dtype=torch.float32, device=pre2d.device)
        _ln_grads_into(pre2d, d_ln2d, dg, db)
    if vg is not None and vb is not None:
        vg.add_(dg)
        vb.add_(db)
        _SINK_DIRTY.update((gid, bid))
        return None, None
    return dg, db


def col_sum(d2d, out=None, accumulate=False):
    """out[j] (+)= sum_r d2d[r, j]."""
    M, n = d2d.shape
    if d2d.stride(1) != 1:
        d2d = d2d.contiguous()
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=d2d.device)
    L.check(L.lib().dv3_col_sum(_raw(d2d), d2d.stride(0), M, n, L.fptr(out), int(accumulate),
                                L.stream_ptr()), "col_sum")
    return out


def _bias_grad(d2d, pid=None):
    """Column sums of a delta [M,n] = the gradient of a Linear bias (into the armed sink, which is
    pre-zeroed, or as a fresh tensor)."""
    view = _sink_of(pid)
    if view is None:
        return col_sum(d2d)
    col_sum(d2d, view, accumulate=True)
    _SINK_DIRTY.add(pid)
    return None


# --------------------------------------------------------------------------------------
# RSSM parameter pack
# --------------------------------------------------------------------------------------
def make_dims(stoch, classes, deter, hidden, actions, embed, unimix):
    return L.RssmDims(stoch, classes, deter, hidden, actions, embed, unimix, LN_EPS)


def pack_rssm(params, planes=None):
    """params: list of 17 tensors ordered as L.RSSM_PARAM_FIELDS -> (struct, keepalive);
    ``planes``: a dv3_rssm_planes struct to attach (kept alive by the caller)."""
    keep = [_c(p.detach()) for p in params]
    st = L.RssmParams()
    for name, t in zip(L.RSSM_PARAM_FIELDS, keep):
        setattr(st, name, L.fptr(t))
    if planes is not None:
        st.planes = C.pointer(planes)
        keep.append(planes)
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
        ctx.set_materialize_grads(False)     # unused outputs: None, not a zero-filled tensor
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
        ctx.pids = [id(q) for q in params]
        ctx.has_state = state_idx is not None
        ctx.save_for_backward(embed, o["first_eff"], o["post_logit"], o["prior_logit"],
                              o["hprev"], o["x_pre"], o["g_pre"], o["y_pre"], o["z_pre"],
                              o["sprev_idx"], o["aprev"], o["x"], o["y"], o["z"], o["deter"],
                              o["init_deter"], o["init_ypre"], o["init_y"], o["init_logit"], *keep)
        ctx.mark_non_differentiable(o["aprev"], o["post_idx"], o["prior_idx"])
        return (o["post_stoch"], o["post_logit"], o["prior_stoch"], o["prior_logit"], o["deter"],
                o["aprev"], o["post_idx"], o["prior_idx"])

    @staticmethod
    def backward(ctx, g_post_stoch, g_post_logit, g_prior_stoch, g_prior_logit, g_deter, *_):
        (embed, first_eff, post_logit, prior_logit, hprev, x_pre, g_pre, y_pre, z_pre, sprev_idx,
         aprev, x, y, z, deter, init_deter, init_ypre, init_y, init_logit, *params) = ctx.saved_tensors
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
        pid = dict(zip(L.RSSM_PARAM_FIELDS, ctx.pids))
        need = dict(zip(L.RSSM_PARAM_FIELDS, ctx.needs_input_grad[8:]))
        G = {k: None for k in L.RSSM_PARAM_FIELDS}
        side = None
        if any(need.values()) and grad_side_enabled() and all(
                _sink_of(pid[k]) is not None for k in L.RSSM_PARAM_FIELDS if need[k]):
            # every RSSM gradient goes into the armed sink and nothing downstream reads it: fork
            side = grad_side_stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            _GRAD_PENDING.append((side, (o, ctx.saved_tensors, ws)))
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
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
                # rows (K = B*T) partitioned over the SMs, accumulated into the optimizer's flat
                # gradient buffer when it is armed (else into fresh tensors); column blocks of one
                # weight are written through strided output views
                def dw(name, d, inp, cols=None):
                    if _sink_gemm(pid[name], d, inp, cols) is not None:
                        G[name] = _SUNK
                    elif cols is None:
                        G[name] = gemm_tc(d, inp, a_t=True, b_t=True, split_k=True)
                    else:
                        if G[name] is None:
                            G[name] = torch.empty(P[name].shape, dtype=torch.float32, device=dev)
                        gemm_tc(d, inp, a_t=True, b_t=True, out=G[name][:, cols[0]:cols[1]], split_k=True)

                def ln(gname, bname, pre, dln):
                    a, b_ = _ln_grads(r2(pre), r2(dln), pid[gname], pid[bname])
                    G[gname], G[bname] = (_SUNK, _SUNK) if a is None else (a, b_)

                def bias(name, d2d):
                    r = _bias_grad(d2d, pid[name])
                    G[name] = _SUNK if r is None else r

                dw("w_in", dx, hot, (0, SC))
                dw("w_in", dx, r2(aprev), (SC, SC + A))
                ln("ln_in_g", "ln_in_b", x_pre, o["d_x_ln"])
                dw("w_gru", dg, r2(x), (0, Hd))
                dw("w_gru", dg, r2(hprev), (Hd, Hd + D))
                ln("ln_gru_g", "ln_gru_b", g_pre, o["d_g_ln"])
                dw("w_out", dy, deters)
                ln("ln_out_g", "ln_out_b", y_pre, o["d_y_ln"])
                dw("w_ims", dprs, r2(y))
                bias("b_ims", dpr)
                dw("w_obs", dz, deters, (0, D))
                dw("w_obs", dz, r2(embed), (D, D + E))
                ln("ln_obs_g", "ln_obs_b", z_pre, o["d_z_ln"])
                dw("w_os", dpos, r2(z))
                bias("b_os", dpo)
                # RSSM.initial (networks.py:99-125): tanh(W) -> prior head -> mode (straight-through
                # on the normalised log-probs).  One row: dv3_rssm_initial_bwd adds its six parameter
                # gradients onto the bulk sums (the armed sink views, or the tensors built above).
                names = ["w_init", "w_out", "ln_out_g", "ln_out_b", "w_ims", "b_ims"]
                bufs = {}
                for k in names:
                    view = _sink_of(pid[k])
                    if view is not None:
                        bufs[k] = view
                        _SINK_DIRTY.add(pid[k])
                        G[k] = _SUNK
                    else:
                        if G[k] is None or G[k] is _SUNK:
                            G[k] = torch.zeros(P[k].shape, dtype=torch.float32, device=dev)
                        bufs[k] = G[k]
                scr = f(int(L.lib().dv3_rssm_initial_bwd_scratch_floats(C.byref(d))))
                L.check(L.lib().dv3_rssm_initial_bwd(
                    C.byref(d), C.byref(pst), L.fptr(init_deter), L.fptr(init_ypre), L.fptr(init_y),
                    L.fptr(init_logit), L.fptr(o["d_init_stoch"]), L.fptr(o["d_init_deter"]),
                    L.fptr(bufs["w_init"]), L.fptr(bufs["w_out"]), L.fptr(bufs["ln_out_g"]),
                    L.fptr(bufs["ln_out_b"]), L.fptr(bufs["w_ims"]), L.fptr(bufs["b_ims"]), L.fptr(scr),
                    L.stream_ptr()), "rssm_initial_bwd")
        grads = [G[k] if need[k] and G[k] is not _SUNK else None for k in L.RSSM_PARAM_FIELDS]
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

    def pack(self, params, planes=None):
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
        if planes is not None:
            a.planes = C.pointer(planes)
        return a, (keep, w, g, b, planes)


class _Imagine(torch.autograd.Function):
    """inputs: start_idx int32 [N,S], start_deter [N,D], act_noise [H,N,A], u_state [H,N,S,C],
    given_action [H-1,N,A] | None, start_logit [N,S,C] | None, H, dims, actor spec | None, the 17
    RSSM params, then the actor params.  outputs: feat [H,N,F] (= [one-hot stoch | deter] of
    state k), logit [H,N,S,C] (row 0 = start_logit or zeros), action [H,N,A], idx int32 [H,N,S]."""

    @staticmethod
    def forward(ctx, start_idx, start_deter, act_noise, u_state, given_action, start_logit, H,
                dims, spec, *params):
        ctx.set_materialize_grads(False)     # unused outputs: None, not a zero-filled tensor
        S, Cc, D, Hd, A, E, unimix = dims
        N = start_idx.shape[0]
        dev = start_deter.device
        SC, Fw = S * Cc, S * Cc + D
        rssm_params, actor_params = params[:17], params[17:]
        d = make_dims(S, Cc, D, Hd, A, E, unimix)
        # weight forms made once per optimizer step (cached per parameter), not inside the call
        rpl = apl = keep_pl = None
        if N >= 64:
            rpl, k1 = rssm_planes(rssm_params)
            apl, k2 = actor_planes(spec, actor_params) if spec is not None else (None, None)
            keep_pl = (rpl, k1, apl, k2)
        ctx.planes = keep_pl
        pst, keep_r = pack_rssm(rssm_params, rpl)
        act_struct, keep_a = (spec.pack(actor_params, apl) if spec is not None else (None, None))
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
        ctx.pids_actor = [id(q) for q in actor_params]
        saved = [o["logit"], o["feat"], o["x_pre"], o["g_pre"], o["y_pre"], act_noise]
        if spec is not None:
            saved += [o["a_pre"], o["a_act"], o["a_mean_raw"]]
            if spec.dist == "normal":
                saved.append(o["a_std_raw"])
        # tf32 planes of the feature buffer: the A operand of every head that reads the rollout
        # (reward, cont, critic, slow critic) and of the actor's first-layer dW in backward
        fsp = split(o["feat"].reshape(H * N, Fw)) if (spec is not None and H * N >= 64) else None
        ctx.has_fsp = fsp is not None
        if fsp is not None:
            saved += [fsp.hi, fsp.lo]
        ctx.n_saved = len(saved)
        ctx.save_for_backward(*saved, *[p.detach() for p in params])
        ctx.mark_non_differentiable(o["idx"])
        # the actor heads' raw outputs at every step are returned as differentiable outputs: the
        # caller builds the policy distribution (entropy, log-prob) from them instead of running
        # the actor MLP a second time over all H*N rows (reference models.py:349 re-evaluates)
        empty = o["action"].new_zeros(0)
        mean_raw = o["a_mean_raw"] if spec is not None else empty
        std_raw = o["a_std_raw"] if spec is not None and spec.dist == "normal" else empty
        fh, fl = (fsp.hi, fsp.lo) if fsp is not None else (empty, empty)
        ctx.mark_non_differentiable(fh, fl)
        return o["feat"], o["logit"], o["action"], o["idx"], mean_raw, std_raw, fh, fl

    @staticmethod
    def backward(ctx, g_feat, g_logit, g_action, _g_idx, g_mean_raw=None, g_std_raw=None, *_):
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
        feat_sp = None
        if ctx.has_fsp:
            feat_sp = Split(saved[ctx.n_saved - 2], saved[ctx.n_saved - 1], H * N, SC + D)
        dev = feat.device
        d = make_dims(S, Cc, D, Hd, A, E, unimix)
        rpl = ctx.planes[0] if ctx.planes is not None else None     # the forward's (weights unchanged)
        pst, keep_r = pack_rssm(rssm_params, rpl)
        act_struct, keep_a = spec.pack(actor_params)
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        # the gradient of the feature buffer is handed over as it is: its two column ranges are the
        # gradients of stoch and deter (row pitch S*C + D), no split copies of 94 MB
        g_state_ld = 0
        g_packed = None
        if g_feat is not None:
            g_feat = _c(_f32(g_feat))
            if os.environ.get("DV3_BWD_PACKED_G") == "1":      # the ABI's packed form (tests)
                g_packed = (g_feat[..., :SC].contiguous(), g_feat[..., SC:].contiguous())
            else:
                g_state_ld = SC + D
        o = dict(d_mean_raw=f(H, N, A), d_x_pre=f(H, N, Hd), d_x_ln=f(H, N, Hd),
                 d_g_pre=f(H, N, 3 * D), d_g_ln=f(H, N, 3 * D), d_y_pre=f(H, N, Hd),
                 d_y_ln=f(H, N, Hd), d_logit=f(H, N, SC))
        if spec.dist == "normal":
            o["d_std_raw"] = f(H, N, A)
        ws = _ws(L.lib().dv3_imagine_bwd_workspace_bytes(C.byref(d), C.byref(act_struct), N, H),
                 dev)
        io = L.fill(L.ImagineBwdIO(), N=N, H=H, logit=logit, feat=feat, x_pre=x_pre, g_pre=g_pre,
                    y_pre=y_pre, a_mean_raw=a_mean_raw, a_std_raw=a_std_raw, act_noise=act_noise,
                    g_stoch=None, g_deter=None, g_logit=_f32(g_logit),
                    g_action=_f32(g_action), d_start_stoch=None, d_start_deter=None,
                    workspace=ws, workspace_bytes=ws.numel(), g_state_ld=g_state_ld, **o)
        if g_packed is not None:
            io.g_stoch, io.g_deter = L.fptr(g_packed[0]), L.fptr(g_packed[1])
        elif g_feat is not None:
            fp = C.POINTER(C.c_float)
            io.g_stoch = C.cast(C.c_void_p(g_feat.data_ptr()), fp)
            io.g_deter = C.cast(C.c_void_p(g_feat.data_ptr() + 4 * SC), fp)
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
        pids = ctx.pids_actor

        def dw(i, d, inp):
            if _sink_gemm(pids[i], d, inp) is None:
                ga[i] = gemm_tc(d, inp, a_t=True, b_t=True, split_k=True)

        tops = split(top)
        dms = split(dm)
        dw(3 * Lr, dms, tops)
        ga[3 * Lr + 1] = _bias_grad(dm, pids[3 * Lr + 1])
        d_act = gemm_tc(dms, split_param_like(actor_params[3 * Lr]), b_t=True)      # dm W_mean
        if spec.dist == "normal":
            ds = o["d_std_raw"].reshape(HN, A)
            if g_std_raw is not None and g_std_raw.numel():
                ds = ds + _f32(g_std_raw).reshape(HN, A)
            dss = split(ds)
            dw(3 * Lr + 2, dss, tops)
            ga[3 * Lr + 3] = _bias_grad(ds, pids[3 * Lr + 3])
            gemm_tc(dss, split_param_like(actor_params[3 * Lr + 2]), b_t=True, out=d_act,
                    accumulate=True)                                                # + ds W_std
        for i in range(Lr - 1, -1, -1):
            pre = a_pre[i].reshape(HN, U)
            d_pre, d_ln, dps = ln_silu_bwd(pre, actor_params[3 * i + 1], actor_params[3 * i + 2],
                                           _c(d_act), with_split=True)
            if i == 0:
                inp = feat_sp if feat_sp is not None else feat.reshape(HN, -1)
            else:
                inp = a_act[i - 1].reshape(HN, U)
            dw(3 * i, dps, inp)
            ga[3 * i + 1], ga[3 * i + 2] = _ln_grads(pre, d_ln, pids[3 * i + 1], pids[3 * i + 2])
            if i > 0:
                d_act = gemm_tc(dps, split(actor_params[3 * i]), b_t=True)
        need = ctx.needs_input_grad[9 + 17:]
        ga = [g if n else None for g, n in zip(ga, need)]
        return (None, None, None, None, None, None, None, None, None, *([None] * 17), *ga)


def imagine_full(start_idx, start_deter, act_noise, u_state, given_action, H, dims, spec,
                 rssm_params, actor_params, start_logit=None):
    """-> feat, logit, action, idx, actor mean_raw [H,N,A], actor std_raw [H,N,A] (empty for a
    one-hot actor / no actor; both differentiable w.r.t. the actor parameters), Split of feat
    viewed [H*N, F] (None for fewer than 64 rows / no actor)."""
    out = _Imagine.apply(start_idx, start_deter, act_noise, u_state, given_action, start_logit, H,
                         dims, spec, *rssm_params, *actor_params)
    feat = out[0]
    sp = None
    if out[6].numel():
        sp = Split(out[6], out[7], feat.shape[0] * feat.shape[1], feat.shape[2])
    return out[:6] + (sp,)


def imagine(start_idx, start_deter, act_noise, u_state, given_action, H, dims, spec, rssm_params,
            actor_params, start_logit=None):
    """-> feat [H,N,F], logit [H,N,S,C], action [H,N,A], idx int32 [H,N,S]."""
    return imagine_full(start_idx, start_deter, act_noise, u_state, given_action, H, dims, spec,
                        rssm_params, actor_params, start_logit)[:4]


# --------------------------------------------------------------------------------------
# step tail (dv3_tail.cu): fused element-wise / reduction chains around the rollouts
# --------------------------------------------------------------------------------------
def symlog(x):
    """tools.symlog (reference tools.py:22-23) as one kernel; no gradient (observation inputs)."""
    x = _f32(x)
    out = torch.empty_like(x)
    L.check(L.lib().dv3_symlog(L.fptr(x), x.numel(), L.fptr(out), L.stream_ptr()), "symlog")
    return out


class _SqerrLogprob(torch.autograd.Function):
    """log_prob of the squared-error heads: tools.SymlogDist (reference tools.py:546-572) when
    ``use_symlog`` else tools.MSEDist (520-543).  mode / value [..., *event]; the event axes
    (everything after the first two) are summed."""

    @staticmethod
    def forward(ctx, mode, value, use_symlog, tol):
        lead = tuple(mode.shape[:2])
        R = lead[0] * lead[1]
        m2, v2 = _f32(mode).reshape(R, -1), _f32(value).reshape(R, -1)
        n = m2.shape[1]
        out = _empty(R, like=m2)
        L.check(L.lib().dv3_sqerr_logprob_fwd(L.fptr(m2), L.fptr(v2), R, n, int(use_symlog), tol,
                                              L.fptr(out), L.stream_ptr()), "sqerr_logprob_fwd")
        ctx.save_for_backward(m2, v2)
        ctx.cfg = (int(use_symlog), tol, mode.shape)
        return out.reshape(lead)

    @staticmethod
    def backward(ctx, g):
        m2, v2 = ctx.saved_tensors
        use_symlog, tol, shape = ctx.cfg
        R, n = m2.shape
        d = torch.empty_like(m2)
        L.check(L.lib().dv3_sqerr_logprob_bwd(L.fptr(m2), L.fptr(v2), L.fptr(_f32(g).reshape(R)), R, n,
                                              use_symlog, tol, L.fptr(d), L.stream_ptr()),
                "sqerr_logprob_bwd")
        return d.reshape(shape), None, None, None


def sqerr_logprob(mode, value, use_symlog, tol=1e-8):
    return _SqerrLogprob.apply(mode, value, bool(use_symlog), float(tol))


class _BernoulliLogprob(torch.autograd.Function):
    """tools.Bernoulli.log_prob (reference tools.py:604-628), element-wise."""

    @staticmethod
    def forward(ctx, logits, x):
        lg, xv = _f32(logits).reshape(-1), _f32(x).reshape(-1)
        out = torch.empty_like(lg)
        L.check(L.lib().dv3_bernoulli_logprob_fwd(L.fptr(lg), L.fptr(xv), lg.numel(), L.fptr(out),
                                                  L.stream_ptr()), "bernoulli_logprob_fwd")
        ctx.save_for_backward(lg, xv)
        ctx.shape = logits.shape
        return out.reshape(logits.shape)

    @staticmethod
    def backward(ctx, g):
        lg, xv = ctx.saved_tensors
        d = torch.empty_like(lg)
        L.check(L.lib().dv3_bernoulli_logprob_bwd(L.fptr(lg), L.fptr(xv), L.fptr(_f32(g).reshape(-1)),
                                                  lg.numel(), L.fptr(d), L.stream_ptr()),
                "bernoulli_logprob_bwd")
        return d.reshape(ctx.shape), None


def bernoulli_logprob(logits, x):
    return _BernoulliLogprob.apply(logits, x)


class _LossMean(torch.autograd.Function):
    """mean_r sum_i scales[i] * terms[i][r] -> 0-dim tensor (reference models.py:140-152: the
    scaled per-head losses plus the KL loss, averaged over the batch)."""

    @staticmethod
    def forward(ctx, scales, *terms):
        ts = [_f32(t).reshape(-1) for t in terms]
        R = ts[0].numel()
        if any(t.numel() != R for t in ts):
            raise L.Dv3Error("loss_mean: terms differ in size")
        out = _empty(1, like=ts[0])
        neg = torch.empty(len(ts), R, dtype=torch.float32, device=ts[0].device)
        sc = (C.c_float * len(ts))(*scales)
        ptrs = L.float_ptr_array(ts)
        L.check(L.lib().dv3_loss_mean_fwd(ptrs, sc, len(ts), R, L.fptr(out), L.fptr(neg),
                                          L.stream_ptr()), "loss_mean_fwd")
        ctx.cfg = (tuple(scales), R, [t.shape for t in terms])
        ctx.like = ts[0]
        ctx.mark_non_differentiable(neg)
        return out.reshape(()), neg

    @staticmethod
    def backward(ctx, g, _gneg):
        scales, R, shapes = ctx.cfg
        grads = torch.empty(len(scales), R, dtype=torch.float32, device=ctx.like.device)
        sc = (C.c_float * len(scales))(*scales)
        L.check(L.lib().dv3_loss_mean_bwd(L.fptr(_f32(g).reshape(1)), sc, len(scales), R, L.fptr(grads),
                                          L.stream_ptr()), "loss_mean_bwd")
        return (None, *[grads[i].reshape(s) for i, s in enumerate(shapes)])


def loss_mean(terms, scales):
    """-> (mean 0-dim, neg [n_terms, R] = -terms, no gradient)."""
    return _LossMean.apply(tuple(float(s) for s in scales), *terms)


class _DiscountWeights(torch.autograd.Function):
    """discount = gamma * sigmoid(cont_logit), weights = cumprod([1, discount[:-1]]) over the time
    axis (reference models.py:620-638); weights carry no gradient."""

    @staticmethod
    def forward(ctx, cont_logit, gamma):
        ctx.set_materialize_grads(False)
        H = cont_logit.shape[0]
        lg = _f32(cont_logit).reshape(H, -1)
        N = lg.shape[1]
        disc, w = torch.empty_like(lg), torch.empty_like(lg)
        L.check(L.lib().dv3_discount_weights_fwd(L.fptr(lg), gamma, H, N, L.fptr(disc), L.fptr(w),
                                                 L.stream_ptr()), "discount_weights_fwd")
        ctx.save_for_backward(lg)
        ctx.cfg = (gamma, cont_logit.shape)
        disc, w = disc.reshape(cont_logit.shape), w.reshape(cont_logit.shape)
        ctx.mark_non_differentiable(w)          # the very tensor object that is returned
        return disc, w

    @staticmethod
    def backward(ctx, g_disc, _gw):
        if g_disc is None:
            return None, None
        (lg,) = ctx.saved_tensors
        gamma, shape = ctx.cfg
        d = torch.empty_like(lg)
        L.check(L.lib().dv3_discount_bwd(L.fptr(lg), L.fptr(_f32(g_disc).reshape(lg.shape)), gamma,
                                         lg.numel(), L.fptr(d), L.stream_ptr()), "discount_bwd")
        return d.reshape(shape), None


def discount_weights(cont_logit, gamma):
    return _DiscountWeights.apply(cont_logit, float(gamma))


REWARD_EMA_MAX = 1 << 20


def reward_ema(x, ema_vals, alpha):
    """RewardEMA.__call__ (reference models.py:11-26): updates ``ema_vals`` [2] in place and returns
    a [2] tensor (offset, scale).  x: any shape, at most REWARD_EMA_MAX elements."""
    xf = _f32(x.detach()).reshape(-1)
    out = _empty(2, like=xf)
    L.check(L.lib().dv3_reward_ema(L.fptr(xf), xf.numel(), float(alpha), L.fptr(ema_vals), L.fptr(out),
                                   L.stream_ptr()), "reward_ema")
    return out


class _ActorLoss(torch.autograd.Function):
    """mean over the (H-1)*N leading elements of -w * actor_target - c_ent * entropy, with
    actor_target = normed_target - normed_base ('dynamics', mode 0) or logp * (target - base)
    ('reinforce', mode 1); reference models.py:393-397, 640-681.  Also returns normed_target."""

    @staticmethod
    def forward(ctx, target, base, weights, entropy, logp, offset_scale, c_ent, mode):
        Hm = target.shape[0]
        H = weights.shape[0]
        tg, bs = _f32(target).reshape(Hm, -1), _f32(base).reshape(Hm, -1)
        N = tg.shape[1]
        w, en = _f32(weights).reshape(H, N), _f32(entropy).reshape(H, N)
        lp = _f32(logp).reshape(H, N) if logp is not None else None
        loss = _empty(1, like=tg)
        normed = torch.empty_like(tg)
        L.check(L.lib().dv3_actor_loss_fwd(L.fptr(tg), L.fptr(bs), L.fptr(w), L.fptr(en), L.fptr(lp),
                                           L.fptr(offset_scale), c_ent, mode, Hm * N, L.fptr(normed),
                                           L.fptr(loss), L.stream_ptr()), "actor_loss_fwd")
        ctx.save_for_backward(tg, bs, w, offset_scale if offset_scale is not None else loss)
        ctx.cfg = (c_ent, mode, Hm, H, N, target.shape, entropy.shape,
                   logp.shape if logp is not None else None, offset_scale is not None)
        normed = normed.reshape(target.shape)
        ctx.mark_non_differentiable(normed)
        return loss.reshape(()), normed

    @staticmethod
    def backward(ctx, g, _gn):
        tg, bs, w, os_ = ctx.saved_tensors
        c_ent, mode, Hm, H, N, tshape, eshape, lshape, has_os = ctx.cfg
        d_t = torch.empty_like(tg) if mode == 0 else None
        d_e = torch.empty(H, N, dtype=torch.float32, device=tg.device)
        d_l = torch.empty(H, N, dtype=torch.float32, device=tg.device) if mode == 1 else None
        L.check(L.lib().dv3_actor_loss_bwd(L.fptr(_f32(g).reshape(1)), L.fptr(tg), L.fptr(bs), L.fptr(w),
                                           L.fptr(os_) if has_os else None, c_ent, mode, Hm * N, H * N,
                                           L.fptr(d_t), L.fptr(d_e), L.fptr(d_l), L.stream_ptr()),
                "actor_loss_bwd")
        return (d_t.reshape(tshape) if d_t is not None else None, None, None, d_e.reshape(eshape),
                d_l.reshape(lshape) if d_l is not None else None, None, None, None)


def actor_loss(target, base, weights, entropy, logp, offset_scale, c_ent, mode):
    """-> (loss 0-dim, normed_target like target).  mode: 'dynamics' | 'reinforce'."""
    return _ActorLoss.apply(target, base, weights, entropy, logp, offset_scale, float(c_ent),
                            {"dynamics": 0, "reinforce": 1}[mode])


class _ValueLoss(torch.autograd.Function):
    """mean_i w_i * (-lp_target_i - lp_slow_i) over [H-1,N] (reference models.py:419-429)."""

    @staticmethod
    def forward(ctx, lp_target, lp_slow, weights):
        a = _f32(lp_target).reshape(-1)
        b = _f32(lp_slow).reshape(-1) if lp_slow is not None else None
        cnt = a.numel()
        w = _f32(weights).reshape(-1)
        if w.numel() < cnt:
            raise L.Dv3Error("value_loss: weights shorter than the log-probs")
        loss = _empty(1, like=a)
        L.check(L.lib().dv3_value_loss_fwd(L.fptr(a), L.fptr(b), L.fptr(w), cnt, L.fptr(loss),
                                           L.stream_ptr()), "value_loss_fwd")
        ctx.save_for_backward(w)
        ctx.cfg = (cnt, lp_target.shape, lp_slow.shape if lp_slow is not None else None)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        cnt, s1, s2 = ctx.cfg
        d = torch.empty(cnt, dtype=torch.float32, device=w.device)
        L.check(L.lib().dv3_value_loss_bwd(L.fptr(_f32(g).reshape(1)), L.fptr(w), cnt, L.fptr(d),
                                           L.stream_ptr()), "value_loss_bwd")
        return d.reshape(s1), (d.reshape(s2) if s2 is not None else None), None


def value_loss(lp_target, lp_slow, weights):
    return _ValueLoss.apply(lp_target, lp_slow, weights)


class _NormalPolicy(torch.autograd.Function):
    """Entropy and log_prob(action) of the 'normal' actor distribution from its raw head outputs
    (reference networks.py:693-700, tools.py:575-601): -> entropy [lead], logp [lead]."""

    @staticmethod
    def forward(ctx, mean_raw, std_raw, action, min_std, max_std, want_logp):
        ctx.set_materialize_grads(False)
        A = mean_raw.shape[-1]
        lead = mean_raw.shape[:-1]
        mr, sr, ac = (_f32(t).reshape(-1, A) for t in (mean_raw, std_raw, action))
        R = mr.shape[0]
        ent = _empty(R, like=mr)
        lp = _empty(R, like=mr) if want_logp else None
        L.check(L.lib().dv3_normal_policy_fwd(L.fptr(mr), L.fptr(sr), L.fptr(ac), min_std, max_std, R, A,
                                              L.fptr(ent), L.fptr(lp), L.stream_ptr()), "normal_policy_fwd")
        ctx.save_for_backward(mr, sr, ac)
        ctx.cfg = (min_std, max_std, want_logp, mean_raw.shape)
        if lp is None:
            lp = ent.new_zeros(0)
            ctx.mark_non_differentiable(lp)
            return ent.reshape(lead), lp
        return ent.reshape(lead), lp.reshape(lead)

    @staticmethod
    def backward(ctx, g_ent, g_lp):
        mr, sr, ac = ctx.saved_tensors
        min_std, max_std, want_logp, shape = ctx.cfg
        R, A = mr.shape
        d_m, d_s = torch.empty_like(mr), torch.empty_like(mr)
        use_lp = want_logp and g_lp is not None
        d_a = torch.empty_like(mr) if (use_lp and ctx.needs_input_grad[2]) else None
        L.check(L.lib().dv3_normal_policy_bwd(
            L.fptr(mr), L.fptr(sr), L.fptr(ac), L.fptr(_f32(g_ent).reshape(R)) if g_ent is not None else None,
            L.fptr(_f32(g_lp).reshape(R)) if use_lp else None, min_std, max_std, R, A, L.fptr(d_m),
            L.fptr(d_s), L.fptr(d_a), L.stream_ptr()), "normal_policy_bwd")
        return (d_m.reshape(shape), d_s.reshape(shape), d_a.reshape(shape) if d_a is not None else None,
                None, None, None)


def normal_policy(mean_raw, std_raw, action, min_std, max_std, want_logp):
    ent, lp = _NormalPolicy.apply(mean_raw, std_raw, action, float(min_std), float(max_std),
                                  bool(want_logp))
    return ent, (lp if want_logp else None)


def tensorstats4(x):
    """-> [4] tensor: mean, unbiased std, min, max (reference tools.py:949-958)."""
    xf = _f32(x.detach()).reshape(-1)
    out = _empty(4, like=xf)
    L.check(L.lib().dv3_tensorstats(L.fptr(xf), xf.numel(), L.fptr(out), L.stream_ptr()), "tensorstats")
    return out


def ema_mix(dst_flat, src_flat, mix):
    """dst = mix * src + (1 - mix) * dst over flat fp32 buffers (slow critic, models.py:683-689)."""
    L.check(L.lib().dv3_ema_mix(L.fptr(dst_flat), L.fptr(src_flat), dst_flat.numel(), float(mix),
                                L.stream_ptr()), "ema_mix")


# --------------------------------------------------------------------------------------
# 4x4 stride-2 conv / transposed conv blocks of the image encoder / decoder (dv3_conv.cu)
# --------------------------------------------------------------------------------------
def im2col_split(x2d, n, H, W, Cc):
    """x2d: [n*H*W, C] channels-last -> Split of the patch matrix [n*H/2*W/2, 16*C]."""
    rows, cols = n * (H // 2) * (W // 2), 16 * Cc
    hi = torch.empty(rows, cols, dtype=torch.float32, device=x2d.device)
    lo = torch.empty(rows, cols, dtype=torch.float32, device=x2d.device)
    L.check(L.lib().dv3_im2col_s2k4(L.fptr(x2d), n, H, W, Cc, None, L.fptr(hi), L.fptr(lo),
                                    L.stream_ptr()), "im2col_s2k4")
    return Split(hi, lo, rows, cols)


def col2im(cols2d, n, H, W, Cc, bias=None, shift=0.0):
    """cols2d [n*H/2*W/2, 16*C] -> [n*H*W, C] (adjoint of the patch gather) + bias + shift."""
    out = torch.empty(n * H * W, Cc, dtype=torch.float32, device=cols2d.device)
    L.check(L.lib().dv3_col2im_s2k4(L.fptr(cols2d), n, H, W, Cc, L.fptr(bias), float(shift),
                                    L.fptr(out), L.stream_ptr()), "col2im_s2k4")
    return out


def _conv_w2(W):
    """Conv2d weight [Cout,Cin,4,4] -> [Cout, (ky,kx,ci)]; ConvTranspose2d weight [Cin,Cout,4,4] ->
    [Cin, (ky,kx,co)]: the same permutation.  Cached per weight epoch like split_param."""
    tag = (_WEIGHT_EPOCH[0], W._version, W.data_ptr(), tuple(W.shape))
    hit = getattr(W, "_dv3_w2", None)
    if hit is not None and hit[0] == tag:
        return hit[1]
    w2 = W.detach().permute(0, 2, 3, 1).reshape(W.shape[0], 16 * W.shape[1]).contiguous()
    sp = split(w2)
    W._dv3_w2 = (tag, sp)
    return sp


def _w2_grad_to_param(dW2, shape):
    """[A, (ky,kx,b)] -> [A, b, 4, 4] (view)."""
    return dW2.view(shape[0], 4, 4, shape[1]).permute(0, 3, 1, 2)


def _sink_w2(pid, dW2, shape):
    """Add a permuted conv-weight gradient into the armed sink; False if there is none."""
    view = _sink_of(pid)
    if view is None:
        return False
    view.add_(_w2_grad_to_param(dW2, shape))
    _SINK_DIRTY.add(pid)
    return True


class _ConvLnSilu(torch.autograd.Function):
    """Conv2dSamePad(k=4, s=2, no bias) -> channel LayerNorm -> SiLU on channels-last rows
    (reference networks.py:463-487): im2col -> tcgen05 GEMM -> LN/SiLU row kernel."""

    @staticmethod
    def forward(ctx, x2d, geom, W, g, b, need_dx):
        n, H, Wd = geom
        Cout, Cin = W.shape[0], W.shape[1]
        cols = im2col_split(_f32(x2d), n, H, Wd, Cin)
        w2 = _conv_w2(W)
        pre = gemm_tc(cols, w2)                                     # [Mo, Cout]
        out, osp = ln_silu_fwd(pre, _c(g.detach()), _c(b.detach()), with_split=True)
        _LAST_OUT_SPLIT[0] = osp
        ctx.save_for_backward(cols.hi, cols.lo, w2.hi, w2.lo, pre, g.detach(), b.detach())
        ctx.cfg = (geom, tuple(W.shape), (id(W), id(g), id(b)), need_dx)
        return out

    @staticmethod
    def backward(ctx, d_out):
        ch, cl, wh, wl, pre, g, b = ctx.saved_tensors
        (n, H, Wd), wshape, (wid, gid, bid), need_dx = ctx.cfg
        Cout, Cin = wshape[0], wshape[1]
        Mo = pre.shape[0]
        cols, w2 = Split(ch, cl, Mo, 16 * Cin), Split(wh, wl, Cout, 16 * Cin)
        d_pre, d_ln, ds = ln_silu_bwd(pre, _c(g), _c(b), _f32(d_out), with_split=True)
        dW = dg = db = dx = None
        if ctx.needs_input_grad[2]:
            dW2 = gemm_tc(ds, cols, a_t=True, b_t=True, split_k=True)     # [Cout, 16 Cin]
            if not _sink_w2(wid, dW2, wshape):
                dW = _w2_grad_to_param(dW2, wshape)
        if ctx.needs_input_grad[3] or ctx.needs_input_grad[4]:
            dg, db = _ln_grads(pre, d_ln, gid, bid)
        if need_dx and ctx.needs_input_grad[0]:
            dcols = gemm_tc(ds, w2, b_t=True)                             # [Mo, 16 Cin]
            dx = col2im(dcols, n, H, Wd, Cin)
        return dx, None, dW, dg, db, None


def conv_ln_silu(x2d, geom, W, g, b, need_dx=True):
    """-> [n*H/2*W/2, Cout] with its tf32 planes attached (the next block's im2col does not need
    them, the decoder's first product does)."""
    out = _ConvLnSilu.apply(x2d, geom, W, g, b, need_dx)
    osp, _LAST_OUT_SPLIT[0] = _LAST_OUT_SPLIT[0], None
    return attach_split(out, osp) if osp is not None else out


class _DeconvBlock(torch.autograd.Function):
    """ConvTranspose2d(k=4, s=2, padding=1) [+ bias] [-> channel LayerNorm -> SiLU] on channels-last
    rows (reference networks.py:533-560): tcgen05 GEMM -> col2im gather -> LN/SiLU row kernel."""

    @staticmethod
    def forward(ctx, x2d, xs, geom, W, g, b, bias, shift):
        n, h, w = geom
        Cin, Cout = W.shape[0], W.shape[1]
        w2 = _conv_w2(W)                                            # [Cin, (ky,kx,co)]
        cols = gemm_tc(xs, w2, b_t=True)                            # [Mi, 16 Cout]
        pre = col2im(cols, n, 2 * h, 2 * w, Cout, None if bias is None else _c(bias.detach()), shift)
        norm = g is not None
        if norm:
            out, osp = ln_silu_fwd(pre, _c(g.detach()), _c(b.detach()), with_split=True)
            _LAST_OUT_SPLIT[0] = osp
            ctx.save_for_backward(xs.hi, xs.lo, w2.hi, w2.lo, pre, g.detach(), b.detach())
        else:
            out = pre
            ctx.save_for_backward(xs.hi, xs.lo, w2.hi, w2.lo)
        ctx.cfg = (geom, tuple(W.shape), (id(W), id(g) if norm else None, id(b) if norm else None,
                                          id(bias) if bias is not None else None), norm,
                   (xs.rows, xs.cols))
        return out

    @staticmethod
    def backward(ctx, d_out):
        (n, h, w), wshape, (wid, gid, bid, biasid), norm, (Mi, Kx) = ctx.cfg
        Cin, Cout = wshape[0], wshape[1]
        if norm:
            xh, xl, wh, wl, pre, g, b = ctx.saved_tensors
            d_pre, d_ln = ln_silu_bwd(pre, _c(g), _c(b), _f32(d_out))
        else:
            xh, xl, wh, wl = ctx.saved_tensors
            d_pre = _c(_f32(d_out)).reshape(n * 4 * h * w, Cout)
        xs, w2 = Split(xh, xl, Mi, Kx), Split(wh, wl, Cin, 16 * Cout)
        dcols = im2col_split(d_pre, n, 2 * h, 2 * w, Cout)          # [Mi, 16 Cout]
        dx = dW = dg = db = dbias = None
        if ctx.needs_input_grad[0]:
            dx = gemm_tc(dcols, w2)                                 # [Mi, Cin]
        if ctx.needs_input_grad[3]:
            dW2 = gemm_tc(xs, dcols, a_t=True, b_t=True, split_k=True)    # [Cin, 16 Cout]
            if not _sink_w2(wid, dW2, wshape):
                dW = _w2_grad_to_param(dW2, wshape)
        if norm and (ctx.needs_input_grad[4] or ctx.needs_input_grad[5]):
            dg, db = _ln_grads(pre, d_ln, gid, bid)
        if biasid is not None and ctx.needs_input_grad[6]:
            dbias = _bias_grad(d_pre, biasid)
        return dx, None, None, dW, dg, db, dbias, None


def deconv_block(x2d, geom, W, g=None, b=None, bias=None, shift=0.0):
    """x2d [n*h*w, Cin] -> [n*2h*2w, Cout]."""
    x2 = _f32(x2d)
    out = _DeconvBlock.apply(x2, split_of(x2, x2d), geom, W, g, b, bias, float(shift))
    if g is not None:
        osp, _LAST_OUT_SPLIT[0] = _LAST_OUT_SPLIT[0], None
        if osp is not None:
            attach_split(out, osp)
    return out
