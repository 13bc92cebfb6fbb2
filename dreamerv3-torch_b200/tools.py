"""Host-side mirror of the reference's ``tools`` surface that sits on the training hot path.

Same call signatures and return conventions as the reference (file:line cited per item), with
the loops and the distribution arithmetic running in the sm_100a kernels of libdv3_b200.so:

* ``lambda_return``          reference tools.py:702-728 (+682-699): returns a tuple of N [T,1]
* ``DiscDist``               reference tools.py:463-517 (symlog two-hot head)
* ``OneHotDist``             reference tools.py:436-460 (unimix categorical, straight-through)
* ``Optimizer``              reference tools.py:731-783, plus the data-parallel gradient allreduce
* ``symlog`` / ``symexp``    reference tools.py:22-27
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import kernels as K


def symlog(x):
    return torch.sign(x) * torch.log(torch.abs(x) + 1.0)


def symexp(x):
    return torch.sign(x) * (torch.exp(torch.abs(x)) - 1.0)


# --------------------------------------------------------------------------------------
# lambda return
# --------------------------------------------------------------------------------------
def lambda_return_stacked(reward, value, pcont, bootstrap, lambda_):
    """Time-major [H,N,1] (or [H,N]) in -> same shape out; one kernel, differentiable."""
    shape = reward.shape
    H = shape[0]
    if isinstance(pcont, (int, float)):
        pcont = pcont * torch.ones_like(reward)
    if bootstrap is None:
        bootstrap = torch.zeros_like(value[-1])
    N = int(np.prod(shape[1:])) if len(shape) > 1 else 1
    ret = K.lambda_return_hn(reward.reshape(H, N), value.reshape(H, N), pcont.reshape(H, N),
                             bootstrap.reshape(N), lambda_)
    return ret.reshape(shape)


def lambda_return(reward, value, pcont, bootstrap, lambda_, axis):
    """Drop-in for the reference call: a tuple of N tensors of shape [T,1] (its callers re-stack
    them with ``torch.stack(target, dim=1)``).  Only the layout the reference produces is
    supported: time on ``axis`` 0 with [T,N,1] operands (models.py:627-634)."""
    if len(reward.shape) != len(value.shape):
        raise AssertionError((reward.shape, value.shape))
    if axis != 0:
        perm = list(range(reward.dim()))
        perm[0], perm[axis] = perm[axis], perm[0]
        reward, value = reward.permute(perm), value.permute(perm)
        if not isinstance(pcont, (int, float)):
            pcont = pcont.permute(perm)
    if reward.dim() != 3 or reward.shape[-1] != 1:
        raise NotImplementedError(f"lambda_return expects [T,N,1] operands, got {tuple(reward.shape)}")
    ret = lambda_return_stacked(reward, value, pcont, bootstrap, lambda_)   # [T,N,1]
    return torch.unbind(ret.permute(1, 0, 2), dim=0)


# --------------------------------------------------------------------------------------
# distributions
# --------------------------------------------------------------------------------------
_BUCKETS = {}


def _buckets(device, low=-20.0, high=20.0, steps=255):
    key = (str(device), low, high, steps)
    if key not in _BUCKETS:
        # built with torch.linspace so the values are the ones the reference's own tensor holds
        _BUCKETS[key] = torch.linspace(low, high, steps=steps, device=device)
    return _BUCKETS[key]


class DiscDist:
    """255-bucket two-hot regression head over symlog space (reference tools.py:463-517)."""

    def __init__(self, logits, low=-20.0, high=20.0, transfwd=symlog, transbwd=symexp, device=None):
        if transfwd is not symlog or transbwd is not symexp:
            raise NotImplementedError("the kernel implements the symlog/symexp transform pair")
        self.logits = logits
        self.buckets = _buckets(logits.device, low, high)

    @property
    def probs(self):
        return torch.softmax(self.logits, -1)

    def mean(self):
        return K.twohot_mean(self.logits, self.buckets)

    def mode(self):
        return K.twohot_mean(self.logits, self.buckets)

    def log_prob(self, x):
        # x is [...] (reward, models.py:140) or [..., 1] (value target, models.py:423)
        lead = tuple(self.logits.shape[:-1])
        if tuple(x.shape) not in (lead, lead + (1,)):
            raise AssertionError((x.shape, self.logits.shape))
        return K.twohot_logprob(self.logits, x.reshape(lead), self.buckets)


class OneHotDist:
    """unimix categorical over the last axis (reference tools.py:436-460).

    ``sample`` draws with supplied or freshly generated uniforms: idx = argmax probs/(-log u),
    the algorithm behind torch.multinomial(n=1)."""

    def __init__(self, logits, unimix_ratio=0.0):
        self._raw = logits
        self._unimix = float(unimix_ratio)
        lg = logits
        if self._unimix > 0.0:
            pr = F.softmax(lg, -1) * (1.0 - self._unimix) + self._unimix / lg.shape[-1]
            lg = torch.log(pr)
        self.logits = lg - torch.logsumexp(lg, -1, keepdim=True)

    @property
    def probs(self):
        return F.softmax(self.logits, -1)

    def mode(self):
        hard = F.one_hot(torch.argmax(self.logits, -1), self.logits.shape[-1]).to(self.logits.dtype)
        return hard.detach() + self.logits - self.logits.detach()

    def sample(self, sample_shape=(), u=None):
        if tuple(sample_shape) != ():
            raise NotImplementedError("sample_shape")
        shape = self._raw.shape
        if u is None:
            u = torch.rand(shape, device=self._raw.device)
        lg3 = self._raw.detach().reshape(-1, 1, shape[-1]).contiguous().float()
        _, hot = K.onehot_sample(lg3, u.reshape(lg3.shape).contiguous().float(), self._unimix)
        probs = self.probs
        return hot.reshape(shape) + (probs - probs.detach())

    def entropy(self):
        lg = torch.clamp(self.logits, min=torch.finfo(self.logits.dtype).min)
        return -(lg * self.probs).sum(-1)

    def log_prob(self, value):
        idx = value.max(-1)[1]
        return self.logits.gather(-1, idx[..., None])[..., 0]


class IndependentOneHot:
    """Independent(OneHotDist, 1): sums entropy / log_prob over the group axis."""

    def __init__(self, logits, unimix_ratio):
        self.base = OneHotDist(logits, unimix_ratio)

    def mode(self):
        return self.base.mode()

    def sample(self, u=None):
        lg = self.base._raw
        flat = OneHotDist(lg.reshape(-1, lg.shape[-1]), self.base._unimix)
        return flat.sample(u=None if u is None else u.reshape(-1, lg.shape[-1])).reshape(lg.shape)

    def entropy(self):
        return self.base.entropy().sum(-1)

    def log_prob(self, value):
        return self.base.log_prob(value).sum(-1)


class NormalTanhMean:
    """actor 'normal': Normal(tanh(mean), (max-min)*sigmoid(std+2)+min) with the absmax clip on
    samples (reference networks.py:693-700 + tools.py:575-601)."""

    def __init__(self, mean_raw, std_raw, min_std, max_std, absmax=None):
        self.mean = torch.tanh(mean_raw)
        self.std = (max_std - min_std) * torch.sigmoid(std_raw + 2.0) + min_std
        self.absmax = absmax

    def _clip(self, out):
        if self.absmax is None:
            return out
        return out * (self.absmax / torch.clip(torch.abs(out), min=self.absmax)).detach()

    def mode(self):
        return self._clip(self.mean)

    def sample(self, sample_shape=(), eps=None):
        if eps is None:
            eps = torch.randn_like(self.mean)
        return self._clip(self.mean + self.std * eps)

    def entropy(self):
        return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(self.std)).sum(-1)

    def log_prob(self, x):
        var = self.std ** 2
        return (-((x - self.mean) ** 2) / (2 * var) - torch.log(self.std)
                - math.log(math.sqrt(2 * math.pi))).sum(-1)


class Bernoulli:
    """cont head (reference tools.py:604-628)."""

    def __init__(self, logits):
        self.logits = logits

    @property
    def mean(self):
        return torch.sigmoid(self.logits)

    def mode(self):
        mean = self.mean
        m = torch.round(mean)
        return m.detach() + mean - mean.detach()

    def log_prob(self, x):
        if self.logits.is_cuda and self.logits.shape[-1] == 1 and x.shape == self.logits.shape:
            return K.bernoulli_logprob(self.logits, x)[..., 0]      # the sum runs over one element
        return torch.sum(-F.softplus(self.logits) * (1 - x) - F.softplus(-self.logits) * x, -1)

    def entropy(self):
        p = self.mean
        return (F.softplus(self.logits) - p * self.logits).sum(-1)


class MSEDist:
    """image decoder head (reference tools.py:520-543)."""

    def __init__(self, mode):
        self._mode = mode

    def mode(self):
        return self._mode

    def mean(self):
        return self._mode

    def log_prob(self, value):
        if self._mode.shape != value.shape:
            raise AssertionError((self._mode.shape, value.shape))
        if self._mode.is_cuda and self._mode.dim() >= 3:
            return K.sqerr_logprob(self._mode, value, False)
        return -((self._mode - value) ** 2).flatten(2).sum(-1)


class SymlogDist:
    """vector decoder head (reference tools.py:546-572), 'mse' distance, 'sum' aggregation."""

    def __init__(self, mode, tol=1e-8):
        self._mode = mode
        self._tol = tol

    def mode(self):
        return symexp(self._mode)

    def mean(self):
        return symexp(self._mode)

    def log_prob(self, value):
        if self._mode.shape != value.shape:
            raise AssertionError((self._mode.shape, value.shape))
        if self._mode.is_cuda and self._mode.dim() >= 3:
            return K.sqerr_logprob(self._mode, value, True, self._tol)
        dist = (self._mode - symlog(value)) ** 2.0
        dist = torch.where(dist < self._tol, torch.zeros_like(dist), dist)
        return -dist.flatten(2).sum(-1)


# --------------------------------------------------------------------------------------
# training utilities
# --------------------------------------------------------------------------------------
class RequiresGrad:
    def __init__(self, model):
        self._model = model

    def __enter__(self):
        self._model.requires_grad_(True)

    def __exit__(self, *exc):
        self._model.requires_grad_(False)


def tensorstats(tensor, prefix=None):
    """mean/std/min/max as device scalars (the reference syncs four times here,
    tools.py:949-958; the host copy happens once per step in ``to_host``)."""
    t = tensor.detach().float()
    if t.is_cuda and t.numel() > 0:
        s4 = K.tensorstats4(t)                      # one kernel; the entries are views
        out = {"mean": s4[0], "std": s4[1], "min": s4[2], "max": s4[3]}
    else:
        out = {"mean": t.mean(), "std": t.std(), "min": t.min(), "max": t.max()}
    return {f"{prefix}_{k}" if prefix else k: v for k, v in out.items()}


def mean_scalar(x):
    """Mean of all elements as a 0-dim device tensor (one library kernel on CUDA)."""
    x = x.detach()
    if x.is_cuda and x.numel() > 0:
        return K.tensorstats4(x)[0]
    return torch.mean(x.float())


def to_host(metrics):
    """One device->host transfer for all scalar metrics; arrays go individually."""
    scalars = {k: v for k, v in metrics.items() if torch.is_tensor(v) and v.numel() == 1}
    out = {}
    if scalars:
        flat = torch.stack([v.detach().reshape(()).float() for v in scalars.values()]).cpu().numpy()
        for k, x in zip(scalars.keys(), flat):
            out[k] = np.asarray(x)
    for k, v in metrics.items():
        if k in out:
            continue
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else v
    return {k: out[k] for k in metrics}


class GradSync:
    """Data-parallel gradient averaging: one flat bucket per optimizer, allreduce(sum)/world
    over NCCL (gloo in the CPU tests) between backward and clip_grad_norm_, where the reference's
    single-process Optimizer has nothing (tools.py:765-768)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def _avg(self, t):
        import torch.distributed as dist
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)     # no divide pass
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)

    def flat(self, flat_grad, segments=None):
        """Average one flat gradient buffer in place (tools.Optimizer's CUDA path).

        ``segments`` = [(start, end, stream)]: ranges of the buffer whose gradients were produced
        on a side stream (the world model's heads run their backward on their own streams).  Each
        is all-reduced on that stream, i.e. as soon as its producer branch is done and
        concurrently with the rest of the backward pass (the persistent observe backward); the
        remainder follows on the current stream.  Every rank issues the collectives in the same
        order."""
        if self.world == 1:
            return
        if not segments:
            self._avg(flat_grad)
            return
        cur = torch.cuda.current_stream()
        segs = sorted(segments, key=lambda s: s[0])
        for a, b, st in segs:
            with torch.cuda.stream(st):
                self._avg(flat_grad[a:b])
        pos = 0
        for a, b, _ in segs + [(flat_grad.numel(), flat_grad.numel(), None)]:
            if a > pos:
                self._avg(flat_grad[pos:a])
            pos = max(pos, b)
        for _, _, st in segs:
            cur.wait_stream(st)

    def __call__(self, params):
        if self.world == 1:
            return
        import torch.distributed as dist
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(self.world)
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n


class FlatAdam(torch.optim.Optimizer):
    """``torch.optim.Optimizer`` face of the flat fused Adam, so that the reference's checkpoint
    helpers work unchanged on the drop-in path: ``tools.recursively_collect_optim_state_dict``
    (reference tools.py:975-1002) only collects ``torch.optim.Optimizer`` instances -- found at
    ``..._model_opt._opt`` -- and ``recursively_load_optim_state_dict`` (1005-1011) calls
    ``load_state_dict`` on that attribute.  ``state_dict`` / ``load_state_dict`` speak
    torch.optim.Adam's layout (step, exp_avg, exp_avg_sq per parameter index) and map onto the
    owner's flat m / v / step buffers; a reference ``latest.pt`` resumes as is."""

    def __init__(self, owner):
        super().__init__(owner._params, dict(lr=owner._lr, betas=Optimizer.BETAS, eps=owner._eps,
                                             weight_decay=0, amsgrad=False, maximize=False,
                                             foreach=None, capturable=True, differentiable=False,
                                             fused=True))
        self._owner_ref = [owner]        # in a list: keep nn.Module / recursion helpers out of it

    def state_dict(self):
        return self._owner_ref[0]._flat_state_dict()

    def load_state_dict(self, sd):
        self._owner_ref[0]._load_flat_state_dict(sd)

    def step(self, closure=None):
        raise RuntimeError("FlatAdam is stepped by tools.Optimizer.__call__ (dv3_adam_clip_step)")

    def zero_grad(self, set_to_none=True):
        for p in self._owner_ref[0]._params:
            p.grad = None


class Optimizer:
    """Adam + global-norm clip, the reference's tools.Optimizer call sequence
    (tools.py:760-776): zero_grad -> backward -> [DP allreduce] -> clip -> (wd) -> step.

    On CUDA the parameters of one optimizer are re-homed into ONE flat fp32 buffer (each
    parameter becomes a view; ``state_dict`` names and values are unchanged), the gradients are
    gathered into a second flat buffer by one batched copy after backward, and clip + Adam run
    as the fused ``dv3_adam_clip_step`` (three launches; step counter on the device, so the call
    is graph-capturable).  The data-parallel allreduce then runs on the flat gradient buffer
    directly.  CPU parameters (the gloo tests) keep the torch.optim.Adam path."""

    BETAS = (0.9, 0.999)

    def __init__(self, name, parameters, lr, eps=1e-4, clip=None, wd=None, wd_pattern=r".*",
                 opt="adam", use_amp=False, grad_sync=None):
        if use_amp:
            raise NotImplementedError("fp32 is the contract of the B200 path (precision: 32)")
        if opt != "adam":
            raise NotImplementedError(opt)
        if wd_pattern != r".*":
            raise NotImplementedError("wd_pattern")
        self._name = name
        self._params = list(parameters)
        self._clip = clip
        self._wd = wd
        self._lr, self._eps = float(lr), float(eps)
        self._sync = grad_sync
        self._flat = (bool(self._params) and
                      all(p.is_cuda and p.dtype == torch.float32 for p in self._params))
        if self._flat:
            self._build_flat()
            self._opt = FlatAdam(self)
        else:
            self._opt = torch.optim.Adam(self._params, lr=lr, eps=eps)

    # ---- flat buffers ------------------------------------------------------------------
    def _build_flat(self):
        dev = self._params[0].device
        self._offsets, total = [], 0
        for p in self._params:
            self._offsets.append(total)
            total += (p.numel() + 3) & ~3              # every parameter starts 16-byte aligned
        z = lambda: torch.zeros(total, dtype=torch.float32, device=dev)
        self._fp, self._fg, self._fm, self._fv = z(), z(), z(), z()
        self._step = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ctl = torch.zeros(4, dtype=torch.float32, device=dev)
        self._scratch = torch.empty(1024, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(self._params, self._offsets):
                n = p.numel()
                view = self._fp[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = None
        self._gviews = self._views(self._fg)
        # tf32 operand planes of the parameters, written by the Adam kernel itself: matrices whose
        # row length is a multiple of 4 read them directly as GEMM operands (kernels.split_param)
        self._fhi, self._flo = z(), z()
        self._plane_views = []
        for p, h, l in zip(self._params, self._views(self._fhi), self._views(self._flo)):
            ok = p.dim() == 2 and p.shape[1] % 4 == 0 and p.shape[0] * p.shape[1] >= 4096
            self._plane_views.append(K.Split(h, l, p.shape[0], p.shape[1]) if ok else None)

    def _views(self, flat):
        return [flat[off:off + p.numel()].view(p.shape) for p, off in zip(self._params, self._offsets)]

    def set_grad_sync(self, sync):
        self._sync = sync

    def param_range(self, params):
        """[start, end) of the flat buffers covered by ``params`` (a contiguous run of this
        optimizer's parameters), or None."""
        if not self._flat:
            return None
        index = {id(p): i for i, p in enumerate(self._params)}
        ids = sorted(index[id(p)] for p in params if id(p) in index)
        if not ids or ids != list(range(ids[0], ids[-1] + 1)):
            return None
        last = self._params[ids[-1]]
        return self._offsets[ids[0]], self._offsets[ids[-1]] + ((last.numel() + 3) & ~3)

    def set_segments(self, segments):
        """[(start, end, stream)] for GradSync.flat; None = one collective over the whole buffer."""
        self._segments = segments

    def state_dict(self):
        """torch.optim.Adam's layout (what the reference checkpoints, dreamer.py:502-506)."""
        return self._opt.state_dict()

    def load_state_dict(self, sd):
        self._opt.load_state_dict(sd)

    def _flat_state_dict(self):
        state = {}
        if float(self._step) > 0:
            ms, vs = self._views(self._fm), self._views(self._fv)
            for i in range(len(self._params)):
                state[i] = {"step": self._step.clone().reshape(()), "exp_avg": ms[i].clone(),
                            "exp_avg_sq": vs[i].clone()}
        group = dict(lr=self._lr, betas=self.BETAS, eps=self._eps, weight_decay=0, amsgrad=False,
                     maximize=False, foreach=None, capturable=True, differentiable=False,
                     fused=True, params=list(range(len(self._params))))
        return {"state": state, "param_groups": [group]}

    def _load_flat_state_dict(self, sd):
        ms, vs = self._views(self._fm), self._views(self._fv)
        with torch.no_grad():
            for i, st in sd.get("state", {}).items():
                ms[int(i)].copy_(st["exp_avg"])
                vs[int(i)].copy_(st["exp_avg_sq"])
                self._step.fill_(float(st["step"]))
        if sd.get("param_groups"):
            self._lr = float(sd["param_groups"][0].get("lr", self._lr))
            self._eps = float(sd["param_groups"][0].get("eps", self._eps))

    # ---- one optimisation step ---------------------------------------------------------
    def __call__(self, loss, params=None, retain_graph=False):
        return self.step(self.backward(loss, retain_graph=retain_graph))

    def backward(self, loss, retain_graph=False, sync=False):
        """First half of a step: zero_grad -> backward (gradients into the flat buffer).  Returns
        the state ``step`` needs.  Split from ``step`` so that a caller can put work that only
        READS the parameters between the two (graphs.TrainStepGraph, pipelined schedule).
        ``sync``: issue the data-parallel all-reduce here instead of in ``step`` (collectives run in
        issue order: the pipelined schedule wants the world model's ahead of the behaviour's)."""
        if loss.dim() != 0:
            raise AssertionError(loss.shape)
        params = self._params
        metrics = {f"{self._name}_loss": loss.detach()}
        if not self._flat:
            self._opt.zero_grad(set_to_none=True)
            loss.backward(retain_graph=retain_graph)
            return dict(metrics=metrics)
        # The flat gradient buffer is zeroed once and armed as the gradient sink: the library's
        # backward Functions accumulate dW / LayerNorm / bias gradients straight into its views
        # (kernels.arm_grad_sink); gradients that still arrive through autograd (torch modules:
        # the conv stacks) are copied in per parameter afterwards.
        for p in params:
            p.grad = None
        self._fg.zero_()
        K.arm_grad_sink(params, self._gviews)
        try:
            loss.backward(retain_graph=retain_graph)
        finally:
            written = K.disarm_grad_sink(params)
        if self._sync is None or not getattr(self, "_segments", None):
            K.join_grad_streams()                # (data parallel: joined behind the segments' all-reduce)
        skipped = []
        with torch.no_grad():
            for i, (p, view, w) in enumerate(zip(params, self._gviews, written)):
                if p.grad is not None:
                    if w:
                        view.add_(p.grad)
                    else:
                        view.copy_(p.grad)
                    p.grad = None
                elif not w:
                    skipped.append(i)
        synced = False
        if sync and self._sync is not None:
            self._sync.flat(self._fg, getattr(self, "_segments", None))
            K.join_grad_streams()
            synced = True
        return dict(metrics=metrics, skipped=skipped, synced=synced)

    def step(self, state):
        """Second half: [DP all-reduce] -> clip -> (weight decay) -> Adam; returns the metrics."""
        params = self._params
        metrics = state["metrics"]
        if not self._flat:
            if self._sync is not None:
                self._sync(params)
            norm = nn.utils.clip_grad_norm_(params, self._clip)
            if self._wd:
                with torch.no_grad():
                    for p in params:
                        p.mul_(1 - self._wd)
            self._opt.step()
            K.invalidate_weight_splits()
            self._opt.zero_grad(set_to_none=True)
            metrics[f"{self._name}_grad_norm"] = norm.detach()
            return metrics
        skipped = state["skipped"]
        # torch.optim.Adam leaves a parameter that received no gradient untouched (no moment
        # decay, no move on stale momentum): keep that by restoring it after the fused update.
        # (The step counter stays global: a parameter that is skipped in some steps gets the
        # optimizer's bias correction, not a private one.)
        keep = None
        if skipped:
            pv, mv, vv = self._views(self._fp), self._views(self._fm), self._views(self._fv)
            keep = [(pv[i], pv[i].clone(), mv[i], mv[i].clone(), vv[i], vv[i].clone()) for i in skipped]
        if self._sync is not None and not state.get("synced"):
            self._sync.flat(self._fg, getattr(self, "_segments", None))
            K.join_grad_streams()
        L_ = K.L
        L_.check(L_.lib().dv3_adam_clip_step_planes(
            L_.fptr(self._fp), L_.fptr(self._fg), L_.fptr(self._fm), L_.fptr(self._fv),
            self._fp.numel(), self._lr, self.BETAS[0], self.BETAS[1], self._eps,
            float(self._clip) if self._clip else 0.0, 1.0 - float(self._wd) if self._wd else 1.0,
            L_.fptr(self._step), L_.fptr(self._ctl), L_.fptr(self._scratch), L_.fptr(self._fhi),
            L_.fptr(self._flo), L_.stream_ptr()), "adam_clip_step_planes")
        if keep is not None:
            with torch.no_grad():
                for p_, p0, m_, m0, v_, v0 in keep:
                    p_.copy_(p0)
                    m_.copy_(m0)
                    v_.copy_(v0)
        skip = set(skipped)
        for i, (p, sp) in enumerate(zip(params, self._plane_views)):
            if sp is not None:
                p._dv3_flat_split = None if i in skip else ((p._version, p.data_ptr()), sp)
        K.invalidate_weight_splits()
        metrics[f"{self._name}_grad_norm"] = self._ctl[0].clone()
        return metrics


# --------------------------------------------------------------------------------------
# initialisers (reference tools.py:890-946): truncated-normal variance scaling / uniform outscale
# --------------------------------------------------------------------------------------
def _fans(m):
    if isinstance(m, nn.Linear):
        return m.in_features, m.out_features
    space = m.kernel_size[0] * m.kernel_size[1]
    return space * m.in_channels, space * m.out_channels


def weight_init(m):
    if isinstance(m, (nn.Linear, nn.Conv2d, nn.ConvTranspose2d)):
        fin, fout = _fans(m)
        std = math.sqrt(2.0 / (fin + fout)) / 0.87962566103423978
        nn.init.trunc_normal_(m.weight.data, mean=0.0, std=std, a=-2.0 * std, b=2.0 * std)
        if m.bias is not None:
            m.bias.data.zero_()
    elif isinstance(m, nn.LayerNorm):
        m.weight.data.fill_(1.0)
        if m.bias is not None:
            m.bias.data.zero_()


def uniform_weight_init(given_scale):
    def init(m):
        if isinstance(m, (nn.Linear, nn.Conv2d, nn.ConvTranspose2d)):
            fin, fout = _fans(m)
            limit = math.sqrt(3 * given_scale * 2.0 / (fin + fout))
            nn.init.uniform_(m.weight.data, a=-limit, b=limit)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.LayerNorm):
            m.weight.data.fill_(1.0)
            if m.bias is not None:
                m.bias.data.zero_()
    return init
