"""CPU oracle for the DreamerV3 training hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A plain torch-CPU (fp32 / fp64) restatement of the reference algorithm for the
path BASELINE.json's north_star names.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this file; the
product package (``dreamerv3-torch_b200/``) never does and has no CPU fallback.

Style: purely functional.  Parameters travel as a flat ``dict[str, Tensor]`` keyed by
the reference's own ``state_dict`` names (so a reference checkpoint drops in), noise is
always an explicit argument (the "supplied uniforms" contract of SURVEY.md 8c):

* categorical draw:  ``idx = argmax_k( probs_k / (-log u_k) )``, one ``u ~ U(0,1)`` per
  class -- ATen's n=1 ``torch.multinomial`` algorithm with the exponential variate
  written as ``-log u``;
* normal draw: ``eps ~ N(0,1)`` supplied directly.

Parity pinning: the reference ships no tests / golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, run in the build container
(``oracle/pin_against_reference.py``; fixtures in ``tests/golden/`` are produced by
``tests/golden/make_golden.py`` from the live reference modules).

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

LN_EPS = 1e-3          # networks.py:55,66,76 / 628 / 755  (every LayerNorm uses eps=1e-03)
UPDATE_BIAS = -1.0     # networks.py:743 (GRUCell update_bias)


# --------------------------------------------------------------------------------------
# scalar transforms                                                       tools.py:22-27
# --------------------------------------------------------------------------------------
def symlog(x: Tensor) -> Tensor:
    return torch.sign(x) * torch.log(torch.abs(x) + 1.0)


def symexp(x: Tensor) -> Tensor:
    return torch.sign(x) * (torch.exp(torch.abs(x)) - 1.0)


def _sub(p: Params, prefix: str) -> Params:
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------------------
# dense blocks
# --------------------------------------------------------------------------------------
def ln(x: Tensor, g: Tensor, b: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), g, b, LN_EPS)


def dense_ln_silu(x: Tensor, w: Tensor, g: Tensor, b: Tensor) -> Tensor:
    """Linear(no bias) -> LayerNorm(eps 1e-3) -> SiLU        networks.py:48-58,62-78,623-632"""
    return F.silu(ln(x @ w.t(), g, b))


def gru_cell(p: Params, x: Tensor, h: Tensor) -> Tensor:
    """LayerNorm GRU cell.                                           networks.py:760-768

    parts = LN(W [x,h]); reset, cand, update = split(parts); the update gate carries a
    constant -1 bias; new = u*tanh(r*cand) + (1-u)*h.
    """
    parts = ln(torch.cat([x, h], -1) @ p["_cell.layers.GRU_linear.weight"].t(),
               p["_cell.layers.GRU_norm.weight"], p["_cell.layers.GRU_norm.bias"])
    d = h.shape[-1]
    r = torch.sigmoid(parts[..., :d])
    c = torch.tanh(r * parts[..., d:2 * d])
    u = torch.sigmoid(parts[..., 2 * d:] + UPDATE_BIAS)
    return u * c + (1.0 - u) * h


# --------------------------------------------------------------------------------------
# unimix one-hot categorical                        tools.py:436-460 + torch Categorical
# --------------------------------------------------------------------------------------
def unimix_logits(logit: Tensor, unimix: float) -> Tuple[Tensor, Tensor]:
    """(normalised log-probs, probs) exactly as OneHotDist.__init__ + Categorical build them.

    tools.py:439-442: p = softmax(l); p = (1-r)p + r/K; l' = log p
    torch Categorical.__init__: l'' = l' - logsumexp(l'); probs = softmax(l'') (lazy).
    """
    if unimix > 0.0:
        pr = F.softmax(logit, dim=-1)
        pr = pr * (1.0 - unimix) + unimix / pr.shape[-1]
        logit = torch.log(pr)
    norm = logit - torch.logsumexp(logit, dim=-1, keepdim=True)
    return norm, F.softmax(norm, dim=-1)


def categorical_index(probs: Tensor, u: Tensor) -> Tensor:
    """Index drawn by torch.multinomial(probs, 1, True) when its Exp(1) variates are -log u."""
    q = -torch.log(u)
    return torch.argmax(probs / q, dim=-1)


def onehot_sample(logit: Tensor, u: Tensor, unimix: float) -> Tuple[Tensor, Tensor]:
    """Straight-through one-hot sample.  tools.py:452-460.  Returns (sample, index)."""
    _, probs = unimix_logits(logit, unimix)
    idx = categorical_index(probs.detach(), u)
    hard = F.one_hot(idx, probs.shape[-1]).to(probs.dtype)
    return hard + (probs - probs.detach()), idx


def onehot_mode(logit: Tensor, unimix: float) -> Tensor:
    """tools.py:446-450: one_hot(argmax(norm logits)) + norm - sg(norm)."""
    norm, _ = unimix_logits(logit, unimix)
    hard = F.one_hot(torch.argmax(norm, dim=-1), norm.shape[-1]).to(norm.dtype)
    return hard.detach() + norm - norm.detach()


def onehot_entropy(logit: Tensor, unimix: float) -> Tensor:
    """Independent(OneHotDist,1).entropy(): -sum p log p over classes, summed over groups."""
    norm, probs = unimix_logits(logit, unimix)
    norm = torch.clamp(norm, min=torch.finfo(norm.dtype).min)
    return -(norm * probs).sum(-1).sum(-1)


def onehot_logprob(logit: Tensor, value: Tensor, unimix: float) -> Tensor:
    """OneHotCategorical.log_prob for a single group axis (actor 'onehot'): pick by argmax."""
    norm, _ = unimix_logits(logit, unimix)
    idx = value.max(-1)[1]
    return norm.gather(-1, idx[..., None])[..., 0]


# --------------------------------------------------------------------------------------
# RSSM                                                                networks.py:13-290
# --------------------------------------------------------------------------------------
class RSSMDims:
    def __init__(self, stoch=32, classes=32, deter=512, hidden=512, actions=6, embed=1024,
                 unimix=0.01):
        self.stoch, self.classes, self.deter, self.hidden = stoch, classes, deter, hidden
        self.actions, self.embed, self.unimix = actions, embed, unimix

    @property
    def flat(self):
        return self.stoch * self.classes


def stat_logits(p: Params, which: str, x: Tensor, d: RSSMDims) -> Tensor:
    """networks.py:241-250 (discrete branch)."""
    name = {"ims": "_imgs_stat_layer", "obs": "_obs_stat_layer"}[which]
    out = x @ p[name + ".weight"].t() + p[name + ".bias"]
    return out.reshape(list(out.shape[:-1]) + [d.stoch, d.classes])


def prior_head(p: Params, deter: Tensor, d: RSSMDims) -> Tensor:
    y = dense_ln_silu(deter, p["_img_out_layers.0.weight"], p["_img_out_layers.1.weight"],
                      p["_img_out_layers.1.bias"])
    return stat_logits(p, "ims", y, d)


def rssm_initial(p: Params, batch: int, d: RSSMDims) -> Dict[str, Tensor]:
    """networks.py:99-125 ('learned'): deter=tanh(W) tiled, stoch=mode(prior(deter)), logit=0."""
    deter = torch.tanh(p["W"]).repeat(batch, 1)
    stoch = onehot_mode(prior_head(p, deter, d), d.unimix)
    return {"stoch": stoch, "deter": deter,
            "logit": torch.zeros(batch, d.stoch, d.classes, dtype=deter.dtype, device=deter.device)}


def img_step(p: Params, state: Dict[str, Tensor], action: Tensor, u: Optional[Tensor],
             d: RSSMDims) -> Dict[str, Tensor]:
    """Prior step.  networks.py:208-233.  u=None -> mode instead of sample."""
    flat = state["stoch"].reshape(list(state["stoch"].shape[:-2]) + [d.flat])
    x = dense_ln_silu(torch.cat([flat, action], -1), p["_img_in_layers.0.weight"],
                      p["_img_in_layers.1.weight"], p["_img_in_layers.1.bias"])
    deter = gru_cell(p, x, state["deter"])
    logit = prior_head(p, deter, d)
    stoch = onehot_mode(logit, d.unimix) if u is None else onehot_sample(logit, u, d.unimix)[0]
    return {"stoch": stoch, "deter": deter, "logit": logit}


def obs_step(p: Params, prev: Optional[Dict[str, Tensor]], action: Tensor, embed: Tensor,
             is_first: Tensor, u_prior: Optional[Tensor], u_post: Optional[Tensor],
             d: RSSMDims):
    """Posterior step.  networks.py:174-206.

    Three is_first cases (all / some / none) as in the reference; the 'some' branch mixes
    row-wise ``val*(1-m)+init*m`` and zeroes the action rows (the reference does the latter
    in place on the caller's tensor, networks.py:184 -- here it is a fresh tensor).
    """
    n = is_first.shape[0]
    if prev is None or bool(is_first.sum() == n):
        prev = rssm_initial(p, n, d)
        action = torch.zeros(n, d.actions, dtype=embed.dtype, device=embed.device)
    elif bool(is_first.sum() > 0):
        m = is_first[:, None]
        action = action * (1.0 - m)
        init = rssm_initial(p, n, d)
        mixed = {}
        for key, val in prev.items():
            mk = m.reshape(m.shape + (1,) * (val.dim() - m.dim()))
            mixed[key] = val * (1.0 - mk) + init[key] * mk
        prev = mixed
    prior = img_step(p, prev, action, u_prior, d)
    z = dense_ln_silu(torch.cat([prior["deter"], embed], -1), p["_obs_out_layers.0.weight"],
                      p["_obs_out_layers.1.weight"], p["_obs_out_layers.1.bias"])
    logit = stat_logits(p, "obs", z, d)
    stoch = onehot_mode(logit, d.unimix) if u_post is None else onehot_sample(logit, u_post, d.unimix)[0]
    post = {"stoch": stoch, "deter": prior["deter"], "logit": logit}
    return post, prior


def observe(p: Params, embed: Tensor, action: Tensor, is_first: Tensor, u_prior: Tensor,
            u_post: Tensor, d: RSSMDims, state: Optional[Dict[str, Tensor]] = None):
    """T-step posterior rollout.  networks.py:127-143 + tools.py:806-850.

    embed [B,T,E], action [B,T,A], is_first [B,T]; u_* are TIME-major [T,B,S,C].
    Returns batch-major (post, prior), each {stoch,deter,logit}.
    """
    T = embed.shape[1]
    posts, priors = [], []
    prev = state
    for t in range(T):
        post, prior = obs_step(p, prev, action[:, t], embed[:, t], is_first[:, t],
                               u_prior[t], u_post[t], d)
        posts.append(post)
        priors.append(prior)
        prev = post
    stack = lambda seq: {k: torch.stack([s[k] for s in seq], 1) for k in seq[0]}
    return stack(posts), stack(priors)


def imagine_with_action(p: Params, action: Tensor, state: Dict[str, Tensor], u: Tensor,
                        d: RSSMDims) -> Dict[str, Tensor]:
    """networks.py:145-152.  action [B,T,A] batch-major, u [T,B,S,C]."""
    outs = []
    for t in range(action.shape[1]):
        state = img_step(p, state, action[:, t], u[t], d)
        outs.append(state)
    return {k: torch.stack([s[k] for s in outs], 1) for k in outs[0]}


def get_feat(state: Dict[str, Tensor]) -> Tensor:
    """networks.py:154-159."""
    s = state["stoch"]
    return torch.cat([s.reshape(list(s.shape[:-2]) + [-1]), state["deter"]], -1)


def kl_balance(post_logit: Tensor, prior_logit: Tensor, free: float, dyn_scale: float,
               rep_scale: float, unimix: float):
    """networks.py:272-290 with torch's _kl_categorical_categorical.  -> loss,value,dyn,rep."""

    def kl(lp, lq):
        np_, pp = unimix_logits(lp, unimix)
        nq, pq = unimix_logits(lq, unimix)
        t = pp * (np_ - nq)
        t = torch.where(pq == 0, torch.full_like(t, math.inf), t)
        t = torch.where(pp == 0, torch.zeros_like(t), t)
        return t.sum(-1).sum(-1)

    rep = value = kl(post_logit, prior_logit.detach())
    dyn = kl(post_logit.detach(), prior_logit)
    rep = torch.clip(rep, min=free)
    dyn = torch.clip(dyn, min=free)
    return dyn_scale * dyn + rep_scale * rep, value, dyn, rep


# --------------------------------------------------------------------------------------
# lambda return                                                        tools.py:682-728
# --------------------------------------------------------------------------------------
def lambda_return(reward: Tensor, value: Tensor, pcont: Tensor, bootstrap: Tensor,
                  lambda_: float) -> Tensor:
    """Time-major [H,N,1] in, [H,N,1] out (the reference unbinds this into a tuple of N
    [H,1] tensors, tools.py:696-699; its caller re-stacks with dim=1 -- same values)."""
    nxt = torch.cat([value[1:], bootstrap[None]], 0)
    inputs = reward + pcont * nxt * (1.0 - lambda_)
    last = bootstrap
    outs = []
    for t in reversed(range(reward.shape[0])):
        last = inputs[t] + pcont[t] * lambda_ * last
        outs.append(last)
    return torch.stack(list(reversed(outs)), 0)


# --------------------------------------------------------------------------------------
# symlog two-hot discrete regression head                              tools.py:463-517
# --------------------------------------------------------------------------------------
def buckets(dtype=torch.float32, device=None) -> Tensor:
    # built on the CPU like the reference's config-time linspace, then moved (same bits everywhere)
    return torch.linspace(-20.0, 20.0, steps=255).to(dtype).to(device)


def twohot_mean(logits: Tensor) -> Tensor:
    """DiscDist.mean / mode: symexp(sum softmax(l) * buckets).  tools.py:481-487."""
    b = buckets(logits.dtype, logits.device)
    return symexp(torch.sum(torch.softmax(logits, -1) * b, dim=-1, keepdim=True))


def twohot_logprob(logits: Tensor, x: Tensor) -> Tensor:
    """DiscDist.log_prob.  tools.py:490-513.  logits [...,255], x [...,1] -> [...]."""
    b = buckets(logits.dtype, logits.device)
    x = symlog(x)
    below = torch.sum((b <= x[..., None]).to(torch.int32), dim=-1) - 1
    above = len(b) - torch.sum((b > x[..., None]).to(torch.int32), dim=-1)
    below = torch.clip(below, 0, len(b) - 1)
    above = torch.clip(above, 0, len(b) - 1)
    equal = below == above
    d_below = torch.where(equal, torch.ones_like(x), torch.abs(b[below] - x))
    d_above = torch.where(equal, torch.ones_like(x), torch.abs(b[above] - x))
    total = d_below + d_above
    w_below, w_above = d_above / total, d_below / total
    target = (F.one_hot(below, len(b)) * w_below[..., None]
              + F.one_hot(above, len(b)) * w_above[..., None]).squeeze(-2)
    log_pred = logits - torch.logsumexp(logits, -1, keepdim=True)
    return (target * log_pred).sum(-1)


# --------------------------------------------------------------------------------------
# MLP heads                                                          networks.py:588-739
# --------------------------------------------------------------------------------------
def mlp_trunk(p: Params, name: str, x: Tensor, layers: int, symlog_inputs: bool = False) -> Tensor:
    """networks.py:657-661: [Linear(no bias)+LN+SiLU] x layers.  p keyed 'layers.<name>_linear<i>.weight'."""
    if symlog_inputs:
        x = symlog(x)
    for i in range(layers):
        x = dense_ln_silu(x, p[f"layers.{name}_linear{i}.weight"],
                          p[f"layers.{name}_norm{i}.weight"], p[f"layers.{name}_norm{i}.bias"])
    return x


def mlp_head(p: Params, x: Tensor, key: str = "mean_layer") -> Tensor:
    return x @ p[key + ".weight"].t() + p[key + ".bias"]


def actor_normal_stats(p: Params, feat: Tensor, layers: int, min_std=0.1, max_std=1.0):
    """dist 'normal'.  networks.py:693-700: mean=tanh(m), std=(max-min)*sigmoid(s+2)+min."""
    h = mlp_trunk(p, "Actor", feat, layers)
    mean = torch.tanh(mlp_head(p, h, "mean_layer"))
    std = (max_std - min_std) * torch.sigmoid(mlp_head(p, h, "std_layer") + 2.0) + min_std
    return mean, std


def contdist_sample(mean: Tensor, std: Tensor, eps: Tensor, absmax: float = 1.0) -> Tensor:
    """ContDist.sample: rsample then clip-by-rescale with a detached factor.  tools.py:594-598."""
    out = mean + std * eps
    return out * (absmax / torch.clip(torch.abs(out), min=absmax)).detach()


def normal_entropy(std: Tensor) -> Tensor:
    return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)).sum(-1)


def normal_logprob(mean: Tensor, std: Tensor, x: Tensor) -> Tensor:
    var = std ** 2
    return (-((x - mean) ** 2) / (2 * var) - torch.log(std) - math.log(math.sqrt(2 * math.pi))).sum(-1)


def bernoulli_mean(logit: Tensor) -> Tensor:
    return torch.sigmoid(logit)


def bernoulli_logprob(logit: Tensor, x: Tensor) -> Tensor:
    """tools.py:622-627."""
    return torch.sum(-F.softplus(logit) * (1 - x) + -F.softplus(-logit) * x, -1)


# --------------------------------------------------------------------------------------
# imagination rollout                                                  models.py:448-548
# --------------------------------------------------------------------------------------
def imagine(p_rssm: Params, p_actor: Params, start: Dict[str, Tensor], horizon: int,
            act_noise: Tensor, u_state: Tensor, d: RSSMDims, actor_layers: int,
            actor_dist: str = "normal", actor_unimix: float = 0.01):
    """start: {stoch [N,S,C], deter [N,D], logit [N,S,C]} (already flattened, models.py:450-451).

    act_noise [H,N,A]: N(0,1) draws (normal actor) or uniforms (onehot actor);
    u_state [H,N,S,C] uniforms.  Returns time-major (feats [H,N,F] detached,
    states {k:[H,N,...]} = [start, succ_0..succ_{H-2}], actions [H,N,A]).
    """
    state = start
    feats, actions, succs = [], [], []
    for k in range(horizon):
        feat = get_feat(state).detach()                                  # models.py:514
        if actor_dist == "normal":
            mean, std = actor_normal_stats(p_actor, feat, actor_layers)
            action = contdist_sample(mean, std, act_noise[k])
        elif actor_dist == "onehot":
            h = mlp_trunk(p_actor, "Actor", feat, actor_layers)
            action, _ = onehot_sample(mlp_head(p_actor, h), act_noise[k], actor_unimix)
        else:
            raise NotImplementedError(actor_dist)
        succ = img_step(p_rssm, state, action, u_state[k], d)           # models.py:516
        feats.append(feat)
        actions.append(action)
        succs.append(succ)
        state = succ
    states = {k: torch.stack([start[k]] + [s[k] for s in succs[:-1]], 0) for k in start}  # models.py:546
    return torch.stack(feats, 0), states, torch.stack(actions, 0)
