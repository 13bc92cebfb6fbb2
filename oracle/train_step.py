"""CPU oracle of one full DreamerV3 train step -- TEST INFRASTRUCTURE, NOT PRODUCT.

Functional restatement (torch CPU, autograd) of ``Dreamer._train`` = ``WorldModel._train`` +
``ImagBehavior._train`` (reference dreamer.py:192-200, models.py:108-171, 327-446, 620-681) and
``tools.Optimizer.__call__`` (tools.py:760-776), built from the per-op functions of
``dv3_oracle``.  Parameters are flat dicts keyed by the reference's ``state_dict`` names, noise is
supplied.  Used by tests (parity of the whole step), bench.py's cpu_baseline / ``--impl
reference`` arm.  Pinned against the live reference by oracle/pin_against_reference.py
(``pin_train_step``) and the fixtures under tests/golden/.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch

import dv3_oracle as O

DEFAULTS = dict(  # configs.yaml:64-113 (defaults) -- only what the step reads
    dyn_stoch=32, dyn_discrete=32, units=512, unimix_ratio=0.01,
    enc_layers=5, enc_units=1024, dec_layers=5, dec_units=1024,
    reward_layers=2, cont_layers=2, critic_layers=2, actor_layers=2, actor_dist="normal",
    actor_entropy=3e-4, actor_min_std=0.1, actor_max_std=1.0, actor_unimix=0.01,
    kl_free=1.0, dyn_scale=0.5, rep_scale=0.1, reward_scale=1.0, cont_scale=1.0,
    discount=0.997, discount_lambda=0.95, imag_horizon=15, imag_gradient="dynamics",
    reward_EMA=True, slow_target=True, slow_target_update=1, slow_target_fraction=0.02,
    model_lr=1e-4, model_eps=1e-8, model_clip=1000.0,
    actor_lr=3e-5, actor_eps=1e-5, actor_clip=100.0,
    critic_lr=3e-5, critic_eps=1e-5, critic_clip=100.0,
    mlp_keys=("orientations", "height", "velocity"),
)


def make_cfg(**kw):
    d = dict(DEFAULTS)
    d.update(kw)
    return SimpleNamespace(**d)


def sub(p, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


def preprocess(data, cfg):
    """models.py:174-190 on numpy / tensor dict -> fp32 tensors (+ cont, scaled discount)."""
    out = {k: torch.as_tensor(v).to(torch.float32) for k, v in data.items()}
    if "image" in out:
        out["image"] = out["image"] / 255.0
    if "discount" in out:
        out["discount"] = (out["discount"] * cfg.discount).unsqueeze(-1)
    out["cont"] = (1.0 - out["is_terminal"]).unsqueeze(-1)
    return out


def wm_loss(P, data, noise, cfg, d):
    """-> (scalar loss, post, aux).  P: WorldModel.state_dict()-keyed tensors.  noise=(u_prior,
    u_post) time-major."""
    obs = torch.cat([data[k] for k in cfg.mlp_keys], -1)
    embed = O.mlp_trunk(sub(P, "encoder._mlp."), "Encoder", obs, cfg.enc_layers, symlog_inputs=True)
    post, prior = O.observe(sub(P, "dynamics."), embed, data["action"], data["is_first"], noise[0],
                            noise[1], d)
    kl_loss, kl_value, dyn, rep = O.kl_balance(post["logit"], prior["logit"], cfg.kl_free,
                                               cfg.dyn_scale, cfg.rep_scale, d.unimix)
    feat = O.get_feat(post)
    losses = {}
    dec = sub(P, "heads.decoder._mlp.")
    trunk = O.mlp_trunk(dec, "Decoder", feat, cfg.dec_layers)
    for k in cfg.mlp_keys:
        mode = O.mlp_head(dec, trunk, f"mean_layer.{k}")
        dist = (mode - O.symlog(data[k])) ** 2.0                      # tools.py:558-562
        dist = torch.where(dist < 1e-8, torch.zeros_like(dist), dist)
        losses[k] = dist.sum(-1)
    rew = sub(P, "heads.reward.")
    r_logits = O.mlp_head(rew, O.mlp_trunk(rew, "Reward", feat, cfg.reward_layers))
    losses["reward"] = -O.twohot_logprob(r_logits, data["reward"][..., None])
    con = sub(P, "heads.cont.")
    c_logit = O.mlp_head(con, O.mlp_trunk(con, "Cont", feat, cfg.cont_layers))
    losses["cont"] = -O.bernoulli_logprob(c_logit, data["cont"])
    scales = {"reward": cfg.reward_scale, "cont": cfg.cont_scale}
    model_loss = sum(v * scales.get(k, 1.0) for k, v in losses.items()) + kl_loss
    aux = dict(embed=embed, feat=feat, prior=prior, losses=losses, kl_value=kl_value, dyn=dyn,
               rep=rep)
    return model_loss.mean(), post, aux


def behavior_losses(P_wm, P_actor, P_value, P_slow, ema_vals, start, noise, cfg, d):
    """models.py:327-429 up to (not including) the optimizer calls.
    start: detached posterior dict [B,T,...]; noise=(act_noise [H,N,A], u_state [H,N,S,C]).
    -> actor_loss, value_loss, new ema_vals, aux"""
    flat = {k: v.detach().reshape([-1] + list(v.shape[2:])) for k, v in start.items()}
    H = cfg.imag_horizon
    feats, states, actions = O.imagine(sub(P_wm, "dynamics."), P_actor, flat, H, noise[0], noise[1],
                                       d, cfg.actor_layers, cfg.actor_dist, cfg.actor_unimix)
    rew = sub(P_wm, "heads.reward.")
    sfeat = O.get_feat(states)
    reward = O.twohot_mean(O.mlp_head(rew, O.mlp_trunk(rew, "Reward", sfeat, cfg.reward_layers)))
    # actor re-evaluated on the detached features (models.py:391, 649-650)
    if cfg.actor_dist == "normal":
        mean, std = O.actor_normal_stats(P_actor, feats, cfg.actor_layers, cfg.actor_min_std,
                                         cfg.actor_max_std)
        actor_ent = O.normal_entropy(std)
    else:
        a_logits = O.mlp_head(P_actor, O.mlp_trunk(P_actor, "Actor", feats, cfg.actor_layers))
        actor_ent = O.onehot_entropy(a_logits[..., None, :], cfg.actor_unimix)
    con = sub(P_wm, "heads.cont.")
    discount = cfg.discount * O.bernoulli_mean(
        O.mlp_head(con, O.mlp_trunk(con, "Cont", sfeat, cfg.cont_layers)))
    v_logits = O.mlp_head(P_value, O.mlp_trunk(P_value, "Value", feats, cfg.critic_layers))
    value = O.twohot_mean(v_logits)
    target = O.lambda_return(reward[1:], value[:-1], discount[1:], value[-1], cfg.discount_lambda)
    weights = torch.cumprod(torch.cat([torch.ones_like(discount[:1]), discount[:-1]], 0), 0).detach()
    base = value[:-1]
    if cfg.reward_EMA:                                                # models.py:19-26, 654-659
        q = torch.quantile(target.detach().flatten(), torch.tensor([0.05, 0.95], device=target.device))
        ema_vals = 0.01 * q + 0.99 * ema_vals
        scale = torch.clip(ema_vals[1] - ema_vals[0], min=1.0)
        offset = ema_vals[0]
        adv = (target - offset) / scale - (base - offset) / scale
    else:
        adv = target - base
    if cfg.imag_gradient == "dynamics":
        actor_target = adv
    elif cfg.imag_gradient == "reinforce":
        if cfg.actor_dist == "normal":
            logp = O.normal_logprob(mean, std, actions)
        else:
            logp = O.onehot_logprob(a_logits, actions, cfg.actor_unimix)
        actor_target = logp[:-1][:, :, None] * (target - value[:-1]).detach()
    else:
        raise NotImplementedError(cfg.imag_gradient)
    actor_loss = -weights[:-1] * actor_target - cfg.actor_entropy * actor_ent[:-1, ..., None]
    actor_loss = actor_loss.mean()
    # value loss (models.py:419-429)
    vl = O.mlp_head(P_value, O.mlp_trunk(P_value, "Value", feats[:-1].detach(), cfg.critic_layers))
    value_loss = -O.twohot_logprob(vl, target.detach())
    if cfg.slow_target:
        sl = O.mlp_head(P_slow, O.mlp_trunk(P_slow, "Value", feats[:-1].detach(), cfg.critic_layers))
        value_loss = value_loss - O.twohot_logprob(vl, O.twohot_mean(sl).detach())
    value_loss = torch.mean(weights[:-1] * value_loss[:, :, None])
    aux = dict(feats=feats, states=states, actions=actions, reward=reward, target=target,
               weights=weights, value=value, actor_ent=actor_ent)
    return actor_loss, value_loss, ema_vals.detach(), aux


class Adam:
    """torch.optim.Adam(lr, eps) single-tensor update + clip_grad_norm_ (tools.py:760-776)."""

    def __init__(self, params, lr, eps, clip):
        self.params = params          # dict name -> leaf tensor
        self.lr, self.eps, self.clip = lr, eps, clip
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = 0

    def step(self, loss):
        names = [k for k, v in self.params.items() if v.requires_grad]
        grads = torch.autograd.grad(loss, [self.params[k] for k in names], allow_unused=True)
        grads = {k: (g if g is not None else torch.zeros_like(self.params[k]))
                 for k, g in zip(names, grads)}
        norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
        coef = torch.clamp(self.clip / (norm + 1e-6), max=1.0)
        self.t += 1
        b1, b2 = 0.9, 0.999
        with torch.no_grad():
            for k in names:
                g = grads[k] * coef
                self.m[k].mul_(b1).add_(g, alpha=1 - b1)
                self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (self.v[k].sqrt() / math.sqrt(1 - b2 ** self.t)).add_(self.eps)
                self.params[k].addcdiv_(self.m[k], denom, value=-self.lr / (1 - b1 ** self.t))
        return norm, grads


class Agent:
    """Holds the three parameter sets + optimizer state; ``train_step`` = Dreamer._train."""

    def __init__(self, P_wm, P_actor, P_value, cfg, d):
        leaf = lambda p: {k: v.detach().clone().float().requires_grad_(True) for k, v in p.items()}
        self.P_wm, self.P_actor, self.P_value = leaf(P_wm), leaf(P_actor), leaf(P_value)
        self.P_slow = {k: v.detach().clone() for k, v in self.P_value.items()}
        self.cfg, self.d = cfg, d
        self.ema_vals = torch.zeros(2, device=next(iter(self.P_wm.values())).device)
        self.updates = 0
        self.opt_wm = Adam(self.P_wm, cfg.model_lr, cfg.model_eps, cfg.model_clip)
        self.opt_actor = Adam(self.P_actor, cfg.actor_lr, cfg.actor_eps, cfg.actor_clip)
        self.opt_value = Adam(self.P_value, cfg.critic_lr, cfg.critic_eps, cfg.critic_clip)

    def train_step(self, data, noise, apply=True):
        """noise: dict(u_prior, u_post, act_noise, u_state).  -> metrics dict of python floats /
        tensors (losses, grad norms, grads when apply=False)."""
        cfg, d = self.cfg, self.d
        data = preprocess(data, cfg)
        out = {}
        loss, post, aux = wm_loss(self.P_wm, data, (noise["u_prior"], noise["u_post"]), cfg, d)
        out["model_loss"] = loss.detach()
        for k, v in aux["losses"].items():
            out[f"{k}_loss"] = v.detach()
        out["kl"] = aux["kl_value"].mean().detach()
        norm, g_wm = self.opt_wm.step(loss) if apply else self._grads(loss, self.P_wm, cfg.model_clip)
        out["model_grad_norm"] = norm
        start = {k: v.detach() for k, v in post.items()}
        if cfg.slow_target:                                           # models.py:683-689
            if self.updates % cfg.slow_target_update == 0:
                mix = cfg.slow_target_fraction
                for k in self.P_slow:
                    self.P_slow[k] = mix * self.P_value[k].detach() + (1 - mix) * self.P_slow[k]
            self.updates += 1
        frozen = {k: v.detach() for k, v in self.P_wm.items()}
        a_loss, v_loss, ema, baux = behavior_losses(
            frozen, self.P_actor, self.P_value, self.P_slow, self.ema_vals, start,
            (noise["act_noise"], noise["u_state"]), cfg, d)
        self.ema_vals = ema
        out["actor_loss"], out["value_loss"] = a_loss.detach(), v_loss.detach()
        if apply:
            out["actor_grad_norm"], g_a = self.opt_actor.step(a_loss)
            out["value_grad_norm"], g_v = self.opt_value.step(v_loss)
        else:
            out["actor_grad_norm"], g_a = self._grads(a_loss, self.P_actor, cfg.actor_clip)
            out["value_grad_norm"], g_v = self._grads(v_loss, self.P_value, cfg.critic_clip)
        out["grads"] = dict(wm=g_wm, actor=g_a, value=g_v)
        out["post"] = start
        out["imag"] = baux
        return out

    @staticmethod
    def _grads(loss, params, clip):
        names = list(params)
        gs = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
        gs = {k: (g if g is not None else torch.zeros_like(params[k])) for k, g in zip(names, gs)}
        norm = torch.sqrt(sum((g.double() ** 2).sum() for g in gs.values())).float()
        return norm, gs
