"""Generate the golden fixtures from the LIVE reference (build container only).

    python tests/golden/make_golden.py        # writes tests/golden/*.pt

Runs the unmodified reference modules from /root/reference on CPU (shims and the noise tape of
oracle/ref_harness.py) at reduced widths so the fixtures stay small, and stores inputs,
parameters, supplied noise and the reference's outputs.  The committed fixtures are what the
GPU box checks against (it has no /root/reference).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as H      # noqa: E402

TINY = dict(dyn_hidden=32, dyn_deter=48, dyn_stoch=8, dyn_discrete=8, units=32, imag_horizon=4,
            batch_size=3, batch_length=5)


def uniforms(g, *shape):
    return torch.rand(*shape, generator=g).clamp_(1e-30, 1.0)


def perturb(modules, g, scale=0.05):
    with torch.no_grad():
        for m in modules:
            for p in m.parameters():
                p.add_(scale * torch.randn(p.shape, generator=g))


def sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def make_agent(actor_dist="normal", seed=0, imag_gradient="dynamics", A=3):
    over = dict(TINY)
    over["num_actions"] = A
    cfg = H.reference_config(("dmc_proprio",), **over)
    cfg.encoder = dict(cfg.encoder, mlp_units=40, mlp_layers=2)
    cfg.decoder = dict(cfg.decoder, mlp_units=40, mlp_layers=2)
    if actor_dist == "onehot":
        cfg.actor = dict(cfg.actor, dist="onehot", std="none")
    cfg.imag_gradient = imag_gradient
    wm, beh = H.build_agent(cfg, H.PROPRIO_SHAPES, seed)
    g = torch.Generator().manual_seed(seed + 11)
    perturb([wm, beh.actor, beh.value], g)
    beh._slow_value.load_state_dict(beh.value.state_dict())
    return cfg, wm, beh


def golden_ops(seed=0):
    """lambda_return, DiscDist, kl_loss: outputs + gradients of the reference."""
    rtools, rnet, _ = H.load_reference()
    g = torch.Generator().manual_seed(seed)
    out = {}
    Hh, N = 6, 10
    r, v = torch.randn(Hh, N, 1, generator=g), torch.randn(Hh, N, 1, generator=g)
    c, b = torch.rand(Hh, N, 1, generator=g), torch.randn(N, 1, generator=g)
    w = torch.randn(Hh, N, 1, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (r, v, c, b)]
    ret = torch.stack(rtools.lambda_return(*leaves, 0.95, axis=0), dim=1)
    grads = torch.autograd.grad((ret * w).sum(), leaves)
    out["lambda_return"] = dict(reward=r, value=v, pcont=c, bootstrap=b, w=w, ret=ret.detach(),
                                grads=[x.detach() for x in grads])
    logits = (3 * torch.randn(Hh, N, 255, generator=g)).requires_grad_(True)
    x = torch.cat([30 * torch.randn(Hh, N - 4, 1, generator=g),
                   torch.tensor([0.0, 1e9, -1e9, 5.0]).repeat(Hh, 1)[..., None]], 1)
    dist = rtools.DiscDist(logits, device="cpu")
    lp, mean = dist.log_prob(x), dist.mean()
    w1, w2 = torch.randn(Hh, N, generator=g), torch.randn(Hh, N, 1, generator=g)
    out["twohot"] = dict(logits=logits.detach(), x=x, log_prob=lp.detach(), mean=mean.detach(), w1=w1,
                         w2=w2,
                         d_log_prob=torch.autograd.grad((lp * w1).sum(), logits, retain_graph=True)[0],
                         d_mean=torch.autograd.grad((mean * w2).sum(), logits)[0])
    return out


def golden_rollouts(seed=0):
    """RSSM.observe (+kl_loss, gradients) and ImagBehavior._imagine (+actor gradients)."""
    out = {}
    for dist in ("normal", "onehot"):
        cfg, wm, beh = make_agent(dist, seed, A=3 if dist == "normal" else 5)
        dyn = wm.dynamics
        A = cfg.num_actions
        g = torch.Generator().manual_seed(seed + 1)
        B, T, E = 3, 5, 40
        embed = torch.randn(B, T, E, generator=g)
        action = torch.rand(B, T, A, generator=g) * 2 - 1
        is_first = torch.zeros(B, T)
        is_first[:, 0] = 1
        is_first[1, 2] = 1
        up, uq = uniforms(g, T, B, 8, 8), uniforms(g, T, B, 8, 8)
        tape = []
        for t in range(T):
            tape += [("u", up[t]), ("u", uq[t])]
        dyn.requires_grad_(True)
        e = embed.clone().requires_grad_(True)
        with H.NoiseTape(tape):
            post, prior = dyn.observe(e, action.clone(), is_first.clone())
            kl = dyn.kl_loss(post, prior, 1.0, 0.5, 0.1)
        w = torch.randn(B, T, 64 + 48, generator=g)
        ((dyn.get_feat(post) * w).sum() + 20 * kl[0].mean()).backward()
        rec = dict(params=sd(dyn), embed=embed, action=action, is_first=is_first, u_prior=up, u_post=uq, w=w,
                   post={k: v.detach() for k, v in post.items()},
                   prior={k: v.detach() for k, v in prior.items()},
                   kl=[x.detach() for x in kl], d_embed=e.grad.clone(),
                   grads={k: p.grad.clone() for k, p in dyn.named_parameters()})
        dyn.zero_grad()
        dyn.requires_grad_(False)
        # imagination from the posterior
        Hh, N = 4, B * T
        start = {k: v.detach() for k, v in post.items()}
        noise_a = (torch.randn(Hh, N, A, generator=g) if dist == "normal" else uniforms(g, Hh, N, A))
        us = uniforms(g, Hh, N, 8, 8)
        tape = []
        for k in range(Hh):
            tape += [("n" if dist == "normal" else "u", noise_a[k]), ("u", us[k])]
        beh.actor.requires_grad_(True)
        with H.NoiseTape(tape):
            feats, states, actions = beh._imagine(start, beh.actor, Hh)
        w2 = torch.randn(Hh, N, 64 + 48, generator=g)
        (dyn.get_feat(states) * w2).sum().backward()
        rec["imagine"] = dict(actor=sd(beh.actor), act_noise=noise_a, u_state=us, w=w2, feats=feats.detach(),
                              states={k: v.detach() for k, v in states.items()}, actions=actions.detach(),
                              grads={k: p.grad.clone() for k, p in beh.actor.named_parameters()})
        beh.actor.zero_grad()
        out[dist] = rec
    return out


def golden_train(seed=0, steps=2):
    """Two full Dreamer._train steps (WorldModel._train + ImagBehavior._train, Adam included)."""
    out = {}
    for dist, grad in (("normal", "dynamics"), ("onehot", "reinforce")):
        A = 3 if dist == "normal" else 5
        cfg, wm, beh = make_agent(dist, seed + 5, grad, A=A)
        B, T, Hh, N = 3, 5, 4, 15
        rec = dict(wm=sd(wm), actor=sd(beh.actor), value=sd(beh.value), steps=[])
        reward_fn = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
        for i in range(steps):
            data = H.synthetic_batch(B, T, A, seed + i, onehot_action=(dist == "onehot"),
                                     resets=((1, 2),))
            data.pop("image")
            g = torch.Generator().manual_seed(seed + 20 + i)
            noise = dict(u_prior=uniforms(g, T, B, 8, 8), u_post=uniforms(g, T, B, 8, 8),
                         act_noise=(torch.randn(Hh, N, A, generator=g) if dist == "normal"
                                    else uniforms(g, Hh, N, A)),
                         u_state=uniforms(g, Hh, N, 8, 8))
            tape = []
            for t in range(T):
                tape += [("u", noise["u_prior"][t]), ("u", noise["u_post"][t])]
            for k in range(Hh):
                tape += [("n" if dist == "normal" else "u", noise["act_noise"][k]), ("u", noise["u_state"][k])]
            feed = {k: v.copy() for k, v in data.items()}
            feed["image"] = np.zeros((B, T, 2, 2, 3), np.uint8)     # preprocess divides it unconditionally
            with H.NoiseTape(tape) as tp, H.quiet():
                post, context, m1 = wm._train(feed)
                _, _, _, _, m2 = beh._train(post, reward_fn)
            assert tp.pos == len(tape)
            rec["steps"].append(dict(
                data=data, noise=noise,
                # np.array(copy): on CPU the reference's to_np() aliases live buffers (ema_vals)
                metrics={k: torch.tensor(np.array(v, copy=True)) for k, v in {**m1, **m2}.items()},
                post={k: v.detach().clone() for k, v in post.items()},
                wm_after=sd(wm), actor_after=sd(beh.actor), value_after=sd(beh.value),
                slow_after=sd(beh._slow_value), ema_after=beh.ema_vals.clone()))
        rec["cfg"] = dict(actor_dist=dist, imag_gradient=grad, num_actions=A, **TINY)
        out[dist] = rec
    return out


def main():
    assert H.available(), "needs /root/reference"
    torch.set_num_threads(4)
    torch.save(golden_ops(), os.path.join(HERE, "ops.pt"))
    torch.save(golden_rollouts(), os.path.join(HERE, "rollouts.pt"))
    torch.save(golden_train(), os.path.join(HERE, "train.pt"))
    for f in ("ops.pt", "rollouts.pt", "train.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
