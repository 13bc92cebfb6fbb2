"""Pin the oracle against the live reference (build container only).

Runs the reference modules from /root/reference on CPU with supplied noise and checks
that ``dv3_oracle`` reproduces them: categorical indices bit-exact, floats to ~1e-6
(both are fp32 torch-CPU, differences come only from op ordering), gradients included.

    python oracle/pin_against_reference.py            # prints a table, exits non-zero on mismatch
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dv3_oracle as O          # noqa: E402
import ref_harness as H         # noqa: E402


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def uniforms(gen, *shape):
    return torch.rand(*shape, generator=gen).clamp_(1e-30, 1.0)


def pin_observe(results, B=4, T=6, seed=0):
    cfg = H.reference_config()
    wm, _ = H.build_agent(cfg, H.PROPRIO_SHAPES, seed)
    dyn = wm.dynamics
    d = O.RSSMDims(actions=6, embed=1024)
    g = torch.Generator().manual_seed(seed + 1)
    embed = torch.randn(B, T, 1024, generator=g)
    action = torch.rand(B, T, 6, generator=g) * 2 - 1
    is_first = torch.zeros(B, T)
    is_first[:, 0] = 1
    is_first[1, 3] = 1
    is_first[2, 4] = 1
    up, uq = uniforms(g, T, B, 32, 32), uniforms(g, T, B, 32, 32)
    # perturb params so LN affine / biases / W are non-trivial
    with torch.no_grad():
        for prm in dyn.parameters():
            prm.add_(0.05 * torch.randn(prm.shape, generator=g))
    params = {k: v.detach().clone().requires_grad_(True) for k, v in dyn.state_dict().items()}
    tape = []
    for t in range(T):
        tape += [("u", up[t]), ("u", uq[t])]
    dyn.requires_grad_(True)
    e_ref = embed.clone().requires_grad_(True)
    with H.NoiseTape(tape):
        post_r, prior_r = dyn.observe(e_ref, action.clone(), is_first.clone())
        loss_r, val_r, dyn_r, rep_r = dyn.kl_loss(post_r, prior_r, 1.0, 0.5, 0.1)
    e_or = embed.clone().requires_grad_(True)
    post_o, prior_o = O.observe(params, e_or, action, is_first, up, uq, d)
    loss_o, val_o, dyn_o, rep_o = O.kl_balance(post_o["logit"], prior_o["logit"], 1.0, 0.5, 0.1, 0.01)
    for k in ("stoch", "deter", "logit"):
        results.append((f"observe post.{k}", _rel(post_o[k], post_r[k])))
        results.append((f"observe prior.{k}", _rel(prior_o[k], prior_r[k])))
    results.append(("observe post idx exact", float((post_o["stoch"].argmax(-1) != post_r["stoch"].argmax(-1)).sum())))
    results.append(("observe prior idx exact", float((prior_o["stoch"].argmax(-1) != prior_r["stoch"].argmax(-1)).sum())))
    for n, a, b in (("kl loss", loss_o, loss_r), ("kl value", val_o, val_r), ("kl dyn", dyn_o, dyn_r), ("kl rep", rep_o, rep_r)):
        results.append((n, _rel(a, b)))
    w = torch.randn(B, T, 1536, generator=g)

    def scalar(post, loss):
        return (O.get_feat(post) * w).sum() + loss.mean()

    scalar(post_r, loss_r).backward()
    scalar(post_o, loss_o).backward()
    results.append(("observe d embed", _rel(e_or.grad, e_ref.grad)))
    ref_grads = dict(dyn.named_parameters())
    for k, v in params.items():
        results.append((f"observe d {k}", _rel(v.grad, ref_grads[k].grad)))


def pin_imagine(results, N=8, Hh=5, seed=0):
    cfg = H.reference_config()
    wm, beh = H.build_agent(cfg, H.PROPRIO_SHAPES, seed)
    dyn = wm.dynamics
    d = O.RSSMDims(actions=6, embed=1024)
    g = torch.Generator().manual_seed(seed + 2)
    with torch.no_grad():
        for prm in list(dyn.parameters()) + list(beh.actor.parameters()):
            prm.add_(0.05 * torch.randn(prm.shape, generator=g))
    idx = torch.randint(0, 32, (1, N, 32), generator=g)
    start = {"stoch": torch.nn.functional.one_hot(idx, 32).float(),
             "deter": torch.tanh(torch.randn(1, N, 512, generator=g)),
             "logit": torch.randn(1, N, 32, 32, generator=g)}
    eps = torch.randn(Hh, N, 6, generator=g)
    us = uniforms(g, Hh, N, 32, 32)
    tape = []
    for k in range(Hh):
        tape += [("n", eps[k]), ("u", us[k])]
    beh.actor.requires_grad_(True)
    with H.NoiseTape(tape):
        feats_r, states_r, actions_r = beh._imagine(start, beh.actor, Hh)
    p_rssm = {k: v.detach().clone() for k, v in dyn.state_dict().items()}
    p_act = {k: v.detach().clone().requires_grad_(True) for k, v in beh.actor.state_dict().items()}
    flat = {k: v.reshape([-1] + list(v.shape[2:])) for k, v in start.items()}
    feats_o, states_o, actions_o = O.imagine(p_rssm, p_act, flat, Hh, eps, us, d, actor_layers=2)
    results.append(("imagine feats", _rel(feats_o, feats_r)))
    results.append(("imagine actions", _rel(actions_o, actions_r)))
    for k in ("stoch", "deter", "logit"):
        results.append((f"imagine states.{k}", _rel(states_o[k], states_r[k])))
    results.append(("imagine idx exact", float((states_o["stoch"].argmax(-1) != states_r["stoch"].argmax(-1)).sum())))
    w = torch.randn(Hh, N, 1536, generator=g)
    (dyn.get_feat(states_r) * w).sum().backward()
    (O.get_feat(states_o) * w).sum().backward()
    ref_grads = dict(beh.actor.named_parameters())
    for k, v in p_act.items():
        results.append((f"imagine d actor.{k}", _rel(v.grad, ref_grads[k].grad)))


def pin_losses(results, seed=0):
    rtools, _, _ = H.load_reference()
    g = torch.Generator().manual_seed(seed + 3)
    Hh, N = 14, 64
    reward = torch.randn(Hh, N, 1, generator=g)
    value = torch.randn(Hh, N, 1, generator=g)
    pcont = torch.rand(Hh, N, 1, generator=g)
    boot = torch.randn(N, 1, generator=g)
    ref = torch.stack(rtools.lambda_return(reward, value, pcont, boot, 0.95, axis=0), dim=1)
    results.append(("lambda_return", _rel(O.lambda_return(reward, value, pcont, boot, 0.95), ref)))
    logits = (3 * torch.randn(Hh, N, 255, generator=g)).requires_grad_(True)
    x = torch.cat([30 * torch.randn(Hh, N - 4, 1, generator=g),
                   torch.tensor([0.0, 1e9, -1e9, 5.0]).repeat(Hh, 1)[..., None]], 1)
    dist = rtools.DiscDist(logits, device="cpu")
    lp_r = dist.log_prob(x)
    lp_o = O.twohot_logprob(logits, x)
    results.append(("twohot log_prob", _rel(lp_o, lp_r)))
    results.append(("twohot mean", _rel(O.twohot_mean(logits), dist.mean())))
    gr = torch.autograd.grad(lp_r.sum(), logits)[0]
    go = torch.autograd.grad(lp_o.sum(), logits)[0]
    results.append(("twohot d logits", _rel(go, gr)))


def main():
    if not H.available():
        print("reference tree not present; nothing to pin against")
        return 0
    torch.set_num_threads(os.cpu_count())
    results = []
    pin_losses(results)
    pin_observe(results)
    pin_imagine(results)
    bad = 0
    for name, err in results:
        exact = name.endswith("exact")
        ok = (err == 0) if exact else (err < 2e-5)
        bad += not ok
        print(f"{'ok ' if ok else 'BAD'} {name:55s} {err:.3e}")
    print("PINNED" if not bad else f"{bad} MISMATCHES")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
