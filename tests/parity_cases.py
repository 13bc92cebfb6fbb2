"""Parity cases shared by the ``-m gpu`` tests and ``__graft_entry__.smoke()``: run the CPU
oracle (oracle/dv3_oracle.py) and the CUDA path (through the C ABI via the package's autograd
bindings) on the same seeded inputs and report the differences.

Tolerances (BASELINE.json north_star): categorical indices bit-exact; latents, losses and
gradients within 1e-4 relative, element-wise with an absolute floor (see ``rel``).
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dv3_oracle as O      # noqa: E402
import synth                # noqa: E402

RTOL = 1e-4


ELEM_FLOOR = 0.05


def rel(a, b):
    """Element-wise relative error with an absolute floor: max_i |a_i - b_i| / max(|b_i|,
    ELEM_FLOOR * max|b|).  Elements down to 5 % of the tensor's largest magnitude are held to
    their own 1e-4; smaller ones (whose fp32 rounding error is set by the magnitude of the terms
    summed, not by the result) to 1e-4 of that floor.  Always >= the tensor-max measure
    max|a-b| / max|b| that round 1 used."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0
    floor = ELEM_FLOOR * float(b.abs().max()) + 1e-12
    return float(((a - b).abs() / torch.clamp(b.abs(), min=floor)).max())


def rel_max(a, b):
    """Tensor-max measure (reported next to ``rel`` where useful)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def to_dev(p, device, grad=False):
    return {k: v.detach().clone().to(device).requires_grad_(grad) for k, v in p.items()}


def rssm_list(pkg, p):
    return [p[pkg._lib.RSSM_STATE_KEYS[f]] for f in pkg._lib.RSSM_PARAM_FIELDS]


def actor_list(p, layers, dist):
    out = []
    for i in range(layers):
        out += [p[f"layers.Actor_linear{i}.weight"], p[f"layers.Actor_norm{i}.weight"],
                p[f"layers.Actor_norm{i}.bias"]]
    out += [p["mean_layer.weight"], p["mean_layer.bias"]]
    if dist == "normal":
        out += [p["std_layer.weight"], p["std_layer.bias"]]
    return out


def kdims(d):
    return (d.stoch, d.classes, d.deter, d.hidden, d.actions, d.embed, float(d.unimix))


# --------------------------------------------------------------------------------------
# loss kernels
# --------------------------------------------------------------------------------------
def lambda_return_case(pkg, device, H=14, N=1024, seed=0):
    g = torch.Generator().manual_seed(seed)
    r, v = torch.randn(H, N, 1, generator=g), torch.randn(H, N, 1, generator=g)
    c = torch.rand(H, N, 1, generator=g)
    b = torch.randn(N, 1, generator=g)
    w = torch.randn(H, N, 1, generator=g)
    cpu = [t.clone().requires_grad_(True) for t in (r, v, c, b)]
    ref = O.lambda_return(*cpu, 0.95)
    (ref * w).sum().backward()
    dev = [t.clone().to(device).requires_grad_(True) for t in (r, v, c, b)]
    out = pkg.tools.lambda_return_stacked(*dev, 0.95)
    (out * w.to(device)).sum().backward()
    res = {"ret_maxabs": float((out.detach().cpu() - ref.detach()).abs().max())}
    for name, a, bb in zip(("d_reward", "d_value", "d_pcont", "d_bootstrap"), dev, cpu):
        res[name] = rel(a.grad, bb.grad)
    tup = pkg.tools.lambda_return(dev[0], dev[1], dev[2], dev[3], 0.95, axis=0)
    res["tuple_len"] = len(tup)
    res["tuple_shape"] = tuple(tup[0].shape)
    res["tuple_maxabs"] = float((torch.stack(tup, dim=1).detach().cpu() - ref.detach()).abs().max())
    return res


def twohot_case(pkg, device, H=14, N=256, seed=0):
    g = torch.Generator().manual_seed(seed)
    logits = 3 * torch.randn(H, N, 255, generator=g)
    x = torch.cat([30 * torch.randn(H, N - 6, 1, generator=g),
                   torch.tensor([0.0, 1e9, -1e9, 5.0, 4.85165e8, -1.0]).repeat(H, 1)[..., None]], 1)
    # exact bucket hits: symexp of a bucket value
    bk = O.buckets()
    x[0, :8, 0] = O.symexp(bk[100:108])
    lc = logits.clone().requires_grad_(True)
    lp_ref = O.twohot_logprob(lc, x)
    mean_ref = O.twohot_mean(lc)
    w1, w2 = torch.randn(H, N, generator=g), torch.randn(H, N, 1, generator=g)
    g_lp = torch.autograd.grad((lp_ref * w1).sum(), lc, retain_graph=True)[0]
    g_mean = torch.autograd.grad((mean_ref * w2).sum(), lc)[0]
    ld = logits.clone().to(device).requires_grad_(True)
    dist = pkg.tools.DiscDist(ld)
    lp = dist.log_prob(x.to(device))
    mean = dist.mode()
    d_lp = torch.autograd.grad((lp * w1.to(device)).sum(), ld, retain_graph=True)[0]
    d_mean = torch.autograd.grad((mean * w2.to(device)).sum(), ld)[0]
    return {"log_prob": rel(lp, lp_ref), "mean": rel(mean, mean_ref), "d_log_prob": rel(d_lp, g_lp),
            "d_mean": rel(d_mean, g_mean)}


def kl_case(pkg, device, R=(16, 64), S=32, C=32, seed=0, scale=3.0):
    g = torch.Generator().manual_seed(seed)
    post = scale * torch.randn(*R, S, C, generator=g)
    prior = scale * torch.randn(*R, S, C, generator=g)
    # make some rows fall below the free-nats clip
    prior[0, :8] = post[0, :8] + 0.01 * torch.randn(8, S, C, generator=g)
    w = torch.rand(*R, generator=g)
    pc, qc = post.clone().requires_grad_(True), prior.clone().requires_grad_(True)
    loss, value, dyn, rep = O.kl_balance(pc, qc, 1.0, 0.5, 0.1, 0.01)
    (loss * w).sum().backward()
    pd, qd = post.clone().to(device).requires_grad_(True), prior.clone().to(device).requires_grad_(True)
    l2, v2, d2, r2, pe, qe = pkg.kernels.kl_balance(pd, qd, 1.0, 0.5, 0.1, 0.01)
    (l2 * w.to(device)).sum().backward()
    return {"loss": rel(l2, loss), "value": rel(v2, value), "dyn": rel(d2, dyn), "rep": rel(r2, rep),
            "post_ent": rel(pe, O.onehot_entropy(post, 0.01)),
            "prior_ent": rel(qe, O.onehot_entropy(prior, 0.01)),
            "d_post": rel(pd.grad, pc.grad), "d_prior": rel(qd.grad, qc.grad),
            "clipped_rows": int((value < 1.0).sum())}


def sample_case(pkg, device, M=2048, S=32, C=32, seed=0, unimix=0.01):
    g = torch.Generator().manual_seed(seed)
    logits = 4 * torch.randn(M, S, C, generator=g)
    u = synth.uniforms(g, M, S, C)
    _, idx_ref = O.onehot_sample(logits, u, unimix)
    idx, hot = pkg.kernels.onehot_sample(logits.to(device), u.to(device), unimix)
    mode_ref = torch.argmax(O.unimix_logits(logits, unimix)[0], -1)
    midx, _ = pkg.kernels.onehot_sample(logits.to(device), None, unimix)
    # straight-through backward
    lc = logits.clone().requires_grad_(True)
    s_ref, _ = O.onehot_sample(lc, u, unimix)
    gs = torch.randn(M, S, C, generator=g)
    g_ref = torch.autograd.grad((s_ref * gs).sum(), lc)[0]
    d = pkg.kernels.onehot_st_bwd(logits.to(device), gs.to(device), None, unimix)
    return {"idx_mismatch": int((idx.cpu().long() != idx_ref).sum()),
            "onehot_ok": bool((hot.argmax(-1).cpu() == idx.cpu().long()).all()
                              and float(hot.sum()) == M * S),
            "mode_mismatch": int((midx.cpu().long() != mode_ref).sum()),
            "d_logits": rel(d, g_ref)}


# --------------------------------------------------------------------------------------
# observe
# --------------------------------------------------------------------------------------
def observe_case(pkg, device, config="dmc_proprio", B=16, T=64, seed=0, backward=True,
                 with_state=False):
    d = synth.dims_of(config)
    p = synth.rssm_params(d, seed)
    embed, action, is_first, up, uq = synth.observe_inputs(d, B, T, seed)
    state = None
    if with_state:
        start, _, _ = synth.imagine_inputs(d, B, 1, seed)
        state = start
        is_first[:, 0] = 0.0
        is_first[0, 0] = 1.0
    pc = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ec = embed.clone().requires_grad_(True)
    post_r, prior_r = O.observe(pc, ec, action, is_first, up, uq, d,
                                None if state is None else {k: v.clone() for k, v in state.items()})
    pd = to_dev(p, device, grad=True)
    ed = embed.clone().to(device).requires_grad_(True)
    sidx = sdet = None
    if state is not None:
        sidx = state["stoch"].argmax(-1).to(torch.int32).to(device)
        sdet = state["deter"].to(device)
    outs = pkg.kernels.observe(ed, action.to(device), is_first.to(device), up.to(device),
                               uq.to(device), sidx, sdet, kdims(d), rssm_list(pkg, pd))
    post_stoch, post_logit, prior_stoch, prior_logit, deter, aprev, post_idx, prior_idx = outs
    res = {
        "post_idx_mismatch": int((post_idx.cpu().long() != post_r["stoch"].argmax(-1)).sum()),
        "prior_idx_mismatch": int((prior_idx.cpu().long() != prior_r["stoch"].argmax(-1)).sum()),
        "post_stoch_maxabs": float((post_stoch.cpu() - post_r["stoch"].detach()).abs().max()),
        "prior_stoch_maxabs": float((prior_stoch.cpu() - prior_r["stoch"].detach()).abs().max()),
        "deter": rel(deter, post_r["deter"]),
        "post_logit": rel(post_logit, post_r["logit"]),
        "prior_logit": rel(prior_logit, prior_r["logit"]),
    }
    if backward:
        g = torch.Generator().manual_seed(seed + 7)
        F_ = d.flat + d.deter
        w_feat = torch.randn(B, T, F_, generator=g)
        w_prior = 0.1 * torch.randn(B, T, d.stoch, d.classes, generator=g)
        w_pl = 0.1 * torch.randn(B, T, d.stoch, d.classes, generator=g)

        def scalar(post, prior, kl, dev):
            feat = torch.cat([post["stoch"].reshape(B, T, -1), post["deter"]], -1)
            return ((feat * w_feat.to(dev)).sum() + (prior["stoch"] * w_prior.to(dev)).sum()
                    + (prior["logit"] * w_pl.to(dev)).sum() + kl.mean() * 50.0)

        kl_r = O.kl_balance(post_r["logit"], prior_r["logit"], 1.0, 0.5, 0.1, d.unimix)[0]
        scalar(post_r, prior_r, kl_r, "cpu").backward()
        kl_d = pkg.kernels.kl_balance(post_logit, prior_logit, 1.0, 0.5, 0.1, d.unimix)[0]
        scalar({"stoch": post_stoch, "deter": deter}, {"stoch": prior_stoch, "logit": prior_logit},
               kl_d, device).backward()
        res["d_embed"] = rel(ed.grad, ec.grad)
        for k in p:
            res[f"d_{k}"] = rel(pd[k].grad, pc[k].grad)
    return res


def obs_step_teacher_forced_case(pkg, device, config="dmc_proprio", B=16, T=6, seed=0):
    """Per-step parity with the oracle's own previous state fed back (the north-star criterion)."""
    d = synth.dims_of(config)
    p = synth.rssm_params(d, seed)
    embed, action, is_first, up, uq = synth.observe_inputs(d, B, T, seed)
    pd = to_dev(p, device)
    res = {"idx_mismatch": 0, "deter": 0.0, "post_logit": 0.0, "prior_logit": 0.0}
    prev = None
    with torch.no_grad():
        for t in range(T):
            post, prior = O.obs_step(p, prev, action[:, t], embed[:, t], is_first[:, t], up[t],
                                     uq[t], d)
            sidx = sdet = None
            first = is_first[:, t:t + 1].clone()
            if prev is not None:
                sidx = prev["stoch"].argmax(-1).to(torch.int32).to(device)
                sdet = prev["deter"].to(device)
            outs = pkg.kernels.observe(embed[:, t:t + 1].contiguous().to(device),
                                       action[:, t:t + 1].contiguous().to(device),
                                       first.to(device), up[t:t + 1].to(device),
                                       uq[t:t + 1].to(device), sidx, sdet, kdims(d),
                                       rssm_list(pkg, pd))
            res["idx_mismatch"] += int((outs[6][:, 0].cpu().long() != post["stoch"].argmax(-1)).sum())
            res["idx_mismatch"] += int((outs[7][:, 0].cpu().long() != prior["stoch"].argmax(-1)).sum())
            res["deter"] = max(res["deter"], rel(outs[4][:, 0], post["deter"]))
            res["post_logit"] = max(res["post_logit"], rel(outs[1][:, 0], post["logit"]))
            res["prior_logit"] = max(res["prior_logit"], rel(outs[3][:, 0], prior["logit"]))
            prev = post
    return res


# --------------------------------------------------------------------------------------
# imagine
# --------------------------------------------------------------------------------------
def imagine_case(pkg, device, config="dmc_proprio", N=1024, H=15, seed=0, backward=True):
    c = synth.CONFIGS[config]
    d = synth.dims_of(config)
    dist, layers = c["actor_dist"], c["actor_layers"]
    p = synth.rssm_params(d, seed)
    pa = synth.actor_params(config, seed + 1)
    start, act_noise, u_state = synth.imagine_inputs(d, N, H, seed, dist)
    pac = {k: v.clone().requires_grad_(True) for k, v in pa.items()}
    feats_r, states_r, actions_r = O.imagine(p, pac, start, H, act_noise, u_state, d, layers, dist)
    pd = to_dev(p, device)
    pad = to_dev(pa, device, grad=True)
    spec = pkg.kernels.ActorSpec(layers, c["units"], dist, 0.1, 1.0, 0.01)
    feat, logit, action, idx = pkg.kernels.imagine(
        start["stoch"].argmax(-1).to(torch.int32).to(device), start["deter"].to(device),
        act_noise.to(device), u_state.to(device), None, H, kdims(d), spec, rssm_list(pkg, pd),
        actor_list(pad, layers, dist), start_logit=start["logit"].to(device))
    SC = d.flat
    res = {
        "idx_mismatch": int((idx.cpu().long() != states_r["stoch"].argmax(-1)).sum()),
        "feat": rel(feat, feats_r),
        "deter": rel(feat[..., SC:], states_r["deter"]),
        "logit": rel(logit, states_r["logit"]),
        "action": rel(action, actions_r),
    }
    if dist == "onehot":
        res["action_mismatch"] = int((action.argmax(-1).cpu() != actions_r.argmax(-1)).sum())
    if backward:
        g = torch.Generator().manual_seed(seed + 9)
        w_feat = torch.randn(H, N, SC + d.deter, generator=g)
        w_log = 0.1 * torch.randn(H, N, d.stoch, d.classes, generator=g)
        w_act = 0.1 * torch.randn(H, N, d.actions, generator=g)
        ((O.get_feat(states_r) * w_feat).sum() + (states_r["logit"] * w_log).sum()
         + (actions_r * w_act).sum()).backward()
        ((feat * w_feat.to(device)).sum() + (logit * w_log.to(device)).sum()
         + (action * w_act.to(device)).sum()).backward()
        for k in pa:
            res[f"d_actor.{k}"] = rel(pad[k].grad, pac[k].grad)
    return res


def imagine_with_action_case(pkg, device, config="dmc_proprio", B=6, T=8, seed=0):
    d = synth.dims_of(config)
    p = synth.rssm_params(d, seed)
    start, _, u = synth.imagine_inputs(d, B, T + 1, seed)
    g = torch.Generator().manual_seed(seed + 3)
    action = torch.rand(B, T, d.actions, generator=g) * 2 - 1
    with torch.no_grad():
        ref = O.imagine_with_action(p, action, start, u[:T], d)
        pd = to_dev(p, device)
        feat, logit, _, idx = pkg.kernels.imagine(
            start["stoch"].argmax(-1).to(torch.int32).to(device), start["deter"].to(device), None,
            u.to(device), action.permute(1, 0, 2).contiguous().to(device), T + 1, kdims(d), None,
            rssm_list(pkg, pd), [])
    return {"idx_mismatch": int((idx[1:].permute(1, 0, 2).cpu().long()
                                 != ref["stoch"].argmax(-1)).sum()),
            "deter": rel(feat[1:, :, d.flat:].permute(1, 0, 2), ref["deter"]),
            "logit": rel(logit[1:].permute(1, 0, 2, 3), ref["logit"])}


def policy_walk_case(pkg, device, config="dmc_proprio", n=5, steps=5, seed=0):
    """Dreamer._policy's call order (reference dreamer.py:117-190) through the PUBLIC methods:
    obs_step(None, None, embed, is_first) on the first call, then obs_step(latent, action, ...)
    with a mid-episode reset, get_feat, actor(feat).sample() / .mode() / .log_prob(action) --
    against the oracle fed the same uniforms / normals; the state fed back is the product's own
    (free-running, as in acting)."""
    c = synth.CONFIGS[config]
    d = synth.dims_of(config)
    p = synth.rssm_params(d, seed)
    pa = synth.actor_params(config, seed + 1)
    cfgs = pkg.configs
    cfg = cfgs.make_config("dmc_proprio", device=device, num_actions=d.actions, dyn_stoch=d.stoch,
                           dyn_discrete=d.classes, dyn_deter=d.deter, dyn_hidden=d.hidden,
                           units=c["units"])
    dyn = pkg.networks.RSSM(cfg.dyn_stoch, cfg.dyn_deter, cfg.dyn_hidden, 1, cfg.dyn_discrete,
                            cfg.act, cfg.norm, cfg.dyn_mean_act, cfg.dyn_std_act, cfg.dyn_min_std,
                            cfg.unimix_ratio, cfg.initial, d.actions, d.embed, device).to(device)
    dyn.load_state_dict({k: v.to(device) for k, v in p.items()}, strict=True)
    actor = pkg.networks.MLP(d.flat + d.deter, (d.actions,), c["actor_layers"], c["units"], "SiLU",
                             True, "normal", "learned", 0.1, 1.0, absmax=1.0, name="Actor").to(device)
    actor.load_state_dict({k: v.to(device) for k, v in pa.items()}, strict=True)
    g = torch.Generator().manual_seed(seed + 11)
    res = {"idx_mismatch": 0, "deter": 0.0, "post_logit": 0.0, "prior_logit": 0.0, "feat": 0.0,
           "action": 0.0, "mode": 0.0, "logprob": 0.0, "mode_state_mismatch": 0}
    latent = action = None        # product state
    o_prev = o_act = None         # oracle state
    with torch.no_grad():
        for t in range(steps):
            embed = torch.randn(n, d.embed, generator=g)
            is_first = torch.zeros(n)
            if t == 0:
                is_first[:] = 1.0
            if t == 2:
                is_first[1] = 1.0
            if t == 3:
                is_first[:] = 1.0               # the all-rows-reset branch (networks.py:176)
            up, uq = synth.uniforms(g, n, d.stoch, d.classes), synth.uniforms(g, n, d.stoch, d.classes)
            eps = torch.randn(n, d.actions, generator=g)
            sample = t != 4                     # last step: obs_step(sample=False) -> mode
            post_o, prior_o = O.obs_step(p, o_prev, o_act, embed, is_first, up,
                                         uq if sample else None, d)
            post, prior = dyn.obs_step(latent, action, embed.to(device), is_first.to(device),
                                       sample=sample, noise=(up.to(device), uq.to(device)))
            res["idx_mismatch"] += int((post["stoch"].argmax(-1).cpu() != post_o["stoch"].argmax(-1)).sum())
            res["idx_mismatch"] += int((prior["stoch"].argmax(-1).cpu() != prior_o["stoch"].argmax(-1)).sum())
            res["deter"] = max(res["deter"], rel(post["deter"], post_o["deter"]))
            res["post_logit"] = max(res["post_logit"], rel(post["logit"], post_o["logit"]))
            res["prior_logit"] = max(res["prior_logit"], rel(prior["logit"], prior_o["logit"]))
            feat = dyn.get_feat(post)
            feat_o = O.get_feat(post_o)
            res["feat"] = max(res["feat"], rel(feat, feat_o))
            dist = actor(feat)
            mean_o, std_o = O.actor_normal_stats(pa, feat_o, c["actor_layers"])
            act = dist.sample(eps=eps.to(device))
            act_o = O.contdist_sample(mean_o, std_o, eps)
            res["action"] = max(res["action"], rel(act, act_o))
            res["mode"] = max(res["mode"], rel(dist.mode(), O.contdist_sample(mean_o, std_o, torch.zeros_like(eps))))
            res["logprob"] = max(res["logprob"], rel(dist.log_prob(act), O.normal_logprob(mean_o, std_o, act_o)))
            latent = {k: v.detach() for k, v in post.items()}
            action = act.detach()
            o_prev, o_act = post_o, act_o
        # img_step public method, sample and mode
        u = synth.uniforms(g, n, d.stoch, d.classes)
        nxt = dyn.img_step(latent, action, noise=u.to(device))
        nxt_o = O.img_step(p, o_prev, o_act, u, d)
        res["idx_mismatch"] += int((nxt["stoch"].argmax(-1).cpu() != nxt_o["stoch"].argmax(-1)).sum())
        res["img_step_logit"] = rel(nxt["logit"], nxt_o["logit"])
        nm = dyn.img_step(latent, action, sample=False)
        nm_o = O.img_step(p, o_prev, o_act, None, d)
        res["mode_state_mismatch"] = int((nm["stoch"].argmax(-1).cpu() != nm_o["stoch"].argmax(-1)).sum())
    return res


# --------------------------------------------------------------------------------------
def check(res, exact=(), tol=RTOL, skip=()):
    """-> list of failure strings."""
    bad = []
    for k, v in res.items():
        if k in skip or isinstance(v, (tuple, bool)):
            continue
        if k.endswith("mismatch") or k in exact:
            if v != 0:
                bad.append(f"{k}={v} (must be 0)")
        elif k.endswith("maxabs"):
            if v != 0.0:
                bad.append(f"{k}={v:.3e} (must be exactly 0)")
        elif isinstance(v, float) and not (v <= tol):
            bad.append(f"{k}={v:.3e} > {tol}")
    return bad


def run_smoke(pkg, device="cuda:0"):
    lines = []
    for name, fn, kw in (
        ("lambda_return", lambda_return_case, dict(H=14, N=256)),
        ("twohot", twohot_case, dict(H=4, N=64)),
        ("kl_balance", kl_case, dict(R=(4, 8))),
        ("observe", observe_case, dict(config="dmc_proprio", B=4, T=6)),
        ("imagine", imagine_case, dict(config="dmc_proprio", N=32, H=5)),
    ):
        res = fn(pkg, device, **kw)
        bad = check(res, skip=("clipped_rows", "tuple_len"))
        worst = max([v for k, v in res.items() if isinstance(v, float)] + [0.0])
        lines.append(f"{name}: worst rel err {worst:.2e}" + (f"  FAIL {bad}" if bad else ""))
        if bad:
            raise AssertionError(f"smoke {name}: {bad}")
    return lines


# --------------------------------------------------------------------------------------
# whole train step: WorldModel._train + ImagBehavior._train forward/backward vs the oracle
# --------------------------------------------------------------------------------------
def build_product_agent(pkg, device, config, P, Pa, Pv, **cfg_over):
    cfgs = pkg.configs
    suite = {"dmc_proprio": "dmc_proprio", "tiny": "dmc_proprio"}.get(config, config)
    c = synth.CONFIGS[config]
    d = synth.dims_of(config)
    over = dict(device=device, device_metrics=True, num_actions=d.actions, dyn_stoch=d.stoch,
                dyn_discrete=d.classes, dyn_deter=d.deter, dyn_hidden=d.hidden, units=c["units"])
    over.update(cfg_over)
    cfg = cfgs.make_config(suite, **over)
    wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
    wm.load_state_dict({k: v.to(device) for k, v in P.items()}, strict=True)
    beh = pkg.models.ImagBehavior(cfg, wm)
    beh.actor.load_state_dict({k: v.to(device) for k, v in Pa.items()}, strict=True)
    beh.value.load_state_dict({k: v.to(device) for k, v in Pv.items()}, strict=True)
    beh._slow_value.load_state_dict({k: v.to(device) for k, v in Pv.items()}, strict=True)
    return cfg, wm, beh


def train_step_case(pkg, device, config="dmc_proprio", B=16, T=64, H=15, seed=0):
    import train_step as TS
    d = synth.dims_of(config)
    c = synth.CONFIGS[config]
    enc_units = d.embed        # the MLP encoder width is the RSSM embed width
    P, Pa, Pv = synth.agent_params(config, seed, enc_units=enc_units)
    ocfg = TS.make_cfg(dyn_stoch=d.stoch, dyn_discrete=d.classes, units=c["units"],
                       enc_units=enc_units, dec_units=enc_units, imag_horizon=H,
                       actor_layers=c["actor_layers"], actor_dist=c["actor_dist"])
    agent = TS.Agent(P, Pa, Pv, ocfg, d)
    data = synth.replay_batch(d, B, T, seed, resets=((1, 3), (2, T // 2)))
    noise = synth.train_noise(d, B, T, H, seed, c["actor_dist"])
    ref = agent.train_step({k: v.copy() for k, v in data.items()}, noise, apply=False)

    over = dict(imag_horizon=H)
    over["encoder"] = dict(mlp_units=enc_units)
    over["decoder"] = dict(mlp_units=enc_units)
    cfg, wm, beh = build_product_agent(pkg, device, config, P, Pa, Pv, **over)
    nd = {k: v.to(device) for k, v in noise.items()}
    pdata = wm.preprocess(data)
    with pkg.tools.RequiresGrad(wm):
        loss, post, aux = wm.loss(pdata, (nd["u_prior"], nd["u_post"]))
        names = [k for k, _ in wm.named_parameters()]
        grads = torch.autograd.grad(loss, list(wm.parameters()), allow_unused=True)
    res = {"model_loss": rel(loss, ref["model_loss"]),
           "post_idx_mismatch": int((post["stoch"].argmax(-1).cpu()
                                     != ref["post"]["stoch"].argmax(-1)).sum())}
    for k, v in aux["losses"].items():
        res[f"{k}_loss"] = rel(v, ref[f"{k}_loss"])
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads if g is not None)).float()
    res["model_grad_norm"] = rel(gn, ref["model_grad_norm"])
    for k, g in zip(names, grads):
        r = ref["grads"]["wm"][k]
        if g is None:
            res[f"dwm.{k}_maxabs"] = float(r.abs().max())
        else:
            res[f"dwm.{k}"] = rel(g, r)
    start = {k: v.detach() for k, v in post.items()}
    beh._update_slow_target()
    reward_fn = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
    a_loss, v_loss, roll, mets, baux = beh.losses(start, reward_fn, (nd["act_noise"], nd["u_state"]))
    with pkg.tools.RequiresGrad(beh):
        ga = torch.autograd.grad(a_loss, list(beh.actor.parameters()), allow_unused=True)
        gv = torch.autograd.grad(v_loss, list(beh.value.parameters()), allow_unused=True)
    imag = ref["imag"]
    res["imag_idx_mismatch"] = int((roll[1]["stoch"].argmax(-1).cpu()
                                    != imag["states"]["stoch"].argmax(-1)).sum())
    res["imag_feat"] = rel(roll[0], imag["feats"])
    res["imag_action"] = rel(roll[2], imag["actions"])
    res["imag_reward"] = rel(baux["reward"], imag["reward"])
    res["target"] = rel(baux["target"], imag["target"])
    res["weights"] = rel(roll[3], imag["weights"])
    res["actor_loss"] = rel(a_loss, ref["actor_loss"])
    res["value_loss"] = rel(v_loss, ref["value_loss"])
    for (k, _), g in zip(beh.actor.named_parameters(), ga):
        res[f"dactor.{k}"] = rel(g, ref["grads"]["actor"][k])
    for (k, _), g in zip(beh.value.named_parameters(), gv):
        res[f"dvalue.{k}"] = rel(g, ref["grads"]["value"][k])
    return res
