"""-m gpu: the CUDA path against the golden fixtures = outputs of the live reference itself
(tests/golden/make_golden.py), with no oracle in between."""
import os

import pytest
import torch

import parity_cases as pc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4          # BASELINE.json north_star: 1e-4 relative, indices bit-exact


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def dims(A):
    return pc.O.RSSMDims(stoch=8, classes=8, deter=48, hidden=32, actions=A, embed=40)


def test_lambda_return_and_twohot_vs_reference(pkg, device):
    ops = load("ops.pt")
    g = ops["lambda_return"]
    leaves = [g[k].clone().to(device).requires_grad_(True) for k in ("reward", "value", "pcont", "bootstrap")]
    ret = pkg.tools.lambda_return_stacked(*leaves, 0.95)
    assert torch.equal(ret.detach().cpu(), g["ret"])        # bit-identical to the reference
    grads = torch.autograd.grad((ret * g["w"].to(device)).sum(), leaves)
    for a, b in zip(grads, g["grads"]):
        assert pc.rel(a, b) < TOL
    t = ops["twohot"]
    logits = t["logits"].clone().to(device).requires_grad_(True)
    dist = pkg.tools.DiscDist(logits)
    lp, mean = dist.log_prob(t["x"].to(device)), dist.mean()
    assert pc.rel(lp, t["log_prob"]) < TOL and pc.rel(mean, t["mean"]) < TOL
    assert pc.rel(torch.autograd.grad((lp * t["w1"].to(device)).sum(), logits, retain_graph=True)[0],
                  t["d_log_prob"]) < TOL
    assert pc.rel(torch.autograd.grad((mean * t["w2"].to(device)).sum(), logits)[0], t["d_mean"]) < TOL


@pytest.mark.parametrize("dist", ["normal", "onehot"])
def test_observe_imagine_vs_reference(pkg, device, dist):
    g = load("rollouts.pt")[dist]
    A = g["action"].shape[-1]
    d = dims(A)
    pd = pc.to_dev(g["params"], device, grad=True)
    e = g["embed"].clone().to(device).requires_grad_(True)
    outs = pkg.kernels.observe(e, g["action"].to(device), g["is_first"].to(device), g["u_prior"].to(device),
                               g["u_post"].to(device), None, None, pc.kdims(d), pc.rssm_list(pkg, pd))
    post_stoch, post_logit, prior_stoch, prior_logit, deter = outs[:5]
    assert torch.equal(post_stoch.detach().cpu(), g["post"]["stoch"])
    assert torch.equal(prior_stoch.detach().cpu(), g["prior"]["stoch"])
    assert pc.rel(deter, g["post"]["deter"]) < TOL
    assert pc.rel(post_logit, g["post"]["logit"]) < TOL and pc.rel(prior_logit, g["prior"]["logit"]) < TOL
    kl = pkg.kernels.kl_balance(post_logit, prior_logit, 1.0, 0.5, 0.1, 0.01)
    for a, b in zip(kl[:4], g["kl"]):
        assert pc.rel(a, b) < TOL
    B, T = g["is_first"].shape
    feat = torch.cat([post_stoch.reshape(B, T, -1), deter], -1)
    ((feat * g["w"].to(device)).sum() + 20 * kl[0].mean()).backward()
    assert pc.rel(e.grad, g["d_embed"]) < TOL
    for k in g["grads"]:
        assert pc.rel(pd[k].grad, g["grads"][k]) < TOL, k
    im = g["imagine"]
    layers = 2
    pa = pc.to_dev(im["actor"], device, grad=True)
    spec = pkg.kernels.ActorSpec(layers, 32, dist, 0.1, 1.0, 0.01)
    start = {k: v.reshape([-1] + list(v.shape[2:])) for k, v in g["post"].items()}
    p0 = pc.to_dev(g["params"], device)
    feat, logit, action, idx = pkg.kernels.imagine(
        start["stoch"].argmax(-1).to(torch.int32).to(device), start["deter"].to(device),
        im["act_noise"].to(device), im["u_state"].to(device), None, 4, pc.kdims(d), spec,
        pc.rssm_list(pkg, p0), pc.actor_list(pa, layers, dist), start_logit=start["logit"].to(device))
    assert torch.equal(idx.cpu().long(), im["states"]["stoch"].argmax(-1))
    assert pc.rel(feat, im["feats"]) < TOL and pc.rel(action, im["actions"]) < TOL
    assert pc.rel(logit, im["states"]["logit"]) < TOL
    (feat * im["w"].to(device)).sum().backward()
    for k in im["grads"]:
        assert pc.rel(pa[k].grad, im["grads"][k]) < TOL, k


@pytest.mark.parametrize("dist", ["normal", "onehot"])
def test_train_steps_vs_reference(pkg, device, dist):
    """WorldModel._train + ImagBehavior._train, two steps with Adam, from the reference's own
    state_dict: metrics and every updated parameter against what the reference produced."""
    g = load("train.pt")[dist]
    c = g["cfg"]
    cfgs = pkg.configs
    over = dict(device=device, num_actions=c["num_actions"], dyn_stoch=8, dyn_discrete=8, dyn_deter=48,
                dyn_hidden=32, units=32, imag_horizon=4, imag_gradient=c["imag_gradient"],
                encoder=dict(mlp_units=40, mlp_layers=2), decoder=dict(mlp_units=40, mlp_layers=2))
    if dist == "onehot":
        over["actor"] = dict(dist="onehot", std="none")
    cfg = cfgs.make_config("dmc_proprio", **over)
    wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
    beh = pkg.models.ImagBehavior(cfg, wm)
    wm.load_state_dict(g["wm"], strict=True)           # the reference's checkpoint loads as is
    beh.actor.load_state_dict(g["actor"], strict=True)
    beh.value.load_state_dict(g["value"], strict=True)
    beh._slow_value.load_state_dict(g["value"], strict=True)
    reward_fn = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
    for i, st in enumerate(g["steps"]):
        n = {k: v.to(device) for k, v in st["noise"].items()}
        post, _, m1 = wm._train(st["data"], noise=(n["u_prior"], n["u_post"]))
        _, _, _, _, m2 = beh._train(post, reward_fn, noise=(n["act_noise"], n["u_state"]))
        m = {**m1, **m2}
        for k in ("model_loss", "model_grad_norm", "actor_loss", "actor_grad_norm", "value_loss",
                  "value_grad_norm", "kl", "reward_loss", "cont_loss", "dyn_loss", "rep_loss", "post_ent",
                  "prior_ent", "actor_entropy", "EMA_005", "EMA_095"):
            assert pc.rel(torch.as_tensor(m[k]), st["metrics"][k]) < TOL, (i, k)
        assert torch.equal(post["stoch"].cpu(), st["post"]["stoch"])
        # Adam turns a gradient into ~lr*sign(g) on the first steps, so elements whose gradient is
        # at fp32-noise level may move by up to lr in either direction: bound the worst case by
        # 2*lr per step and require that (almost) every element agrees to 5e-6.
        for mod, ref, lr in ((wm, st["wm_after"], 1e-4), (beh.actor, st["actor_after"], 3e-5),
                             (beh.value, st["value_after"], 3e-5), (beh._slow_value, st["slow_after"], 3e-5)):
            sd = mod.state_dict()
            for k in ref:
                diff = (sd[k].cpu() - ref[k]).abs()
                assert float(diff.max()) <= 2 * lr * (i + 1) + 1e-6, (i, k, float(diff.max()))
                assert float((diff > 5e-6).float().mean()) < 2e-3, (i, k)
