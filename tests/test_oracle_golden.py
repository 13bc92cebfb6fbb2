"""CPU (-m "not gpu"): the oracle (oracle/dv3_oracle.py, oracle/train_step.py) against the golden
fixtures written by tests/golden/make_golden.py from the live reference.  Indices bit-exact,
floats to 2e-5 relative (both sides are fp32 torch-CPU; only op order differs)."""
import os

import pytest
import torch

import dv3_oracle as O
import train_step as TS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5


def rel(a, b):
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(b).detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def tiny_dims(A, embed=40):
    return O.RSSMDims(stoch=8, classes=8, deter=48, hidden=32, actions=A, embed=embed)


def test_lambda_return_golden():
    g = load("ops.pt")["lambda_return"]
    leaves = [g[k].clone().requires_grad_(True) for k in ("reward", "value", "pcont", "bootstrap")]
    ret = O.lambda_return(*leaves, 0.95)
    assert torch.equal(ret.detach(), g["ret"])          # same op order -> bit-identical
    grads = torch.autograd.grad((ret * g["w"]).sum(), leaves)
    for a, b in zip(grads, g["grads"]):
        assert rel(a, b) < TOL


def test_twohot_golden():
    g = load("ops.pt")["twohot"]
    logits = g["logits"].clone().requires_grad_(True)
    lp, mean = O.twohot_logprob(logits, g["x"]), O.twohot_mean(logits)
    assert rel(lp, g["log_prob"]) < TOL and rel(mean, g["mean"]) < TOL
    assert rel(torch.autograd.grad((lp * g["w1"]).sum(), logits, retain_graph=True)[0], g["d_log_prob"]) < TOL
    assert rel(torch.autograd.grad((mean * g["w2"]).sum(), logits)[0], g["d_mean"]) < TOL


@pytest.mark.parametrize("dist", ["normal", "onehot"])
def test_observe_and_imagine_golden(dist):
    g = load("rollouts.pt")[dist]
    A = g["action"].shape[-1]
    d = tiny_dims(A)
    p = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    e = g["embed"].clone().requires_grad_(True)
    post, prior = O.observe(p, e, g["action"], g["is_first"], g["u_prior"], g["u_post"], d)
    for name, mine, ref in (("post", post, g["post"]), ("prior", prior, g["prior"])):
        assert torch.equal(mine["stoch"].argmax(-1), ref["stoch"].argmax(-1)), name
        assert torch.equal(mine["stoch"].detach(), ref["stoch"]), name     # forward value exactly one-hot
        assert rel(mine["deter"], ref["deter"]) < TOL and rel(mine["logit"], ref["logit"]) < TOL
    kl = O.kl_balance(post["logit"], prior["logit"], 1.0, 0.5, 0.1, 0.01)
    for a, b in zip(kl, g["kl"]):
        assert rel(a, b) < TOL
    ((O.get_feat(post) * g["w"]).sum() + 20 * kl[0].mean()).backward()
    assert rel(e.grad, g["d_embed"]) < TOL
    for k in p:
        assert rel(p[k].grad, g["grads"][k]) < TOL, k
    im = g["imagine"]
    pa = {k: v.clone().requires_grad_(True) for k, v in im["actor"].items()}
    start = {k: v.reshape([-1] + list(v.shape[2:])) for k, v in g["post"].items()}
    pr = {k: v.detach() for k, v in g["params"].items()}
    feats, states, actions = O.imagine(pr, pa, start, 4, im["act_noise"], im["u_state"], d, 2, dist)
    assert torch.equal(states["stoch"].argmax(-1), im["states"]["stoch"].argmax(-1))
    assert rel(feats, im["feats"]) < TOL and rel(actions, im["actions"]) < TOL
    for k in ("deter", "logit"):
        assert rel(states[k], im["states"][k]) < TOL
    (O.get_feat(states) * im["w"]).sum().backward()
    for k in pa:
        assert rel(pa[k].grad, im["grads"][k]) < TOL, k


@pytest.mark.parametrize("dist", ["normal", "onehot"])
def test_full_train_steps_golden(dist):
    """Two consecutive Dreamer._train steps incl. Adam: losses, grad norms, every updated
    parameter, the slow critic and the RewardEMA state."""
    g = load("train.pt")[dist]
    c = g["cfg"]
    d = tiny_dims(c["num_actions"])
    cfg = TS.make_cfg(dyn_stoch=8, dyn_discrete=8, units=32, enc_layers=2, enc_units=40, dec_layers=2,
                      dec_units=40, imag_horizon=4, actor_dist=c["actor_dist"],
                      imag_gradient=c["imag_gradient"])
    agent = TS.Agent(g["wm"], g["actor"], g["value"], cfg, d)
    for i, st in enumerate(g["steps"]):
        out = agent.train_step(st["data"], st["noise"])
        m = st["metrics"]
        for k in ("model_loss", "model_grad_norm", "actor_loss", "actor_grad_norm", "value_loss",
                  "value_grad_norm", "kl", "reward_loss", "cont_loss"):
            assert rel(out[k], m[k]) < 5e-5, (i, k, float(torch.as_tensor(out[k]).mean()), float(m[k].mean()))
        assert torch.equal(out["post"]["stoch"].argmax(-1), st["post"]["stoch"].argmax(-1))
        # Adam normalises the step, so tiny gradient differences are amplified where |g| ~ eps:
        # compare the update relative to the largest update of the tensor's optimizer group.
        for name, mine, ref in (("wm", agent.P_wm, st["wm_after"]), ("actor", agent.P_actor, st["actor_after"]),
                                ("value", agent.P_value, st["value_after"])):
            for k in ref:
                assert float((mine[k].detach() - ref[k]).abs().max()) < 2e-6, (i, name, k)
        for k in st["slow_after"]:
            assert float((agent.P_slow[k] - st["slow_after"][k]).abs().max()) < 2e-6
        assert rel(agent.ema_vals, st["ema_after"]) < 1e-5
