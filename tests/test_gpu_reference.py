"""-m gpu: the CUDA path against THE REFERENCE ITSELF at the real BASELINE configs.

``oracle/_ref`` holds the unmodified reference modules (staged by oracle/build_ref.py, carried to
the GPU box by the snapshot).  One full ``Dreamer._train`` step -- ``WorldModel._train`` followed by
``ImagBehavior._train`` on a 16x64 replay batch with H=15 imagination, Adam updates included --
is run by the reference on the host CPU (supplied noise through ref_harness.NoiseTape) and by the
product on the GPU from the same state_dict, batch and noise, for dmc_proprio (configs[0]),
dmc_vision (configs[1]: conv encoder / decoder) and atari100k (configs[2]: one-hot actor,
reinforce gradient).  Posterior / imagined class indices must be bit-exact, every metric within
1e-4 relative, every updated parameter within the Adam sign-noise bound.

A free-running 64-step rollout draws ~2 M categorical samples; two fp32 implementations with
different summation orders can legitimately disagree on a draw whose two best scores tie to
~1e-6 (after which that sequence diverges).  The test therefore tries up to three noise seeds and
passes on the first one without such a tie; the seeds tried are reported on failure.
"""
import numpy as np
import pytest
import torch

import parity_cases as pc
import ref_harness as H

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _uniforms(g, *shape):
    return torch.rand(*shape, generator=g).clamp_(1e-30, 1.0)


def _perturb(modules, g, scale=0.02):
    with torch.no_grad():
        for m in modules:
            for p in m.parameters():
                p.add_(scale * torch.randn(p.shape, generator=g))


def _sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _reference_step(suite, seed, B=16, T=64):
    cfg = H.reference_config((suite,), num_actions=H.SUITE_ACTIONS[suite], device="cpu")
    wm, beh = H.build_agent(cfg, H.suite_shapes(suite), 0)
    g = torch.Generator().manual_seed(11)
    _perturb([wm, beh.actor, beh.value], g)
    beh._slow_value.load_state_dict(beh.value.state_dict())
    before = dict(wm=_sd(wm), actor=_sd(beh.actor), value=_sd(beh.value))
    A, S, C, Hh, N = cfg.num_actions, cfg.dyn_stoch, cfg.dyn_discrete, cfg.imag_horizon, B * T
    onehot = cfg.actor["dist"] == "onehot"
    data = H.suite_batch(suite, B, T, seed, resets=((1, 3), (2, T // 2)))
    g = torch.Generator().manual_seed(seed + 20)
    noise = dict(u_prior=_uniforms(g, T, B, S, C), u_post=_uniforms(g, T, B, S, C),
                 act_noise=_uniforms(g, Hh, N, A) if onehot else torch.randn(Hh, N, A, generator=g),
                 u_state=_uniforms(g, Hh, N, S, C))
    tape = []
    for t in range(T):
        tape += [("u", noise["u_prior"][t]), ("u", noise["u_post"][t])]
    for k in range(Hh):
        tape += [("u" if onehot else "n", noise["act_noise"][k]), ("u", noise["u_state"][k])]
    reward_fn = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
    feed = {k: v.copy() for k, v in data.items()}
    with H.NoiseTape(tape) as tp, H.quiet():
        post, _, m1 = wm._train(feed)
        feat, state, action, _, m2 = beh._train(post, reward_fn)
    assert tp.pos == len(tape)
    metrics = {k: torch.tensor(np.array(v, copy=True)) for k, v in {**m1, **m2}.items()}
    after = dict(wm=_sd(wm), actor=_sd(beh.actor), value=_sd(beh.value), slow=_sd(beh._slow_value))
    return dict(cfg=cfg, data=data, noise=noise, before=before, after=after, metrics=metrics,
                post={k: v.detach().clone() for k, v in post.items()},
                imag_idx=state["stoch"].argmax(-1), imag_action=action.detach().clone())


def _product_step(pkg, device, suite, ref):
    """WorldModel._train from the reference's initial state_dict; then -- so that the behaviour
    half is compared on identical weights rather than through the Adam sign noise of the WM update
    (an element whose gradient is at fp32-noise level moves by +-lr) -- the reference's updated WM
    weights and posterior are loaded before ImagBehavior._train."""
    cfgs = pkg.configs
    cfg = cfgs.make_config(suite, device=device)
    shapes = cfgs.PROPRIO_SHAPES if suite == "dmc_proprio" else cfgs.VISION_SHAPES
    wm = pkg.models.WorldModel(cfgs.ObsSpace(shapes), None, 0, cfg)
    beh = pkg.models.ImagBehavior(cfg, wm)
    wm.load_state_dict(ref["before"]["wm"], strict=True)          # the reference's own state_dict
    beh.actor.load_state_dict(ref["before"]["actor"], strict=True)
    beh.value.load_state_dict(ref["before"]["value"], strict=True)
    beh._slow_value.load_state_dict(ref["before"]["value"], strict=True)
    reward_fn = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
    n = {k: v.to(device) for k, v in ref["noise"].items()}
    data = {k: v.copy() for k, v in ref["data"].items()}
    if suite == "dmc_proprio":
        data.pop("image")
    post, _, m1 = wm._train(data, noise=(n["u_prior"], n["u_post"]))
    wm_after = {k: v.detach().clone() for k, v in wm.state_dict().items()}
    wm.load_state_dict(ref["after"]["wm"], strict=True)
    start = {k: v.to(device) for k, v in ref["post"].items()}
    _, state, action, _, m2 = beh._train(start, reward_fn, noise=(n["act_noise"], n["u_state"]))
    return wm_after, beh, post, state, action, {**m1, **m2}


METRICS = ("model_loss", "model_grad_norm", "actor_loss", "actor_grad_norm", "value_loss",
           "value_grad_norm", "kl", "reward_loss", "cont_loss", "dyn_loss", "rep_loss", "post_ent",
           "prior_ent", "actor_entropy", "EMA_005", "EMA_095", "value_mean", "target_mean",
           "imag_reward_mean")


@pytest.mark.parametrize("suite", ["dmc_proprio", "dmc_vision", "atari100k"])
def test_whole_train_step_vs_reference_itself(pkg, device, suite):
    if not H.available():
        pytest.skip("oracle/_ref not staged (run oracle/build_ref.py in the build container)")
    torch.set_num_threads(max(1, (torch.get_num_threads())))
    tried = []
    for seed in (0, 1, 2):
        ref = _reference_step(suite, seed)
        wm_after, beh, post, state, action, m = _product_step(pkg, device, suite, ref)
        flips = int((post["stoch"].argmax(-1).cpu() != ref["post"]["stoch"].argmax(-1)).sum())
        flips_im = int((state["stoch"].argmax(-1).cpu() != ref["imag_idx"]).sum())
        tried.append((seed, flips, flips_im))
        if flips or flips_im:
            continue
        assert torch.equal(post["stoch"].cpu(), ref["post"]["stoch"])
        assert pc.rel(post["deter"], ref["post"]["deter"]) < TOL
        assert pc.rel(post["logit"], ref["post"]["logit"]) < TOL
        assert pc.rel(action, ref["imag_action"]) < TOL
        bad = {}
        for k in METRICS + (("image_loss",) if suite != "dmc_proprio" else ()):
            if k in ref["metrics"]:
                e = pc.rel(torch.as_tensor(m[k]), ref["metrics"][k])
                if not e < TOL:
                    bad[k] = e
        assert not bad, (suite, seed, bad)
        # Adam turns a gradient into ~lr*sign(g) on the first step: elements whose gradient is at
        # fp32-noise level may move by up to lr in either direction
        for sd, key, lr in ((wm_after, "wm", 1e-4), (beh.actor.state_dict(), "actor", 3e-5),
                            (beh.value.state_dict(), "value", 3e-5),
                            (beh._slow_value.state_dict(), "slow", 3e-5)):
            for k, r in ref["after"][key].items():
                diff = (sd[k].cpu() - r).abs()
                assert float(diff.max()) <= 2 * lr + 1e-6, (suite, key, k, float(diff.max()))
                assert float((diff > 5e-6).float().mean()) < 2e-3, (suite, key, k)
        return
    pytest.fail(f"{suite}: categorical near-tie flips with every noise seed (seed, post, imagined): {tried}")


class _Holder:
    """Stand-in for the reference's Dreamer agent object: the checkpoint helpers walk __dict__."""


def test_reference_checkpoint_helpers_roundtrip(pkg, device):
    """dreamer.py:502-506 / 563-567: ``tools.recursively_collect_optim_state_dict(agent)`` and
    ``recursively_load_optim_state_dict`` -- the reference's OWN helpers -- must find and restore
    the fused flat Adam state through the ``_opt`` attribute, and a reference-format
    ``optims_state_dict`` (torch.optim.Adam layout) must load."""
    if not H.available():
        pytest.skip("oracle/_ref not staged")
    rtools = H.load_reference("cpu")[0]
    d = pc.synth.dims_of("tiny")
    P, Pa, Pv = pc.synth.agent_params("tiny", enc_units=d.embed)
    kw = dict(device_metrics=False, encoder=dict(mlp_units=d.embed), decoder=dict(mlp_units=d.embed),
              imag_horizon=5)

    def make():
        cfg, wm, beh = pc.build_product_agent(pkg, device, "tiny", P, Pa, Pv, **kw)
        a = _Holder()
        a._wm, a._task_behavior = wm, beh
        return a

    a = make()
    data = pc.synth.replay_batch(d, 4, 6, resets=((1, 3),))
    reward_fn = lambda f, s, x: a._wm.heads["reward"](a._wm.dynamics.get_feat(s)).mode()
    for _ in range(2):
        post, _, _ = a._wm._train(data)
        a._task_behavior._train(post, reward_fn)
    sd = rtools.recursively_collect_optim_state_dict(a)
    paths = {"_wm._model_opt._opt", "_task_behavior._actor_opt._opt", "_task_behavior._value_opt._opt"}
    assert paths <= set(sd), sorted(sd)
    for p in paths:
        st = sd[p]["state"]
        assert len(st) > 0 and float(st[0]["step"]) == 2.0
        assert {"step", "exp_avg", "exp_avg_sq"} <= set(st[0])
    b = make()
    rtools.recursively_load_optim_state_dict(b, {k: sd[k] for k in paths})
    for name in ("_model_opt",):
        oa, ob = a._wm._model_opt, b._wm._model_opt
        assert torch.equal(oa._fm, ob._fm) and torch.equal(oa._fv, ob._fv)
        assert float(ob._step) == 2.0
    for name in ("_actor_opt", "_value_opt"):
        oa, ob = getattr(a._task_behavior, name), getattr(b._task_behavior, name)
        assert torch.equal(oa._fm, ob._fm) and torch.equal(oa._fv, ob._fv) and float(ob._step) == 2.0
    # a state_dict written by the reference's stock torch.optim.Adam (CPU tensors, python-float lr)
    ps = [torch.nn.Parameter(p.detach().cpu().clone()) for p in b._task_behavior.actor.parameters()]
    stock = torch.optim.Adam(ps, lr=3e-5, eps=1e-5)
    for p in ps:
        p.grad = torch.randn_like(p)
    stock.step()
    b._task_behavior._actor_opt._opt.load_state_dict(stock.state_dict())
    ms = b._task_behavior._actor_opt._views(b._task_behavior._actor_opt._fm)
    for i, p in enumerate(ps):
        assert torch.allclose(ms[i].cpu(), stock.state[p]["exp_avg"])
    assert float(b._task_behavior._actor_opt._step) == 1.0
