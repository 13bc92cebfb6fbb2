"""-m gpu: the whole-train-step CUDA graph replays the same training as the eager calls."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _agent(pkg, device, seed=0, full=False):
    cfgs = pkg.configs
    torch.manual_seed(seed)
    if full:        # configs.yaml defaults: the sizes the benchmark runs
        cfg = cfgs.make_config("dmc_proprio", device=device)
    else:
        cfg = cfgs.make_config("dmc_proprio", device=device, dyn_stoch=8, dyn_discrete=8, dyn_deter=64,
                               dyn_hidden=64, units=64, imag_horizon=5,
                               encoder=dict(mlp_units=64, mlp_layers=2), decoder=dict(mlp_units=64, mlp_layers=2))
    wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
    beh = pkg.models.ImagBehavior(cfg, wm)
    return cfg, wm, beh


def _batch(rs, B, T, A):
    host = {k: rs.randn(B, T, n).astype(np.float32) for k, n in (("orientations", 14), ("height", 1), ("velocity", 9))}
    host["action"] = rs.uniform(-1, 1, size=(B, T, A)).astype(np.float32)
    host["reward"] = rs.randn(B, T).astype(np.float32)
    host["discount"] = np.ones((B, T), np.float32)
    host["is_terminal"] = np.zeros((B, T), np.float32)
    host["is_first"] = np.zeros((B, T), np.float32)
    host["is_first"][:, 0] = 1
    host["is_first"][1, T // 2] = 1
    return host


def test_graph_step_matches_eager(pkg, device):
    B, T = 6, 10
    cfg, wm_a, beh_a = _agent(pkg, device)
    _, wm_b, beh_b = _agent(pkg, device)
    wm_b.load_state_dict(wm_a.state_dict())
    beh_b.load_state_dict(beh_a.state_dict())
    A, S, C, H = cfg.num_actions, cfg.dyn_stoch, cfg.dyn_discrete, cfg.imag_horizon
    N = B * T
    graph = pkg.graphs.TrainStepGraph(wm_b, beh_b, warmup=2)
    reward = lambda f, s, a: wm_a.heads["reward"](wm_a.dynamics.get_feat(s)).mode()
    rs = np.random.RandomState(3)
    gen = torch.Generator().manual_seed(5)
    for step in range(5):
        data = _batch(rs, B, T, A)
        noise = dict(u_prior=torch.rand(T, B, S, C, generator=gen).clamp_(1e-30, 1.0),
                     u_post=torch.rand(T, B, S, C, generator=gen).clamp_(1e-30, 1.0),
                     act_noise=torch.randn(H, N, A, generator=gen),
                     u_state=torch.rand(H, N, S, C, generator=gen).clamp_(1e-30, 1.0))
        nd = {k: v.to(device) for k, v in noise.items()}
        post, _, m1 = wm_a._train(data, noise=(nd["u_prior"], nd["u_post"]))
        _, _, _, _, m2 = beh_a._train(post, reward, noise=(nd["act_noise"], nd["u_state"]))
        out = graph(data, noise=noise)
        assert graph.captured == (step >= 2)
        assert torch.equal(out["post"]["stoch"], post["stoch"]), step
        for k in ("model_loss", "model_grad_norm", "kl"):
            assert abs(float(out["wm_metrics"][k]) - float(m1[k])) <= 2e-5 * abs(float(m1[k])) + 1e-7, (step, k)
        for k in ("actor_loss", "value_loss", "actor_grad_norm", "value_grad_norm"):
            assert abs(float(out["beh_metrics"][k]) - float(m2[k])) <= 2e-5 * abs(float(m2[k])) + 1e-7, (step, k)
    for (k, a), b in zip(wm_a.state_dict().items(), wm_b.state_dict().values()):
        assert float((a - b).abs().max()) <= 2e-6, k
    for (k, a), b in zip(beh_a.state_dict().items(), beh_b.state_dict().values()):
        assert float((a.float() - b.float()).abs().max()) <= 2e-6, k


def test_eager_forwards_between_replays_see_updated_weights(pkg, device):
    """The acting path runs eager forwards (encoder, actor) between graph replays
    (INTEGRATION.md flow: TrainStepGraph inside Dreamer._train, Dreamer._policy between steps).
    Their cached tf32 weight planes must follow the weights the replays update: compare against
    a twin agent trained by eager calls (whose optimizer invalidates the cache itself)."""
    B, T = 6, 10
    cfg, wm_a, beh_a = _agent(pkg, device)
    _, wm_b, beh_b = _agent(pkg, device)
    wm_b.load_state_dict(wm_a.state_dict())
    beh_b.load_state_dict(beh_a.state_dict())
    A, S, C, H = cfg.num_actions, cfg.dyn_stoch, cfg.dyn_discrete, cfg.imag_horizon
    N = B * T
    graph = pkg.graphs.TrainStepGraph(wm_b, beh_b, warmup=2)
    reward = lambda f, s, a: wm_a.heads["reward"](wm_a.dynamics.get_feat(s)).mode()
    rs = np.random.RandomState(3)
    gen = torch.Generator().manual_seed(5)
    feat = torch.randn(64, S * C + cfg.dyn_deter, generator=gen).to(device)
    obs = {k: torch.randn(64, n, generator=gen).to(device)
           for k, n in (("orientations", 14), ("height", 1), ("velocity", 9))}

    def policy_outputs(wm, beh):
        with torch.no_grad():
            dist = beh.actor(feat)                      # eager MLP forward: caches weight planes
            return dist.mean.clone(), dist.std.clone(), wm.encoder(obs).clone()

    for step in range(6):
        data = _batch(rs, B, T, A)
        noise = dict(u_prior=torch.rand(T, B, S, C, generator=gen).clamp_(1e-30, 1.0),
                     u_post=torch.rand(T, B, S, C, generator=gen).clamp_(1e-30, 1.0),
                     act_noise=torch.randn(H, N, A, generator=gen),
                     u_state=torch.rand(H, N, S, C, generator=gen).clamp_(1e-30, 1.0))
        nd = {k: v.to(device) for k, v in noise.items()}
        post, _, _ = wm_a._train(data, noise=(nd["u_prior"], nd["u_post"]))
        beh_a._train(post, reward, noise=(nd["act_noise"], nd["u_state"]))
        # eager forwards on the graph agent BEFORE the step (also right before the capture call)
        policy_outputs(wm_b, beh_b)
        graph(data, noise=noise)
        ref = policy_outputs(wm_a, beh_a)
        got = policy_outputs(wm_b, beh_b)
        for name, a, b in zip(("mean", "std", "embed"), got, ref):
            assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-6, (step, name)
    assert graph.captured
    # and the weights did move, so a stale cache would have been visible
    with torch.no_grad():
        fresh = _agent(pkg, device)[2].actor(feat).mean
    assert float((fresh - got[0]).abs().max()) > 1e-5


@pytest.mark.parametrize("full,B,T,K", [(False, 6, 10, 6), (True, 16, 64, 4)])
def test_pipelined_schedule_matches_sequential(pkg, device, full, B, T, K):
    """TrainStepGraph(pipeline=True): the behaviour update of batch t runs concurrently with the
    world-model forward + backward of batch t+1 (it only reads the world model; Adam follows the
    join).  Every update sees the inputs it sees in the sequential schedule, so after the same
    batches and noise (+ flush) all parameters equal those of an agent trained by sequential eager
    calls; the noise of call r drives the behaviour update of batch r-1."""
    cfg, wm_a, beh_a = _agent(pkg, device, full=full)
    _, wm_b, beh_b = _agent(pkg, device, full=full)
    wm_b.load_state_dict(wm_a.state_dict())
    beh_b.load_state_dict(beh_a.state_dict())
    A, S, C, H = cfg.num_actions, cfg.dyn_stoch, cfg.dyn_discrete, cfg.imag_horizon
    N = B * T
    graph = pkg.graphs.TrainStepGraph(wm_b, beh_b, warmup=2, pipeline=True)
    reward = lambda f, s, a: wm_a.heads["reward"](wm_a.dynamics.get_feat(s)).mode()
    rs = np.random.RandomState(3)
    gen = torch.Generator().manual_seed(5)
    prev_beh = dict(act_noise=torch.zeros(H, N, A), u_state=torch.ones(H, N, S, C))
    m2_prev = None
    for step in range(K):
        data = _batch(rs, B, T, A)
        wm_noise = dict(u_prior=torch.rand(T, B, S, C, generator=gen).clamp_(1e-30, 1.0),
                        u_post=torch.rand(T, B, S, C, generator=gen).clamp_(1e-30, 1.0))
        beh_noise = dict(act_noise=torch.randn(H, N, A, generator=gen),
                         u_state=torch.rand(H, N, S, C, generator=gen).clamp_(1e-30, 1.0))
        post, _, m1 = wm_a._train(data, noise=(wm_noise["u_prior"].to(device), wm_noise["u_post"].to(device)))
        _, _, _, _, m2 = beh_a._train(post, reward, noise=(beh_noise["act_noise"].to(device),
                                                           beh_noise["u_state"].to(device)))
        out = graph(data, noise=dict(wm_noise, **prev_beh))
        assert graph.captured == (step >= 2)
        assert torch.equal(out["post"]["stoch"], post["stoch"]), step
        for k in ("model_loss", "model_grad_norm", "kl"):
            assert abs(float(out["wm_metrics"][k]) - float(m1[k])) <= 2e-5 * abs(float(m1[k])) + 1e-7, (step, k)
        if step == 0:
            assert not out["beh_metrics"]
        else:                                   # the behaviour metrics of the PREVIOUS batch
            for k in ("actor_loss", "value_loss", "actor_grad_norm", "value_grad_norm"):
                assert abs(float(out["beh_metrics"][k]) - float(m2_prev[k])) <= \
                    2e-5 * abs(float(m2_prev[k])) + 1e-7, (step, k)
        prev_beh, m2_prev = beh_noise, m2
    last = graph.flush(noise=prev_beh)
    assert last is not None
    for k in ("actor_loss", "value_loss"):
        assert abs(float(last[4][k]) - float(m2_prev[k])) <= 2e-5 * abs(float(m2_prev[k])) + 1e-7, k
    # small config: everything reproduces to rounding; full size: the split-K parameter-gradient
    # products add with fp32 atomics (summation order not fixed), and Adam turns a gradient at noise
    # level into +-lr per step -- the same bound as tests/test_gpu_reference.py, per step taken
    tol = 2e-6 if not full else K * 2 * 1e-4 + 1e-6
    frac_tol = 2e-6
    for sd_a, sd_b in ((wm_a.state_dict(), wm_b.state_dict()), (beh_a.state_dict(), beh_b.state_dict())):
        for (k, a), b in zip(sd_a.items(), sd_b.values()):
            d = (a.float() - b.float()).abs()
            assert float(d.max()) <= tol, k
            if full and d.numel() > 1000:      # ... and all but a vanishing fraction agree closely
                assert float((d > frac_tol + 1e-4 * b.float().abs()).float().mean()) < 0.02, k


def test_lazy_metrics_are_the_previous_calls(pkg, device):
    """TrainStepGraph(lazy_metrics=True): host metrics come back one call late (None first, drain()
    last) and equal the metrics an eager-reading twin returns for the same call."""
    B, T = 6, 10
    cfg, wm_a, beh_a = _agent(pkg, device)
    _, wm_b, beh_b = _agent(pkg, device)
    wm_b.load_state_dict(wm_a.state_dict())
    beh_b.load_state_dict(beh_a.state_dict())
    A, S, C, H = cfg.num_actions, cfg.dyn_stoch, cfg.dyn_discrete, cfg.imag_horizon
    N = B * T
    g_a = pkg.graphs.TrainStepGraph(wm_a, beh_a, warmup=2)
    g_b = pkg.graphs.TrainStepGraph(wm_b, beh_b, warmup=2, lazy_metrics=True)
    rs = np.random.RandomState(3)
    gen = torch.Generator().manual_seed(5)
    prev = None
    for step in range(5):
        data = _batch(rs, B, T, A)
        noise = dict(u_prior=torch.rand(T, B, S, C, generator=gen).clamp_(1e-30, 1.0),
                     u_post=torch.rand(T, B, S, C, generator=gen).clamp_(1e-30, 1.0),
                     act_noise=torch.randn(H, N, A, generator=gen),
                     u_state=torch.rand(H, N, S, C, generator=gen).clamp_(1e-30, 1.0))
        ref = g_a(data, noise=noise)
        out = g_b(data, noise=noise)
        if step == 0:
            assert out["wm_metrics"] is None and out["beh_metrics"] is None
        else:
            for grp in ("wm_metrics", "beh_metrics"):
                assert set(out[grp]) == set(prev[grp])
                for k, v in prev[grp].items():
                    np.testing.assert_allclose(np.asarray(out[grp][k], dtype=np.float64),
                                               np.asarray(v, dtype=np.float64), rtol=2e-5, atol=1e-7, err_msg=k)
        prev = ref
    last = g_b.drain()
    assert last is not None and g_b.drain() is None
    for gi, grp in enumerate(("wm_metrics", "beh_metrics")):
        for k, v in prev[grp].items():
            np.testing.assert_allclose(np.asarray(last[gi][k], dtype=np.float64), np.asarray(v, dtype=np.float64),
                                       rtol=2e-5, atol=1e-7, err_msg=k)
