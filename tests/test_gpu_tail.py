"""-m gpu: the step-tail kernels (dv3_tail.cu) against the torch expressions they replace -- the
reference's own formulas (tools.py:22-27, 520-628, 949-958; models.py:11-26, 393-429, 620-681),
evaluated in fp64 / fp32 torch on the same inputs, values and gradients."""
import math

import pytest
import torch
import torch.nn.functional as F

import parity_cases as pc

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _g(seed=0):
    return torch.Generator().manual_seed(seed)


def test_symlog(pkg, device):
    x = (30 * torch.randn(1000, 24, generator=_g())).to(device)
    x[0, :3] = torch.tensor([0.0, -0.0, 1e9])
    ref = torch.sign(x) * torch.log(torch.abs(x) + 1.0)
    assert pc.rel(pkg.kernels.symlog(x), ref) < TOL


@pytest.mark.parametrize("use_symlog,shape", [(True, (16, 64, 14)), (True, (16, 64, 1)), (False, (4, 8, 16, 16, 3)),
                                              (True, (3, 5, 9))])
def test_sqerr_logprob(pkg, device, use_symlog, shape):
    g = _g(1)
    mode = torch.randn(*shape, generator=g).to(device).requires_grad_(True)
    value = (5 * torch.randn(*shape, generator=g)).to(device)
    if use_symlog:      # some exact hits: distance below tol is dropped (value and gradient)
        with torch.no_grad():
            mode.view(-1)[:7] = (torch.sign(value) * torch.log(value.abs() + 1)).view(-1)[:7]
    w = torch.randn(*shape[:2], generator=g).to(device)
    lp = pkg.kernels.sqerr_logprob(mode, value, use_symlog, 1e-8)
    (d,) = torch.autograd.grad((lp * w).sum(), mode)
    m2 = mode.detach().double().requires_grad_(True)
    v2 = value.double()
    if use_symlog:
        dist = (m2 - torch.sign(v2) * torch.log(v2.abs() + 1)) ** 2.0
        dist = torch.where(dist < 1e-8, torch.zeros_like(dist), dist)
        ref = -dist.flatten(2).sum(-1)
    else:
        ref = -((m2 - v2) ** 2).flatten(2).sum(-1)
    (dr,) = torch.autograd.grad((ref * w.double()).sum(), m2)
    assert lp.shape == shape[:2]
    assert pc.rel(lp, ref) < TOL and pc.rel(d, dr) < TOL


def test_bernoulli_logprob(pkg, device):
    g = _g(2)
    l = (6 * torch.randn(16, 64, 1, generator=g)).to(device)
    l[0, :4, 0] = torch.tensor([25.0, -25.0, 0.0, 19.99])
    l.requires_grad_(True)
    x = (torch.rand(16, 64, 1, generator=g) > 0.3).float().to(device)
    w = torch.randn(16, 64, generator=g).to(device)
    lp = pkg.tools.Bernoulli(l).log_prob(x)
    (d,) = torch.autograd.grad((lp * w).sum(), l)
    l2 = l.detach().clone().requires_grad_(True)
    ref = torch.sum(-F.softplus(l2) * (1 - x) - F.softplus(-l2) * x, -1)
    (dr,) = torch.autograd.grad((ref * w).sum(), l2)
    assert lp.shape == (16, 64)
    assert pc.rel(lp, ref) < TOL and pc.rel(d, dr) < TOL


def test_loss_mean(pkg, device):
    g = _g(3)
    terms = [torch.randn(16, 64, generator=g).to(device).requires_grad_(True) for _ in range(6)]
    scales = [-1.0, -1.0, -0.5, -1.0, -2.0, 1.0]
    out, neg = pkg.kernels.loss_mean(terms, scales)
    grads = torch.autograd.grad(out * 3.0, terms)
    ref = torch.mean(sum(s * t.double() for s, t in zip(scales, terms)))
    assert pc.rel(out, ref) < TOL
    for i, (t, gr) in enumerate(zip(terms, grads)):
        assert torch.equal(neg[i].reshape(16, 64), -t.detach())
        assert pc.rel(gr, torch.full_like(t, 3.0 * scales[i] / t.numel())) < TOL


def test_discount_weights(pkg, device):
    g = _g(4)
    H, N = 15, 300
    l = (3 * torch.randn(H, N, 1, generator=g)).to(device).requires_grad_(True)
    wgt = torch.randn(H, N, 1, generator=g).to(device)
    disc, w = pkg.kernels.discount_weights(l, 0.997)
    (d,) = torch.autograd.grad((disc * wgt).sum(), l)
    l2 = l.detach().clone().requires_grad_(True)
    dr_ = 0.997 * torch.sigmoid(l2)
    wr = torch.cumprod(torch.cat([torch.ones_like(dr_[:1]), dr_[:-1]], 0), 0).detach()
    (dr,) = torch.autograd.grad((dr_ * wgt).sum(), l2)
    assert not w.requires_grad
    assert pc.rel(disc, dr_) < TOL and pc.rel(w, wr) < TOL and pc.rel(d, dr) < TOL


@pytest.mark.parametrize("n", [1, 2, 60, 1000, 14336, 16384, 200000])
def test_reward_ema(pkg, device, n):
    g = _g(5 + n)
    ema = torch.tensor([-0.3, 0.9]).to(device)
    ema_ref = ema.clone()
    for it in range(2):
        x = (2 * torch.randn(n, generator=g)).to(device)
        os_ = pkg.kernels.reward_ema(x.reshape(-1, 1), ema, 0.01)
        q = torch.quantile(x, torch.tensor([0.05, 0.95], device=device))
        ema_ref[:] = 0.01 * q + (1 - 0.01) * ema_ref
        scale = torch.clip(ema_ref[1] - ema_ref[0], min=1.0)
        assert float((ema - ema_ref).abs().max()) <= 2e-7 * max(1.0, float(ema_ref.abs().max())), (n, it)
        assert float((os_[0] - ema_ref[0]).abs()) <= 2e-7 and float((os_[1] - scale).abs()) <= 2e-7


@pytest.mark.parametrize("mode", ["dynamics", "reinforce"])
@pytest.mark.parametrize("ema", [True, False])
def test_actor_loss(pkg, device, mode, ema):
    g = _g(6)
    H, N, c_ent = 15, 200, 3e-4
    dev = lambda t: t.to(device)
    target = dev(torch.randn(H - 1, N, 1, generator=g)).requires_grad_(True)
    base = dev(torch.randn(H - 1, N, 1, generator=g))
    w = dev(torch.rand(H, N, 1, generator=g))
    ent = dev(torch.randn(H, N, generator=g)).requires_grad_(True)
    logp = dev(torch.randn(H, N, generator=g)).requires_grad_(True) if mode == "reinforce" else None
    os_ = dev(torch.tensor([-0.2, 1.7])) if ema else None
    loss, normed = pkg.kernels.actor_loss(target, base, w, ent, logp, os_, c_ent, mode)
    leaves = [target, ent] + ([logp] if logp is not None else [])
    grads = torch.autograd.grad(loss * 2.0, leaves, allow_unused=True)
    t2, e2 = target.detach().double().requires_grad_(True), ent.detach().double().requires_grad_(True)
    l2 = logp.detach().double().requires_grad_(True) if logp is not None else None
    off, sc = (os_.double() if ema else torch.tensor([0.0, 1.0], dtype=torch.float64, device=device))
    nt, nb = (t2 - off) / sc, (base.double() - off) / sc
    if mode == "dynamics":
        at = nt - nb
    else:
        at = l2[:-1][:, :, None] * (t2 - base.double()).detach()
    ref = torch.mean(-w.double()[:-1] * at - c_ent * e2[:-1, ..., None])
    rg = torch.autograd.grad(ref * 2.0, [t2, e2] + ([l2] if l2 is not None else []), allow_unused=True)
    assert pc.rel(loss, ref) < TOL and pc.rel(normed, nt) < TOL
    for a, b in zip(grads, rg):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0
        else:
            assert pc.rel(a, b) < TOL


def test_value_loss(pkg, device):
    g = _g(7)
    H, N = 15, 100
    a = torch.randn(H - 1, N, generator=g).to(device).requires_grad_(True)
    b = torch.randn(H - 1, N, generator=g).to(device).requires_grad_(True)
    w = torch.rand(H, N, 1, generator=g).to(device)
    for slow in (b, None):
        loss = pkg.kernels.value_loss(a, slow, w)
        vl = -a.double() - (slow.double() if slow is not None else 0.0)
        ref = torch.mean(w.double()[:-1] * vl[:, :, None])
        assert pc.rel(loss, ref) < TOL
        ga = torch.autograd.grad(loss, [a] + ([slow] if slow is not None else []))
        for gi in ga:
            assert pc.rel(gi, -w[:-1, :, 0] / ((H - 1) * N)) < TOL


@pytest.mark.parametrize("want_logp", [False, True])
def test_normal_policy(pkg, device, want_logp):
    g = _g(8)
    H, N, A = 5, 70, 6
    mr = torch.randn(H, N, A, generator=g).to(device).requires_grad_(True)
    sr = torch.randn(H, N, A, generator=g).to(device).requires_grad_(True)
    x = (torch.rand(H, N, A, generator=g) * 2 - 1).to(device).requires_grad_(True)
    w1, w2 = torch.randn(H, N, generator=g).to(device), torch.randn(H, N, generator=g).to(device)
    ent, lp = pkg.kernels.normal_policy(mr, sr, x, 0.1, 1.0, want_logp)
    obj = (ent * w1).sum() + ((lp * w2).sum() if want_logp else 0.0)
    grads = torch.autograd.grad(obj, [mr, sr, x], allow_unused=True)
    m2, s2, x2 = (t.detach().double().requires_grad_(True) for t in (mr, sr, x))
    mean, std = torch.tanh(m2), 0.9 * torch.sigmoid(s2 + 2.0) + 0.1
    ent_r = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)).sum(-1)
    lp_r = (-((x2 - mean) ** 2) / (2 * std ** 2) - torch.log(std) - math.log(math.sqrt(2 * math.pi))).sum(-1)
    obj_r = (ent_r * w1.double()).sum() + ((lp_r * w2.double()).sum() if want_logp else 0.0)
    rg = torch.autograd.grad(obj_r, [m2, s2, x2], allow_unused=True)
    assert pc.rel(ent, ent_r) < TOL
    if want_logp:
        assert pc.rel(lp, lp_r) < TOL
    for a, b in zip(grads, rg):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0
        elif not want_logp and a is None:
            assert float(b.abs().max()) == 0.0
        else:
            assert pc.rel(a, b) < 2e-5


def test_tensorstats_and_ema_mix(pkg, device):
    g = _g(9)
    x = (3 + 2 * torch.randn(14, 1024, 1, generator=g)).to(device)
    st = pkg.tools.tensorstats(x, "v")
    assert set(st) == {"v_mean", "v_std", "v_min", "v_max"}
    for k, r in (("v_mean", x.double().mean()), ("v_std", x.double().std()), ("v_min", x.min()), ("v_max", x.max())):
        assert abs(float(st[k]) - float(r)) <= 1e-5 * max(1.0, abs(float(r))), k
    a, b = torch.randn(1000, generator=g).to(device), torch.randn(1000, generator=g).to(device)
    ref = 0.02 * b + (1 - 0.02) * a
    pkg.kernels.ema_mix(a, b, 0.02)
    assert float((a - ref).abs().max()) <= 2e-7


def test_optimizer_gradient_sink_matches_torch_adam(pkg, device):
    """tools.Optimizer with the gradient sink (dW / LayerNorm / bias gradients accumulated straight
    into the flat buffer by the backward Functions) == clip_grad_norm_ + torch.optim.Adam on a twin
    whose gradients come from torch.autograd.grad; a module used twice accumulates; a parameter that
    receives no gradient is left untouched, as torch.optim.Adam leaves it."""
    torch.manual_seed(0)
    mk = lambda: pkg.networks.MLP(96, (7,), 3, 64, "SiLU", True, "normal", "learned", 0.1, 1.0,
                                  name="T").to(device)
    a, b = mk(), mk()
    b.load_state_dict(a.state_dict())
    unused_a = torch.nn.Parameter(torch.ones(5, device=device))
    unused_b = torch.nn.Parameter(torch.ones(5, device=device))
    opt = pkg.tools.Optimizer("t", list(a.parameters()) + [unused_a], 1e-3, 1e-8, 1.0, 0.0)
    ref = torch.optim.Adam(list(b.parameters()) + [unused_b], lr=1e-3, eps=1e-8)
    g = _g(10)
    for step in range(3):
        x1 = torch.randn(128, 96, generator=g).to(device)
        x2 = torch.randn(64, 96, generator=g).to(device)

        def loss_of(m):
            d1, d2 = m(x1), m(x2)                      # the module is used twice per step
            return (d1.mean ** 2).mean() + (d2.std * d2.mean).mean() * 3.0

        with pkg.tools.RequiresGrad(a):
            met = opt(loss_of(a))
        with pkg.tools.RequiresGrad(b):
            ref.zero_grad(set_to_none=True)
            loss_of(b).backward()
            norm = torch.nn.utils.clip_grad_norm_(list(b.parameters()), 1.0)
            ref.step()
            pkg.kernels.invalidate_weight_splits()
        assert abs(float(met["t_grad_norm"]) - float(norm)) <= 2e-5 * float(norm), step
        for (k, pa), pb in zip(a.named_parameters(), b.parameters()):
            assert float((pa - pb).abs().max()) <= 2e-6, (step, k)
        assert all(p.grad is None for p in a.parameters())
    assert torch.equal(unused_a.detach(), torch.ones(5, device=device))
    sd = opt.state_dict()
    assert float(sd["state"][0]["step"]) == 3.0
