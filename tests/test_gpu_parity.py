"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.
Indices bit-exact; floats within 1e-4 relative (BASELINE.json north_star)."""
import pytest
import torch

import parity_cases as pc

pytestmark = pytest.mark.gpu


def _assert(res, **kw):
    bad = pc.check(res, **kw)
    assert not bad, f"{bad}\nall: {res}"


def test_device_is_blackwell(pkg, device):
    assert pkg._lib.lib().dv3_device_arch() >= 100


@pytest.mark.parametrize("H,N", [(14, 1024), (1, 7), (15, 33)])
def test_lambda_return(pkg, device, H, N):
    res = pc.lambda_return_case(pkg, device, H=H, N=N)
    assert res["tuple_len"] == N and res["tuple_shape"] == (H, 1)
    _assert(res, skip=("tuple_len",))


def test_twohot(pkg, device):
    _assert(pc.twohot_case(pkg, device))


@pytest.mark.parametrize("scale", [0.5, 3.0, 12.0])
def test_kl_balance(pkg, device, scale):
    res = pc.kl_case(pkg, device, scale=scale)
    _assert(res, skip=("clipped_rows",))


@pytest.mark.parametrize("C,unimix", [(32, 0.01), (18, 0.01), (32, 0.0), (5, 0.01)])
def test_onehot_sample_and_straight_through(pkg, device, C, unimix):
    res = pc.sample_case(pkg, device, M=1024, S=32 if C == 32 else 1, C=C, unimix=unimix)
    assert res["onehot_ok"]
    _assert(res)


@pytest.mark.parametrize("stepwise", ["0", "1"])
@pytest.mark.parametrize("config,B,T", [("tiny", 3, 5), ("dmc_proprio", 16, 64), ("dmc_vision", 16, 16),
                                        ("dmc_proprio", 7, 9)])
def test_observe_fwd_bwd(pkg, device, config, B, T, stepwise, knob):
    """stepwise=0: persistent cooperative kernel for the forward recurrence (B <= 16);
    stepwise=1: the launch-per-phase path (what larger shapes use)."""
    knob("DV3_OBSERVE_STEPWISE", stepwise)
    before = pkg._lib.lib().dv3_launch_count()
    _assert(pc.observe_case(pkg, device, config=config, B=B, T=T, backward=False))
    launches = pkg._lib.lib().dv3_launch_count() - before
    if stepwise == "0":
        assert launches < 40, launches          # one launch for the whole recurrence
    else:
        assert launches > 7 * T
    _assert(pc.observe_case(pkg, device, config=config, B=B, T=T))


def test_observe_with_state(pkg, device):
    _assert(pc.observe_case(pkg, device, config="dmc_proprio", B=5, T=7, with_state=True))


def test_obs_step_teacher_forced(pkg, device):
    _assert(pc.obs_step_teacher_forced_case(pkg, device, config="dmc_proprio", B=16, T=6))


@pytest.mark.parametrize("config,N,H", [("tiny", 9, 4), ("tiny_onehot", 9, 4), ("dmc_proprio", 1024, 15),
                                        ("atari100k", 256, 15)])
def test_imagine_fwd_bwd(pkg, device, config, N, H):
    _assert(pc.imagine_case(pkg, device, config=config, N=N, H=H))


@pytest.mark.parametrize("config,N,H", [("dmc_proprio", 1024, 15), ("atari100k", 256, 15),
                                        ("dmc_proprio", 128, 3), ("wide_l1", 256, 4), ("wide_l3", 384, 5),
                                        ("dmc_proprio", 128, 1)])
def test_imagine_persistent_kernel(pkg, device, config, N, H, knob):
    """DV3_IMAGINE_PERSISTENT=1: the whole rollout as one cooperative kernel (groups of 16 CTAs per
    128 rows, every GEMM on tcgen05, LayerNorm / gates / draws in the epilogues) against the same
    oracle and to the same bar as the stepwise launches: indices exact, floats 1e-4."""
    knob("DV3_IMAGINE_PERSISTENT", "1")
    before = pkg._lib.lib().dv3_launch_count()
    res = pc.imagine_case(pkg, device, config=config, N=N, H=H, backward=False)
    launches = pkg._lib.lib().dv3_launch_count() - before
    assert launches < 20, launches              # state-0 set-up + one launch for all H steps
    _assert(res)
    _assert(pc.imagine_case(pkg, device, config=config, N=N, H=H))


def test_imagine_bwd_packed_and_strided_state_gradients(pkg, device, monkeypatch):
    """dv3_imagine_bwd takes the upstream state gradients packed ([H,N,S*C] and [H,N,D]) or as the
    two column ranges of one feature-gradient buffer (g_state_ld = S*C + D, what the autograd
    wrapper passes): same oracle, same bar, both forms."""
    monkeypatch.setenv("DV3_BWD_PACKED_G", "1")
    _assert(pc.imagine_case(pkg, device, config="dmc_proprio", N=256, H=6))
    monkeypatch.delenv("DV3_BWD_PACKED_G")
    _assert(pc.imagine_case(pkg, device, config="dmc_proprio", N=256, H=6))


def test_policy_walk_public_methods(pkg, device):
    """Dreamer._policy (dreamer.py:117-190) through RSSM.obs_step(None, None, ...) / obs_step /
    img_step / get_feat / actor(feat): the acting path's public-method surface."""
    _assert(pc.policy_walk_case(pkg, device))


def test_imagine_large_config(pkg, device):
    """BASELINE configs[3] (reference configs.yaml:158-174): 1024 starts x H=15, dyn_deter 4096,
    dyn_hidden / units 1024, 5-layer one-hot actor, 17 actions -- forward + actor gradients.  The
    block-per-row GRU gate kernels (D = 4096), the 5-layer actor loop and every GEMM tiling that
    the `large_imagination` bench numbers come from."""
    _assert(pc.imagine_case(pkg, device, config="large", N=1024, H=15))


def test_observe_large_config(pkg, device):
    """The same widths through observe (stepwise path; E = 12288), forward + every gradient."""
    _assert(pc.observe_case(pkg, device, config="large", B=16, T=6))


def test_imagine_with_action(pkg, device):
    _assert(pc.imagine_with_action_case(pkg, device))


def test_empty_inputs(pkg, device):
    import torch
    d = pc.synth.dims_of("tiny")
    p = pc.to_dev(pc.synth.rssm_params(d), device)
    z = lambda *s: torch.zeros(*s, device=device)
    outs = pkg.kernels.observe(z(0, 4, d.embed), z(0, 4, d.actions), z(0, 4), z(4, 0, d.stoch, d.classes),
                               z(4, 0, d.stoch, d.classes), None, None, pc.kdims(d), pc.rssm_list(pkg, p))
    assert outs[0].shape == (0, 4, d.stoch, d.classes)
    out = pkg.tools.lambda_return_stacked(z(0, 5, 1), z(0, 5, 1), z(0, 5, 1), z(5, 1), 0.95)
    assert out.shape == (0, 5, 1)


def test_bad_shapes_raise(pkg, device):
    import torch
    with pytest.raises(pkg._lib.Dv3Error):
        pkg.kernels.onehot_sample(torch.zeros(2, 2, 40, device=device), None, 0.01)   # classes > 32
    with pytest.raises(pkg._lib.Dv3Error):
        pkg.kernels.ln_silu_fwd(torch.zeros(2, 8), torch.ones(8), torch.zeros(8))      # CPU tensor


@pytest.mark.parametrize("config,B,T,H", [("tiny", 4, 6, 5), ("dmc_proprio", 16, 64, 15)])
def test_whole_train_step(pkg, device, config, B, T, H):
    """WorldModel.loss + ImagBehavior.losses (forward + backward, every parameter gradient)
    against oracle/train_step.py with the same parameters, batch and noise."""
    res = pc.train_step_case(pkg, device, config=config, B=B, T=T, H=H)
    _assert(res)


def test_train_api_runs_and_updates(pkg, device):
    """_train drop-in surface: return tuples, metric keys, parameters move, action mutated."""
    import numpy as np
    import torch
    d = pc.synth.dims_of("tiny")
    P, Pa, Pv = pc.synth.agent_params("tiny", enc_units=d.embed)
    cfg, wm, beh = pc.build_product_agent(pkg, device, "tiny", P, Pa, Pv, device_metrics=False,
                                          encoder=dict(mlp_units=d.embed), decoder=dict(mlp_units=d.embed),
                                          imag_horizon=5)
    data = pc.synth.replay_batch(d, 4, 6, resets=((1, 3),))
    before = {k: v.clone() for k, v in wm.state_dict().items()}
    post, context, metrics = wm._train(data)
    assert set(post) == {"stoch", "deter", "logit"} and post["stoch"].shape == (4, 6, d.stoch, d.classes)
    assert set(context) == {"embed", "feat", "kl", "postent"}
    for k in ("model_loss", "model_grad_norm", "reward_loss", "cont_loss", "kl_free", "dyn_scale", "rep_scale",
              "dyn_loss", "rep_loss", "kl", "prior_ent", "post_ent"):
        assert k in metrics, k
    assert isinstance(metrics["model_loss"], np.ndarray)
    moved = sum(float((wm.state_dict()[k] - before[k]).abs().max()) > 0 for k in before)
    assert moved == len(before)
    reward_fn = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
    feat, state, action, weights, m2 = beh._train(post, reward_fn)
    assert feat.shape == (5, 24, d.flat + d.deter) and not feat.requires_grad
    assert action.shape == (5, 24, d.actions) and weights.shape == (5, 24, 1)
    for k in ("actor_loss", "actor_grad_norm", "value_loss", "value_grad_norm", "actor_entropy", "EMA_005",
              "EMA_095", "value_mean", "target_std", "imag_reward_min", "imag_action_max", "normed_target_mean"):
        assert k in m2, k
    assert all(np.isfinite(np.asarray(v)).all() for v in m2.values())


@pytest.mark.gpu
@pytest.mark.parametrize("M,n", [(1024, 512), (15360, 512), (1024, 1536), (37, 40), (5, 2000)])
def test_ln_param_grads(pkg, device, M, n):
    """LN weight/bias gradient kernel against the torch expression it replaces."""
    import torch
    g = torch.Generator().manual_seed(M + n)
    pre = torch.randn(M, n, generator=g).to(device) * 2 + 0.3
    d_ln = torch.randn(M, n, generator=g).to(device)
    dg, db = pkg.kernels._ln_grads(pre, d_ln)
    xh = torch.nn.functional.layer_norm(pre.double(), (n,), None, None, 1e-3)
    rg, rb = (d_ln.double() * xh).sum(0), d_ln.double().sum(0)
    assert float((dg.double() - rg).abs().max() / rg.abs().max()) < 1e-5
    assert float((db.double() - rb).abs().max() / rb.abs().max()) < 1e-5


def test_observe_backward_runs_persistent(pkg, device):
    """At dmc sizes both directions of observe are single cooperative launches: count the
    library's launches around a forward + backward (bulk products and row kernels excluded by a
    generous bound -- the stepwise backward alone would be > 800 launches)."""
    import torch
    cfgs = pkg.configs
    torch.manual_seed(0)
    cfg = cfgs.make_config("dmc_proprio", device=device)
    wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
    B, T, A = 16, 64, cfg.num_actions
    action = torch.rand(B, T, A, device=device)
    first = torch.zeros(B, T, device=device)
    first[:, 0] = 1
    lib = pkg._lib.lib()
    e = torch.randn(B, T, 1024, device=device, requires_grad=True)
    with pkg.tools.RequiresGrad(wm.dynamics):
        n0 = lib.dv3_launch_count()
        post, prior = wm.dynamics.observe(e, action, first)
        n1 = lib.dv3_launch_count()
        (post["deter"].sum() + post["stoch"].sum() + prior["logit"].sum() + post["logit"].sum()).backward()
        n2 = lib.dv3_launch_count()
    assert n1 - n0 < 40, n1 - n0
    assert n2 - n1 < 120, n2 - n1


def test_fused_adam_clip_matches_torch(pkg, device):
    """tools.Optimizer's CUDA path (flat buffers + dv3_adam_clip_step) against
    nn.utils.clip_grad_norm_ + torch.optim.Adam, the reference's sequence (tools.py:760-776),
    over 12 steps with and without the clip engaging."""
    import torch
    from torch import nn
    torch.manual_seed(0)
    def net():
        return nn.Sequential(nn.Linear(37, 64), nn.LayerNorm(64), nn.SiLU(), nn.Linear(64, 5)).to(device)
    a, b = net(), net()
    b.load_state_dict(a.state_dict())
    for clip in (1000.0, 0.05):
        opt = pkg.tools.Optimizer("t", a.parameters(), lr=3e-3, eps=1e-5, clip=clip)
        ref = torch.optim.Adam(b.parameters(), lr=3e-3, eps=1e-5)
        g = torch.Generator(device=device).manual_seed(1)
        for step in range(12):
            x = torch.randn(32, 37, device=device, generator=g)
            y = torch.randn(32, 5, device=device, generator=g)
            a.requires_grad_(True)
            m = opt(((a(x) - y) ** 2).mean(), a.parameters())
            ref.zero_grad(set_to_none=True)
            ((b(x) - y) ** 2).mean().backward()
            norm = nn.utils.clip_grad_norm_(b.parameters(), clip)
            ref.step()
            assert abs(float(m["t_grad_norm"]) - float(norm)) <= 1e-5 * float(norm) + 1e-8, step
            for pa, pb in zip(a.parameters(), b.parameters()):
                assert float((pa - pb).abs().max()) <= 2e-6, (clip, step)
        sd = opt.state_dict()
        assert set(sd) == {"state", "param_groups"} and len(sd["state"]) == len(list(a.parameters()))


@pytest.mark.gpu
@pytest.mark.parametrize("M,n", [(5000, 512), (4100, 1024), (4096, 500), (4097, 72), (6000, 1536), (300, 512)])
def test_ln_silu_bulk_rows_match_torch(pkg, device, M, n):
    """LayerNorm+SiLU forward / backward (reference networks.py:657-681 blocks, nn.LayerNorm eps 1e-3)
    on bulk row counts, where the warp-per-row kernels take over (M >= 4096, n <= 1024; (6000, 1536)
    and (300, 512) stay on the block-per-row form): against torch fp32 autograd, and the emitted
    hi/lo planes must be the exact split of the fp32 outputs."""
    K = pkg.kernels
    g = torch.Generator().manual_seed(M + n)
    pre = (torch.randn(M, n, generator=g) * 2 + 0.3).to(device)
    gam = (1 + 0.1 * torch.randn(n, generator=g)).to(device)
    bet = (0.1 * torch.randn(n, generator=g)).to(device)
    dy = torch.randn(M, n, generator=g).to(device)
    x = pre.clone().requires_grad_(True)
    ref = torch.nn.functional.silu(torch.nn.functional.layer_norm(x, (n,), gam, bet, 1e-3))
    ref.backward(dy)
    out, sp = K.ln_silu_fwd(pre, gam, bet, 1e-3, with_split=True)
    assert float((out - ref).abs().max()) <= 1e-5 * (1 + float(ref.abs().max()))
    assert torch.equal(sp.hi[:, :n] + sp.lo[:, :n], out)
    assert int((sp.hi.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert torch.equal(K.ln_silu_fwd(pre, gam, bet, 1e-3), out)
    d_pre, d_ln, dsp = K.ln_silu_bwd(pre, gam, bet, dy, 1e-3, with_split=True)
    scale = 1 + float(x.grad.abs().max())
    assert float((d_pre - x.grad).abs().max()) <= 2e-5 * scale
    assert torch.equal(dsp.hi[:, :n] + dsp.lo[:, :n], d_pre)
    v = torch.nn.functional.layer_norm(pre, (n,), gam, bet, 1e-3)
    sg = torch.sigmoid(v)
    assert float((d_ln - dy * sg * (1 + v * (1 - sg))).abs().max()) <= 1e-5 * (1 + float(dy.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("M,K", [(15360, 512), (1024, 1030), (7, 5), (513, 255), (4096, 1536), (64, 4)])
def test_split_is_exact(pkg, device, M, K):
    """hi keeps the top 11 mantissa bits (13 low bits zero), hi + lo == x bit for bit, padding is
    zero -- on the vector path (16-byte aligned pitch) and the scalar one (views, odd pitches)."""
    g = torch.Generator().manual_seed(M * 3 + K)
    x = (torch.randn(M, K, generator=g) * 10 ** torch.randint(-3, 4, (M, 1), generator=g).float()).to(device)
    for src in (x, torch.cat([x, x], 1)[:, 1:1 + K]):
        sp = pkg.kernels.split(src)
        assert sp.rows == M and sp.cols == K and sp.ld % 4 == 0
        assert torch.equal(sp.hi[:, :K] + sp.lo[:, :K], src)
        assert int((sp.hi.view(torch.int32) & 0x1FFF).abs().max()) == 0
        if sp.ld > K:
            assert float(sp.hi[:, K:].abs().max()) == 0.0 and float(sp.lo[:, K:].abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("unimix", [0.01, 0.0])
@pytest.mark.parametrize("with_u", [True, False])
def test_onehot_sample_group_form_is_bit_identical(pkg, device, knob, unimix, with_u):
    """The thread-per-group sampler (32 classes in registers, sums in the butterfly's association
    order) must pick exactly the indices of the lane-per-class kernel -- which is the one pinned
    against the oracle -- for draws (reference tools.py:436-460 + ATen multinomial) and modes."""
    K = pkg.kernels
    g = torch.Generator().manual_seed(11)
    logits = (torch.randn(4096, 32, 32, generator=g) * 3).to(device)
    logits[5, 3] = 0.0                      # exact ties: first index must win
    logits[6, 0, 7] = float("nan")
    u = torch.rand(4096, 32, 32, generator=g).clamp_(1e-30, 1.0).to(device) if with_u else None
    knob("DV3_SAMPLE_GROUP", "0")
    idx0, hot0 = K.onehot_sample(logits, u, unimix)
    knob("DV3_SAMPLE_GROUP", "1")
    idx1, hot1 = K.onehot_sample(logits, u, unimix)
    assert torch.equal(idx0, idx1) and torch.equal(hot0, hot1)
    if not with_u:
        assert int(idx1[5, 3]) == 0
    assert float(hot1.sum()) == 4096 * 32


@pytest.mark.gpu
@pytest.mark.parametrize("M,D", [(1024, 512), (600, 1024), (1024, 256)])
@pytest.mark.parametrize("warp_form", ["1", "0"])
def test_gru_gates_bwd_matches_autograd(pkg, device, knob, M, D, warp_form):
    """Backward of the LayerNorm-GRU gate block (reference networks.py:760-768) through the C ABI,
    block-per-row and warp-per-row kernels (the latter takes D = 512 / 1024 at M >= 512), against
    torch autograd of the same expression."""
    import ctypes as C
    L = pkg._lib
    knob("DV3_GRU_WARP", warp_form)
    gen = torch.Generator().manual_seed(M + D)
    g_pre = torch.randn(M, 3 * D, generator=gen).to(device).requires_grad_(True)
    gam = (1 + 0.1 * torch.randn(3 * D, generator=gen)).to(device)
    bet = (0.1 * torch.randn(3 * D, generator=gen)).to(device)
    h = torch.tanh(torch.randn(M, D, generator=gen)).to(device).requires_grad_(True)
    d_new = torch.randn(M, D, generator=gen).to(device)
    parts = torch.nn.functional.layer_norm(g_pre, (3 * D,), gam, bet, 1e-3)
    parts.retain_grad()
    r, c, u = parts.split(D, -1)
    r = torch.sigmoid(r); c = torch.tanh(r * c); u = torch.sigmoid(u - 1)
    # only the direct (1-u) path of d_h is the kernel's job: detach h inside the candidate terms
    h_new = u * c + (1 - u) * h
    h_new.backward(d_new)
    d_g_pre = torch.empty(M, 3 * D, device=device); d_g_ln = torch.empty(M, 3 * D, device=device)
    d_h = torch.empty(M, D, device=device)
    L.check(L.lib().dv3_gru_gates_bwd(L.fptr(g_pre.detach()), 3 * D, L.fptr(gam), L.fptr(bet), 1e-3,
                                      L.fptr(h.detach()), D, L.fptr(d_new), D, M, D, L.fptr(d_g_pre),
                                      L.fptr(d_g_ln), 3 * D, L.fptr(d_h), D, L.stream_ptr()), "gru_gates_bwd")
    torch.cuda.synchronize()
    for got, want in ((d_g_pre, g_pre.grad), (d_g_ln, parts.grad), (d_h, h.grad)):
        assert float((got - want).abs().max()) <= 2e-5 * (1 + float(want.abs().max()))


@pytest.mark.gpu
def test_onehot_st_bwd_group_form_matches(pkg, device, knob):
    """Straight-through backward of the unimix categorical (reference tools.py:436-460 through
    autograd): the four-lanes-per-group kernel against the lane-per-class one and torch autograd."""
    K = pkg.kernels
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(2048, 32, 32, generator=g) * 2).to(device)
    gs = torch.randn(2048, 32, 32, generator=g).to(device)
    ext = torch.randn(2048, 32, 32, generator=g).to(device)
    knob("DV3_STBWD_GROUP", "0")
    d0 = K.onehot_st_bwd(logits, gs, ext, 0.01)
    knob("DV3_STBWD_GROUP", "1")
    d1 = K.onehot_st_bwd(logits, gs, ext, 0.01)
    d0 = d0[0] if isinstance(d0, tuple) else d0
    d1 = d1[0] if isinstance(d1, tuple) else d1
    assert float((d0 - d1).abs().max()) <= 1e-6 * (1 + float(d0.abs().max()))
    x = logits.clone().requires_grad_(True)
    p = torch.softmax(x, -1) * 0.99 + 0.01 / 32
    probs = torch.softmax(torch.log(p) - torch.logsumexp(torch.log(p), -1, keepdim=True), -1)
    (probs * gs).sum().backward()
    want = x.grad + ext
    assert float((d1 - want).abs().max()) <= 2e-5 * (1 + float(want.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("n,A,with_add", [(512, 6, False), (1024, 17, True), (500, 0, True)])
def test_onehot_linear_warp_form_matches(pkg, device, knob, n, A, with_add):
    """One-hot Linear + LN + SiLU (reference networks.py:208-221 on cat(one-hot stoch, action)):
    the warp-per-row kernel accumulates in the block-per-row kernel's order, so `pre` must be
    bit-identical and `out` equal to fp32 rounding; both against a dense torch product."""
    L = pkg._lib
    M, S, Cc = 1024, 32, 32
    gen = torch.Generator().manual_seed(n + A)
    idx = torch.randint(0, Cc, (M, S), generator=gen, dtype=torch.int32).to(device)
    act = torch.randn(M, max(A, 1), generator=gen).to(device)[:, :A].contiguous() if A else None
    WT = (torch.randn(S * Cc + A, n, generator=gen) / 6).to(device)
    add = torch.randn(M, n, generator=gen).to(device) if with_add else None
    gam = (1 + 0.1 * torch.randn(n, generator=gen)).to(device)
    bet = (0.1 * torch.randn(n, generator=gen)).to(device)
    res = {}
    for form in ("0", "1"):
        knob("DV3_GATHER_WARP", form)
        pre = torch.empty(M, n, device=device); out = torch.empty(M, n, device=device)
        L.check(L.lib().dv3_onehot_linear_ln_silu(L.iptr(idx), S, Cc, L.fptr(act), A, L.fptr(WT), L.fptr(add),
                                                  L.fptr(gam), L.fptr(bet), 1e-3, M, n, L.fptr(pre), L.fptr(out),
                                                  L.stream_ptr()), "onehot_linear")
        res[form] = (pre, out)
    assert torch.equal(res["0"][0], res["1"][0])
    assert float((res["0"][1] - res["1"][1]).abs().max()) <= 2e-6 * (1 + float(res["0"][1].abs().max()))
    hot = torch.nn.functional.one_hot(idx.long(), Cc).float().reshape(M, S * Cc)
    x = torch.cat([hot, act], 1) if A else hot
    ref = x.double() @ WT.double() + (add.double() if with_add else 0)
    assert float((res["1"][0].double() - ref).abs().max()) <= 1e-5 * (1 + float(ref.abs().max()))
    want = torch.nn.functional.silu(torch.nn.functional.layer_norm(res["1"][0], (n,), gam, bet, 1e-3))
    assert float((res["1"][1] - want).abs().max()) <= 1e-5 * (1 + float(want.abs().max()))
