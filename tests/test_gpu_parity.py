"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.
Indices bit-exact; floats within 1e-4 relative (BASELINE.json north_star)."""
import pytest

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


@pytest.mark.parametrize("config,B,T", [("tiny", 3, 5), ("dmc_proprio", 16, 64), ("dmc_vision", 16, 16)])
def test_observe_fwd_bwd(pkg, device, config, B, T):
    _assert(pc.observe_case(pkg, device, config=config, B=B, T=T))


def test_observe_with_state(pkg, device):
    _assert(pc.observe_case(pkg, device, config="dmc_proprio", B=5, T=7, with_state=True))


def test_obs_step_teacher_forced(pkg, device):
    _assert(pc.obs_step_teacher_forced_case(pkg, device, config="dmc_proprio", B=16, T=6))


@pytest.mark.parametrize("config,N,H", [("tiny", 9, 4), ("tiny_onehot", 9, 4), ("dmc_proprio", 1024, 15),
                                        ("atari100k", 256, 15)])
def test_imagine_fwd_bwd(pkg, device, config, N, H):
    _assert(pc.imagine_case(pkg, device, config=config, N=N, H=H))


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
