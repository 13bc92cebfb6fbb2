"""Seeded synthetic parameters / inputs for the oracle -- TEST INFRASTRUCTURE, NOT PRODUCT.

Shapes and ``state_dict`` key names are the reference's (probed from the live
modules); values are variance-scaled random numbers with perturbed LayerNorm affine terms,
biases and initial-state vector so that no term of the arithmetic is trivially 0 or 1.  This
file does not read /root/reference, so it travels to the GPU box.
"""
from __future__ import annotations

import math

import torch

from dv3_oracle import RSSMDims

CONFIGS = {
    # name: RSSM dims, units, actor layers, actor dist                     (configs.yaml lines)
    "dmc_proprio": dict(dims=dict(deter=512, hidden=512, actions=6, embed=1024), units=512,
                        actor_layers=2, actor_dist="normal"),          # 140-147
    "dmc_vision": dict(dims=dict(deter=512, hidden=512, actions=6, embed=4096), units=512,
                       actor_layers=2, actor_dist="normal"),           # 149-156
    "atari100k": dict(dims=dict(deter=512, hidden=512, actions=18, embed=4096), units=512,
                      actor_layers=2, actor_dist="onehot"),            # 176-190
    "large": dict(dims=dict(deter=4096, hidden=1024, actions=17, embed=12288), units=1024,
                  actor_layers=5, actor_dist="onehot"),                # 165-173, 203-212
    # default widths with 1 / 3 actor layers and 16 continuous / 32 discrete actions: the corner
    # cases of the persistent imagination kernel (test-only shapes, not reference configs)
    "wide_l1": dict(dims=dict(deter=512, hidden=512, actions=16, embed=1024), units=512,
                    actor_layers=1, actor_dist="normal"),
    "wide_l3": dict(dims=dict(deter=512, hidden=512, actions=32, embed=1024), units=512,
                    actor_layers=3, actor_dist="onehot"),
    # reduced widths for fast CPU-side checks
    "tiny": dict(dims=dict(stoch=8, classes=8, deter=64, hidden=48, actions=3, embed=40),
                 units=32, actor_layers=2, actor_dist="normal"),
    "tiny_onehot": dict(dims=dict(stoch=8, classes=8, deter=64, hidden=48, actions=5, embed=40),
                        units=32, actor_layers=3, actor_dist="onehot"),
}


def dims_of(name: str) -> RSSMDims:
    return RSSMDims(**CONFIGS[name]["dims"])


def _lin(gen, out, inp, scale=1.0):
    std = scale * math.sqrt(2.0 / (inp + out)) / 0.87962566103423978
    return torch.randn(out, inp, generator=gen) * std


def _ln(gen, n):
    return 1.0 + 0.1 * torch.randn(n, generator=gen), 0.1 * torch.randn(n, generator=gen)


def rssm_params(d: RSSMDims, seed=0):
    g = torch.Generator().manual_seed(seed)
    SC, D, Hd, A, E = d.flat, d.deter, d.hidden, d.actions, d.embed
    p = {}
    p["_img_in_layers.0.weight"] = _lin(g, Hd, SC + A)
    p["_img_in_layers.1.weight"], p["_img_in_layers.1.bias"] = _ln(g, Hd)
    p["_cell.layers.GRU_linear.weight"] = _lin(g, 3 * D, Hd + D)
    p["_cell.layers.GRU_norm.weight"], p["_cell.layers.GRU_norm.bias"] = _ln(g, 3 * D)
    p["_img_out_layers.0.weight"] = _lin(g, Hd, D)
    p["_img_out_layers.1.weight"], p["_img_out_layers.1.bias"] = _ln(g, Hd)
    p["_obs_out_layers.0.weight"] = _lin(g, Hd, D + E)
    p["_obs_out_layers.1.weight"], p["_obs_out_layers.1.bias"] = _ln(g, Hd)
    p["_imgs_stat_layer.weight"] = _lin(g, SC, Hd, 2.0)
    p["_imgs_stat_layer.bias"] = 0.1 * torch.randn(SC, generator=g)
    p["_obs_stat_layer.weight"] = _lin(g, SC, Hd, 2.0)
    p["_obs_stat_layer.bias"] = 0.1 * torch.randn(SC, generator=g)
    p["W"] = 0.5 * torch.randn(1, D, generator=g)
    return p


def mlp_params(name, inp, units, layers, out, seed=0, std_layer=False, out_scale=1.0):
    """state_dict of a reference networks.MLP: layers.<name>_linear<i>.weight, ..._norm<i>.*,
    mean_layer.*, (std_layer.*)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    d = inp
    for i in range(layers):
        p[f"layers.{name}_linear{i}.weight"] = _lin(g, units, d)
        p[f"layers.{name}_norm{i}.weight"], p[f"layers.{name}_norm{i}.bias"] = _ln(g, units)
        d = units
    p["mean_layer.weight"] = _lin(g, out, d, out_scale)
    p["mean_layer.bias"] = 0.1 * torch.randn(out, generator=g)
    if std_layer:
        p["std_layer.weight"] = _lin(g, out, d, out_scale)
        p["std_layer.bias"] = 0.1 * torch.randn(out, generator=g)
    return p


def actor_params(name, seed=1):
    c = CONFIGS[name]
    d = dims_of(name)
    return mlp_params("Actor", d.flat + d.deter, c["units"], c["actor_layers"], d.actions, seed,
                      std_layer=(c["actor_dist"] == "normal"))


def uniforms(gen, *shape):
    return torch.rand(*shape, generator=gen).clamp_(1e-30, 1.0)


def observe_inputs(d: RSSMDims, B, T, seed=0, resets=((1, 3), (2, 4)), onehot_action=False):
    g = torch.Generator().manual_seed(seed + 100)
    embed = torch.randn(B, T, d.embed, generator=g)
    if onehot_action:
        idx = torch.randint(0, d.actions, (B, T), generator=g)
        action = torch.nn.functional.one_hot(idx, d.actions).float()
    else:
        action = torch.rand(B, T, d.actions, generator=g) * 2 - 1
    is_first = torch.zeros(B, T)
    if T > 0:
        is_first[:, 0] = 1.0
    for b, t in resets:
        if b < B and t < T:
            is_first[b, t] = 1.0
    u_prior = uniforms(g, T, B, d.stoch, d.classes)
    u_post = uniforms(g, T, B, d.stoch, d.classes)
    return embed, action, is_first, u_prior, u_post


def imagine_inputs(d: RSSMDims, N, H, seed=0, actor_dist="normal"):
    g = torch.Generator().manual_seed(seed + 200)
    idx = torch.randint(0, d.classes, (N, d.stoch), generator=g)
    start = {"stoch": torch.nn.functional.one_hot(idx, d.classes).float(),
             "deter": torch.tanh(torch.randn(N, d.deter, generator=g)),
             "logit": torch.randn(N, d.stoch, d.classes, generator=g)}
    if actor_dist == "normal":
        act_noise = torch.randn(H, N, d.actions, generator=g)
    else:
        act_noise = uniforms(g, H, N, d.actions)
    u_state = uniforms(g, H, N, d.stoch, d.classes)
    return start, act_noise, u_state


# --------------------------------------------------------------------------------------
# whole-agent parameter sets (reference state_dict keys of WorldModel / actor / value)
# --------------------------------------------------------------------------------------
PROPRIO_KEYS = {"orientations": 14, "height": 1, "velocity": 9}


def agent_params(config="dmc_proprio", seed=0, enc_units=1024, enc_layers=5, head_layers=2):
    """-> (P_wm, P_actor, P_value) for the proprio suites (MLP encoder / decoder)."""
    c = CONFIGS[config]
    d = dims_of(config)
    assert enc_units == d.embed, "MLP encoder width must equal the RSSM embed width"
    F_ = d.flat + d.deter
    U = c["units"]
    g = torch.Generator().manual_seed(seed + 300)
    P = {}
    obs = sum(PROPRIO_KEYS.values())
    enc = mlp_params("Encoder", obs, enc_units, enc_layers, 1, seed + 1)
    P.update({"encoder._mlp." + k: v for k, v in enc.items() if not k.startswith("mean_layer")})
    P.update({"dynamics." + k: v for k, v in rssm_params(d, seed).items()})
    dec = mlp_params("Decoder", F_, enc_units, enc_layers, 1, seed + 2)
    P.update({"heads.decoder._mlp." + k: v for k, v in dec.items()
              if not k.startswith("mean_layer")})
    for k, n in PROPRIO_KEYS.items():
        P[f"heads.decoder._mlp.mean_layer.{k}.weight"] = _lin(g, n, enc_units)
        P[f"heads.decoder._mlp.mean_layer.{k}.bias"] = 0.1 * torch.randn(n, generator=g)
    P.update({"heads.reward." + k: v
              for k, v in mlp_params("Reward", F_, U, head_layers, 255, seed + 3, out_scale=0.3).items()})
    P.update({"heads.cont." + k: v
              for k, v in mlp_params("Cont", F_, U, head_layers, 1, seed + 4).items()})
    P_actor = actor_params(config, seed + 5)
    P_value = mlp_params("Value", F_, U, head_layers, 255, seed + 6, out_scale=0.3)
    return P, P_actor, P_value


def replay_batch(d: RSSMDims, B=16, T=64, seed=0, resets=(), onehot_action=False):
    """SURVEY.md 8d synthetic replay batch as a numpy dict (what the reference's dataset yields)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    data = {k: rs.randn(B, T, n).astype(np.float32) for k, n in PROPRIO_KEYS.items()}
    if onehot_action:
        data["action"] = np.eye(d.actions, dtype=np.float32)[rs.randint(0, d.actions, size=(B, T))]
    else:
        data["action"] = rs.uniform(-1, 1, size=(B, T, d.actions)).astype(np.float32)
    data["reward"] = rs.randn(B, T).astype(np.float32)
    data["discount"] = np.ones((B, T), np.float32)
    data["is_terminal"] = np.zeros((B, T), np.float32)
    first = np.zeros((B, T), np.float32)
    first[:, 0] = 1.0
    for b, t in resets:
        first[b, t] = 1.0
    data["is_first"] = first
    return data


def train_noise(d: RSSMDims, B, T, H, seed=0, actor_dist="normal"):
    g = torch.Generator().manual_seed(seed + 400)
    N = B * T
    return dict(u_prior=uniforms(g, T, B, d.stoch, d.classes),
                u_post=uniforms(g, T, B, d.stoch, d.classes),
                act_noise=(torch.randn(H, N, d.actions, generator=g) if actor_dist == "normal"
                           else uniforms(g, H, N, d.actions)),
                u_state=uniforms(g, H, N, d.stoch, d.classes))
