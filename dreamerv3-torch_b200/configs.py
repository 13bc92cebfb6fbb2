"""Hyper-parameters of the reference's ``configs.yaml`` that the hot path reads, as plain Python
(the YAML loader / argparse layer of the reference is out of scope).  ``make_config(suite)``
returns the same attribute names ``WorldModel`` / ``ImagBehavior`` expect from the reference's
config namespace; line numbers cite /root/reference/configs.yaml.
"""
from __future__ import annotations

import copy
from types import SimpleNamespace

DEFAULTS = dict(                                            # configs.yaml:1-130 (defaults)
    device="cuda:0", compile=False, precision=32, reward_EMA=True,
    dyn_hidden=512, dyn_deter=512, dyn_stoch=32, dyn_discrete=32, dyn_rec_depth=1,
    dyn_mean_act="none", dyn_std_act="sigmoid2", dyn_min_std=0.1,
    grad_heads=["decoder", "reward", "cont"], units=512, act="SiLU", norm=True,
    encoder=dict(mlp_keys="$^", cnn_keys="image", act="SiLU", norm=True, cnn_depth=32,
                 kernel_size=4, minres=4, mlp_layers=5, mlp_units=1024, symlog_inputs=True),
    decoder=dict(mlp_keys="$^", cnn_keys="image", act="SiLU", norm=True, cnn_depth=32,
                 kernel_size=4, minres=4, mlp_layers=5, mlp_units=1024, cnn_sigmoid=False,
                 image_dist="mse", vector_dist="symlog_mse", outscale=1.0),
    actor=dict(layers=2, dist="normal", entropy=3e-4, unimix_ratio=0.01, std="learned",
               min_std=0.1, max_std=1.0, temp=0.1, lr=3e-5, eps=1e-5, grad_clip=100.0,
               outscale=1.0),
    critic=dict(layers=2, dist="symlog_disc", slow_target=True, slow_target_update=1,
                slow_target_fraction=0.02, lr=3e-5, eps=1e-5, grad_clip=100.0, outscale=0.0),
    reward_head=dict(layers=2, dist="symlog_disc", loss_scale=1.0, outscale=0.0),
    cont_head=dict(layers=2, loss_scale=1.0, outscale=1.0),
    dyn_scale=0.5, rep_scale=0.1, kl_free=1.0, weight_decay=0.0, unimix_ratio=0.01,
    initial="learned",
    batch_size=16, batch_length=64, model_lr=1e-4, opt_eps=1e-8, grad_clip=1000, opt="adam",
    discount=0.997, discount_lambda=0.95, imag_horizon=15, imag_gradient="dynamics",
    imag_gradient_mix=0.0,
    num_actions=6,
    device_metrics=False,   # ours: keep metrics as device tensors (one host sync per step, or none)
)

SUITES = {
    "dmc_proprio": dict(encoder=dict(mlp_keys=".*", cnn_keys="$^"),          # configs.yaml:140-147
                        decoder=dict(mlp_keys=".*", cnn_keys="$^"), num_actions=6),
    "dmc_vision": dict(encoder=dict(mlp_keys="$^", cnn_keys="image"),        # configs.yaml:149-156
                       decoder=dict(mlp_keys="$^", cnn_keys="image"), num_actions=6),
    "atari100k": dict(actor=dict(dist="onehot", std="none"),                 # configs.yaml:176-190
                      imag_gradient="reinforce", num_actions=18),
    "crafter": dict(dyn_hidden=1024, dyn_deter=4096, units=1024,             # configs.yaml:158-174
                    encoder=dict(mlp_keys="$^", cnn_keys="image", cnn_depth=96),
                    decoder=dict(mlp_keys="$^", cnn_keys="image", cnn_depth=96),
                    # the reference's overlay writes `value: {layers: 5}`, a key nothing reads
                    # (the critic is built from `critic`, models.py:246-257): its critic keeps 2 layers
                    actor=dict(layers=5, dist="onehot", std="none"),
                    reward_head=dict(layers=5), cont_head=dict(layers=5),
                    imag_gradient="reinforce", num_actions=17),
}

PROPRIO_SHAPES = {"orientations": (14,), "height": (1,), "velocity": (9,), "image": (64, 64, 3)}
VISION_SHAPES = {"image": (64, 64, 3)}


def _merge(base, upd):
    for k, v in upd.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            _merge(base[k], v)
        else:
            base[k] = v


def make_config(suite="dmc_proprio", **overrides):
    cfg = copy.deepcopy(DEFAULTS)
    _merge(cfg, copy.deepcopy(SUITES[suite]))
    _merge(cfg, overrides)
    return SimpleNamespace(**cfg)


class _Space:
    def __init__(self, shape):
        self.shape = tuple(shape)


class ObsSpace:
    """Stand-in for the gym Dict space: WorldModel only reads ``.spaces[k].shape``
    (reference models.py:35)."""

    def __init__(self, shapes):
        self.spaces = {k: _Space(v) for k, v in shapes.items()}
