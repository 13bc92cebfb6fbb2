"""Drive the LIVE reference (/root/reference) on CPU -- build-container only, test infrastructure.

Used by ``oracle/pin_against_reference.py`` (pins the oracle) and
``tests/golden/make_golden.py`` (writes the committed fixtures).  Nothing here is imported
by the product, by ``-m gpu`` tests, ``smoke()`` or ``bench.py``: /root/reference does not
exist on the GPU box.

Shims -- applied to the imported modules in memory, the reference tree is
never edited:
  1. networks.MLP defaults device="cuda" (networks.py:606) -> patched default "cpu".
  2. configs.yaml is read with PyYAML, which leaves ``1e-4`` style scalars as strings.
Supplied noise: ``torch.multinomial`` and ``torch.distributions.normal._standard_normal``
are swapped for tape readers for the duration of a call (class NoiseTape).
"""
from __future__ import annotations

import contextlib
import io
import os
import re
import sys
import types

import numpy as np
import torch

REF = os.environ.get("DV3_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "networks.py"))


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def load_reference():
    """-> (tools, networks, models) reference modules, MLP device default patched to cpu."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with quiet():
        import tools as rtools          # noqa
        import networks as rnetworks    # noqa
        import models as rmodels        # noqa
    dflt = list(rnetworks.MLP.__init__.__defaults__)
    names = rnetworks.MLP.__init__.__code__.co_varnames[1:rnetworks.MLP.__init__.__code__.co_argcount]
    dflt[len(dflt) - (len(names) - names.index("device"))] = "cpu"
    rnetworks.MLP.__init__.__defaults__ = tuple(dflt)
    return rtools, rnetworks, rmodels


_SCI = re.compile(r"^[+-]?\d+(\.\d*)?[eE][+-]?\d+$")


def _coerce(v):
    if isinstance(v, str) and _SCI.match(v):
        return float(v)
    if isinstance(v, dict):
        return {k: _coerce(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_coerce(x) for x in v]
    return v


def _merge(base, upd):
    for k, v in upd.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            _merge(base[k], v)
        else:
            base[k] = v


def reference_config(overlays=("dmc_proprio",), num_actions=6, **extra):
    import yaml
    with open(os.path.join(REF, "configs.yaml")) as f:
        raw = _coerce(yaml.safe_load(f))
    cfg = dict(raw["defaults"])
    for name in overlays:
        _merge(cfg, raw[name])
    cfg.update(device="cpu", compile=False, causal_world_model=False, num_actions=num_actions,
               precision=32)
    cfg.update(extra)
    return types.SimpleNamespace(**cfg)


class _Space:
    def __init__(self, shape):
        self.shape = tuple(shape)


class ObsSpace:
    def __init__(self, shapes):
        self.spaces = {k: _Space(v) for k, v in shapes.items()}


PROPRIO_SHAPES = {"orientations": (14,), "height": (1,), "velocity": (9,), "image": (64, 64, 3)}
VISION_SHAPES = {"image": (64, 64, 3)}


class NoiseTape:
    """Feed supplied uniforms / normals to the reference's samplers, in consumption order.

    tape: list of ("u", Tensor[...,K]) | ("n", Tensor[...]) entries.
    A "u" entry answers one torch.multinomial(probs_2d, 1, True) call with
    argmax(probs/(-log u)) (ATen's own n=1 algorithm); an "n" entry answers one
    _standard_normal(shape) call.
    """

    def __init__(self, tape):
        self.tape = list(tape)
        self.pos = 0

    def _next(self, kind):
        assert self.pos < len(self.tape), "noise tape exhausted"
        k, t = self.tape[self.pos]
        assert k == kind, f"tape entry {self.pos} is {k}, sampler asked for {kind}"
        self.pos += 1
        return t

    def __enter__(self):
        self._mn = torch.multinomial
        self._sn = torch.distributions.normal._standard_normal

        def multinomial(probs, num_samples, replacement=False, *, generator=None, out=None):
            assert num_samples == 1
            u = self._next("u").reshape(probs.shape).to(probs.dtype)
            return torch.argmax(probs / (-torch.log(u)), dim=-1, keepdim=True)

        def standard_normal(shape, dtype, device):
            return self._next("n").reshape(shape).to(dtype)

        torch.multinomial = multinomial
        torch.distributions.normal._standard_normal = standard_normal
        return self

    def __exit__(self, *exc):
        torch.multinomial = self._mn
        torch.distributions.normal._standard_normal = self._sn
        return False


def synthetic_batch(B=16, T=64, A=6, seed=0, onehot_action=False, resets=(), vision=False):
    """SURVEY.md 8d synthetic replay batch (numpy dict, as the reference's dataset yields)."""
    rs = np.random.RandomState(seed)
    data = {}
    if not vision:
        data["orientations"] = rs.randn(B, T, 14).astype(np.float32)
        data["height"] = rs.randn(B, T, 1).astype(np.float32)
        data["velocity"] = rs.randn(B, T, 9).astype(np.float32)
    data["image"] = rs.randint(0, 255, size=(B, T, 64, 64, 3)).astype(np.uint8)
    if onehot_action:
        idx = rs.randint(0, A, size=(B, T))
        data["action"] = np.eye(A, dtype=np.float32)[idx]
    else:
        data["action"] = rs.uniform(-1, 1, size=(B, T, A)).astype(np.float32)
    data["reward"] = rs.randn(B, T).astype(np.float32)
    data["discount"] = np.ones((B, T), np.float32)
    data["is_terminal"] = np.zeros((B, T), np.float32)
    first = np.zeros((B, T), np.float32)
    first[:, 0] = 1.0
    for b, t in resets:
        first[b, t] = 1.0
    data["is_first"] = first
    return data


def build_agent(cfg, shapes, seed=0):
    """-> (WorldModel, ImagBehavior) reference modules on CPU."""
    rtools, rnetworks, rmodels = load_reference()
    torch.manual_seed(seed)
    with quiet():
        wm = rmodels.WorldModel(ObsSpace(shapes), None, 0, cfg)
        beh = rmodels.ImagBehavior(cfg, wm)
    wm.requires_grad_(False)
    beh.requires_grad_(False)
    return wm, beh
