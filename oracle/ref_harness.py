"""Drive the UNMODIFIED reference modules -- test / measurement infrastructure, never product.

The reference tree is looked up at ``$DV3_REFERENCE_DIR``, then /root/reference (the build
container), then ``oracle/_ref`` (the copy staged by ``oracle/build_ref.py``, which travels to the
GPU box).  Used by ``oracle/pin_against_reference.py`` (pins the oracle), by
``tests/golden/make_golden.py`` (writes the committed fixtures), by ``tests/test_gpu_reference.py``
(CUDA path vs the reference itself at the real configs) and by ``bench.py``'s reference arms
(``--impl reference``, ``cpu_baseline``, the same-GPU eager comparator).  The product package
never imports this file.

Shims -- applied to the imported modules in memory, the reference files are
never edited:
  1. networks.MLP defaults device="cuda" (networks.py:606) -> patched to the device in use.
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
import time
import types

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    for cand in (os.environ.get("DV3_REFERENCE_DIR"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "networks.py")):
            return cand
    return os.path.join(_HERE, "_ref")


REF = _find_reference()


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "networks.py"))


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def load_reference(device="cpu"):
    """-> (tools, networks, models) reference modules, MLP's device default set to ``device``.
    The reference's module names are generic (``tools``, ``networks``, ``models``): they are
    imported with the reference directory first on sys.path and must not be shadowed."""
    if not available():
        raise RuntimeError(f"reference modules not found (looked in {REF}); run oracle/build_ref.py "
                           "in the build container")
    if sys.path[0] != REF:
        sys.path.insert(0, REF)
    for name in ("tools", "networks", "models"):
        mod = sys.modules.get(name)
        if mod is not None and os.path.dirname(os.path.abspath(getattr(mod, "__file__", "") or "")) != REF:
            raise RuntimeError(f"module {name!r} already imported from elsewhere: {mod.__file__}")
    with quiet():
        import tools as rtools          # noqa
        import networks as rnetworks    # noqa
        import models as rmodels        # noqa
    dflt = list(rnetworks.MLP.__init__.__defaults__)
    names = rnetworks.MLP.__init__.__code__.co_varnames[1:rnetworks.MLP.__init__.__code__.co_argcount]
    dflt[len(dflt) - (len(names) - names.index("device"))] = str(device)
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


SUITE_ACTIONS = {"dmc_proprio": 6, "dmc_vision": 6, "atari100k": 18, "crafter": 17}


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
    """-> (WorldModel, ImagBehavior) reference modules on ``cfg.device``."""
    rtools, rnetworks, rmodels = load_reference(cfg.device)
    torch.manual_seed(seed)
    with quiet():
        wm = rmodels.WorldModel(ObsSpace(shapes), None, 0, cfg)
        beh = rmodels.ImagBehavior(cfg, wm)
    # the reference's modules do not place themselves: dreamer.py moves the whole agent with
    # ``Dreamer(...).to(config.device)`` (dreamer.py:~590)
    wm.to(cfg.device)
    beh.to(cfg.device)
    wm.requires_grad_(False)
    beh.requires_grad_(False)
    return wm, beh


def suite_shapes(suite):
    return PROPRIO_SHAPES if suite == "dmc_proprio" else VISION_SHAPES


def suite_batch(suite, B=16, T=64, seed=0, resets=()):
    """The synthetic replay batch of a suite as the reference's dataset yields it (numpy dict);
    proprio batches carry a dummy ``image`` because ``preprocess`` divides it unconditionally
    (models.py:180)."""
    A = SUITE_ACTIONS[suite]
    data = synthetic_batch(B, T, A, seed, onehot_action=(suite in ("atari100k", "crafter")),
                           resets=resets, vision=(suite != "dmc_proprio"))
    if suite == "dmc_proprio":
        data["image"] = np.zeros((B, T, 2, 2, 3), np.uint8)
    return data


def train_rate(suite="dmc_proprio", device="cpu", steps=3, warmup=1, seed=0, B=16, T=64):
    """Train steps/s of the reference's own ``WorldModel._train`` -> ``ImagBehavior._train``
    (the body of Dreamer._train, dreamer.py:194-200) on ``device``: 'cpu' = all host threads,
    'cuda:N' = the reference's stock eager PyTorch-CUDA path.  Sampling noise comes from torch's
    global generator, as in the reference."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = reference_config((suite,), num_actions=SUITE_ACTIONS[suite], device=str(device))
    wm, beh = build_agent(cfg, suite_shapes(suite), seed)
    data = suite_batch(suite, B, T, seed)
    reward_fn = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
    gpu = str(device).startswith("cuda")
    times = []
    for i in range(warmup + steps):
        feed = {k: v.copy() for k, v in data.items()}
        if gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with quiet():
            post, _, _ = wm._train(feed)
            beh._train(post, reward_fn)
        if gpu:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return dict(value=len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores,
                steps=len(times), warmup=warmup, device=str(device))
