"""Replay batcher: windows bit-identical to the reference's
tools.sample_episodes / from_generator (golden fixture made by tests/golden/make_replay_golden.py
from the live reference), and the pinned-memory device feeder delivers them unchanged."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
pkg = importlib.import_module("dreamerv3-torch_b200")
replay = importlib.import_module("dreamerv3-torch_b200.replay")
from make_replay_golden import make_store      # noqa: E402  (pure numpy; no reference import)

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "replay.npz"))
CASES = {"a": (16, 4, 0), "b": (50, 3, 7)}


@pytest.mark.parametrize("name", sorted(CASES))
def test_windows_match_reference(name):
    length, batch, seed = CASES[name]
    gen = replay.from_generator(replay.sample_episodes(make_store(), length, seed), batch)
    for step in range(3):
        got = next(gen)
        keys = sorted(k.split("/")[2] for k in GOLD.files if k.startswith("%s/%d/" % (name, step)))
        assert sorted(got) == keys and "log_extra" not in got
        for k in keys:
            want = GOLD["%s/%d/%s" % (name, step, k)]
            assert got[k].dtype == want.dtype and got[k].shape == want.shape == (batch, length) + want.shape[2:]
            assert np.array_equal(got[k], want), (name, step, k)


def test_windows_mark_episode_joins():
    """Every window starts with is_first, and every splice point inside a window is marked."""
    store = make_store(seed=3)
    gen = replay.sample_episodes(store, 64, seed=1)
    for _ in range(20):
        w = next(gen)
        assert len(w["reward"]) == 64 and w["is_first"][0]
        assert w["is_first"].sum() >= 2        # no episode of the store is 64 long


def test_short_episodes_are_skipped():
    store = {"x": {"reward": np.zeros(1, np.float32), "is_first": np.ones(1, bool)},
             "y": {"reward": np.arange(5, dtype=np.float32), "is_first": np.zeros(5, bool)}}
    w = next(replay.sample_episodes(store, 9, seed=0))
    assert len(w["reward"]) == 9 and set(np.unique(w["reward"])) <= set(range(5))


def test_feeder_needs_cuda():
    with pytest.raises(RuntimeError):
        replay.DeviceFeeder(iter([]), "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [2, 3])
def test_device_feeder_delivers_batches_in_order(depth):
    length, batch, seed = 16, 4, 0
    mk = lambda: replay.from_generator(replay.sample_episodes(make_store(), length, seed), batch)
    feeder = replay.DeviceFeeder(mk(), "cuda:0", depth=depth)
    ref = mk()
    burn = torch.randn(2048, 2048, device="cuda:0")
    sums = []
    for step in range(12):
        dev = next(feeder)
        want = next(ref)
        # keep the consumer stream busy so refills really overlap pending reads
        for _ in range(4):
            burn = torch.tanh(burn @ burn * 1e-3)
        sums.append((dev["vector"].sum() + dev["reward"].sum(), want))
        for k, v in want.items():
            assert dev[k].is_cuda and tuple(dev[k].shape) == v.shape
            assert np.array_equal(dev[k].cpu().numpy(), v), (step, k)
    assert feeder.h2d_bytes_per_batch == sum(v.nbytes for v in want.values())
    # deferred reads (queued before the next __next__) saw the right batch too
    for s, want in sums:
        np.testing.assert_allclose(s.item(), float(want["vector"].sum(dtype=np.float64) +
                                                   want["reward"].sum(dtype=np.float64)), rtol=1e-4, atol=1e-3)


@pytest.mark.gpu
def test_device_feeder_rejects_shape_change():
    def gen():
        yield {"a": np.zeros((2, 3), np.float32)}
        yield {"a": np.zeros((2, 3), np.float32)}
        yield {"a": np.zeros((2, 4), np.float32)}
    feeder = replay.DeviceFeeder(gen(), "cuda:0", depth=2)
    next(feeder)
    with pytest.raises(RuntimeError, match="shape"):
        next(feeder)
