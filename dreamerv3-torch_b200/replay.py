"""Replay windows -> device batches: the caller side of `WorldModel._train`.

`sample_episodes` / `from_generator` produce the same windows, in the same order, from the same
seed as the reference's numpy batcher (reference tools.py:310-371): one `RandomState(seed)`,
episode drawn with `choice(p = length share)`, start drawn with `randint(0, total - 1)` for the
first piece only, later pieces start at 0 and mark `is_first`.  `DeviceFeeder` is the B200 part:
it stages every batch in pinned host buffers and copies it to HBM on its own stream, `depth`
batches ahead, so the train step (a CUDA-graph replay) never waits on the host -- the reference
converts and copies synchronously inside `preprocess` (reference models.py:174-190).
"""
from __future__ import annotations

import collections

import numpy as np
import torch


def sample_episodes(episodes, length, seed=0):
    """Endless generator of windows `{key: array[length, ...]}` cut from `episodes` (a dict of
    `{key: array[T_i, ...]}` dicts).  Keys containing "log_" are dropped.  Random stream and
    window boundaries follow reference tools.py:323-367."""
    rng = np.random.RandomState(seed)
    while True:
        eps = list(episodes.values())
        lens = np.array([len(next(iter(e.values()))) for e in eps])
        share = lens / np.sum(lens)
        window, filled = None, 0
        while filled < length:
            # same draw as choosing among the episode objects themselves
            ep = eps[int(rng.choice(len(eps), p=share))]
            total = len(next(iter(ep.values())))
            if total < 2:
                continue
            if window is None:
                start = int(rng.randint(0, total - 1))
                window = {k: v[start:min(start + length, total)].copy()
                          for k, v in ep.items() if "log_" not in k}
                if "is_first" in window:
                    window["is_first"][0] = True
            else:
                room = length - filled
                window = {k: np.append(window[k], v[0:min(room, total)].copy(), axis=0)
                          for k, v in ep.items() if "log_" not in k}
                if "is_first" in window:
                    window["is_first"][filled] = True
            filled = len(next(iter(window.values())))
        yield window


def from_generator(generator, batch_size):
    """Stack `batch_size` consecutive windows along a new leading axis (reference
    tools.py:310-321)."""
    while True:
        rows = [next(generator) for _ in range(batch_size)]
        yield {k: np.stack([r[k] for r in rows], 0) for k in rows[0].keys()}


class DeviceFeeder:
    """Iterator over device-resident batches, `depth` batches in flight.

    Each slot owns one pinned host buffer and one device buffer per key (allocated from the first
    batch; shapes and dtypes must not change afterwards -- a mismatch raises).  `__next__` returns
    the oldest slot's device tensors after making the consumer's current stream wait on that
    slot's copy event, after handing the previously returned slot back for refilling.  The tensors
    returned by one `__next__` are valid for all work enqueued on the current stream before the
    next `__next__` (the refill copy waits on an event recorded there).
    """

    def __init__(self, batches, device, depth=2, float_keys_to_fp32=True):
        self._it = iter(batches)
        self._dev = torch.device(device)
        if self._dev.type != "cuda":
            raise RuntimeError("DeviceFeeder stages batches for a CUDA device; got %s" % device)
        if depth < 2:
            raise ValueError("depth must be >= 2 (one slot in use, the others in flight)")
        self._cast = float_keys_to_fp32
        self._slots = [None] * depth
        self._ready = collections.deque()      # slot indices whose copy has been issued
        self._free = collections.deque(range(depth))
        self._held = None                      # slot handed out by the previous __next__
        self._stream = torch.cuda.Stream(device=self._dev)
        self.h2d_bytes_per_batch = 0
        while self._free:
            self._issue()

    def _host_array(self, v):
        a = np.asarray(v)
        if self._cast and a.dtype == np.float64:
            a = a.astype(np.float32)
        return a

    def _issue(self):
        host = {k: self._host_array(v) for k, v in next(self._it).items()}
        i = self._free.popleft()
        slot = self._slots[i]
        if slot is None:
            pin = {k: torch.empty(a.shape, dtype=torch.from_numpy(np.empty(0, a.dtype)).dtype,
                                  pin_memory=True) for k, a in host.items()}
            dev = {k: torch.empty_like(t, device=self._dev) for k, t in pin.items()}
            slot = self._slots[i] = {"pin": pin, "dev": dev,
                                     "done": torch.cuda.Event(), "released": None}
            self.h2d_bytes_per_batch = sum(t.numel() * t.element_size() for t in pin.values())
        if set(host) != set(slot["pin"]):
            raise RuntimeError("DeviceFeeder: batch keys changed: %s vs %s"
                               % (sorted(host), sorted(slot["pin"])))
        slot["done"].synchronize()             # the previous copy out of these pinned buffers
        for k, a in host.items():
            p = slot["pin"][k]
            if tuple(a.shape) != tuple(p.shape):
                raise RuntimeError("DeviceFeeder: shape of %r changed: %s vs %s"
                                   % (k, a.shape, tuple(p.shape)))
            p.numpy()[...] = a
        with torch.cuda.stream(self._stream):
            if slot["released"] is not None:
                self._stream.wait_event(slot["released"])     # consumer finished with the slot
            for k, p in slot["pin"].items():
                slot["dev"][k].copy_(p, non_blocking=True)
            slot["done"].record(self._stream)
        self._ready.append(i)

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self._dev)
        if self._held is not None:
            ev = torch.cuda.Event()
            ev.record(cur)                     # everything queued so far may still read the slot
            self._slots[self._held]["released"] = ev
            self._free.append(self._held)
            self._held = None
        i = self._ready.popleft()
        cur.wait_event(self._slots[i]["done"])
        self._held = i
        while self._free:                      # refill what was just handed back: depth-1 ahead
            self._issue()
        return self._slots[i]["dev"]
