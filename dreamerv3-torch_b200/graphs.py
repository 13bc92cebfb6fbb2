"""Whole-train-step CUDA graph: the reference's ``Dreamer._train`` body (dreamer.py:194-200 --
``WorldModel._train`` followed by ``ImagBehavior._train`` on the detached posteriors) captured
once and replayed, so that the ~1400 kernel launches of a step cost one ``cudaGraphLaunch``.

Every call is exactly one training step.  The first ``warmup`` calls run eagerly (lazy CUDA
initialisation, allocator warm-up), the next call captures -- recording only -- and replays.
Inputs are copied into static device tensors before each replay (asynchronously from pinned
host memory); sampling noise is drawn inside the graph from torch's graph-safe generator, or
supplied per call (``noise=dict(u_prior, u_post, act_noise, u_state)``, staged like the batch);
metrics come back as device scalars, or as numpy after one stacked device->host copy.

``pipeline=True`` -- the PIPELINED schedule.  The behaviour update of batch t only READS the world
model (it imagines with the weights the world-model update of batch t produced and trains actor and
critic), and the world-model forward + backward of batch t+1 reads the same weights; so call t+1 runs
``ImagBehavior._train`` of batch t on a second stream, concurrently with ``WorldModel`` forward +
backward of batch t+1, joins, and only then applies the world model's Adam.  Both halves are made of
latency-bound kernels that leave most of the GPU idle on their own.  Every update sees exactly the
inputs it sees in the sequential schedule -- the parameter trajectories are identical (tested bit for
bit) -- but the actor / critic update of the batch passed to call t lands during call t+1 (``flush()``
applies the last one), and the ``beh_metrics`` / ``imag_*`` a call returns belong to the previous batch.
"""
from __future__ import annotations

import torch

from . import kernels as K
from . import tools


class TrainStepGraph:
    def __init__(self, world_model, behavior, objective=None, warmup=3, device_metrics=False,
                 pipeline=False, lazy_metrics=False):
        self.wm, self.beh = world_model, behavior
        self.pipeline = bool(pipeline)
        # lazy_metrics: host metrics come back ONE CALL LATE (None from the first call, drain() for
        # the last): the call packs its metrics on the device, starts one asynchronous copy into
        # pinned memory and returns without waiting for the step, so the host prepares the next call
        # while the GPU still runs this one
        self.lazy_metrics = bool(lazy_metrics)
        self._lazy_pending = None
        self._pin = [None, None]
        self._start = None                   # static start states of the pending behaviour update
        self._pending_noise = None
        if objective is None:
            objective = lambda f, s, a: world_model.heads["reward"](
                world_model.dynamics.get_feat(s)).mode()
        self.objective = objective
        self.warmup = int(warmup)
        self.device_metrics = device_metrics
        self._calls = 0
        self._graph = None
        self._static = None
        self._out = None
        self.library_launches_per_step = 0   # libdv3_b200 kernel nodes recorded in the graph
        # warm-up steps and the capture share one side stream: autograd's gradient accumulators
        # remember the stream they were created on, and the legacy default stream must not be
        # made to wait on a capturing stream
        self._stream = torch.cuda.Stream(device=world_model._config.device)
        self._beh_stream = torch.cuda.Stream(device=world_model._config.device)
        if self.pipeline:
            self.warmup = max(self.warmup, 2)    # the prologue (world model only) + one eager body
        cfg = world_model._config
        if cfg.critic["slow_target"] and cfg.critic["slow_target_update"] != 1:
            raise NotImplementedError("graph capture bakes in the slow-critic update of every "
                                      "step (slow_target_update must be 1)")

    # -- staging -------------------------------------------------------------------------
    def _stage(self, data, noise=None):
        dev = self.wm._config.device
        data = dict(data)
        if noise is not None:
            data.update({"__noise_" + k: v for k, v in noise.items()})
        if self._static is None:
            self._static = {}
            for k, v in data.items():
                t = v if torch.is_tensor(v) else torch.as_tensor(v)
                self._static[k] = torch.empty(t.shape, dtype=t.dtype, device=dev)
        if set(data) != set(self._static):
            raise K.L.Dv3Error("TrainStepGraph: the set of batch / noise entries changed")
        for k, buf in self._static.items():
            v = data[k]
            t = v if torch.is_tensor(v) else torch.as_tensor(v)
            if t.shape != buf.shape or t.dtype != buf.dtype:
                raise K.L.Dv3Error(f"TrainStepGraph: batch entry {k!r} changed shape/dtype "
                                   f"({tuple(t.shape)} {t.dtype} vs {tuple(buf.shape)} {buf.dtype})")
            buf.copy_(t, non_blocking=True)
        return self._static

    def _run(self, staged):
        cfg = self.wm._config
        keep = getattr(cfg, "device_metrics", False)
        cfg.device_metrics = True
        batch = {k: v for k, v in staged.items() if not k.startswith("__noise_")}
        n1 = n2 = None
        if "__noise_u_prior" in staged:
            n1 = (staged["__noise_u_prior"], staged["__noise_u_post"])
            n2 = (staged["__noise_act_noise"], staged["__noise_u_state"])
        try:
            if self.pipeline:
                return self._run_pipelined(batch, n1, n2)
            post, context, m1 = self.wm._train(batch, noise=n1)
            feat, state, action, weights, m2 = self.beh._train(post, self.objective, noise=n2)
        finally:
            cfg.device_metrics = keep
        return dict(post=post, context=context, wm_metrics=m1, imag_feat=feat, imag_state=state,
                    imag_action=action, weights=weights, beh_metrics=m2)

    # -- pipelined schedule ----------------------------------------------------------------
    def _stash(self, post):
        """The posterior of this batch -> the static start-state buffers the NEXT call's behaviour
        update reads (fixed addresses: the captured graph feeds itself through them)."""
        dyn = self.wm.dynamics
        idx = dyn._to_idx(post["stoch"])
        if self._start is None:
            self._start = {k: torch.empty_like(v) for k, v in post.items()}
            self._start_idx = torch.empty_like(idx)
        for k, v in post.items():
            self._start[k].copy_(v)
        self._start_idx.copy_(idx)
        dyn.tag_idx(self._start["stoch"], self._start_idx)

    def _run_pipelined(self, batch, n1, n2):
        if self._start is None:
            # prologue: nothing to train the behaviour on yet
            post, context, m1 = self.wm._train(batch, noise=n1)
            self._stash(post)
            return dict(post=post, context=context, wm_metrics=m1, imag_feat=None, imag_state=None,
                        imag_action=None, weights=None, beh_metrics={})
        cur = torch.cuda.current_stream()
        sb = self._beh_stream
        fork = torch.cuda.Event()
        fork.record(cur)
        # batch t+1: forward + backward, no update yet.  Enqueued first so that, in data-parallel
        # runs, the world model's all-reduces are issued ahead of the behaviour's (collectives
        # execute in issue order); the two branches still start together (the fork event).
        st = self.wm._train_begin(batch, noise=n1, sync=True)
        sb.wait_event(fork)
        with torch.cuda.stream(sb):          # batch t: reads the world model, updates actor / critic
            feat, state, action, weights, m2 = self.beh._train(self._start, self.objective, noise=n2)
        cur.wait_stream(sb)
        post, context, m1 = self.wm._train_end(st)      # ... now the world model may change
        self._stash(post)
        return dict(post=post, context=context, wm_metrics=m1, imag_feat=feat, imag_state=state,
                    imag_action=action, weights=weights, beh_metrics=m2)

    def flush(self, noise=None):
        """Pipelined schedule: apply the behaviour update of the last batch (eagerly).  Returns its
        (feat, state, action, weights, metrics), or None when nothing is pending."""
        if not self.pipeline or self._start is None or getattr(self, "_flushed", False):
            return None
        cfg = self.wm._config
        keep = getattr(cfg, "device_metrics", False)
        cfg.device_metrics = True
        n2 = None
        if noise is not None:
            dev = cfg.device
            n2 = (torch.as_tensor(noise["act_noise"]).to(dev), torch.as_tensor(noise["u_state"]).to(dev))
        try:
            cur = torch.cuda.current_stream()
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                out = self.beh._train(self._start, self.objective, noise=n2)
            cur.wait_stream(self._stream)
        finally:
            cfg.device_metrics = keep
        K.invalidate_weight_splits()
        self._flushed = True                 # the captured graph would train on this batch again
        return out

    # -- one training step ---------------------------------------------------------------
    def __call__(self, data, noise=None):
        if getattr(self, "_flushed", False):
            raise K.L.Dv3Error("TrainStepGraph: flush() ended the pipelined schedule; build a new "
                               "TrainStepGraph to continue training")
        batch = self._stage(data, noise)
        if self._graph is None and self._calls < self.warmup:
            cur = torch.cuda.current_stream()
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                out = self._run(batch)
            cur.wait_stream(self._stream)
        else:
            if self._graph is None:
                torch.cuda.synchronize()
                # planes cached by an eager forward since the last optimizer step must not be
                # reused by the capture: the split kernels have to be recorded in the graph
                K.invalidate_weight_splits()
                self._graph = torch.cuda.CUDAGraph()
                n0 = K.L.lib().dv3_launch_count()
                with torch.cuda.graph(self._graph, stream=self._stream):
                    self._out = self._run(batch)
                self.library_launches_per_step = int(K.L.lib().dv3_launch_count() - n0)
            self._graph.replay()
            # A replay rewrites the flat parameter buffers (fused Adam, slow-critic EMA) without
            # running any Python: neither the epoch nor a version counter moves, so planes cached
            # by eager forwards between steps (Dreamer._policy -> actor(feat), the encoder,
            # video_pred) would go stale -- and planes cached during capture live in the graph's
            # private pool.  Every replay therefore ends the epoch.
            K.invalidate_weight_splits()
            out = self._out
        self._calls += 1
        if self.device_metrics:
            return out
        out = dict(out)
        if self.lazy_metrics:
            out["wm_metrics"], out["beh_metrics"] = self._lazy(out["wm_metrics"], out["beh_metrics"])
            return out
        out["wm_metrics"] = tools.to_host(out["wm_metrics"])
        out["beh_metrics"] = tools.to_host(out["beh_metrics"])
        return out

    # -- metrics one call late -----------------------------------------------------------------
    def _lazy(self, m1, m2):
        groups = (m1 or {}, m2 or {})
        layout, parts = [], []
        for gi, grp in enumerate(groups):
            for k, v in grp.items():
                if torch.is_tensor(v):
                    layout.append((gi, k, tuple(v.shape), v.numel()))
                    parts.append(v.detach().reshape(-1).float())
                else:
                    layout.append((gi, k, None, v))
        n = sum(p.numel() for p in parts)
        slot = self._calls & 1
        if self._pin[slot] is None or self._pin[slot].numel() < n:
            self._pin[slot] = torch.empty(max(n, 1), dtype=torch.float32).pin_memory()
        ev = torch.cuda.Event()
        if parts:
            self._pin[slot][:n].copy_(torch.cat(parts), non_blocking=True)
        ev.record()
        prev, self._lazy_pending = self._lazy_pending, (slot, ev, layout)
        return self._unpack(prev) if prev is not None else (None, None)

    def _unpack(self, pending):
        import numpy as np
        slot, ev, layout = pending
        ev.synchronize()
        host = self._pin[slot].numpy()
        out, pos = ({}, {}), 0
        for gi, k, shape, n in layout:
            if shape is None:
                out[gi][k] = n
            else:
                out[gi][k] = np.array(host[pos:pos + n]).reshape(shape)
                pos += n
        return out

    def drain(self):
        """lazy_metrics: the (wm_metrics, beh_metrics) of the last call, or None."""
        prev, self._lazy_pending = self._lazy_pending, None
        return self._unpack(prev) if prev is not None else None

    @property
    def captured(self):
        return self._graph is not None
