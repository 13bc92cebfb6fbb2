"""Host-side mirror of the reference's ``models`` surface: ``WorldModel`` and ``ImagBehavior``
with the same constructors, ``_train`` / ``_imagine`` signatures, return tuples, metric keys and
``state_dict`` names (reference models.py:29-213, 218-689).

The rollouts (RSSM.observe, the actor-in-the-loop imagination), the KL balance, the two-hot
heads' log-probs / means and the lambda-return run on libdv3_b200.so.  Both ``_train`` methods
take an optional ``noise`` argument so the caller can supply the uniforms / normals (parity
tests); without it they are drawn from torch's generator.
"""
from __future__ import annotations

import contextlib
import copy
import os

import torch
from torch import nn

from . import kernels as K
from . import networks
from . import tools


class RewardEMA:
    """5/95-percentile EMA normaliser (reference models.py:11-26)."""

    def __init__(self, device, alpha=1e-2):
        self.alpha = alpha
        self.range = torch.tensor([0.05, 0.95], device=device)

    def __call__(self, x, ema_vals):
        if x.is_cuda and 0 < x.numel() <= K.REWARD_EMA_MAX and ema_vals.is_contiguous():
            os_ = self.offset_scale(x, ema_vals)
            return os_[0], os_[1]
        q = torch.quantile(x.detach().flatten(), self.range)
        ema_vals[:] = self.alpha * q + (1 - self.alpha) * ema_vals
        scale = torch.clip(ema_vals[1] - ema_vals[0], min=1.0)
        return ema_vals[0].detach(), scale.detach()

    def offset_scale(self, x, ema_vals):
        """One kernel: sort + 5/95 % quantiles + EMA (in place on ``ema_vals``) -> [offset, scale]."""
        return K.reward_ema(x, ema_vals, self.alpha)


class WorldModel(nn.Module):
    def __init__(self, obs_space, act_space, step, config, grad_sync=None):
        super().__init__()
        if config.precision != 32:
            raise NotImplementedError("precision: 32 is the contract of the B200 path")
        self._step = step
        self._config = config
        shapes = {k: tuple(v.shape) for k, v in obs_space.spaces.items()}
        self.encoder = networks.MultiEncoder(shapes, **config.encoder)
        self.embed_size = self.encoder.outdim
        self.dynamics = networks.RSSM(
            config.dyn_stoch, config.dyn_deter, config.dyn_hidden, config.dyn_rec_depth,
            config.dyn_discrete, config.act, config.norm, config.dyn_mean_act, config.dyn_std_act,
            config.dyn_min_std, config.unimix_ratio, config.initial, config.num_actions,
            self.embed_size, config.device)
        feat_size = config.dyn_stoch * config.dyn_discrete + config.dyn_deter
        self.heads = nn.ModuleDict()
        self.heads["decoder"] = networks.MultiDecoder(feat_size, shapes, **config.decoder)
        self.heads["reward"] = networks.MLP(
            feat_size, (255,) if config.reward_head["dist"] == "symlog_disc" else (),
            config.reward_head["layers"], config.units, config.act, config.norm,
            dist=config.reward_head["dist"], outscale=config.reward_head["outscale"],
            name="Reward")
        self.heads["cont"] = networks.MLP(
            feat_size, (), config.cont_head["layers"], config.units, config.act, config.norm,
            dist="binary", outscale=config.cont_head["outscale"], name="Cont")
        for name in config.grad_heads:
            if name not in self.heads:
                raise AssertionError(name)
        self.to(config.device)
        self._model_opt = tools.Optimizer(
            "model", self.parameters(), config.model_lr, config.opt_eps, config.grad_clip,
            config.weight_decay, opt=config.opt, grad_sync=grad_sync)
        self._scales = dict(reward=config.reward_head["loss_scale"],
                            cont=config.cont_head["loss_scale"])
        self.requires_grad_(False)

    def _head_streams(self, ref, n):
        """One CUDA stream per head (none on CPU tensors or with DV3_SIDE_STREAM=0)."""
        if not ref.is_cuda or os.environ.get("DV3_SIDE_STREAM", "1") == "0":
            return []
        pool = getattr(self, "_hstreams", None)
        if pool is None or len(pool) < n:
            pool = [torch.cuda.Stream(device=ref.device) for _ in range(n)]
            object.__setattr__(self, "_hstreams", pool)
        return pool[:n]

    def preprocess(self, obs):
        """numpy / CPU tensors -> fp32 device tensors (reference models.py:174-190); pinned host
        tensors are copied asynchronously."""
        dev = self._config.device
        out = {}
        for k, v in obs.items():
            t = v if torch.is_tensor(v) else torch.as_tensor(v)
            out[k] = t.to(dev, non_blocking=True).to(torch.float32)
        if "image" in out:
            out["image"] = out["image"] / 255.0
        if "discount" in out:
            out["discount"] = (out["discount"] * self._config.discount).unsqueeze(-1)
        if "is_first" not in out or "is_terminal" not in out:
            raise AssertionError("batch needs is_first and is_terminal")
        out["cont"] = (1.0 - out["is_terminal"]).unsqueeze(-1)
        return out

    def loss(self, data, noise=None):
        """Forward of the world-model step -> (mean loss, post, aux); no optimizer."""
        cfg = self._config
        embed = self.encoder(data)
        post, prior = self.dynamics.observe(embed, data["action"], data["is_first"], noise=noise)
        kl_loss, kl_value, dyn_loss, rep_loss, post_ent, prior_ent = \
            self.dynamics.kl_loss_with_entropy(post, prior, cfg.kl_free, cfg.dyn_scale,
                                               cfg.rep_scale)
        if kl_loss.shape != embed.shape[:2]:
            raise AssertionError(kl_loss.shape)
        feat = self.dynamics.get_feat(post)
        # The heads (decoder, reward, cont) are independent chains of 1024-row products that each
        # fill less than half of the SMs: every head runs -- forward here, and therefore its
        # backward too (autograd replays a node on its forward stream) -- on its own stream, a
        # parallel branch of the captured step graph.
        streams = self._head_streams(feat, len(self.heads))
        main = torch.cuda.current_stream() if streams else None
        if streams:
            feat2 = feat.reshape(-1, feat.shape[-1])
            if feat2.shape[0] >= 64:
                K.split_of(feat2, feat)          # shared tf32 planes are made before the fork
        logps = {}
        for i, (name, head) in enumerate(self.heads.items()):
            ctx = contextlib.nullcontext()
            if streams:
                streams[i].wait_stream(main)
                ctx = torch.cuda.stream(streams[i])
            with ctx:
                pred = head(feat if name in cfg.grad_heads else feat.detach())
                preds = pred if isinstance(pred, dict) else {name: pred}
                for key, dist in preds.items():
                    lp = dist.log_prob(data[key])
                    if lp.shape != embed.shape[:2]:
                        raise AssertionError((key, lp.shape))
                    logps[key] = lp
        for st in streams:
            main.wait_stream(st)
        if streams and self._model_opt._sync is not None:
            # data parallel: each head's slice of the flat gradient is all-reduced on the head's
            # stream, overlapping the rest of the backward pass
            segs = []
            for i, head in enumerate(self.heads.values()):
                rng = self._model_opt.param_range(list(head.parameters()))
                if rng is not None:
                    segs.append((rng[0], rng[1], streams[i]))
            if K.grad_side_enabled():
                # the RSSM's parameter gradients are produced on the gradient side stream (a branch
                # under the encoder's backward, kernels._Observe.backward): reduce them there
                rng = self._model_opt.param_range(list(self.dynamics.parameters()))
                if rng is not None:
                    segs.append((rng[0], rng[1], K.grad_side_stream(feat.device)))
            self._model_opt.set_segments(segs or None)
        # model_loss = mean(sum_k scale_k * (-log_prob_k) + kl_loss) (reference models.py:140-152)
        names = list(logps)
        if embed.is_cuda and len(names) < 8:
            mean_loss, neg = K.loss_mean([logps[k] for k in names] + [kl_loss],
                                         [-self._scales.get(k, 1.0) for k in names] + [1.0])
            losses = {k: neg[i].reshape(lp.shape) for i, k in enumerate(names)}
        else:
            losses = {k: -v for k, v in logps.items()}
            mean_loss = torch.mean(sum(v * self._scales.get(k, 1.0) for k, v in losses.items())
                                   + kl_loss)
        aux = dict(embed=embed, feat=feat, prior=prior, losses=losses, kl_value=kl_value,
                   dyn_loss=dyn_loss, rep_loss=rep_loss, post_ent=post_ent, prior_ent=prior_ent)
        return mean_loss, post, aux

    def _train(self, data, noise=None):
        return self._train_end(self._train_begin(data, noise))

    def _train_begin(self, data, noise=None, sync=False):
        """Forward + backward of the world-model step (gradients in the optimizer's buffer, weights
        untouched).  ``_train_end`` applies the update and builds the metrics; work that only reads
        the weights may run in between (graphs.TrainStepGraph, pipelined schedule)."""
        data = self.preprocess(data)
        with tools.RequiresGrad(self):
            loss, post, aux = self.loss(data, noise)
            opt_state = self._model_opt.backward(loss, sync=sync)
        return dict(post=post, aux=aux, opt_state=opt_state)

    def _train_end(self, st):
        cfg = self._config
        post, aux = st["post"], st["aux"]
        metrics = self._model_opt.step(st["opt_state"])
        metrics.update({f"{k}_loss": v.detach() for k, v in aux["losses"].items()})
        metrics["kl_free"] = cfg.kl_free
        metrics["dyn_scale"] = cfg.dyn_scale
        metrics["rep_scale"] = cfg.rep_scale
        metrics["dyn_loss"] = aux["dyn_loss"]
        metrics["rep_loss"] = aux["rep_loss"]
        metrics["kl"] = tools.mean_scalar(aux["kl_value"])
        metrics["prior_ent"] = tools.mean_scalar(aux["prior_ent"])
        metrics["post_ent"] = tools.mean_scalar(aux["post_ent"])
        context = dict(embed=aux["embed"], feat=aux["feat"], kl=aux["kl_value"],
                       postent=aux["post_ent"])
        idx = self.dynamics._to_idx(post["stoch"])
        post = {k: v.detach() for k, v in post.items()}
        self.dynamics.tag_idx(post["stoch"], idx)        # _imagine starts from these indices
        if not getattr(cfg, "device_metrics", False):
            metrics = tools.to_host(metrics)
        return post, context, metrics

    def video_pred(self, data):
        data = self.preprocess(data)
        embed = self.encoder(data)
        states, _ = self.dynamics.observe(embed[:6, :5], data["action"][:6, :5].clone(),
                                          data["is_first"][:6, :5])
        dec = self.heads["decoder"]
        recon = dec(self.dynamics.get_feat(states))["image"].mode()[:6]
        init = {k: v[:, -1] for k, v in states.items()}
        prior = self.dynamics.imagine_with_action(data["action"][:6, 5:], init)
        openl = dec(self.dynamics.get_feat(prior))["image"].mode()
        model = torch.cat([recon[:, :5], openl], 1)
        truth = data["image"][:6]
        return torch.cat([truth, model, (model - truth + 1.0) / 2.0], 2)


class ImagBehavior(nn.Module):
    def __init__(self, config, world_model, future_predictor=None, grad_sync=None):
        super().__init__()
        self._config = config
        self._world_model = world_model
        feat_size = config.dyn_stoch * config.dyn_discrete + config.dyn_deter
        self.actor = networks.MLP(
            feat_size, (config.num_actions,), config.actor["layers"], config.units, config.act,
            config.norm, config.actor["dist"], config.actor["std"], config.actor["min_std"],
            config.actor["max_std"], absmax=1.0, temp=config.actor["temp"],
            unimix_ratio=config.actor["unimix_ratio"], outscale=config.actor["outscale"],
            name="Actor")
        self.value = networks.MLP(
            feat_size, (255,) if config.critic["dist"] == "symlog_disc" else (),
            config.critic["layers"], config.units, config.act, config.norm, config.critic["dist"],
            outscale=config.critic["outscale"], name="Value")
        if config.critic["slow_target"]:
            self._slow_value = copy.deepcopy(self.value)
            self._updates = 0
        if config.reward_EMA:
            self.register_buffer("ema_vals", torch.zeros((2,)))
            self.reward_ema = RewardEMA(device=config.device)
        self.to(config.device)
        kw = dict(wd=config.weight_decay, opt=config.opt, grad_sync=grad_sync)
        self._actor_opt = tools.Optimizer("actor", self.actor.parameters(), config.actor["lr"],
                                          config.actor["eps"], config.actor["grad_clip"], **kw)
        self._value_opt = tools.Optimizer("value", self.value.parameters(), config.critic["lr"],
                                          config.critic["eps"], config.critic["grad_clip"], **kw)
        self.actor.requires_grad_(False)
        self.value.requires_grad_(False)
        self._slow_flat = None
        self._side_stream = None
        if config.critic["slow_target"]:
            self._slow_value.requires_grad_(False)
            self._rehome_slow()

    def _side(self, ref):
        """Side stream of the critic branch (None on CPU tensors or with DV3_SIDE_STREAM=0)."""
        if not ref.is_cuda or os.environ.get("DV3_SIDE_STREAM", "1") == "0":
            return None
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=ref.device)
        return self._side_stream

    def _rehome_slow(self):
        """Give the slow critic the flat layout of the critic's parameter buffer, so that the EMA
        update (reference models.py:683-689) is one kernel over two flat buffers."""
        opt = self._value_opt
        if not getattr(opt, "_flat", False):
            return
        flat = torch.zeros_like(opt._fp)
        with torch.no_grad():
            for p, off in zip(self._slow_value.parameters(), opt._offsets):
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self._slow_flat = flat

    # ---- rollout -------------------------------------------------------------------------
    def _imagine(self, start, policy, horizon, noise=None):
        """start: dict of [B,T,...] posterior tensors -> (feats [H,N,F] detached, states dict of
        [H,N,...], actions [H,N,A]); reference models.py:448-548.  ``noise`` = (act_noise
        [H,N,A], u_state [H,N,S,C])."""
        dyn = self._world_model.dynamics
        S, Cc = dyn._stoch, dyn._discrete
        flat = {k: v.reshape([-1] + list(v.shape[2:])) for k, v in start.items()}
        start_idx = dyn._to_idx(start["stoch"]).reshape(-1, S)
        N = flat["deter"].shape[0]
        dev = flat["deter"].device
        spec = policy.actor_spec()
        if noise is None:
            A = dyn._num_actions
            an = (torch.randn(horizon, N, A, device=dev) if spec.dist == "normal"
                  else torch.rand(horizon, N, A, device=dev))
            noise = (an, torch.rand(horizon, N, S, Cc, device=dev))
        feat, logit, action, idx, mean_raw, std_raw, sp = K.imagine_full(
            start_idx, flat["deter"].detach(), noise[0], noise[1], None, horizon,
            dyn.dims, spec, dyn.kernel_params(), policy.actor_params(),
            start_logit=flat.get("logit"))
        SC = S * Cc
        states = dict(stoch=feat[..., :SC].reshape(horizon, N, S, Cc), deter=feat[..., SC:],
                      logit=logit)
        # feat rows are [one-hot stoch | deter]: get_feat(states) is feat itself, and every head
        # that reads the rollout features (reward, cont, critic, slow critic, actor) shares one
        # tf32 split of it
        for v in (states["stoch"], states["deter"]):
            v._dv3_feat, v._dv3_feat_version = feat, feat._version
        imag_feat = feat.detach()
        # raw head outputs of the in-loop actor (differentiable w.r.t. the actor parameters):
        # losses() builds the policy distribution from them instead of re-running the actor
        action._dv3_policy_raw = (policy, mean_raw, std_raw if std_raw.numel() else None)
        if sp is not None:
            K.attach_split(feat, sp)
            K.attach_split(imag_feat, sp)
        return imag_feat, states, action

    # ---- training step -------------------------------------------------------------------
    def losses(self, start, objective, noise=None):
        """Forward of the behaviour step (reference models.py:327-429) without the optimizer
        calls -> (actor_loss, value_loss, rollout tuple, metrics of device tensors, aux)."""
        cfg = self._config
        metrics = {}
        with tools.RequiresGrad(self.actor):
            imag_feat, imag_state, imag_action = self._imagine(start, self.actor,
                                                               cfg.imag_horizon, noise)
            # One critic forward over all H steps serves the lambda-return target, the baseline
            # and the value loss: the reference evaluates self.value three times on the same
            # (detached) features with unchanged weights (models.py:629, 421, 662).
            # The critic branch (critic + slow-critic forward, value loss, and -- because autograd
            # replays a node on the stream of its forward -- the critic's whole backward) lives on
            # a side stream: in the captured step graph it is a parallel branch that fills the SMs
            # the latency-bound imagination backward leaves idle.
            main = torch.cuda.current_stream() if imag_feat.is_cuda else None
            side = self._side(imag_feat)
            on_side = (lambda: torch.cuda.stream(side)) if side is not None else contextlib.nullcontext
            if side is not None:
                side.wait_stream(main)
            slow_mode = None
            with tools.RequiresGrad(self.value), on_side():
                v_all = self.value(imag_feat)
                v_mode = v_all.mode().detach() if isinstance(v_all, tools.DiscDist) else None
                feat_m1 = imag_feat[:-1].detach()
                sp = K.split_of_attached(imag_feat)
                if sp is not None:
                    K.attach_split(feat_m1, sp.prefix(feat_m1.shape[0] * feat_m1.shape[1]))
                if cfg.critic["slow_target"]:
                    slow_mode = self._slow_value(feat_m1).mode().detach()
            reward = objective(imag_feat, imag_state, imag_action)
            if side is not None:
                main.wait_stream(side)              # v_mode feeds the lambda-return target
            raw = getattr(imag_action, "_dv3_policy_raw", None)
            reuse = raw is not None and raw[0] is self.actor
            grad_mode = cfg.imag_gradient
            fused = (imag_feat.is_cuda and v_mode is not None and grad_mode in ("dynamics", "reinforce")
                     and (not cfg.reward_EMA or (imag_feat.shape[0] - 1) * imag_feat.shape[1]
                          <= K.REWARD_EMA_MAX))
            policy = logp = None
            if reuse and self.actor._dist == "normal" and raw[2] is not None and fused:
                # entropy / log-prob straight from the in-loop actor's raw head outputs: the same
                # numbers the reference gets from self.actor(imag_feat) (models.py:349), one kernel
                actor_ent, logp = K.normal_policy(raw[1], raw[2], imag_action, self.actor._min_std,
                                                  self.actor._max_std, grad_mode == "reinforce")
            else:
                if reuse:
                    std = raw[2] if raw[2] is not None else self.actor._std
                    policy = self.actor.dist(self.actor._dist, raw[1], std, self.actor._shape)
                else:
                    policy = self.actor(imag_feat)
                actor_ent = policy.entropy()
                if fused and grad_mode == "reinforce":
                    logp = policy.log_prob(imag_action)
            target, weights, base = self._compute_target(imag_feat, imag_state, reward,
                                                         value_mode=v_mode)
            if fused:
                os_ = None
                if cfg.reward_EMA:
                    os_ = self.reward_ema.offset_scale(target, self.ema_vals)
                    metrics["EMA_005"] = self.ema_vals[0].detach()   # valid until the next step
                    metrics["EMA_095"] = self.ema_vals[1].detach()
                actor_loss, normed = K.actor_loss(target, base, weights, actor_ent, logp, os_,
                                                  cfg.actor["entropy"], grad_mode)
                if cfg.reward_EMA:
                    metrics.update(tools.tensorstats(normed, "normed_target"))
            else:
                actor_loss, mets = self._compute_actor_loss(imag_feat, imag_action, target, weights,
                                                            base, policy, value_mode=v_mode)
                actor_loss = actor_loss - cfg.actor["entropy"] * actor_ent[:-1, ..., None]
                actor_loss = torch.mean(actor_loss)
                metrics.update(mets)
        if side is not None:
            side.wait_stream(main)                  # target, weights, reward, actions
        with tools.RequiresGrad(self.value), on_side():
            value = (tools.DiscDist(logits=v_all.logits[:-1]) if v_mode is not None
                     else self.value(feat_m1))
            lp_target = value.log_prob(target.detach())
            lp_slow = value.log_prob(slow_mode) if slow_mode is not None else None
            if imag_feat.is_cuda:
                value_loss = K.value_loss(lp_target, lp_slow, weights)
            else:
                value_loss = -lp_target
                if lp_slow is not None:
                    value_loss = value_loss - lp_slow
                value_loss = torch.mean(weights[:-1] * value_loss[:, :, None])
            metrics.update(tools.tensorstats(v_mode[:-1] if v_mode is not None else value.mode(), "value"))
            metrics.update(tools.tensorstats(target, "target"))
            metrics.update(tools.tensorstats(reward, "imag_reward"))
            if cfg.actor["dist"] in ["onehot"]:
                metrics.update(tools.tensorstats(torch.argmax(imag_action, dim=-1).float(),
                                                 "imag_action"))
            else:
                metrics.update(tools.tensorstats(imag_action, "imag_action"))
            metrics["actor_entropy"] = tools.mean_scalar(actor_ent)
        if side is not None:
            main.wait_stream(side)                  # callers continue on the current stream
        aux = dict(reward=reward, target=target, value=value)
        return actor_loss, value_loss, (imag_feat, imag_state, imag_action, weights), metrics, aux

    def _train(self, start, objective, noise=None):
        cfg = self._config
        self._update_slow_target()
        actor_loss, value_loss, roll, metrics, _ = self.losses(start, objective, noise)
        imag_feat, imag_state, imag_action, weights = roll
        side = self._side(imag_feat)
        with tools.RequiresGrad(self):
            if side is None:
                metrics.update(self._actor_opt(actor_loss, self.actor.parameters()))
                metrics.update(self._value_opt(value_loss, self.value.parameters()))
            else:
                # critic backward + Adam on the side stream, concurrently with the actor's
                # (imagination) backward on the current one
                main = torch.cuda.current_stream()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    m_v = self._value_opt(value_loss, self.value.parameters())
                metrics.update(self._actor_opt(actor_loss, self.actor.parameters()))
                metrics.update(m_v)
                main.wait_stream(side)
        if not getattr(cfg, "device_metrics", False):
            metrics = tools.to_host(metrics)
        return imag_feat, imag_state, imag_action, weights, metrics

    def _compute_target(self, imag_feat, imag_state, reward, value_mode=None):
        """reference models.py:620-638; the target comes back stacked [H-1,N,1].  ``value_mode``:
        the critic's mode over all H steps when the caller already evaluated it."""
        cfg = self._config
        wm = self._world_model
        weights = None
        if "cont" in wm.heads:
            inp = wm.dynamics.get_feat(imag_state)
            cont = wm.heads["cont"](inp)
            if inp.is_cuda and isinstance(cont, tools.Bernoulli):
                discount, weights = K.discount_weights(cont.logits, cfg.discount)
            else:
                discount = cfg.discount * cont.mean
        else:
            discount = cfg.discount * torch.ones_like(reward)
        value = value_mode if value_mode is not None else self.value(imag_feat).mode()
        target = tools.lambda_return_stacked(reward[1:], value[:-1], discount[1:], value[-1],
                                             cfg.discount_lambda)
        if weights is None:
            weights = torch.cumprod(torch.cat([torch.ones_like(discount[:1]), discount[:-1]], 0),
                                    0).detach()
        return target, weights, value[:-1]

    def _compute_actor_loss(self, imag_feat, imag_action, target, weights, base, policy=None,
                            value_mode=None):
        cfg = self._config
        metrics = {}
        if policy is None:
            policy = self.actor(imag_feat.detach())
        if torch.is_tensor(target) is False:
            target = torch.stack(target, dim=1)
        if cfg.reward_EMA:
            offset, scale = self.reward_ema(target, self.ema_vals)
            normed_target = (target - offset) / scale
            normed_base = (base - offset) / scale
            adv = normed_target - normed_base
            metrics.update(tools.tensorstats(normed_target, "normed_target"))
            metrics["EMA_005"] = self.ema_vals[0].clone()
            metrics["EMA_095"] = self.ema_vals[1].clone()
        else:
            adv = target - base
        if cfg.imag_gradient == "dynamics":
            actor_target = adv
        elif cfg.imag_gradient in ("reinforce", "both"):
            actor_target = (policy.log_prob(imag_action)[:-1][:, :, None]
                            * (target - (value_mode[:-1] if value_mode is not None
                                         else self.value(imag_feat[:-1]).mode())).detach())
            if cfg.imag_gradient == "both":
                mix = cfg.imag_gradient_mix
                actor_target = mix * target + (1 - mix) * actor_target
                metrics["imag_gradient_mix"] = mix
        else:
            raise NotImplementedError(cfg.imag_gradient)
        return -weights[:-1] * actor_target, metrics

    def _update_slow_target(self):
        cfg = self._config
        if cfg.critic["slow_target"]:
            if self._updates % cfg.critic["slow_target_update"] == 0:
                mix = cfg.critic["slow_target_fraction"]
                with torch.no_grad():
                    if self._slow_flat is not None:
                        K.ema_mix(self._slow_flat, self._value_opt._fp, mix)
                    else:
                        src = list(self.value.parameters())
                        dst = list(self._slow_value.parameters())
                        torch._foreach_mul_(dst, 1 - mix)
                        torch._foreach_add_(dst, src, alpha=mix)
                K.invalidate_weight_splits()
            self._updates += 1
