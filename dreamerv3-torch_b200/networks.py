"""Host-side mirror of the reference's ``networks`` surface (same constructor arguments, method
names, dict keys and ``state_dict`` names, so reference checkpoints load unchanged).

``RSSM.observe`` / ``imagine_with_action`` / ``obs_step`` / ``img_step`` / ``kl_loss`` run on
the sm_100a kernels of libdv3_b200.so; there is no PyTorch fallback for them.  The batched
encoder / decoder / head MLPs and the 4x4 stride-2 conv / transposed-conv stacks of the image
encoder / decoder run on the same tcgen05 GEMM + row kernels (CPU tensors and other conv
geometries fall back to torch ops).

Reference lines: RSSM networks.py:13-290, GRUCell 742-768, MLP 588-739, MultiEncoder /
MultiDecoder 293-441, ConvEncoder / ConvDecoder 444-585, Conv2dSamePad / ImgChLayerNorm 771-810.
"""
from __future__ import annotations

import math
import re

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from . import kernels as K
from . import tools

LN_EPS = 1e-3

# fp32 is the contract of this path (precision: 32, parity to 1e-4 with bit-exact categorical
# draws).  cuDNN convolutions default to TF32 on sm_80+ (torch.backends.cudnn.allow_tf32 = True),
# which puts ~1e-3 of error into the conv encoder's embedding -- enough to flip posterior draws
# against the reference's fp32 CPU path (tests/test_gpu_reference.py found 15 flips per 16x64
# batch).  Forward and backward convolutions both read this process-wide flag.
torch.backends.cudnn.allow_tf32 = False


class GRUCell(nn.Module):
    """Parameter holder for the LayerNorm GRU (names match reference networks.py:742-758); the
    arithmetic lives in the kernels (dv3_gru_gates_*)."""

    def __init__(self, inp_size, size, norm=True, act=torch.tanh, update_bias=-1):
        super().__init__()
        if not norm or update_bias != -1:
            raise NotImplementedError("kernel path implements norm=True, update_bias=-1")
        self._size = size
        self.layers = nn.Sequential()
        self.layers.add_module("GRU_linear", nn.Linear(inp_size + size, 3 * size, bias=False))
        self.layers.add_module("GRU_norm", nn.LayerNorm(3 * size, eps=LN_EPS))


def _dense_block(inp, out):
    seq = nn.Sequential(nn.Linear(inp, out, bias=False), nn.LayerNorm(out, eps=LN_EPS), nn.SiLU())
    seq.apply(tools.weight_init)
    return seq


class RSSM(nn.Module):
    def __init__(self, stoch=30, deter=200, hidden=200, rec_depth=1, discrete=False, act="SiLU",
                 norm=True, mean_act="none", std_act="softplus", min_std=0.1, unimix_ratio=0.01,
                 initial="learned", num_actions=None, embed=None, device=None):
        super().__init__()
        if not discrete:
            raise NotImplementedError("kernel path implements discrete latents (dyn_discrete > 0), "
                                      "the only mode the reference configs use")
        if discrete > 32:
            raise NotImplementedError("dyn_discrete <= 32 (one class per lane)")
        if rec_depth != 1 or act != "SiLU" or not norm or initial != "learned":
            raise NotImplementedError("kernel path implements rec_depth=1, SiLU, norm=True, "
                                      "initial='learned'")
        self._stoch, self._deter, self._hidden = stoch, deter, hidden
        self._discrete, self._unimix_ratio = discrete, unimix_ratio
        self._num_actions, self._embed, self._device = num_actions, embed, device
        flat = stoch * discrete
        self._img_in_layers = _dense_block(flat + num_actions, hidden)
        self._cell = GRUCell(hidden, deter)
        self._cell.apply(tools.weight_init)
        self._img_out_layers = _dense_block(deter, hidden)
        self._obs_out_layers = _dense_block(deter + embed, hidden)
        self._imgs_stat_layer = nn.Linear(hidden, flat)
        self._imgs_stat_layer.apply(tools.uniform_weight_init(1.0))
        self._obs_stat_layer = nn.Linear(hidden, flat)
        self._obs_stat_layer.apply(tools.uniform_weight_init(1.0))
        self.W = nn.Parameter(torch.zeros((1, deter), device=device), requires_grad=True)

    # ---- kernel plumbing -------------------------------------------------------------
    @property
    def dims(self):
        return (self._stoch, self._discrete, self._deter, self._hidden, self._num_actions,
                self._embed, float(self._unimix_ratio))

    def kernel_params(self):
        named = dict(self.named_parameters())
        return [named[L.RSSM_STATE_KEYS[f]] for f in L.RSSM_PARAM_FIELDS]

    def _uniforms(self, *shape):
        return torch.rand(*shape, self._stoch, self._discrete, device=self.W.device)

    @staticmethod
    def _to_idx(stoch):
        """Class indices of a one-hot stoch tensor.  Tensors produced by ``observe`` carry the
        int32 indices the kernel sampled (``tag_idx``), so no argmax pass is needed for them."""
        hit = getattr(stoch, "_dv3_idx", None)
        if hit is not None and hit[0] == stoch._version and hit[1].shape == stoch.shape[:-1]:
            return hit[1]
        return torch.argmax(stoch, -1).to(torch.int32).contiguous()

    @staticmethod
    def tag_idx(stoch, idx):
        stoch._dv3_idx = (stoch._version, idx)
        return stoch

    # ---- reference API ---------------------------------------------------------------
    def initial(self, batch_size):
        """networks.py:99-125: deter = tanh(W) tiled, stoch = mode(prior(deter)), logit = 0."""
        deter = torch.tanh(self.W).repeat(batch_size, 1)
        stoch = self.get_stoch(deter)
        logit = torch.zeros(batch_size, self._stoch, self._discrete, device=deter.device)
        return dict(logit=logit, stoch=stoch, deter=deter)

    def get_stoch(self, deter):
        y = self._img_out_layers(deter)
        logit = self._imgs_stat_layer(y).reshape(list(y.shape[:-1]) + [self._stoch, self._discrete])
        return self.get_dist({"logit": logit}).mode()

    def observe(self, embed, action, is_first, state=None, noise=None):
        """embed [B,T,E], action [B,T,A], is_first [B,T] -> (post, prior) dicts of batch-major
        stoch / deter / logit.  ``noise`` = (u_prior, u_post), time-major [T,B,S,C] uniforms; drawn
        from torch's generator when omitted.  Like the reference (networks.py:184) the rows of
        ``action`` at is_first positions are zeroed in place."""
        B, T = embed.shape[:2]
        if noise is None:
            noise = (self._uniforms(T, B), self._uniforms(T, B))
        sidx = sdet = None
        if state is not None:
            sidx, sdet = self._to_idx(state["stoch"]), state["deter"]
        (post_stoch, post_logit, prior_stoch, prior_logit, deter, aprev, post_idx, prior_idx) = \
            K.observe(embed, action, is_first, noise[0], noise[1], sidx, sdet, self.dims,
                      self.kernel_params())
        self.tag_idx(post_stoch, post_idx)
        self.tag_idx(prior_stoch, prior_idx)
        if action.shape == aprev.shape and not action.requires_grad:
            action.copy_(aprev)
        post = dict(stoch=post_stoch, deter=deter, logit=post_logit)
        prior = dict(stoch=prior_stoch, deter=deter, logit=prior_logit)
        return post, prior

    def imagine_with_action(self, action, state, noise=None):
        """action [B,T',A], state dict [B,...] -> prior dict [B,T',...] (networks.py:145-152)."""
        B, Tn = action.shape[:2]
        H = Tn + 1
        if noise is None:
            noise = self._uniforms(H, B)
        given = action.permute(1, 0, 2).contiguous()
        feat, logit, _, _ = K.imagine(self._to_idx(state["stoch"]), state["deter"], None, noise,
                                      given, H, self.dims, None, self.kernel_params(), [])
        SC = self._stoch * self._discrete
        stoch = feat[1:, :, :SC].reshape(Tn, B, self._stoch, self._discrete)
        swap = lambda x: x.permute([1, 0] + list(range(2, x.dim())))
        return dict(stoch=swap(stoch), deter=swap(feat[1:, :, SC:]), logit=swap(logit[1:]))

    def _mode_of(self, logit):
        """one-hot mode of the unimix categorical (tools.py:448-450), value only."""
        lead = logit.shape[:-2]
        lg = logit.detach().reshape(-1, self._stoch, self._discrete).contiguous().float()
        _, hot = K.onehot_sample(lg, None, float(self._unimix_ratio))
        return hot.reshape(tuple(lead) + (self._stoch, self._discrete))

    def obs_step(self, prev_state, prev_action, embed, is_first, sample=True, noise=None):
        """networks.py:174-206.  ``prev_state is None`` (the acting path's first call,
        dreamer.py:117-127) starts every row from ``initial()`` with a zero action, exactly like
        rows whose ``is_first`` is set; ``sample=False`` returns the posterior mode (the prior draw
        inside ``img_step`` is still a sample, as in the reference).  ``noise`` = (u_prior, u_post)
        uniforms [n,S,C]."""
        n = embed.shape[0]
        first = is_first.reshape(n, 1).to(torch.float32)
        if prev_state is None:
            prev_action = torch.zeros(n, self._num_actions, device=embed.device)
        elif prev_action is None:
            raise L.Dv3Error("obs_step: prev_action is None but prev_state is given")
        if noise is not None:
            noise = (noise[0].reshape(1, n, self._stoch, self._discrete),
                     noise[1].reshape(1, n, self._stoch, self._discrete))
        post, prior = self.observe(embed[:, None], prev_action[:, None].to(torch.float32).clone(),
                                   first, prev_state, noise)
        post = {k: v[:, 0] for k, v in post.items()}
        prior = {k: v[:, 0] for k, v in prior.items()}
        if not sample:
            post["stoch"] = self._mode_of(post["logit"])
        return post, prior

    def img_step(self, prev_state, prev_action, sample=True, noise=None):
        """networks.py:208-233; ``noise`` = uniforms [n,S,C] for the prior draw."""
        n = prev_action.shape[0]
        if noise is not None:
            noise = torch.cat([noise.reshape(1, n, self._stoch, self._discrete)] * 2, 0)
        out = self.imagine_with_action(prev_action[:, None], prev_state, noise)
        out = {k: v[:, 0] for k, v in out.items()}
        if not sample:
            out["stoch"] = self._mode_of(out["logit"])
        return out

    def get_feat(self, state):
        """cat(flat(stoch), deter) (reference networks.py:154-159).  The imagination kernel already
        lays its states out as [one-hot stoch | deter] rows of one buffer; when ``stoch`` and
        ``deter`` are the two column ranges of that buffer it is returned as is (no copy)."""
        s = state["stoch"]
        base = getattr(s, "_dv3_feat", None)
        if base is not None and getattr(state["deter"], "_dv3_feat", None) is base \
                and base._version == s._dv3_feat_version:
            return base
        return torch.cat([s.reshape(list(s.shape[:-2]) + [self._stoch * self._discrete]),
                          state["deter"]], -1)

    def get_dist(self, state, dtype=None):
        return tools.IndependentOneHot(state["logit"], self._unimix_ratio)

    def kl_loss(self, post, prior, free, dyn_scale, rep_scale):
        """networks.py:272-290 -> loss, value, dyn_loss, rep_loss (each [B,T])."""
        loss, value, dyn, rep, _, _ = K.kl_balance(post["logit"], prior["logit"], free, dyn_scale,
                                                   rep_scale, self._unimix_ratio)
        return loss, value, dyn, rep

    def kl_loss_with_entropy(self, post, prior, free, dyn_scale, rep_scale):
        """Same kernel call; also returns the posterior / prior entropies the reference computes
        separately for its metrics (models.py:156-168)."""
        return K.kl_balance(post["logit"], prior["logit"], free, dyn_scale, rep_scale,
                            self._unimix_ratio)


# --------------------------------------------------------------------------------------
# MLP heads / encoder / decoder (torch ops; "next" rows of the scope table)
# --------------------------------------------------------------------------------------
class MLP(nn.Module):
    def __init__(self, inp_dim, shape, layers, units, act="SiLU", norm=True, dist="normal", std=1.0,
                 min_std=0.1, max_std=1.0, absmax=None, temp=0.1, unimix_ratio=0.01, outscale=1.0,
                 symlog_inputs=False, device=None, name="NoName"):
        super().__init__()
        if act != "SiLU" or not norm:
            raise NotImplementedError("act=SiLU, norm=True")
        self._shape = (shape,) if isinstance(shape, int) else shape
        if self._shape is not None and not isinstance(self._shape, dict) and len(self._shape) == 0:
            self._shape = (1,)
        self._dist, self._std = dist, std
        self._min_std, self._max_std, self._absmax = min_std, max_std, absmax
        self._unimix_ratio, self._symlog_inputs = unimix_ratio, symlog_inputs
        self._name, self._layers, self._units = name, layers, units
        self.layers = nn.Sequential()
        d = inp_dim
        for i in range(layers):
            self.layers.add_module(f"{name}_linear{i}", nn.Linear(d, units, bias=False))
            self.layers.add_module(f"{name}_norm{i}", nn.LayerNorm(units, eps=LN_EPS))
            self.layers.add_module(f"{name}_act{i}", nn.SiLU())
            d = units
        self.layers.apply(tools.weight_init)
        if isinstance(self._shape, dict):
            self.mean_layer = nn.ModuleDict({k: nn.Linear(d, int(np.prod(v)))
                                             for k, v in self._shape.items()})
            self.mean_layer.apply(tools.uniform_weight_init(outscale))
            if std == "learned":
                raise NotImplementedError("learned std with dict outputs")
        elif self._shape is not None:
            self.mean_layer = nn.Linear(d, int(np.prod(self._shape)))
            self.mean_layer.apply(tools.uniform_weight_init(outscale))
            if std == "learned":
                self.std_layer = nn.Linear(units, int(np.prod(self._shape)))
                self.std_layer.apply(tools.uniform_weight_init(outscale))

    def trunk(self, features):
        """[Linear(no bias) -> LayerNorm -> SiLU] x layers (networks.py:657-661) on the tensor-core
        GEMM + LN/SiLU row kernels."""
        if not features.is_cuda:
            raise L.Dv3Error("MLP forward needs CUDA tensors: the B200 path has no CPU fallback")
        x = features
        if self._symlog_inputs:
            x = tools.symlog(features) if features.requires_grad else K.symlog(features)
        for i in range(self._layers):
            lin = getattr(self.layers, f"{self._name}_linear{i}")
            nrm = getattr(self.layers, f"{self._name}_norm{i}")
            x = K.dense_ln_silu(x, lin.weight, nrm.weight, nrm.bias)
        return x

    @staticmethod
    def _head(layer, x):
        return K.linear_bias(x, layer.weight, layer.bias)

    def forward(self, features, dtype=None):
        out = self.trunk(features)
        if self._shape is None:
            return out
        if isinstance(self._shape, dict):
            return {k: self.dist(self._dist, self._head(self.mean_layer[k], out), self._std, shp)
                    for k, shp in self._shape.items()}
        std = self._head(self.std_layer, out) if self._std == "learned" else self._std
        return self.dist(self._dist, self._head(self.mean_layer, out), std, self._shape)

    def dist(self, dist, mean, std, shape):
        if dist == "normal":
            return tools.NormalTanhMean(mean, std, self._min_std, self._max_std, self._absmax)
        if dist == "onehot":
            return tools.OneHotDist(mean, unimix_ratio=self._unimix_ratio)
        if dist == "binary":
            return tools.Bernoulli(mean)
        if dist == "symlog_disc":
            return tools.DiscDist(logits=mean)
        if dist == "symlog_mse":
            return tools.SymlogDist(mean)
        raise NotImplementedError(dist)

    # flat parameter view for the imagination kernel
    def actor_spec(self):
        if self._dist not in ("normal", "onehot"):
            raise NotImplementedError(f"actor dist {self._dist}")
        if self._dist == "normal" and self._std != "learned":
            raise NotImplementedError("normal actor needs std='learned'")
        return K.ActorSpec(self._layers, self._units, self._dist, self._min_std, self._max_std,
                           float(self._unimix_ratio))

    def actor_params(self):
        ps = []
        for i in range(self._layers):
            lin = getattr(self.layers, f"{self._name}_linear{i}")
            nrm = getattr(self.layers, f"{self._name}_norm{i}")
            ps += [lin.weight, nrm.weight, nrm.bias]
        ps += [self.mean_layer.weight, self.mean_layer.bias]
        if self._dist == "normal":
            ps += [self.std_layer.weight, self.std_layer.bias]
        return ps


class Conv2dSamePad(nn.Conv2d):
    def forward(self, x):
        ih, iw = x.shape[-2:]
        pads = []
        for i, k, s, dl in ((iw, self.kernel_size[1], self.stride[1], self.dilation[1]),
                            (ih, self.kernel_size[0], self.stride[0], self.dilation[0])):
            p = max((math.ceil(i / s) - 1) * s + (k - 1) * dl + 1 - i, 0)
            pads += [p // 2, p - p // 2]
        if any(pads):
            x = F.pad(x, pads)
        return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation,
                        self.groups)


class ImgChLayerNorm(nn.Module):
    def __init__(self, ch, eps=LN_EPS):
        super().__init__()
        self.norm = nn.LayerNorm(ch, eps=eps)

    def forward(self, x):
        return self.norm(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class ConvEncoder(nn.Module):
    def __init__(self, input_shape, depth=32, act="SiLU", norm=True, kernel_size=4, minres=4):
        super().__init__()
        h, w, ch = input_shape
        stages = int(np.log2(h) - np.log2(minres))
        mods, cin, cout = [], ch, depth
        for _ in range(stages):
            mods.append(Conv2dSamePad(cin, cout, kernel_size, stride=2, bias=False))
            if norm:
                mods.append(ImgChLayerNorm(cout))
            mods.append(getattr(nn, act)())
            cin, cout = cout, cout * 2
            h, w = h // 2, w // 2
        self.outdim = cin * h * w
        self.layers = nn.Sequential(*mods)
        self.layers.apply(tools.weight_init)

    def _blocks(self):
        """[(conv, norm)] when the stack is the one the kernels implement: 4x4 stride-2 'same'
        convolutions without bias, each followed by channel LayerNorm and SiLU."""
        mods = list(self.layers)
        if len(mods) % 3:
            return None
        out = []
        for i in range(0, len(mods), 3):
            conv, norm, act = mods[i:i + 3]
            if not (isinstance(conv, Conv2dSamePad) and isinstance(norm, ImgChLayerNorm)
                    and isinstance(act, nn.SiLU) and conv.kernel_size == (4, 4)
                    and conv.stride == (2, 2) and conv.bias is None and conv.dilation == (1, 1)):
                return None
            out.append((conv, norm))
        return out

    def forward(self, obs):
        lead = obs.shape[:-3]
        h, w, c = obs.shape[-3:]
        if not obs.is_cuda:
            raise L.Dv3Error("ConvEncoder forward needs CUDA tensors: the B200 path has no CPU fallback")
        blocks = self._blocks()
        if blocks is not None and h % (1 << len(blocks)) == 0 and w % (1 << len(blocks)) == 0:
            # channels-last rows [n*h*w, c]: every stage is im2col -> tcgen05 GEMM -> LN/SiLU rows
            n = int(np.prod(lead)) if len(lead) else 1
            x = (obs - 0.5).reshape(n * h * w, c)
            for i, (conv, norm) in enumerate(blocks):
                x = K.conv_ln_silu(x, (n, h, w), conv.weight, norm.norm.weight, norm.norm.bias,
                                   need_dx=i > 0)
                h, w = h // 2, w // 2
            # the reference flattens NCHW: (c, y, x) order
            x = x.view(n, h * w, x.shape[1]).transpose(1, 2)
            return x.reshape(list(lead) + [-1])
        x = (obs - 0.5).reshape((-1,) + tuple(obs.shape[-3:])).permute(0, 3, 1, 2)
        x = self.layers(x)
        return x.reshape(list(lead) + [-1])


class ConvDecoder(nn.Module):
    def __init__(self, feat_size, shape=(3, 64, 64), depth=32, act="SiLU", norm=True, kernel_size=4,
                 minres=4, outscale=1.0, cnn_sigmoid=False):
        super().__init__()
        self._shape, self._minres, self._cnn_sigmoid = tuple(shape), minres, cnn_sigmoid
        n = int(np.log2(shape[1]) - np.log2(minres))
        self._embed_size = minres ** 2 * depth * 2 ** (n - 1)
        self._linear_layer = nn.Linear(feat_size, self._embed_size)
        self._linear_layer.apply(tools.uniform_weight_init(outscale))
        val = (kernel_size - 1) - 2 + 1
        pad = math.ceil(val / 2)
        outpad = pad * 2 - val
        mods = []
        cin = self._embed_size // (minres ** 2)
        for i in range(n):
            last = i == n - 1
            cout = self._shape[0] if last else cin // 2
            mods.append(nn.ConvTranspose2d(cin, cout, kernel_size, 2, padding=(pad, pad),
                                           output_padding=(outpad, outpad), bias=last))
            if not last:
                if norm:
                    mods.append(ImgChLayerNorm(cout))
                mods.append(getattr(nn, act)())
            cin = cout
        for m in mods[:-1]:
            m.apply(tools.weight_init)
        mods[-1].apply(tools.uniform_weight_init(outscale))
        self.layers = nn.Sequential(*mods)

    def _blocks(self):
        """[(deconv, norm | None)] when the stack is the one the kernels implement: 4x4 stride-2
        padding-1 transposed convolutions, channel LayerNorm + SiLU between them, bias on the last."""
        mods = list(self.layers)
        out, i = [], 0
        while i < len(mods):
            dc = mods[i]
            if not (isinstance(dc, nn.ConvTranspose2d) and dc.kernel_size == (4, 4)
                    and dc.stride == (2, 2) and dc.padding == (1, 1) and dc.output_padding == (0, 0)
                    and dc.dilation == (1, 1) and dc.groups == 1):
                return None
            if i + 2 < len(mods) and isinstance(mods[i + 1], ImgChLayerNorm) \
                    and isinstance(mods[i + 2], nn.SiLU) and dc.bias is None:
                out.append((dc, mods[i + 1]))
                i += 3
            elif i == len(mods) - 1:
                out.append((dc, None))
                i += 1
            else:
                return None
        return out

    def forward(self, features, dtype=None):
        if not features.is_cuda:
            raise L.Dv3Error("ConvDecoder forward needs CUDA tensors: the B200 path has no CPU fallback")
        blocks = self._blocks()
        if blocks is not None:
            # the Linear output viewed [n, minres, minres, C] is already channels-last rows
            x = K.linear_bias(features, self._linear_layer.weight, self._linear_layer.bias)
            ch = self._embed_size // self._minres ** 2
            n = x.numel() // self._embed_size
            h = w = self._minres
            x = x.reshape(n * h * w, ch)
            for dc, norm in blocks:
                if norm is not None:
                    x = K.deconv_block(x, (n, h, w), dc.weight, norm.norm.weight, norm.norm.bias)
                else:
                    x = K.deconv_block(x, (n, h, w), dc.weight, bias=dc.bias,
                                       shift=0.0 if self._cnn_sigmoid else 0.5)
                h, w = 2 * h, 2 * w
            # rows are (n, y, x) with the channels last: the reference's permuted output
            mean = x.reshape(tuple(features.shape[:-1]) + (h, w, self._shape[0]))
            return torch.sigmoid(mean) if self._cnn_sigmoid else mean
        x = self._linear_layer(features)
        x = x.reshape(-1, self._minres, self._minres, self._embed_size // self._minres ** 2)
        x = self.layers(x.permute(0, 3, 1, 2))
        mean = x.reshape(tuple(features.shape[:-1]) + self._shape).permute(0, 1, 3, 4, 2)
        return torch.sigmoid(mean) if self._cnn_sigmoid else mean + 0.5


def _split_shapes(shapes, mlp_keys, cnn_keys, excluded):
    shapes = {k: tuple(v) for k, v in shapes.items()
              if k not in excluded and not k.startswith("log_")}
    cnn = {k: v for k, v in shapes.items() if len(v) == 3 and re.match(cnn_keys, k)}
    mlp = {k: v for k, v in shapes.items() if len(v) in (1, 2) and re.match(mlp_keys, k)}
    return cnn, mlp


class MultiEncoder(nn.Module):
    def __init__(self, shapes, mlp_keys, cnn_keys, act, norm, cnn_depth, kernel_size, minres,
                 mlp_layers, mlp_units, symlog_inputs):
        super().__init__()
        self.cnn_shapes, self.mlp_shapes = _split_shapes(
            shapes, mlp_keys, cnn_keys, ("is_first", "is_last", "is_terminal", "reward"))
        self.outdim = 0
        if self.cnn_shapes:
            ch = sum(v[-1] for v in self.cnn_shapes.values())
            hw = tuple(self.cnn_shapes.values())[0][:2]
            self._cnn = ConvEncoder(hw + (ch,), cnn_depth, act, norm, kernel_size, minres)
            self.outdim += self._cnn.outdim
        if self.mlp_shapes:
            size = sum(sum(v) for v in self.mlp_shapes.values())
            self._mlp = MLP(size, None, mlp_layers, mlp_units, act, norm,
                            symlog_inputs=symlog_inputs, name="Encoder")
            self.outdim += mlp_units

    def forward(self, obs):
        outs = []
        if self.cnn_shapes:
            outs.append(self._cnn(torch.cat([obs[k] for k in self.cnn_shapes], -1)))
        if self.mlp_shapes:
            outs.append(self._mlp(torch.cat([obs[k] for k in self.mlp_shapes], -1)))
        return torch.cat(outs, -1)


class MultiDecoder(nn.Module):
    def __init__(self, feat_size, shapes, mlp_keys, cnn_keys, act, norm, cnn_depth, kernel_size,
                 minres, mlp_layers, mlp_units, cnn_sigmoid, image_dist, vector_dist, outscale):
        super().__init__()
        self.cnn_shapes, self.mlp_shapes = _split_shapes(
            shapes, mlp_keys, cnn_keys, ("is_first", "is_last", "is_terminal"))
        if image_dist != "mse":
            raise NotImplementedError(image_dist)
        if self.cnn_shapes:
            some = list(self.cnn_shapes.values())[0]
            shape = (sum(v[-1] for v in self.cnn_shapes.values()),) + some[:-1]
            self._cnn = ConvDecoder(feat_size, shape, cnn_depth, act, norm, kernel_size, minres,
                                    outscale=outscale, cnn_sigmoid=cnn_sigmoid)
        if self.mlp_shapes:
            self._mlp = MLP(feat_size, self.mlp_shapes, mlp_layers, mlp_units, act, norm,
                            vector_dist, outscale=outscale, name="Decoder")

    def forward(self, features):
        dists = {}
        if self.cnn_shapes:
            out = self._cnn(features)
            parts = torch.split(out, [v[-1] for v in self.cnn_shapes.values()], -1)
            dists.update({k: tools.MSEDist(p) for k, p in zip(self.cnn_shapes, parts)})
        if self.mlp_shapes:
            dists.update(self._mlp(features))
        return dists
