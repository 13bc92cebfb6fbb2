"""ctypes binding of libdv3_b200.so -- the C ABI declared in include/dv3_b200.h.

There is no CPU fallback: if the shared library is missing or a call fails, this module raises.
Structures mirror the header field for field (tests/test_abi.py parses the header and checks).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdv3_b200.so")

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int32)
_pf = C.POINTER(_f)
_v = C.c_void_p


def _fields(spec):
    kinds = {"f": _f, "i": _i, "pf": _pf, "v": _v, "i32": C.c_int32, "f32": C.c_float,
             "sz": C.c_size_t}
    return [(name, kinds[kind]) for name, kind in spec]


class RssmDims(C.Structure):
    _fields_ = _fields([("stoch", "i32"), ("classes", "i32"), ("deter", "i32"), ("hidden", "i32"),
                        ("actions", "i32"), ("embed", "i32"), ("unimix", "f32"), ("ln_eps", "f32")])


RSSM_PARAM_FIELDS = ["w_in", "ln_in_g", "ln_in_b", "w_gru", "ln_gru_g", "ln_gru_b", "w_out",
                     "ln_out_g", "ln_out_b", "w_ims", "b_ims", "w_obs", "ln_obs_g", "ln_obs_b",
                     "w_os", "b_os", "w_init"]

# C field -> reference state_dict key of networks.RSSM
RSSM_STATE_KEYS = {
    "w_in": "_img_in_layers.0.weight", "ln_in_g": "_img_in_layers.1.weight",
    "ln_in_b": "_img_in_layers.1.bias", "w_gru": "_cell.layers.GRU_linear.weight",
    "ln_gru_g": "_cell.layers.GRU_norm.weight", "ln_gru_b": "_cell.layers.GRU_norm.bias",
    "w_out": "_img_out_layers.0.weight", "ln_out_g": "_img_out_layers.1.weight",
    "ln_out_b": "_img_out_layers.1.bias", "w_ims": "_imgs_stat_layer.weight",
    "b_ims": "_imgs_stat_layer.bias", "w_obs": "_obs_out_layers.0.weight",
    "ln_obs_g": "_obs_out_layers.1.weight", "ln_obs_b": "_obs_out_layers.1.bias",
    "w_os": "_obs_stat_layer.weight", "b_os": "_obs_stat_layer.bias", "w_init": "W",
}


class TcOperand(C.Structure):
    _fields_ = _fields([("hi", "f"), ("lo", "f"), ("ld", "i32"), ("mn_major", "i32")])


class RssmPlanes(C.Structure):
    """dv3_rssm_planes: caller-derived forms of the RSSM weights (made once per optimizer step)."""
    _fields_ = [("w_gru", TcOperand), ("w_out", TcOperand), ("w_ims", TcOperand), ("w_in_t", _f),
                ("w_in_t_sp", TcOperand)]


class RssmParams(C.Structure):
    _fields_ = _fields([(n, "f") for n in RSSM_PARAM_FIELDS]) + [("planes", C.POINTER(RssmPlanes))]


class ActorPlanes(C.Structure):
    """dv3_actor_planes"""
    _fields_ = [("w", TcOperand * 16), ("w0_t", _f)]


class ObserveIO(C.Structure):
    _fields_ = _fields([
        ("B", "i32"), ("T", "i32"),
        ("embed", "f"), ("action", "f"), ("is_first", "f"), ("u_prior", "f"), ("u_post", "f"),
        ("state_idx", "i"), ("state_deter", "f"),
        ("post_stoch", "f"), ("post_logit", "f"), ("prior_stoch", "f"), ("prior_logit", "f"),
        ("deter", "f"),
        ("post_idx", "i"), ("prior_idx", "i"), ("first_eff", "f"), ("sprev_idx", "i"),
        ("hprev", "f"), ("aprev", "f"), ("x_pre", "f"), ("x", "f"), ("g_pre", "f"),
        ("y_pre", "f"), ("y", "f"), ("z_pre", "f"), ("z", "f"),
        ("init_deter", "f"), ("init_ypre", "f"), ("init_y", "f"), ("init_logit", "f"),
        ("init_idx", "i"),
        ("workspace", "v"), ("workspace_bytes", "sz")])


class ObserveBwdIO(C.Structure):
    _fields_ = _fields([
        ("B", "i32"), ("T", "i32"),
        ("first_eff", "f"), ("post_logit", "f"), ("prior_logit", "f"), ("hprev", "f"),
        ("x_pre", "f"), ("g_pre", "f"), ("y_pre", "f"), ("z_pre", "f"),
        ("g_post_stoch", "f"), ("g_post_logit", "f"), ("g_prior_stoch", "f"),
        ("g_prior_logit", "f"), ("g_deter", "f"),
        ("d_embed", "f"), ("d_x_pre", "f"), ("d_x_ln", "f"), ("d_g_pre", "f"), ("d_g_ln", "f"),
        ("d_y_pre", "f"), ("d_y_ln", "f"), ("d_z_pre", "f"), ("d_z_ln", "f"),
        ("d_post_logit", "f"), ("d_prior_logit", "f"), ("d_init_stoch", "f"),
        ("d_init_deter", "f"), ("d_state_deter", "f"), ("d_state_stoch", "f"),
        ("workspace", "v"), ("workspace_bytes", "sz")])


class Actor(C.Structure):
    _fields_ = _fields([
        ("layers", "i32"), ("units", "i32"), ("dist", "i32"), ("min_std", "f32"),
        ("max_std", "f32"), ("unimix", "f32"), ("w", "pf"), ("ln_g", "pf"), ("ln_b", "pf"),
        ("w_mean", "f"), ("b_mean", "f"), ("w_std", "f"), ("b_std", "f")]) + [
        ("planes", C.POINTER(ActorPlanes))]


class ImagineIO(C.Structure):
    _fields_ = _fields([
        ("N", "i32"), ("H", "i32"),
        ("start_idx", "i"), ("start_deter", "f"), ("act_noise", "f"), ("u_state", "f"),
        ("given_action", "f"),
        ("feat", "f"), ("logit", "f"), ("action", "f"), ("idx", "i"),
        ("x_pre", "f"), ("x", "f"), ("g_pre", "f"), ("y_pre", "f"), ("y", "f"),
        ("a_pre", "f"), ("a_act", "f"), ("a_mean_raw", "f"), ("a_std_raw", "f"),
        ("workspace", "v"), ("workspace_bytes", "sz")])


class ImagineBwdIO(C.Structure):
    _fields_ = _fields([
        ("N", "i32"), ("H", "i32"),
        ("logit", "f"), ("feat", "f"), ("x_pre", "f"), ("g_pre", "f"), ("y_pre", "f"),
        ("a_mean_raw", "f"), ("a_std_raw", "f"), ("act_noise", "f"),
        ("g_stoch", "f"), ("g_deter", "f"), ("g_logit", "f"), ("g_action", "f"),
        ("d_mean_raw", "f"), ("d_std_raw", "f"), ("d_x_pre", "f"), ("d_x_ln", "f"),
        ("d_g_pre", "f"), ("d_g_ln", "f"), ("d_y_pre", "f"), ("d_y_ln", "f"), ("d_logit", "f"),
        ("d_start_stoch", "f"), ("d_start_deter", "f"),
        ("workspace", "v"), ("workspace_bytes", "sz"), ("g_state_ld", "i32")])


STRUCTS = {"dv3_tc_operand": TcOperand, "dv3_rssm_dims": RssmDims, "dv3_rssm_params": RssmParams,
           "dv3_rssm_planes": RssmPlanes, "dv3_actor_planes": ActorPlanes, "dv3_observe_io": ObserveIO,
           "dv3_observe_bwd_io": ObserveBwdIO, "dv3_actor": Actor, "dv3_imagine_io": ImagineIO,
           "dv3_imagine_bwd_io": ImagineBwdIO}

_P = C.POINTER
_i32, _f32, _dbl = C.c_int32, C.c_float, C.c_double

# name -> (restype, argtypes); every function include/dv3_b200.h declares
SIGNATURES = {
    "dv3_version": (C.c_int, []),
    "dv3_last_error": (C.c_char_p, []),
    "dv3_device_arch": (C.c_int, []),
    "dv3_reload_env": (None, []),
    "dv3_launch_count": (C.c_longlong, []),
    "dv3_prof_enable": (None, [C.c_int]),
    "dv3_prof_read": (C.c_int, [_P(C.c_double), _P(C.c_double), _P(C.c_longlong)]),
    "dv3_observe_workspace_bytes": (C.c_size_t, [_P(RssmDims), _i32, _i32]),
    "dv3_observe_fwd": (C.c_int, [_P(RssmDims), _P(RssmParams), _P(ObserveIO), _v]),
    "dv3_obs_step_fwd": (C.c_int, [_P(RssmDims), _P(RssmParams), _P(ObserveIO), _v]),
    "dv3_observe_bwd_workspace_bytes": (C.c_size_t, [_P(RssmDims), _i32, _i32]),
    "dv3_observe_bwd": (C.c_int, [_P(RssmDims), _P(RssmParams), _P(ObserveBwdIO), _v]),
    "dv3_imagine_workspace_bytes": (C.c_size_t, [_P(RssmDims), _P(Actor), _i32, _i32]),
    "dv3_imagine_fwd": (C.c_int, [_P(RssmDims), _P(RssmParams), _P(Actor), _P(ImagineIO), _v]),
    "dv3_img_step_fwd": (C.c_int, [_P(RssmDims), _P(RssmParams), _P(ImagineIO), _v]),
    "dv3_imagine_bwd_workspace_bytes": (C.c_size_t, [_P(RssmDims), _P(Actor), _i32, _i32]),
    "dv3_imagine_bwd": (C.c_int, [_P(RssmDims), _P(RssmParams), _P(Actor), _P(ImagineBwdIO), _v]),
    "dv3_lambda_return_fwd": (C.c_int, [_f, _f, _f, _f, _dbl, _i32, _i32, _f, _v]),
    "dv3_lambda_return_bwd": (C.c_int, [_f, _f, _f, _f, _f, _dbl, _i32, _i32, _f, _f, _f, _f, _v]),
    "dv3_twohot_logprob_fwd": (C.c_int, [_f, _f, _f, _i32, _i32, _f, _v]),
    "dv3_twohot_logprob_bwd": (C.c_int, [_f, _f, _f, _f, _i32, _i32, _f, _v]),
    "dv3_twohot_mean_fwd": (C.c_int, [_f, _f, _i32, _i32, _f, _v]),
    "dv3_twohot_mean_bwd": (C.c_int, [_f, _f, _f, _i32, _i32, _f, _v]),
    "dv3_kl_balance_fwd": (C.c_int, [_f, _f, _i32, _i32, _i32, _f32, _f32, _f32, _f32,
                                     _f, _f, _f, _f, _f, _f, _v]),
    "dv3_kl_balance_bwd": (C.c_int, [_f, _f, _f, _i32, _i32, _i32, _f32, _f32, _f32, _f32,
                                     _f, _f, _v]),
    "dv3_linear_fwd": (C.c_int, [_f, _i32, _f, _i32, _i32, _f, _i32, _f, _i32, _i32, _f, _f, _i32,
                                 _f, _i32, _i32, _i32, _i32, _v]),
    "dv3_linear_tc_scratch_bytes": (C.c_size_t, [_i32, _i32, _i32]),
    "dv3_linear_tc_fwd": (C.c_int, [_f, _i32, _i32, _f, _i32, _i32, _f, _f, _i32, _f, _i32, _i32,
                                    _i32, _i32, _v, C.c_size_t, _v]),
    "dv3_gemm_tc": (C.c_int, [_P(TcOperand), _i32, _P(TcOperand), _i32, _P(TcOperand), _f, _f, _i32,
                              _f, _i32, _i32, _i32, _i32, _v]),
    "dv3_gemm_tc_rawa": (C.c_int, [_f, _i32, _i32, _f, _i32, _i32, _P(TcOperand), _f, _f, _i32, _f, _i32,
                                   _i32, _i32, _v]),
    "dv3_split_tf32": (C.c_int, [_f, _i32, _i32, _i32, _f, _f, _i32, _v]),
    "dv3_linear_tc2_fwd": (C.c_int, [_f, _i32, _i32, _f, _i32, _i32, _f, _i32, _f, _f, _i32, _f,
                                     _i32, _i32, _i32, _i32, _v]),
    "dv3_transpose": (C.c_int, [_f, _i32, _i32, _i32, _f, _v]),
    "dv3_ln_silu_fwd": (C.c_int, [_f, _i32, _f, _f, _f32, _i32, _i32, _f, _i32, _v]),
    "dv3_ln_silu_bwd": (C.c_int, [_f, _i32, _f, _f, _f32, _f, _i32, _i32, _i32, _f, _f, _i32, _v]),
    "dv3_adam_clip_step": (C.c_int, [_f, _f, _f, _f, C.c_longlong, _f32, _f32, _f32, _f32, _f32, _f32,
                                     _f, _f, _f, _v]),
    "dv3_adam_clip_step_planes": (C.c_int, [_f, _f, _f, _f, C.c_longlong, _f32, _f32, _f32, _f32, _f32,
                                            _f32, _f, _f, _f, _f, _f, _v]),
    "dv3_im2col_s2k4": (C.c_int, [_f, _i32, _i32, _i32, _i32, _f, _f, _f, _v]),
    "dv3_col2im_s2k4": (C.c_int, [_f, _i32, _i32, _i32, _i32, _f, _f32, _f, _v]),
    "dv3_debug_observe_timing": (C.c_int, [_P(C.c_ulonglong), _i32]),
    "dv3_debug_imagine_timing": (C.c_int, [_P(C.c_ulonglong), _i32]),
    "dv3_ln_silu_fwd_split": (C.c_int, [_f, _i32, _f, _f, _f32, _i32, _i32, _f, _i32, _f, _f, _i32,
                                        _v]),
    "dv3_ln_silu_bwd_split": (C.c_int, [_f, _i32, _f, _f, _f32, _f, _i32, _i32, _i32, _f, _f, _i32,
                                        _f, _f, _i32, _v]),
    "dv3_ln_param_grads": (C.c_int, [_f, _i32, _f, _i32, _f32, _i32, _i32, _f, _f, _v]),
    "dv3_ln_param_grads_acc": (C.c_int, [_f, _i32, _f, _i32, _f32, _i32, _i32, _f, _f, _v]),
    "dv3_gru_gates_fwd": (C.c_int, [_f, _i32, _f, _f, _f32, _f, _i32, _i32, _i32, _f, _i32, _v]),
    "dv3_gru_gates_bwd": (C.c_int, [_f, _i32, _f, _f, _f32, _f, _i32, _f, _i32, _i32, _i32, _f, _f,
                                    _i32, _f, _i32, _v]),
    "dv3_onehot_linear_ln_silu": (C.c_int, [_i, _i32, _i32, _f, _i32, _f, _f, _f, _f, _f32, _i32,
                                            _i32, _f, _f, _v]),
    "dv3_onehot_sample": (C.c_int, [_f, _f, _f32, _i32, _i32, _i32, _i, _f, _i32, _v]),
    "dv3_onehot_st_bwd": (C.c_int, [_f, _f, _f, _f32, _i32, _i32, _i32, _f, _v]),
    "dv3_idx_to_onehot": (C.c_int, [_i, _i32, _i32, _i32, _f, _i32, _v]),
    "dv3_symlog": (C.c_int, [_f, C.c_longlong, _f, _v]),
    "dv3_sqerr_logprob_fwd": (C.c_int, [_f, _f, _i32, _i32, _i32, _f32, _f, _v]),
    "dv3_sqerr_logprob_bwd": (C.c_int, [_f, _f, _f, _i32, _i32, _i32, _f32, _f, _v]),
    "dv3_bernoulli_logprob_fwd": (C.c_int, [_f, _f, C.c_longlong, _f, _v]),
    "dv3_bernoulli_logprob_bwd": (C.c_int, [_f, _f, _f, C.c_longlong, _f, _v]),
    "dv3_loss_mean_fwd": (C.c_int, [_pf, _f, _i32, _i32, _f, _f, _v]),
    "dv3_loss_mean_bwd": (C.c_int, [_f, _f, _i32, _i32, _f, _v]),
    "dv3_discount_weights_fwd": (C.c_int, [_f, _f32, _i32, _i32, _f, _f, _v]),
    "dv3_discount_bwd": (C.c_int, [_f, _f, _f32, C.c_longlong, _f, _v]),
    "dv3_reward_ema": (C.c_int, [_f, _i32, _dbl, _f, _f, _v]),
    "dv3_actor_loss_fwd": (C.c_int, [_f, _f, _f, _f, _f, _f, _f32, _i32, _i32, _f, _f, _v]),
    "dv3_actor_loss_bwd": (C.c_int, [_f, _f, _f, _f, _f, _f32, _i32, _i32, _i32, _f, _f, _f, _v]),
    "dv3_value_loss_fwd": (C.c_int, [_f, _f, _f, _i32, _f, _v]),
    "dv3_value_loss_bwd": (C.c_int, [_f, _f, _i32, _f, _v]),
    "dv3_normal_policy_fwd": (C.c_int, [_f, _f, _f, _f32, _f32, _i32, _i32, _f, _f, _v]),
    "dv3_normal_policy_bwd": (C.c_int, [_f, _f, _f, _f, _f, _f32, _f32, _i32, _i32, _f, _f, _f, _v]),
    "dv3_rssm_initial_bwd_scratch_floats": (C.c_size_t, [_P(RssmDims)]),
    "dv3_rssm_initial_bwd": (C.c_int, [_P(RssmDims), _P(RssmParams), _f, _f, _f, _f, _f, _f, _f, _f, _f, _f,
                                       _f, _f, _f, _v]),
    "dv3_col_sum": (C.c_int, [_f, _i32, _i32, _i32, _f, _i32, _v]),
    "dv3_tensorstats": (C.c_int, [_f, C.c_longlong, _f, _v]),
    "dv3_ema_mix": (C.c_int, [_f, _f, C.c_longlong, _dbl, _v]),
}

_lib = None


class Dv3Error(RuntimeError):
    pass


def lib():
    """The loaded shared library.  Raises if it was not built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise Dv3Error(f"{LIB_PATH} is missing: build it with `python -c 'import "
                           "__graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no "
                           "CPU/PyTorch fallback for the DreamerV3 hot path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().dv3_last_error().decode("utf-8", "replace")
        raise Dv3Error(f"{what} failed ({rc}): {msg}")


def stream_ptr() -> C.c_void_p:
    """The current stream of the CURRENT device; ``fptr`` / ``iptr`` refuse tensors that live on
    another device, so a launch can never land on the wrong GPU's stream."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_device(t):
    if t.device.index != torch.cuda.current_device():
        raise Dv3Error(f"tensor on {t.device} but the current CUDA device is "
                       f"cuda:{torch.cuda.current_device()}: wrap the call in torch.cuda.device(...)")


def fptr(t):
    """float* of a contiguous fp32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise Dv3Error(f"expected contiguous fp32 CUDA tensor, got {t.dtype} "
                       f"{'contig' if t.is_contiguous() else 'strided'} on {t.device}")
    _check_device(t)
    return C.cast(C.c_void_p(t.data_ptr()), _f)


def iptr(t):
    if t is None:
        return None
    if t.dtype != torch.int32 or not t.is_contiguous() or not t.is_cuda:
        raise Dv3Error(f"expected contiguous int32 CUDA tensor, got {t.dtype} on {t.device}")
    _check_device(t)
    return C.cast(C.c_void_p(t.data_ptr()), _i)


def fill(struct, **kw):
    """Set struct fields from tensors / scalars by field type."""
    kinds = dict(struct._fields_)
    for name, val in kw.items():
        kind = kinds[name]
        if kind is _f:
            setattr(struct, name, fptr(val))
        elif kind is _i:
            setattr(struct, name, iptr(val))
        elif kind is _v:
            setattr(struct, name, C.c_void_p(val.data_ptr()) if val is not None else None)
        else:
            setattr(struct, name, val)
    return struct


def float_ptr_array(tensors):
    arr = (_f * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = fptr(t)
    return arr
