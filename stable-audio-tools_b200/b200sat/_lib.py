"""ctypes binding of libb200sat.so (the C ABI declared in include/b200sat.h).

No torch types cross the boundary: tensors are passed as raw device pointers + sizes + a cudaStream_t.
There is no fallback: if the library is missing or a call fails, a Python exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200sat.so")

c_void_p, c_int, c_float, c_char_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_char_p
c_ull = ctypes.c_ulonglong
c_fp = ctypes.c_void_p  # const float*
c_long = ctypes.c_long

# name -> (restype, argtypes); kept in sync with include/b200sat.h (tests/test_abi.py checks every symbol).
SIGNATURES = {
    "b200sat_last_error": (c_char_p, []),
    "b200sat_version": (c_int, []),
    "b200sat_num_sms": (c_int, []),
    "b200sat_set_sm_limit": (c_int, [c_int]),
    "b200sat_launch_count": (c_ull, []),
    "b200sat_gemm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_fp, c_void_p, c_int, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_fp, c_void_p, c_int, c_fp, c_fp, c_float, c_fp, c_int, c_void_p]),
    "b200sat_attention_fwd": (c_int, [c_void_p] * 4 + [c_fp] + [c_int] * 5 + [c_long] * 12 + [c_int, c_float, c_void_p]),
    "b200sat_layernorm_fwd": (c_int, [c_void_p, c_long, c_fp, c_fp, c_fp, c_fp, c_long, c_int, c_void_p, c_long, c_int, c_int,
                                      c_float, c_void_p]),
    "b200sat_small_linear": (c_int, [c_void_p, c_long, c_void_p, c_long, c_fp, c_void_p, c_long, c_void_p, c_long,
                                     c_int, c_int, c_int, c_int, c_int, c_int, c_fp, c_long, c_void_p]),
    "b200sat_fourier_features": (c_int, [c_fp, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "b200sat_dit_concat": (c_int, [c_fp, c_fp, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_fp, c_void_p, c_void_p]),
    "b200sat_dit_pre": (c_int, [c_fp, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_fp, c_void_p, c_void_p]),
    "b200sat_dit_post": (c_int, [c_void_p, c_long, c_int, c_void_p, c_fp, c_int, c_int, c_int, c_int, c_float, c_float,
                                 c_void_p]),
    "b200sat_sampler_update": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_void_p, c_long, c_int, c_void_p]),
    "b200sat_step_set": (c_int, [c_void_p, c_int, c_void_p]),
    "b200sat_conv1d_fwd": (c_int, [c_void_p] * 4 + [c_fp] + [c_void_p] * 6 + [c_fp, c_fp] + [c_int] * 10 + [c_void_p]),
    "b200sat_residual_unit_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_fp, c_fp, c_fp, c_void_p, c_fp, c_fp, c_fp, c_void_p, c_void_p,
                                          c_int, c_int, c_int, c_int, c_void_p]),
    "b200sat_wn_pack": (c_int, [c_fp, c_fp, c_fp, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sat_snake_prep": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_void_p]),
    "b200sat_conv_in": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp] + [c_void_p] * 4 + [c_int] * 6 + [c_void_p]),
    "b200sat_conv_out": (c_int, [c_void_p, c_void_p, c_fp, c_fp, c_fp] + [c_int] * 7 + [c_void_p]),
    "b200sat_to_planes": (c_int, [c_fp, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200sat_attention_bwd": (c_int, [c_void_p] * 5 + [c_fp, c_fp] + [c_void_p] * 3 + [c_int] * 5 + [ctypes.POINTER(ctypes.c_long), c_int,
                                      c_float, c_fp, c_fp, c_void_p]),
    "b200sat_layernorm_bwd": (c_int, [c_void_p, c_long, c_void_p, c_long, c_fp, c_void_p, c_long, c_void_p, c_long, c_fp, c_int, c_int,
                                      c_float, c_void_p]),
    "b200sat_colsum": (c_int, [c_void_p, c_long, c_fp, c_int, c_int, c_void_p]),
    "b200sat_stft_prefilter": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sat_stft_loss_accumulate": (c_int, [c_fp, c_fp, c_void_p, c_fp, c_fp, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "b200sat_stft_loss_backward": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "b200sat_stft_prefilter_backward": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sat_conv_wgrad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_fp, c_int, c_int,
                                   c_void_p]),
    "b200sat_layernorm_mod_bwd": (c_int, [c_void_p, c_long, c_void_p, c_long, c_fp, c_fp, c_long, c_int, c_void_p, c_long, c_void_p, c_long, c_fp,
                                          c_int, c_int, c_float, c_void_p]),
    "b200sat_gate_bwd": (c_int, [c_void_p, c_long, c_void_p, c_long, c_fp, c_void_p, c_long, c_fp, c_int, c_int, c_int, c_void_p]),
    "b200sat_adamw_ema_step": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_void_p, c_long, c_long, c_float, c_float, c_float, c_float, c_float, c_int,
                                       c_float, c_float, c_int, c_void_p]),
    "b200sat_conv2d_flat": (c_int, [c_void_p, c_void_p, c_fp, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_float,
                                    c_void_p]),
    "b200sat_disc_stft": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sat_disc_conv0": (c_int, [c_fp, c_fp, c_fp, c_void_p, c_void_p, c_fp, c_int, c_int, c_int, c_float, c_void_p]),
    "b200sat_disc_convpost": (c_int, [c_void_p, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_void_p]),
    "b200sat_disc_hinge_sums": (c_int, [c_fp, c_fp, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200sat_disc_l1_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_void_p]),
    "b200sat_disc_logit_grad": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "b200sat_disc_act_bwd": (c_int, [c_void_p, c_fp, c_fp, c_void_p, c_void_p, c_float, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200sat_disc_spec_pack": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sat_disc_conv0_wgrad": (c_int, [c_void_p, c_fp, c_fp, c_int, c_int, c_int, c_void_p]),
    "b200sat_disc_convpost_wgrad": (c_int, [c_fp, c_void_p, c_fp, c_fp, c_int, c_int, c_int, c_void_p]),
    "b200sat_conv_wgrad_taps": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_fp, c_int, c_void_p]),
    "b200sat_conv_wgrad_taps_win": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_fp, c_int, c_void_p]),
    "b200sat_conv_wgrad_taps_cat": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_fp, c_int, c_void_p]),
    "b200sat_snake_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_fp, c_fp, c_void_p, c_fp, c_fp, c_fp, c_long, c_int, c_void_p]),
    "b200sat_wn_pack_dgrad": (c_int, [c_fp, c_fp, c_fp, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sat_wn_bwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_void_p]),
    "b200sat_edge_wgrad": (c_int, [c_void_p, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_long, c_long, c_void_p]),
    "b200sat_vae_sample_bwd": (c_int, [c_void_p, c_void_p, c_fp, c_fp, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200sat_vae_sample": (c_int, [c_void_p, c_void_p, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_void_p]),
}

_lib = None


class B200SatError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200SatError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (no CPU/PyTorch fallback exists)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc == 0:
        return
    msg = lib().b200sat_last_error().decode(errors="replace")
    if rc == -2:
        raise NotImplementedError(f"b200sat {what}: unsupported: {msg}")
    if rc < 0:
        raise ValueError(f"b200sat {what}: invalid argument: {msg}")
    raise B200SatError(f"b200sat {what}: CUDA error {rc}: {msg}")
