"""ctypes binding of ``libcrnerf_b200.so`` (C ABI in ``include/crnerf_b200.h``).

This is the only place the shared library is loaded.  There is no fallback: if
the library has not been built (``python -c 'import __graft_entry__ as g; g.build()'``
or ``bash cr-nerf-pytorch_b200/csrc/build.sh``) importing the ops raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CRNERF_B200_LIB selects another build of the same ABI (A/B timing of kernel variants, tools/ab_kernel.sh)
LIB_PATH = os.environ.get("CRNERF_B200_LIB") or os.path.join(_HERE, "libcrnerf_b200.so")

c_float_p = C.c_void_p  # device pointers travel as integers


class MlpWeights(C.Structure):
    """crnerf_mlp_weights"""
    _fields_ = [("weight", C.c_void_p * 12), ("bias", C.c_void_p * 12),
                ("e_xyz", C.c_int32), ("e_dir", C.c_int32)]


class RenderOpts(C.Structure):
    """crnerf_render_opts"""
    _fields_ = [("xyz_jitter", C.c_void_p), ("channel_partials", C.c_void_p), ("overflow_flag", C.c_void_p)]


class CnnWeights(C.Structure):
    """crnerf_cnn_weights"""
    _fields_ = [("conv_w", C.c_void_p * 3), ("conv_b", C.c_void_p * 3),
                ("fc_w", C.c_void_p), ("fc_b", C.c_void_p)]


class StyleWeights(C.Structure):
    """crnerf_style_weights"""
    _fields_ = [("cnet", CnnWeights), ("snet", CnnWeights),
                ("compress_w", C.c_void_p), ("compress_b", C.c_void_p),
                ("unzip_w", C.c_void_p), ("unzip_b", C.c_void_p),
                ("rgb_w", C.c_void_p), ("rgb_b", C.c_void_p)]


class EncoderWeights(C.Structure):
    """crnerf_encoder_weights"""
    _fields_ = [("weight", C.c_void_p * 7), ("bias", C.c_void_p * 7)]


# name -> (restype, argtypes); must list every symbol include/crnerf_b200.h declares
SIGNATURES = {
    "crnerf_last_error": (C.c_char_p, []),
    "crnerf_abi_version": (C.c_int, []),
    "crnerf_device_ok": (C.c_int, []),
    "crnerf_launch_count": (C.c_uint64, []),
    "crnerf_mlp_packed_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "crnerf_mlp_packed_bytes_op": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "crnerf_render_partial_rows": (C.c_int, [C.c_int, C.c_int]),
    "crnerf_render_pass_opts": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.POINTER(RenderOpts), C.c_void_p]),
    "crnerf_mlp_pack": (C.c_int, [C.POINTER(MlpWeights), C.c_int, C.c_void_p, C.c_size_t,
                                  C.c_void_p, C.c_void_p]),
    "crnerf_render_pass": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "crnerf_render_acts_bytes": (C.c_size_t, [C.c_int64]),
    "crnerf_render_pass_train": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crnerf_render_pass_train_opts": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RenderOpts),
                                                C.c_void_p]),
    "crnerf_render_backward_weights_bytes": (C.c_size_t, [C.c_int]),
    "crnerf_render_backward_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "crnerf_render_backward": (C.c_int, [C.POINTER(MlpWeights), C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_int] +
                               [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "crnerf_composite_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "crnerf_relu_bias_grad_scratch_floats": (C.c_size_t, [C.c_int]),
    "crnerf_relu_bias_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "crnerf_mlp_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "crnerf_generate_rays": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p]),
    "crnerf_grid_patch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_float] + [C.c_void_p] * 7),
    "crnerf_rgb_to_u8": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "crnerf_encoder_packed_bytes": (C.c_size_t, []),
    "crnerf_encoder_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "crnerf_encoder_pack": (C.c_int, [C.POINTER(EncoderWeights), C.c_void_p, C.c_size_t, C.c_void_p]),
    "crnerf_encoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_size_t, C.c_void_p]),
    "crnerf_encoder_tape_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "crnerf_encoder_backward_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "crnerf_encoder_forward_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_size_t, C.c_void_p]),
    "crnerf_encoder_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.POINTER(EncoderWeights), C.c_void_p, C.c_void_p,
                                          C.c_size_t, C.c_void_p]),
    "crnerf_ray_loss_forward": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_float, C.c_float, C.c_float,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "crnerf_ray_loss_backward": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_float, C.c_float, C.c_float] +
                                 [C.c_void_p] * 5),
    "crnerf_pair_loss_forward": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_float),
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "crnerf_pair_loss_backward": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_float),
                                            C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                            C.c_void_p]),
    "crnerf_loss_scratch_floats": (C.c_size_t, []),
    "crnerf_mask_sample_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "crnerf_mask_sample_backward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "crnerf_pos_embed": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "crnerf_coarse_z": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    "crnerf_sample_pdf_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                          C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "crnerf_sample_pdf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                    C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "crnerf_style_scratch_floats": (C.c_size_t, [C.c_int64]),
    "crnerf_style_forward": (C.c_int, [C.POINTER(StyleWeights), C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crnerf_style_forward_sums": (C.c_int, [C.POINTER(StyleWeights), C.c_void_p, C.c_int64, C.c_int64,
                                            C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                            C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crnerf_style_aux_floats": (C.c_size_t, []),
    "crnerf_style_backward_grads_floats": (C.c_size_t, []),
    "crnerf_style_backward_scratch_floats": (C.c_size_t, [C.c_int64, C.c_int64]),
    "crnerf_style_backward_layout": (None, [C.POINTER(C.c_int64)]),
    "crnerf_style_forward_train": (C.c_int, [C.POINTER(StyleWeights), C.c_void_p, C.c_int64, C.c_int64,
                                             C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crnerf_style_backward": (C.c_int, [C.POINTER(StyleWeights), C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64] +
                              [C.c_void_p] * 7),
    "crnerf_sum_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "crnerf_cnn_forward": (C.c_int, [C.POINTER(CnnWeights), C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crnerf_style_stats1": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "crnerf_style_stats2": (C.c_int, [C.POINTER(StyleWeights), C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crnerf_style_apply": (C.c_int, [C.POINTER(StyleWeights), C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "crnerf_adam_step": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_double, C.c_int, C.c_void_p]),
    "crnerf_debug_set": (C.c_int, [C.c_void_p, C.c_int]),
    "crnerf_debug_program": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int]),
}

_lib = None


class CrnerfError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it is missing - no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library has not been built. Run "
            "`bash cr-nerf-pytorch_b200/csrc/build.sh` (or __graft_entry__.build()). "
            "crnerf_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so and header drift apart
        fn.restype = res
        fn.argtypes = args
    if lib.crnerf_abi_version() != 1:
        raise ImportError("libcrnerf_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().crnerf_last_error().decode("utf-8", "replace")
        raise CrnerfError(f"crnerf_b200 error {rc}: {msg}")
