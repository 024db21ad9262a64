"""``torch.ops.crnerf.*`` - the library's kernels registered as torch operators.

BASELINE.json's north star asks for the hot path "bound as torch extensions"; the build
contract asks for a C-ABI boundary with no torch types in its signatures.  Both hold: the
kernels live behind ``include/crnerf_b200.h`` (``libcrnerf_b200.so``, bound with ctypes in
``ops.py``), and this module registers them with the dispatcher through ``torch.library``
(the Python twin of ``TORCH_LIBRARY(crnerf, m)``): schema, a ``CUDA`` kernel
that calls the C ABI on the current stream, a ``Meta`` kernel for shape propagation, and a
``CPU`` kernel that raises - there is no fallback.  ``models/rendering.py`` calls the
operators, so ``torch.ops.crnerf.render_pass`` & co. are what the reference-shaped API runs.

Operator list (reference lines each one replaces are in ``include/crnerf_b200.h``):

  coarse_z(rays, t_steps, perturb_rand?, use_disp) -> z
  render_pass(packed, operand, rays, z_vals, noise?, view_dir?, n_freq_xyz, n_freq_dir,
              xyz_jitter?, overflow_ptr, channel_partials) -> (weights, feature, depth, partials)
  sample_pdf(bins, weights, u, n_importance, eps) -> samples
  sample_pdf_merge(z_coarse, weights_coarse, u, n_importance, eps) -> z_fine
  pos_embed(x, n_freqs) -> emb
  mlp_forward(packed, operand, e_xyz, e_dir, x, sigma_only) -> out
  style_forward(content, style?, params[22], channel_sums?) -> rgb
  rgb_to_u8(rgb) -> u8
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import CrnerfError

__all__ = ["STYLE_KEYS", "style_params"]

_lib = torch.library.Library("crnerf", "DEF")

# order of the 22 style_net tensors handed to crnerf::style_forward (reference state_dict keys,
# models/linearStyleTransfer.py:7-25,44-56,279-283)
STYLE_KEYS = tuple(
    [f"multi_net.{net}.{k}" for net in ("snet", "cnet")
     for k in ("convs.0.weight", "convs.0.bias", "convs.2.weight", "convs.2.bias", "convs.4.weight",
               "convs.4.bias", "fc.weight", "fc.bias")] +
    ["multi_net.compress.weight", "multi_net.compress.bias", "multi_net.unzip.weight",
     "multi_net.unzip.bias", "decoder.feat_2_rgb_list.0.weight", "decoder.feat_2_rgb_list.0.bias"])


def style_params(module) -> list:
    """The ``params`` argument of ``crnerf::style_forward`` for a ``style_net`` module."""
    named = dict(module.named_parameters())
    return [named[k] for k in STYLE_KEYS]


def _packed(buf, operand, e_xyz, e_dir):
    return ops.PackedMLP(buf, int(operand), int(e_xyz), int(e_dir))


# ---- CUDA kernels -------------------------------------------------------------------------
def _coarse_z(rays, t_steps, perturb_rand, use_disp):
    return ops.coarse_z(rays, t_steps, perturb_rand, bool(use_disp))


def _render_pass(packed, operand, rays, z_vals, noise, view_dir, n_freq_xyz, n_freq_dir, xyz_jitter,
                 overflow_ptr, channel_partials):
    p = _packed(packed, operand, 3 + 6 * n_freq_xyz, 3 + 6 * n_freq_dir)
    out = ops.render_pass(p, rays, z_vals, noise, view_dir, n_freq_xyz, n_freq_dir, xyz_jitter,
                          bool(channel_partials), overflow_ptr=int(overflow_ptr))
    if channel_partials:
        return out
    return out + (rays.new_empty((0, 64)),)


def _sample_pdf(bins, weights, u, n_importance, eps):
    return ops.sample_pdf(bins, weights, u, n_importance, eps)


def _sample_pdf_merge(z_coarse, weights_coarse, u, n_importance, eps):
    return ops.sample_pdf_merge(z_coarse, weights_coarse, u, n_importance, eps)


def _pos_embed(x, n_freqs):
    return ops.pos_embed(x, n_freqs)


def _mlp_forward(packed, operand, e_xyz, e_dir, x, sigma_only):
    return ops.mlp_forward(_packed(packed, operand, e_xyz, e_dir), x, bool(sigma_only))


_style_cache = {"key": None, "sw": None}     # pointer struct of the last parameter set seen


def _style_forward(content, style, params, channel_sums=None):
    if len(params) != len(STYLE_KEYS):
        raise ValueError(f"params must hold the {len(STYLE_KEYS)} style_net tensors in STYLE_KEYS order")
    key = tuple(t.data_ptr() for t in params)
    if _style_cache["key"] != key:
        _style_cache["sw"] = ops.StyleWeightsRef(dict(zip(STYLE_KEYS, params)))
        _style_cache["key"] = key
    return ops.style_forward(_style_cache["sw"], content, style, channel_sums=channel_sums)


def _rgb_to_u8(rgb):
    return ops.rgb_to_u8(rgb)


# ---- Meta kernels (shapes only) -------------------------------------------------------------
def _m_coarse_z(rays, t_steps, perturb_rand, use_disp):
    return rays.new_empty((rays.shape[0], t_steps.shape[0]))


def _m_render_pass(packed, operand, rays, z_vals, noise, view_dir, n_freq_xyz, n_freq_dir, xyz_jitter,
                   overflow_ptr, channel_partials):
    n, s = z_vals.shape
    rows = 2 * 148 if channel_partials else 0        # upper bound; the CUDA kernel sizes it per device
    return (z_vals.new_empty((n, s)), z_vals.new_empty((n, 64)), z_vals.new_empty((n,)),
            z_vals.new_empty((rows, 64)))


def _m_sample_pdf(bins, weights, u, n_importance, eps):
    return bins.new_empty((bins.shape[0], n_importance))


def _m_sample_pdf_merge(z_coarse, weights_coarse, u, n_importance, eps):
    return z_coarse.new_empty((z_coarse.shape[0], z_coarse.shape[1] + n_importance))


def _m_pos_embed(x, n_freqs):
    return x.new_empty((x.shape[0], 3 + 6 * n_freqs))


def _m_mlp_forward(packed, operand, e_xyz, e_dir, x, sigma_only):
    return x.new_empty((x.shape[0], 1 if sigma_only else 65))


def _m_style_forward(content, style, params, channel_sums=None):
    return content.new_empty((1, 3, content.shape[2], content.shape[3]))


def _m_rgb_to_u8(rgb):
    if rgb.dim() == 4:
        return rgb.new_empty((rgb.shape[2], rgb.shape[3], 3), dtype=torch.uint8)
    return rgb.new_empty((rgb.shape[1], 3), dtype=torch.uint8)


def _no_cpu(name):
    def raise_(*args, **kwargs):
        raise CrnerfError(f"crnerf::{name} got CPU tensors: crnerf_b200 runs on CUDA (sm_100) only and has "
                          "no CPU fallback")
    return raise_


_OPS = (
    ("coarse_z", "(Tensor rays, Tensor t_steps, Tensor? perturb_rand, bool use_disp) -> Tensor",
     _coarse_z, _m_coarse_z),
    ("render_pass", "(Tensor packed, int operand, Tensor rays, Tensor z_vals, Tensor? noise, Tensor? view_dir, "
                    "int n_freq_xyz, int n_freq_dir, Tensor? xyz_jitter, int overflow_ptr, bool channel_partials) "
                    "-> (Tensor, Tensor, Tensor, Tensor)", _render_pass, _m_render_pass),
    ("sample_pdf", "(Tensor bins, Tensor weights, Tensor u, int n_importance, float eps) -> Tensor",
     _sample_pdf, _m_sample_pdf),
    ("sample_pdf_merge", "(Tensor z_coarse, Tensor weights_coarse, Tensor u, int n_importance, float eps) -> Tensor",
     _sample_pdf_merge, _m_sample_pdf_merge),
    ("pos_embed", "(Tensor x, int n_freqs) -> Tensor", _pos_embed, _m_pos_embed),
    ("mlp_forward", "(Tensor packed, int operand, int e_xyz, int e_dir, Tensor x, bool sigma_only) -> Tensor",
     _mlp_forward, _m_mlp_forward),
    ("style_forward", "(Tensor content, Tensor? style, Tensor[] params, Tensor? channel_sums=None) -> Tensor", _style_forward,
     _m_style_forward),
    ("rgb_to_u8", "(Tensor rgb) -> Tensor", _rgb_to_u8, _m_rgb_to_u8),
)

for _name, _schema, _cuda, _meta in _OPS:
    _lib.define(_name + _schema)
    _lib.impl(_name, _cuda, "CUDA")
    _lib.impl(_name, _meta, "Meta")
    _lib.impl(_name, _no_cpu(_name), "CPU")

OP_NAMES = tuple(n for n, *_ in _OPS)
