"""Autograd for the fused render pass (the training step of train_mask_grid_sample.py:186-197).

Forward: ``crnerf_render_pass_train`` - the same fused embed -> MLP -> composite kernel as
inference, which additionally stores each layer's 16-bit activations (in the backward kernels'
operand layout) and the per-point ``[features | sigma]``: nothing is recomputed in the backward.

Backward: ``crnerf_render_backward`` - composite backward, then per layer the weight gradient
``dW += G^T X`` and the input gradient ``G' = (G W) * relu'(X)`` as tcgen05 GEMMs over the saved
activations (csrc/backward_gemm.cu): no library GEMM, no fp32 (points x width) tensor.  The
gradient operands are 16-bit in the forward's operand format (fp16 with a per-pass power-of-two
scale, or bf16 for ``args.crnerf_operand = 'bf16'`` models - BASELINE configs[4]), accumulation is
fp32.  Parameter gradients land in ``param.grad`` of the unchanged module tree, so Adam / DDP work
as in the reference.  Rays, depths and noise get no gradient (the reference detaches the
importance samples, models/rendering.py:184).
"""
from __future__ import annotations

import torch

from . import ops

BACKWARD_MATMUL = "native"            # the only backward there is; kept as a name for tools that print it
DEFAULT_BACKWARD_MATMUL = BACKWARD_MATMUL


class RenderPassFn(torch.autograd.Function):
    """(weights, feature, depth) = render_pass(rays, z, noise; 12 weights, 12 biases)."""

    @staticmethod
    def forward(ctx, packed, rays, z_vals, noise, view_dir, n_fx, n_fd, jitter, *params):
        w, f, d, acts, raw = ops.render_pass_train(packed, rays, z_vals, noise, view_dir, n_fx, n_fd, jitter)
        ctx.save_for_backward(z_vals, noise, acts, raw, *params)
        ctx.operand, ctx.e_xyz, ctx.e_dir = packed.operand, packed.e_xyz, packed.e_dir
        ctx.set_materialize_grads(False)   # outputs the loss does not use arrive as None, not as zeros
        return w, f, d

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_w, g_f, g_d):
        z, noise, acts, raw, *params = ctx.saved_tensors
        if g_w is None and g_f is None and g_d is None:
            return (None,) * (8 + len(params))
        W, B = params[:12], params[12:]
        gW, gB = ops.render_backward(W, B, ctx.operand, ctx.e_xyz, ctx.e_dir, acts, raw, z, noise, g_f, g_w, g_d)
        grads = [g.to(w.dtype) for g, w in zip(gW, W)] + [g.to(b.dtype) for g, b in zip(gB, B)]
        return (None, None, None, None, None, None, None, None, *grads)


def render_pass(model, rays, z_vals, noise, view_dir, n_fx, n_fd, jitter=None):
    """Differentiable fused pass for a ``models.nerf.NeRF_sigma``.  ``jitter`` (n_points, 3): the
    ``args.pertubeCord`` displacement of the sample positions (no gradient, as in the reference)."""
    lin = model._linears()
    params = [m.weight for m in lin] + [m.bias for m in lin]
    return RenderPassFn.apply(model.packed(), rays, z_vals, noise, view_dir, n_fx, n_fd, jitter, *params)


class StyleNetFn(torch.autograd.Function):
    """``rgb = style_net(content, style)`` under autograd (the training step's decode(),
    train_mask_grid_sample.py:127-149): forward = the inference kernels of the cross-ray block
    (tensor-core Gram statistics at fp32-class accuracy) plus a small ``aux`` record, backward =
    csrc/style_backward.cu - fp32, deterministic, no library GEMM.  ``params`` are the 22 style_net
    tensors in ``ops.STYLE_GRAD_KEYS`` order."""

    @staticmethod
    def forward(ctx, sw, content, style, channel_sums, *params):
        rgb, aux = ops.style_forward_train(sw, content, style, channel_sums)
        ctx.sw = sw
        ctx.save_for_backward(content, style, aux)
        ctx.needs = (content.requires_grad, style.requires_grad)
        return rgb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_rgb):
        content, style, aux = ctx.saved_tensors
        g_c, g_s, grads = ops.style_backward(ctx.sw, content, style, aux, g_rgb)
        return (None, g_c if ctx.needs[0] else None, g_s if ctx.needs[1] else None, None,
                *[grads[k] for k in ops.STYLE_GRAD_KEYS])


def style_net_forward(module, sw, content, style, channel_sums=None):
    named = dict(module.named_parameters())
    return StyleNetFn.apply(sw, content, style, channel_sums, *[named[k] for k in ops.STYLE_GRAD_KEYS])


class Fp32Region(torch.autograd.Function):
    """``fn(x)`` as ONE node of the outer graph, for the modules that stay on library convolutions
    under autograd (``Context_Guided_Network``, ``encoder_sameoutputsize``).  The inner graph is
    recorded in ``forward`` and differentiated in ``backward``, both under
    ``cudnn.flags(allow_tf32=False)`` - a plain ``with`` block around the forward would leave the
    backward convolutions, which run later on the autograd thread, on TF32 (1e-3 off the reference's
    fp32 results), and a global switch would change the caller's process.  ``params`` are passed as
    inputs only so that their gradients are routed; ``fn`` reads them from the module."""

    @staticmethod
    def forward(ctx, fn, x, *params):
        with torch.enable_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            leaf = x.detach().requires_grad_(x.requires_grad)
            out = fn(leaf)
        ctx.leaf, ctx.out, ctx.params = leaf, out, params
        return out.detach()

    @staticmethod
    def backward(ctx, g):
        wanted = [t for t in (ctx.leaf,) + tuple(ctx.params) if t.requires_grad]
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            grads = iter(torch.autograd.grad(ctx.out, wanted, g, allow_unused=True))
        res = [next(grads) if t.requires_grad else None for t in (ctx.leaf,) + tuple(ctx.params)]
        ctx.leaf = ctx.out = ctx.params = None
        return (None, *res)


class EncoderFn(torch.autograd.Function):
    """``encoder_sameoutputsize.forward`` under autograd on the library's own kernels (the training
    step back-propagates through ``enc_a``: reference train_mask_grid_sample.py, models/
    linearStyleTransfer.py:250-276): forward = the inference tensor-core kernels with the activation
    planes kept, backward = csrc/encoder_train.cuh (input-gradient convolutions on the forward kernel,
    tcgen05 weight-gradient GEMMs with K = pixels, fp16 hi/lo operands, deterministic, no library
    convolution).  ``params`` = conv1.weight, conv1.bias, ..., conv7.bias."""

    @staticmethod
    def forward(ctx, packed, x, *params):
        out, tape = ops.encoder_forward_train(packed, x)
        ctx.packed, ctx.tape = packed, tape
        ctx.save_for_backward(x, out)
        ctx.want_x = x.requires_grad
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, out = ctx.saved_tensors
        gw, gb, gx = ops.encoder_backward(ctx.packed, x, out, ctx.tape, g, ctx.want_x)
        ctx.packed = ctx.tape = None
        return (None, gx, *[t for pair in zip(gw, gb) for t in pair])
