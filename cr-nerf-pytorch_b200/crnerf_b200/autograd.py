"""Autograd for the fused render pass (the training step of train_mask_grid_sample.py:186-197).

Forward: ``crnerf_render_pass_train`` - the same fused embed -> MLP -> composite kernel as
inference, which additionally stores each layer's 16-bit activations and the per-point
``[features | sigma]`` (nothing is recomputed in the backward).

Backward: ``crnerf_composite_backward`` (our kernel) turns the gradients of
``feature`` / ``weights`` / ``depth`` into the gradients of the MLP's pre-activation outputs;
the twelve dgrad/wgrad pairs that follow are plain dense GEMMs over the saved activations and
go to cuBLAS through ``torch.matmul`` (``BACKWARD_MATMUL``: "fp32" default, "tf32", or "bf16" = bf16
operands with fp32 accumulation and output - the pairing for ``args.crnerf_operand = 'bf16'`` models,
BASELINE configs[4], whose saved activations are then consumed without any conversion),
with the ReLU masks applied elementwise.  Parameter gradients land in ``param.grad`` of the
unchanged module tree, so Adam / DDP work as in the reference.  Rays, depths and noise get no
gradient (the reference detaches the importance samples, models/rendering.py:184).
"""
from __future__ import annotations

import contextlib

import torch

from . import ops

# "fp32" (default: the reference trains with torch's default fp32 matmuls) | "tf32" | "bf16"
# (explicit opt-ins; "bf16" is the pairing for args.crnerf_operand = 'bf16' models)
BACKWARD_MATMUL = "fp32"
DEFAULT_BACKWARD_MATMUL = BACKWARD_MATMUL


@contextlib.contextmanager
def _matmul_mode():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = BACKWARD_MATMUL == "tf32"
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def _mm(a, b):
    if BACKWARD_MATMUL == "bf16":
        # bf16 operands, fp32 accumulate AND fp32 output (cuBLAS through torch.mm's out_dtype); with
        # operand='bf16' models the saved activations already are bf16 and are used as they are
        return torch.mm(a.to(torch.bfloat16), b.to(torch.bfloat16), out_dtype=torch.float32)
    return a.float() @ b.float()


class RenderPassFn(torch.autograd.Function):
    """(weights, feature, depth) = render_pass(rays, z, noise; 12 weights, 12 biases)."""

    @staticmethod
    def forward(ctx, packed, rays, z_vals, noise, view_dir, n_fx, n_fd, *params):
        w, f, d, acts, raw = ops.render_pass_train(packed, rays, z_vals, noise, view_dir, n_fx, n_fd)
        ctx.save_for_backward(rays, z_vals, noise, view_dir, acts, raw, *params)
        ctx.n_fx, ctx.n_fd = n_fx, n_fd
        ctx.set_materialize_grads(False)   # outputs the loss does not use arrive as None, not as zeros
        return w, f, d

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_w, g_f, g_d):
        rays, z, noise, view_dir, acts, raw, *params = ctx.saved_tensors
        if g_w is None and g_f is None and g_d is None:
            return (None,) * (7 + len(params))
        W, B = params[:12], params[12:]
        n, s = z.shape
        P = n * s
        d_rgb, d_sig = ops.composite_backward(raw, z, noise, g_f, g_w, g_d)
        trunk, dir_out = ops.split_acts(acts, P)          # 16-bit views, no copies
        # embeddings as the forward saw them (recomputed: 120 columns, cheap)
        xyz = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]).reshape(P, 3)
        emb_xyz = ops.pos_embed(xyz, ctx.n_fx)
        emb_dir = ops.pos_embed((rays[:, 3:6] if view_dir is None else view_dir).contiguous(), ctx.n_fd)
        e_xyz = emb_xyz.shape[1]
        gW, gB = [None] * 12, [None] * 12
        with _matmul_mode():
            # static_rgb: sigmoid already folded into d_rgb (index 10)
            gW[10] = _mm(d_rgb.t(), dir_out)
            gB[10] = ops.relu_bias_grad(d_rgb, None)
            g = _mm(d_rgb, W[10])
            gB[9] = ops.relu_bias_grad(g, dir_out)        # dir_encoding ReLU mask + bias grad (index 9)
            fin = trunk[8]
            gW[9] = torch.cat([_mm(g.t(), fin), _mm(g.view(n, s, 128).sum(1).t(), emb_dir)], dim=1)
            g = _mm(g, W[9][:, :256])                     # xyz_encoding_final (index 8), no activation
            h8 = trunk[7]
            gW[8] = _mm(g.t(), h8)
            gB[8] = ops.relu_bias_grad(g, None)
            g = _mm(g, W[8])
            # static_sigma (index 11) joins at h8
            gW[11] = _mm(d_sig[None, :], h8)
            gB[11] = d_sig.sum().reshape(1)
            g = g + d_sig[:, None] * W[11].float()
            for l in range(7, -1, -1):                    # xyz_encoding_{l+1}: ReLU mask + bias grad
                gB[l] = ops.relu_bias_grad(g, trunk[l])
                if l == 0:
                    gW[l] = _mm(g.t(), emb_xyz)
                    break
                x_prev = trunk[l - 1]
                if l == 4:                                # skip layer: input = [xyz_emb | h4]
                    gW[l] = torch.cat([_mm(g.t(), emb_xyz), _mm(g.t(), x_prev)], dim=1)
                    g = _mm(g, W[l][:, e_xyz:])
                else:
                    gW[l] = _mm(g.t(), x_prev)
                    g = _mm(g, W[l])
        grads = [gw.to(w.dtype) for gw, w in zip(gW, W)] + [gb.to(b.dtype) for gb, b in zip(gB, B)]
        return (None, None, None, None, None, None, None, *grads)


def render_pass(model, rays, z_vals, noise, view_dir, n_fx, n_fd):
    """Differentiable fused pass for a ``models.nerf.NeRF_sigma``."""
    lin = model._linears()
    params = [m.weight for m in lin] + [m.bias for m in lin]
    return RenderPassFn.apply(model.packed(), rays, z_vals, noise, view_dir, n_fx, n_fd, *params)
