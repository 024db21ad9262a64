"""``models.nerf`` of the reference (models/nerf.py), backed by sm_100a kernels.

Exports the reference's names: ``PosEmbedding`` (nerf.py:4-30) and ``NeRF_sigma``
(nerf.py:115-182) are the live classes on the hot path and run on the CUDA
library; ``NeRF`` (nerf.py:33-113) and ``NeRF_sigma_tanh`` (nerf.py:184-264) are
never instantiated by the reference's scripts and are importable placeholders.

Module structure, parameter names and construction order equal the reference's,
so ``state_dict`` keys (``xyz_encoding_1.0.weight`` ... ``static_rgb.0.bias``),
``load_ckpt(model, path, model_name='nerf_fine')`` and seeded default
initialisation are interchangeable.
"""
import torch
from torch import nn

from crnerf_b200 import ops


# Process-wide "some optimizer stepped" counter, part of every packed-weight cache key: optimizer
# steps that write through ``p.data`` leave ``p._version`` untouched (reference utils/__init__.py:33-36
# offers such optimizers), the global post-step hook sees them all.
_weights_epoch = [0]


def _bump_weights_epoch(optimizer, args, kwargs):
    _weights_epoch[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook
    _reg_post_hook(_bump_weights_epoch)
except ImportError:      # very old torch: fall back to the documented invalidate_packed()
    pass


class PosEmbedding(nn.Module):
    def __init__(self, max_logscale, N_freqs, logscale=True):
        """x -> (x, sin(2^k x), cos(2^k x), ...), reference models/nerf.py:5-15."""
        super().__init__()
        self.N_freqs = N_freqs
        self.logscale = logscale
        if logscale:
            self.freqs = 2 ** torch.linspace(0, max_logscale, N_freqs)
        else:
            self.freqs = torch.linspace(1, 2 ** max_logscale, N_freqs)
        expect = 2.0 ** torch.arange(N_freqs, dtype=self.freqs.dtype)
        self._pow2 = bool(N_freqs == 0 or torch.equal(self.freqs, expect))

    def forward(self, x):
        """(B, 3) -> (B, 6*N_freqs+3), reference models/nerf.py:17-30."""
        if not self._pow2:
            raise NotImplementedError(
                "crnerf_b200 PosEmbedding supports the power-of-two bands PosEmbedding(L-1, L) "
                "every reference script builds; other frequency sets have no kernel")
        return ops.pos_embed(x, self.N_freqs)


class NeRF_sigma(nn.Module):
    def __init__(self, typ, args,
                 D=8, W=256, skips=[4],
                 in_channels_xyz=63, in_channels_dir=27,
                 encode_appearance=False, in_channels_a=48,
                 encode_random=False):
        """Same constructor contract as reference models/nerf.py:116-154
        (``in_channels_a`` is stored and unused there too)."""
        super().__init__()
        self.typ = typ
        self.D = D
        self.W = W
        self.skips = skips
        self.in_channels_xyz = in_channels_xyz
        self.in_channels_dir = in_channels_dir
        self.encode_appearance = False if typ == 'coarse' else encode_appearance
        self.in_channels_a = in_channels_a if encode_appearance else 0
        self.encode_random = False if typ == 'coarse' else encode_random
        self.out_dim = args.nerf_out_dim

        for i in range(D):
            if i == 0:
                layer = nn.Linear(in_channels_xyz, W)
            elif i in skips:
                layer = nn.Linear(W + in_channels_xyz, W)
            else:
                layer = nn.Linear(W, W)
            setattr(self, f"xyz_encoding_{i+1}", nn.Sequential(layer, nn.ReLU(inplace=True)))
        self.xyz_encoding_final = nn.Linear(W, W)
        self.static_sigma = nn.Sequential(nn.Linear(W, 1), nn.Softplus())
        self.dir_encoding = nn.Sequential(nn.Linear(W + in_channels_dir, W // 2),
                                          nn.ReLU(inplace=True))
        self.static_rgb = nn.Sequential(nn.Linear(W // 2, args.nerf_out_dim), nn.Sigmoid())

        self.operand = getattr(args, 'crnerf_operand', 'fp16')
        self._packed = None
        self._packed_key = None
        self._epoch = 0
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())

    # -- kernel plumbing ---------------------------------------------------
    def _linears(self):
        mods = [getattr(self, f"xyz_encoding_{i+1}")[0] for i in range(self.D)]
        return mods + [self.xyz_encoding_final, self.dir_encoding[0], self.static_rgb[0],
                       self.static_sigma[0]]

    def _check_architecture(self):
        if self.D != 8 or self.W != 256 or list(self.skips) != [4] or self.out_dim != 64:
            raise NotImplementedError(
                "the fused sm_100a MLP kernel is specialised for the configuration every "
                "reference script uses (D=8, W=256, skips=[4], nerf_out_dim=64); got "
                f"D={self.D}, W={self.W}, skips={self.skips}, out_dim={self.out_dim}")

    def packed(self):
        """Tensor-core-ready weight image.

        Training (``wants_grad()``): re-packed on EVERY call - one 10 us kernel - so no cache key
        can go stale whatever updated the weights (optimizers that write through ``p.data``, as
        torch_optimizer's RAdam / Ranger do, never bump ``p._version``).  The fp16 range verdict of
        the previous image is consumed here without a host sync (``PackedMLP.poll_range``).

        Inference: cached; the key holds every parameter's ``(data_ptr, _version)`` plus two
        epochs that catch what ``_version`` misses: a process-wide one bumped by a post-step hook
        on every ``torch.optim.Optimizer`` (any optimizer, any update style) and a per-module one
        bumped by ``load_state_dict`` / ``.to()`` and by ``invalidate_packed()`` - which is what
        to call after writing weights by hand through ``.data`` (EMA, weight surgery)."""
        self._check_architecture()
        lin = self._linears()
        training = self.wants_grad()
        if training:
            # inside a CUDA-graph capture (crnerf_b200.graphs.GraphedTrainStep) nothing may touch
            # the host: no verdict read-back there (run a few eager steps first, as the helper does)
            capturing = torch.cuda.is_current_stream_capturing()
            if self._packed is not None and not capturing:
                self._packed.poll_range()
            self._packed = ops.pack_mlp([m.weight for m in lin], [m.bias for m in lin],
                                        self.in_channels_xyz, self.in_channels_dir, self.operand,
                                        check_range=False if capturing else "deferred")
            self._packed_key = None
            return self._packed
        key = (self.operand, _weights_epoch[0], self._epoch) + tuple(
            (m.weight.data_ptr(), m.weight._version, m.bias.data_ptr(), m.bias._version) for m in lin)
        if self._packed is None or key != self._packed_key:
            # inference verifies that every weight fits the fp16 operand range (one host sync per
            # weight version)
            self._packed = ops.pack_mlp([m.weight for m in lin], [m.bias for m in lin],
                                        self.in_channels_xyz, self.in_channels_dir, self.operand,
                                        check_range=True)
            self._packed_key = key
        return self._packed

    def invalidate_packed(self):
        """Force a re-pack at the next call (after in-place weight edits that bypass autograd's
        version counter)."""
        self._epoch += 1

    def _apply(self, fn, *a, **k):
        self._epoch += 1
        return super()._apply(fn, *a, **k)

    def wants_grad(self):
        """True when the call must be recorded for autograd (training step)."""
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _no_autograd(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError(
                "NeRF_sigma.forward on pre-embedded rows is inference-only (torch.no_grad()); the "
                "differentiable entry is models.rendering.render_rays_cross_ray, which is what the "
                "reference's training step calls")

    def forward(self, x, sigma_only=False, output_random=True):
        """(B, in_channels_xyz+in_channels_dir) -> (B, nerf_out_dim+1) = [features | sigma],
        or (B, 1) if ``sigma_only``; reference models/nerf.py:157-182 (``output_random`` is
        ignored there as well)."""
        self._no_autograd(x)
        return ops.mlp_forward(self.packed(), x, sigma_only=sigma_only)


class _NotOnHotPath(nn.Module):
    _why = ""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError(self._why)


class NeRF(_NotOnHotPath):
    """Importable placeholder for reference models/nerf.py:33-113."""
    _why = ("models.nerf.NeRF is never instantiated by the reference's train/eval scripts "
            "(both models are NeRF_sigma, train_mask_grid_sample.py:39,56); crnerf_b200 does not "
            "provide it")


class NeRF_sigma_tanh(_NotOnHotPath):
    """Importable placeholder for reference models/nerf.py:184-264."""
    _why = ("models.nerf.NeRF_sigma_tanh is never instantiated by the reference's scripts; "
            "crnerf_b200 does not provide it")
