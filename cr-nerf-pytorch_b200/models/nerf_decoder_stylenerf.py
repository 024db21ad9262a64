"""``models.nerf_decoder_stylenerf`` of the reference: the live "conv decoder".

``style_net`` builds ``NeuralRenderer(img_size=img_wh, featmap_size=img_wh)``
(reference models/linearStyleTransfer.py:283), so ``n_blocks = log2(1) = 0`` and
the forward (reference models/nerf_decoder_stylenerf.py:279-291) is one 1x1 conv
64->3 followed by a sigmoid; the up-sampling loop has zero iterations and
``Blur``/kornia is never executed.  That is the configuration
implemented here, as the streaming apply kernel of csrc/crossray.cu (inside
``style_net`` it is folded into the cross-ray map, so the feature map is read
once for fusion + decode).

The module tree keeps the reference's state_dict keys, including the unused
``rgb_upsample.1.f`` buffer of ``Blur`` (:105-109), so reference checkpoints
load with ``load_ckpt(models['decoder'], path, 'decoder')``.
"""
from math import log2

import torch
import torch.nn as nn

from crnerf_b200 import ops


class _Blur(nn.Module):
    """State-dict stand-in for the reference's ``Blur`` (:105-114): holds the
    [1,2,1] buffer ``f``; never executed because n_blocks == 0."""

    def __init__(self):
        super().__init__()
        self.register_buffer('f', torch.Tensor([1, 2, 1]))

    def forward(self, x):
        raise NotImplementedError("Blur is unreachable at n_blocks == 0")


class NeuralRenderer(nn.Module):

    def __init__(self, bg_type="white", feat_nc=128, out_dim=3, final_actvn=True, min_feat=32,
                 featmap_size=(32, 32), img_size=(256, 256), **kwargs):
        """Constructor contract of reference nerf_decoder_stylenerf.py:229-242."""
        super().__init__()
        self.bg_type = bg_type
        self.featmap_size = featmap_size
        self.final_actvn = final_actvn
        self.n_feat = feat_nc
        self.out_dim = out_dim
        self.n_blocks = int(log2(img_size[0] / featmap_size[0]))
        self.min_feat = min_feat
        if self.n_blocks != 0:
            raise NotImplementedError(
                "NeuralRenderer with featmap_size != img_size (n_blocks > 0) is unreachable from the "
                "reference's scripts - its forward reads an undefined variable there "
                "(nerf_decoder_stylenerf.py:282) - and has no kernel")
        if feat_nc != 64 or out_dim != 3:
            raise NotImplementedError("the decoder kernel is specialised for 64 -> 3 channels "
                                      "(nerf_out_dim == 64)")
        # same members, in the same order, as the reference's _make_layer (:257-275)
        self.feat_upsample_list = nn.ModuleList([])
        self.rgb_upsample = nn.Sequential(
            nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False), _Blur())
        self.feat_2_rgb_list = nn.ModuleList([nn.Conv2d(self.n_feat, self.out_dim, 1, 1, padding=0)])
        self.feat_layers = nn.ModuleList([])
        self.actvn = nn.LeakyReLU(0.2, inplace=True)

    def forward(self, x):
        """(1,64,H,W) -> (1,3,H,W) = sigmoid(conv1x1(x)), reference :279-291."""
        conv = self.feat_2_rgb_list[0]
        if not self.final_actvn:
            raise NotImplementedError("final_actvn=False leaves `rgbs` undefined in the reference (:289-291)")
        if torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad):
            # training step (32x32 patches): differentiable tensor ops, see linearStyleTransfer.py
            return torch.sigmoid(conv(x))
        params = {"decoder.feat_2_rgb_list.0.weight": conv.weight,
                  "decoder.feat_2_rgb_list.0.bias": conv.bias}
        key = ops.StyleWeightsRef.version_key(params)
        if getattr(self, '_sw_key', None) != key:
            self._sw = ops.StyleWeightsRef(params)
            self._sw_key = key
        return ops.style_forward(self._sw, x, None)


def get_renderer(args):
    """Reference nerf_decoder_stylenerf.py:452-458; only reached when ``encode_a`` is
    off, which ``opt.py:84`` makes impossible (default=True, store_true)."""
    if args.model_mode == '1-1':
        return NeuralRenderer(img_size=(args.img_wh[0], args.img_wh[1]),
                              featmap_size=(args.img_wh[0], args.img_wh[1]),
                              feat_nc=args.nerf_out_dim, out_dim=3, args_here=args)
    raise NotImplementedError("model_mode '1-4-1' (NeuralRenderer_11_tanh) is not used by any "
                              "reference command and has no kernel")
