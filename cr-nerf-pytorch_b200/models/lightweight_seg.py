"""``models.lightweight_seg`` of the reference (models/lightweight_seg.py): the Context Guided
Network that predicts the transient-object mask from the whole training photo
(train_mask_grid_sample.py:114, 170-176: ``Context_Guided_Network(classes=1, M=2, N=2,
input_channel=3)``).

SURVEY.md 8f-3 puts this network on library convolutions (it is a stack of stride-2 / depth-wise /
dilated 3x3 convs with BatchNorm + PReLU on one ~0.1 MPixel photo per step - launch-bound, not
tensor- or HBM-bound), so the layers are ``torch.nn`` modules running on cuDNN; what this mirror
owns is

* the module tree: attribute names, registration order and initialisers reproduce the reference's,
  so ``state_dict()`` keys / shapes, seeded default init (construction draws, then the
  ``kaiming_normal_`` pass over ``modules()``, lightweight_seg.py:307-317) and checkpoints are
  interchangeable (tests/test_cgnet.py pins all three against the unmodified file);
* fp32 convolutions in the forward AND in the backward (the reference's CPU / fp32 results are the
  parity bar, 1e-4; cuDNN's default TF32 path is off by 1e-3).  The switch is scoped, not global:
  the conv stack runs inside one autograd node (``_Fp32Region``) that builds the inner graph under
  ``cudnn.flags(allow_tf32=False)`` and differentiates it under the same flags, so the caller's
  process-wide setting is never touched;
* a capture-safe forward / backward (no host sync, no data-dependent shapes), so the whole training
  step including this network replays as one CUDA graph (``crnerf_b200.graphs.GraphedTrainStep``);
* ``mask_rows``: the caller's tail (train_mask_grid_sample.py:171-175 - second bilinear resize,
  ``'1 n h w -> (h w) n'``, ``[rgb_idx]``) evaluated only at the sampled pixels by
  ``crnerf_mask_sample_*`` instead of materialising the resized mask.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from crnerf_b200.autograd import Fp32Region as _Fp32Region

__all__ = ["Context_Guided_Network"]

_BN_EPS = 1e-3


def _conv(n_in, n_out, k, stride=1, dilation=1, groups=1):
    pad = ((k - 1) // 2) * dilation
    return nn.Conv2d(n_in, n_out, (k, k), stride=stride, padding=(pad, pad), dilation=dilation, groups=groups,
                     bias=False)


class _Wrapped(nn.Module):
    """A single convolution registered as ``.conv`` (the reference wraps every bare conv in a
    one-attribute module: Conv, ChannelWiseConv, ChannelWiseDilatedConv, lightweight_seg.py:83-168)."""

    def __init__(self, conv):
        super().__init__()
        self.conv = conv

    def forward(self, x):
        return self.conv(x)


class ConvBNPReLU(nn.Module):                      # lightweight_seg.py:13-37
    def __init__(self, nIn, nOut, kSize, stride=1):
        super().__init__()
        self.conv = _conv(nIn, nOut, kSize, stride)
        self.bn = nn.BatchNorm2d(nOut, eps=_BN_EPS)
        self.act = nn.PReLU(nOut)

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class BNPReLU(nn.Module):                          # lightweight_seg.py:40-58
    def __init__(self, nOut):
        super().__init__()
        self.bn = nn.BatchNorm2d(nOut, eps=_BN_EPS)
        self.act = nn.PReLU(nOut)

    def forward(self, x):
        return self.act(self.bn(x))


class FGlo(nn.Module):                             # lightweight_seg.py:170-188: squeeze-and-excite gate
    def __init__(self, channel, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel), nn.Sigmoid())

    def forward(self, x):
        gate = self.fc(self.avg_pool(x).flatten(1))
        return x * gate[:, :, None, None]


class ContextGuidedBlock_Down(nn.Module):          # lightweight_seg.py:190-224: (H,W,C) -> (H/2,W/2,nOut)
    def __init__(self, nIn, nOut, dilation_rate=2, reduction=16):
        super().__init__()
        self.conv1x1 = ConvBNPReLU(nIn, nOut, 3, 2)
        self.F_loc = _Wrapped(_conv(nOut, nOut, 3, groups=nOut))
        self.F_sur = _Wrapped(_conv(nOut, nOut, 3, dilation=dilation_rate, groups=nOut))
        self.bn = nn.BatchNorm2d(2 * nOut, eps=_BN_EPS)
        self.act = nn.PReLU(2 * nOut)
        self.reduce = _Wrapped(_conv(2 * nOut, nOut, 1))
        self.F_glo = FGlo(nOut, reduction)

    def forward(self, x):
        y = self.conv1x1(x)
        joint = self.act(self.bn(torch.cat([self.F_loc(y), self.F_sur(y)], 1)))
        return self.F_glo(self.reduce(joint))


class ContextGuidedBlock(nn.Module):               # lightweight_seg.py:227-257
    def __init__(self, nIn, nOut, dilation_rate=2, reduction=16, add=True):
        super().__init__()
        n = int(nOut / 2)
        self.conv1x1 = ConvBNPReLU(nIn, n, 1, 1)
        self.F_loc = _Wrapped(_conv(n, n, 3, groups=n))
        self.F_sur = _Wrapped(_conv(n, n, 3, dilation=dilation_rate, groups=n))
        self.bn_prelu = BNPReLU(nOut)
        self.add = add
        self.F_glo = FGlo(nOut, reduction)

    def forward(self, x):
        y = self.conv1x1(x)
        out = self.F_glo(self.bn_prelu(torch.cat([self.F_loc(y), self.F_sur(y)], 1)))
        return x + out if self.add else out


class InputInjection(nn.Module):                   # lightweight_seg.py:259-268: image pyramid by 3x3/2 average pools
    def __init__(self, downsamplingRatio):
        super().__init__()
        self.pool = nn.ModuleList(nn.AvgPool2d(3, stride=2, padding=1) for _ in range(downsamplingRatio))

    def forward(self, x):
        for pool in self.pool:
            x = pool(x)
        return x


class Context_Guided_Network(nn.Module):
    """CGNet (lightweight_seg.py:271-368).  ``forward(img (B,C,H,W)) -> sigmoid mask (B,classes,H,W)``."""

    def __init__(self, classes=19, M=3, N=21, input_channel=64, dropout_flag=False):
        super().__init__()
        c = input_channel
        self.level1_0 = ConvBNPReLU(c, 32, 3, 2)
        self.level1_1 = ConvBNPReLU(32, 32, 3, 1)
        self.level1_2 = ConvBNPReLU(32, 32, 3, 1)
        self.sample1 = InputInjection(1)
        self.sample2 = InputInjection(2)
        self.b1 = BNPReLU(32 + c)
        self.level2_0 = ContextGuidedBlock_Down(32 + c, 64, dilation_rate=2, reduction=8)
        self.level2 = nn.ModuleList(ContextGuidedBlock(64, 64, dilation_rate=2, reduction=8) for _ in range(M - 1))
        self.bn_prelu_2 = BNPReLU(128 + c)
        self.level3_0 = ContextGuidedBlock_Down(128 + c, 128, dilation_rate=4, reduction=16)
        self.level3 = nn.ModuleList(ContextGuidedBlock(128, 128, dilation_rate=4, reduction=16) for _ in range(N - 1))
        self.bn_prelu_3 = BNPReLU(256)
        head = _Wrapped(_conv(256, classes, 1))
        self.classifier = nn.Sequential(nn.Dropout2d(0.1, False), head) if dropout_flag else nn.Sequential(head)
        self.sigmoid = nn.Sigmoid()
        self.conv_precision = "fp32"               # or "tf32": cuDNN's default path (see encoder_sameoutputsize)
        for m in self.modules():                   # :307-317 (same visiting order -> same RNG consumption)
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)

    def _stack(self, x):
        s1 = self.level1_2(self.level1_1(self.level1_0(x)))
        inp1 = self.sample1(x)
        inp2 = self.sample2(x)
        s2_0 = self.level2_0(self.b1(torch.cat([s1, inp1], 1)))
        s2 = s2_0
        for blk in self.level2:
            s2 = blk(s2)
        s3_0 = self.level3_0(self.bn_prelu_2(torch.cat([s2, s2_0, inp2], 1)))
        s3 = s3_0
        for blk in self.level3:
            s3 = blk(s3)
        return self.classifier(self.bn_prelu_3(torch.cat([s3_0, s3], 1)))

    def logits(self, x):
        """Everything up to the 1x1 classifier: (B, classes, H/8, W/8), fp32 convolutions."""
        if not x.is_cuda or self.conv_precision != "fp32":
            return self._stack(x)
        if not torch.is_grad_enabled():
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return self._stack(x)
        return _Fp32Region.apply(self._stack, x, *self.parameters())

    def forward(self, input):
        z = self.logits(input)
        up = F.interpolate(z, input.size()[2:], mode='bilinear', align_corners=False)     # :365
        return self.sigmoid(up)

    def mask_rows(self, whole_img, hw_whole, rgb_idx=None):
        """``pred_mask`` rows as the training step consumes them (train_mask_grid_sample.py:171-175):
        forward -> bilinear resize to ``hw_whole`` -> ``(h w) n`` -> ``[rgb_idx]``; the resize and the
        gather are one kernel that touches only the sampled pixels (differentiable)."""
        from crnerf_b200 import loss as crnerf_loss
        return crnerf_loss.mask_sample(self.forward(whole_img), tuple(hw_whole), rgb_idx)
