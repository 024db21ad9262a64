"""``models.linearStyleTransfer`` of the reference, backed by sm_100a kernels.

Hot-path classes (reference file:line): ``CNN`` (:6-37), ``MulLayer`` (:43-94) and
``style_net`` (:278-291) hold the same parameters under the same names
(``multi_net.{snet,cnet}.convs.{0,2,4}``, ``.fc``, ``multi_net.compress``,
``multi_net.unzip``, ``decoder.feat_2_rgb_list.0``) and run the cross-ray fusion +
decoder as four launches with two streaming passes over the feature map (csrc/crossray.cu,
csrc/gram_tc.cu) instead of ~40 library launches; under autograd ``style_net.forward`` uses the
same kernels forward and csrc/style_backward.cu backward (crnerf_b200.autograd.StyleNetFn).  The feature map is read in place whether it is
contiguous NCHW or the transposed view of the renderer's (N,64) rows that the
reference's callers build.

``encoder_sameoutputsize`` (:208-276, the style/content encoder ``enc_a``; SURVEY.md 8f
"next" row 1) keeps the reference's parameter names; its inference forward runs
csrc/encoder.cu, its training forward/backward stays on library ops.
"""
import torch
import torch.nn as nn

from crnerf_b200 import ops, torch_ops
from models.nerf_decoder_stylenerf import NeuralRenderer


import torch.nn.functional as F


# "native": style_net under autograd runs the library's own forward / backward kernels;
# "torch": the differentiable tensor-op restatement below (kept for MulLayer / CNN used on their own
# and as the yardstick of tests/test_gpu_style_backward.py)
TRAIN_BACKEND = "native"


def _wants_grad(module, *tensors):
    """True when the call must be recorded for autograd (the training step's decode())."""
    return torch.is_grad_enabled() and (any(t is not None and t.requires_grad for t in tensors) or
                                        any(p.requires_grad for p in module.parameters()))


# Tensor-op restatement of the cross-ray block under autograd.  ``style_net.forward`` itself - the
# decode() of the training step (train_mask_grid_sample.py:127-149, 32x32 patches) - uses the
# library's own kernels in both directions (crnerf_b200.autograd.StyleNetFn); these functions serve
# ``MulLayer`` / ``CNN`` called on their own under autograd, and the tests.
def _cnn_torch(cnn, x):
    """CNN.forward (reference linearStyleTransfer.py:28-37) as differentiable tensor ops."""
    y = cnn.convs(x)
    b, c, h, w = y.shape
    y = y.reshape(b, c, h * w)
    gram = torch.bmm(y, y.transpose(1, 2)) / (h * w)
    return cnn.fc(gram.reshape(b, -1))


def _mul_layer_torch(m, cF, sF):
    """MulLayer.forward with trans=True (reference linearStyleTransfer.py:58-90)."""
    cb, cc, ch, cw = cF.shape
    c_mean = cF.reshape(cb, cc, -1).mean(dim=2).reshape(cb, cc, 1, 1)
    cFc = cF - c_mean
    sb, sc, _, _ = sF.shape
    s_mean = sF.reshape(sb, sc, -1).mean(dim=2).reshape(sb, sc, 1, 1)
    sFc = sF - s_mean
    comp = m.compress(cFc)
    comp = comp.reshape(cb, comp.shape[1], -1)
    c_mat = _cnn_torch(m.cnet, cFc).reshape(cb, m.matrixSize, m.matrixSize)
    s_mat = _cnn_torch(m.snet, sFc).reshape(sb, m.matrixSize, m.matrixSize)
    trans = torch.bmm(s_mat, c_mat)
    y = torch.bmm(trans, comp).reshape(cb, m.matrixSize, ch, cw)
    return m.unzip(y) + s_mean, trans


class _StyleParamsMixin:
    """Caches the C struct of parameter pointers; rebuilt when a parameter moves or changes."""

    def _style_ref(self, prefix_map):
        params = {}
        for name, p in self.named_parameters():
            for src, dst in prefix_map:
                if name.startswith(src):
                    params[dst + name[len(src):]] = p
                    break
        key = ops.StyleWeightsRef.version_key(params)
        if getattr(self, '_sw_key', None) != key:
            self._sw = ops.StyleWeightsRef(params)
            self._sw_key = key
        return self._sw


class CNN(nn.Module, _StyleParamsMixin):
    def __init__(self, matrixSize=32, in_channel=64):
        """Three 1x1 convs + Gram + fc, reference linearStyleTransfer.py:7-25."""
        super(CNN, self).__init__()
        self.convs = nn.Sequential(nn.Conv2d(in_channel, 128, 1, 1, 0),
                                   nn.LeakyReLU(0.2, inplace=True),
                                   nn.Conv2d(128, 64, 1, 1, 0),
                                   nn.LeakyReLU(0.2, inplace=True),
                                   nn.Conv2d(64, matrixSize, 1, 1, 0))
        self.fc = nn.Linear(matrixSize * matrixSize, matrixSize * matrixSize)

    def forward(self, x):
        """(1,64,H,W) -> (1,1024), reference linearStyleTransfer.py:28-37."""
        if _wants_grad(self, x):
            return _cnn_torch(self, x)
        return ops.cnn_forward(self._style_ref([("", "multi_net.cnet.")]), "cnet", x)


class MulLayer(nn.Module, _StyleParamsMixin):
    def __init__(self, matrixSize=32, in_channel=64):
        """Reference linearStyleTransfer.py:44-56 (snet/cnet use the default in_channel=64)."""
        super(MulLayer, self).__init__()
        self.snet = CNN(matrixSize)
        self.cnet = CNN(matrixSize)
        self.matrixSize = matrixSize
        self.compress = nn.Conv2d(in_channel, matrixSize, 1, 1, 0)
        self.unzip = nn.Conv2d(matrixSize, in_channel, 1, 1, 0)
        self.transmatrix = None

    def forward(self, cF, sF, trans=True):
        """content (1,64,H,W), style (1,64,h,w) -> (fused (1,64,H,W), transmatrix (1,32,32)),
        reference linearStyleTransfer.py:58-90."""
        if not trans:
            # the reference's trans=False branch returns None (bare `return`, :91-93)
            return None
        if _wants_grad(self, cF, sF):
            return _mul_layer_torch(self, cF, sF)
        sw = self._style_ref([("", "multi_net.")])
        if not sw.has_decoder:
            # stand-alone MulLayer: the kernel still needs a 64->3 head to write; use zeros
            dev = cF.device
            params = dict(sw.tensors)
            params["decoder.feat_2_rgb_list.0.weight"] = torch.zeros(3, 64, 1, 1, device=dev)
            params["decoder.feat_2_rgb_list.0.bias"] = torch.zeros(3, device=dev)
            sw = ops.StyleWeightsRef(params)
        _, trans_m, fused = ops.style_forward(sw, cF, sF, want_trans=True, want_fused=True)
        return fused, trans_m


class style_net(nn.Module, _StyleParamsMixin):
    def __init__(self, args, residual_blocks=2):
        """Reference linearStyleTransfer.py:279-283."""
        super(style_net, self).__init__()
        nerf_channel = args.nerf_out_dim
        self.multi_net = MulLayer(in_channel=nerf_channel)
        self.decoder = NeuralRenderer(img_size=(args.img_wh[0], args.img_wh[1]),
                                      featmap_size=(args.img_wh[0], args.img_wh[1]),
                                      feat_nc=args.nerf_out_dim, out_dim=3, args_here=args)
        self._style_args = None

    def forward(self, content_feature, style_feature, type=None, channel_sums=None):
        """content (1,64,H,W), style (1,64,32,32) or None -> rgb (1,3,H,W),
        reference linearStyleTransfer.py:284-291.  One extension: ``channel_sums`` (rows, 64), the
        ``chansum_*`` entry of ``render_rays_cross_ray(..., channel_sums=True)`` for the same
        features - the cross-ray block then reads the feature map twice instead of three times."""
        if _wants_grad(self, content_feature, style_feature):
            if style_feature is None and type == "content":
                return self.decoder(content_feature)
            if content_feature.is_cuda and TRAIN_BACKEND == "native":
                # decode() of the training step: the inference kernels forward, csrc/style_backward.cu backward
                from crnerf_b200 import autograd as crnerf_autograd
                return crnerf_autograd.style_net_forward(self, self._style_ref([("", "")]), content_feature,
                                                         style_feature, channel_sums)
            fused, _ = _mul_layer_torch(self.multi_net, content_feature, style_feature)
            return self.decoder(fused)
        # torch.ops.crnerf.style_forward (crnerf_b200/torch_ops.py) -> crnerf_style_forward
        if self._style_args is None:
            self._style_args = torch_ops.style_params(self)
        if style_feature is None and type == "content":
            return torch.ops.crnerf.style_forward(content_feature, None, self._style_args)
        return torch.ops.crnerf.style_forward(content_feature, style_feature, self._style_args, channel_sums)

    def _apply(self, fn, *a, **k):
        self._style_args = None            # .to()/.cuda() may replace the parameter tensors
        return super()._apply(fn, *a, **k)


class encoder_sameoutputsize(nn.Module):
    """Style/content encoder ``enc_a`` / ``enc_cont`` (reference :208-276): six
    reflection-padded 3x3 convs with LeakyReLU(0.2), two 2x2 max-pools, adaptive
    average pool to 32x32 and a 1x1 conv.

    CUDA input, batch 1, ``out_channel == 64``: csrc/encoder.cu - the 3x3 convolutions as tcgen05
    implicit GEMMs with fp16 hi/lo split operands (fp32-class accuracy).  Under autograd (the training
    step back-propagates through ``enc_a``) the same kernels run with the activation planes kept and
    csrc/encoder_train.cuh computes the gradients (``crnerf_b200.autograd.EncoderFn``);
    ``train_backend = "library"`` selects differentiable library ops instead."""

    def __init__(self, out_channel=64):
        super(encoder_sameoutputsize, self).__init__()
        self.conv1 = nn.Conv2d(3, 3, 1, 1, 0)
        self.reflecPad1 = nn.ReflectionPad2d((1, 1, 1, 1))
        self.conv2 = nn.Conv2d(3, 64, 3, 1, 0)
        self.relu2 = nn.LeakyReLU(0.2, inplace=True)
        self.reflecPad3 = nn.ReflectionPad2d((1, 1, 1, 1))
        self.conv3 = nn.Conv2d(64, 64, 3, 1, 0)
        self.relu3 = nn.LeakyReLU(0.2, inplace=True)
        self.maxPool = nn.MaxPool2d(kernel_size=2, stride=2, return_indices=True)
        self.reflecPad4 = nn.ReflectionPad2d((1, 1, 1, 1))
        self.conv4 = nn.Conv2d(64, 128, 3, 1, 0)
        self.relu4 = nn.LeakyReLU(0.2, inplace=True)
        self.reflecPad5 = nn.ReflectionPad2d((1, 1, 1, 1))
        self.conv5 = nn.Conv2d(128, 128, 3, 1, 0)
        self.relu5 = nn.LeakyReLU(0.2, inplace=True)
        self.maxPool2 = nn.MaxPool2d(kernel_size=2, stride=2, return_indices=True)
        self.reflecPad6 = nn.ReflectionPad2d((1, 1, 1, 1))
        self.conv6 = nn.Conv2d(128, 128, 3, 1, 0)
        self.relu6 = nn.LeakyReLU(0.2, inplace=True)
        self.adppool = nn.AdaptiveAvgPool2d(32)
        self.conv7 = nn.Conv2d(128, out_channel, 1, 1, 0)
        self.relu7 = nn.LeakyReLU(0.2, inplace=True)
        self._packed = None
        self._packed_key = None
        # precision of the library convolutions under autograd: "fp32" (the reference's CPU / fp32
        # results to 1e-4; default) or "tf32" (cuDNN's default path, what the reference itself gets on
        # a GPU: ~3e-3 relative, about 3x faster at photo sizes - tools/bench_train_full.py)
        self.conv_precision = "fp32"
        # under autograd: "native" = csrc/encoder.cu forward + csrc/encoder_train.cuh backward (no library
        # convolution); "library" = differentiable torch ops at ``conv_precision`` (kept for A/B checks)
        self.train_backend = "native"

    def _convs(self):
        return [self.conv1, self.conv2, self.conv3, self.conv4, self.conv5, self.conv6, self.conv7]

    def packed(self):
        """Weight image for the kernels, rebuilt when a parameter changes."""
        ps = [p for c in self._convs() for p in (c.weight, c.bias)]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._packed is None or key != self._packed_key:
            self._packed = ops.pack_encoder(ps[0::2], ps[1::2])
            self._packed_key = key
        return self._packed

    def forward(self, x):
        if x.is_cuda and self.conv7.out_channels == 64 and not _wants_grad(self, x) and x.shape[0] == 1 \
                and min(x.shape[2:]) >= 8:
            return ops.encoder_forward(self.packed(), x)
        if x.is_cuda and torch.is_grad_enabled() and self.train_backend == "native" \
                and self.conv7.out_channels == 64 and x.shape[0] == 1 and min(x.shape[2:]) >= 8:
            # training step: the inference kernels forward (activation planes kept), csrc/encoder_train.cuh backward
            from crnerf_b200.autograd import EncoderFn
            self._packed = None    # weights move every step, and p.data updates do not bump _version: always re-pack
            return EncoderFn.apply(self.packed(), x, *[p for c in self._convs() for p in (c.weight, c.bias)])
        if x.is_cuda and torch.is_grad_enabled() and self.conv_precision == "fp32":
            # library convolutions, fp32 in forward and backward (scoped, not global)
            from crnerf_b200.autograd import Fp32Region
            return Fp32Region.apply(self._stack, x, *self.parameters())
        return self._stack(x)

    def _stack(self, x):
        h = self.relu2(self.conv2(self.reflecPad1(self.conv1(x))))
        h = self.relu3(self.conv3(self.reflecPad3(h)))
        h, _ = self.maxPool(h)
        h = self.relu4(self.conv4(self.reflecPad4(h)))
        h = self.relu5(self.conv5(self.reflecPad5(h)))
        h, _ = self.maxPool2(h)
        h = self.relu6(self.conv6(self.reflecPad6(h)))
        return self.relu7(self.conv7(self.adppool(h)))


class _NotOnHotPath(nn.Module):
    _why = ""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError(self._why)


class encoder3(_NotOnHotPath):
    """Importable placeholder for reference linearStyleTransfer.py:97-150 (unused there)."""
    _why = "models.linearStyleTransfer.encoder3 is imported but never instantiated by the reference"


class decoder3(_NotOnHotPath):
    """Importable placeholder for reference linearStyleTransfer.py:152-206 (unused there)."""
    _why = "models.linearStyleTransfer.decoder3 is never instantiated by the reference"
