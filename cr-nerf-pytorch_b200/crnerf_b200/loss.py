"""Host side of csrc/loss.cu: the loss + mask tail of the reference's training step
(losses.py:50-89, train_mask_grid_sample.py:171-175) as three autograd Functions, each
one forward launch and one backward launch.  CUDA fp32 only; no CPU fallback."""
import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .ops import CrnerfError, _c, _need, _p, _stream, check

_scratch = {}

MODE_SQUARE, MODE_ABS_DIFF, MODE_SQ_DIFF = 0, 1, 2


def _loss_scratch(dev: torch.device) -> torch.Tensor:
    """One zero-initialised scratch per (device, stream): the kernels leave its counters zero."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _scratch.get(key)
    if buf is None:
        n = int(_lib.load().crnerf_loss_scratch_floats())
        if n == 0:
            raise CrnerfError("crnerf_b200 needs an sm_100 device")
        buf = _scratch[key] = torch.zeros(n, dtype=torch.float32, device=dev)
    return buf


def _rows3(t: torch.Tensor, name: str, n: Optional[int] = None) -> torch.Tensor:
    _need(t, name, 2)
    if t.shape[1] != 3 or (n is not None and t.shape[0] != n):
        raise ValueError(f"{name} must be ({'n' if n is None else n}, 3), got {tuple(t.shape)}")
    return _c(t)


class RayLossFn(torch.autograd.Function):
    """out (4,) = [c_l, f_l, r_ms, r_md] of CRNeRFLoss (reference losses.py:62-76)."""

    @staticmethod
    def forward(ctx, rgb_coarse, rgb_fine, targets, mask, coef, size_delta, digit_delta):
        lib = _lib.load()
        coarse = _rows3(rgb_coarse, "rgb_coarse")
        n = coarse.shape[0]
        if n == 0:
            raise ValueError("empty batch")
        fine = None if rgb_fine is None else _rows3(rgb_fine, "rgb_fine", n)
        tgt = _rows3(targets, "targets", n)
        m = None
        if mask is not None:
            _need(mask, "mask")
            if mask.numel() != n:
                raise ValueError(f"mask must hold one value per ray ({n}), got {tuple(mask.shape)}")
            m = _c(mask).reshape(n)
        dev = coarse.device
        out = torch.empty(4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.crnerf_ray_loss_forward(_p(coarse), _p(fine), _p(tgt), _p(m), n, coef, size_delta, digit_delta,
                                              _p(out), _p(_loss_scratch(dev)), _stream(dev)))
        ctx.save_for_backward(coarse, fine, tgt, m)
        ctx.k = (coef, size_delta, digit_delta)
        ctx.mask_shape = None if mask is None else mask.shape
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        lib = _lib.load()
        coarse, fine, tgt, m = ctx.saved_tensors
        n, dev = coarse.shape[0], coarse.device
        need_c, need_f, _, need_m = ctx.needs_input_grad[:4]
        g_c = torch.empty_like(coarse) if need_c else None
        g_f = torch.empty_like(fine) if (need_f and fine is not None) else None
        g_m = torch.empty(n, dtype=torch.float32, device=dev) if (need_m and m is not None) else None
        go = _c(go.float())
        with torch.cuda.device(dev):
            check(lib.crnerf_ray_loss_backward(_p(coarse), _p(fine), _p(tgt), _p(m), n, *ctx.k, _p(go), _p(g_c),
                                               _p(g_f), _p(g_m), _stream(dev)))
        if g_m is not None:
            g_m = g_m.reshape(ctx.mask_shape)
        return g_c, g_f, None, g_m, None, None, None


def ray_loss(rgb_coarse, rgb_fine, targets, mask, coef=1.0, size_delta=0.0, digit_delta=0.0) -> torch.Tensor:
    return RayLossFn.apply(rgb_coarse, rgb_fine, targets, mask, float(coef), float(size_delta), float(digit_delta))


def _term_arrays(modes, scales, tensors):
    k = len(modes)
    arr_p = C.c_void_p * k
    a = arr_p(*[tensors[2 * i].data_ptr() for i in range(k)])
    b = arr_p(*[None if tensors[2 * i + 1] is None else tensors[2 * i + 1].data_ptr() for i in range(k)])
    n = (C.c_int64 * k)(*[tensors[2 * i].numel() for i in range(k)])
    mode = (C.c_int * k)(*modes)
    scale = (C.c_float * k)(*scales)
    return a, b, n, mode, scale


class PairLossFn(torch.autograd.Function):
    """out (k,) with out[i] = scale_i * mean(f_i(a_i, b_i)) - the embedding terms kl_a,
    rec_a_random, content_constraint of CRNeRFLoss (reference losses.py:52-58, :67-68)."""

    @staticmethod
    def forward(ctx, modes, scales, *tensors):
        lib = _lib.load()
        k = len(modes)
        if not 1 <= k <= 4 or len(tensors) != 2 * k or len(scales) != k:
            raise ValueError("1..4 terms, two tensors (a, b or None) per term")
        flat = []
        for i in range(k):
            a, b = tensors[2 * i], tensors[2 * i + 1]
            _need(a, f"a[{i}]")
            if modes[i] == MODE_SQUARE:
                b = None
            else:
                _need(b, f"b[{i}]")
                if b.shape != a.shape:
                    raise ValueError(f"term {i}: shapes {tuple(a.shape)} vs {tuple(b.shape)}")
            if a.numel() == 0:
                raise ValueError(f"term {i} is empty")
            flat += [_c(a), None if b is None else _c(b)]
        dev = flat[0].device
        out = torch.empty(k, dtype=torch.float32, device=dev)
        a, b, n, mode, scale = _term_arrays(modes, scales, flat)
        with torch.cuda.device(dev):
            check(lib.crnerf_pair_loss_forward(k, a, b, n, mode, scale, _p(out), _p(_loss_scratch(dev)),
                                               _stream(dev)))
        ctx.save_for_backward(*[t for t in flat if t is not None])
        ctx.has_b = [t is not None for t in flat[1::2]]
        ctx.modes, ctx.scales = list(modes), list(scales)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        k = len(ctx.modes)
        flat = []
        for i in range(k):
            flat.append(saved.pop(0))
            flat.append(saved.pop(0) if ctx.has_b[i] else None)
        dev = flat[0].device
        needs = ctx.needs_input_grad[2:]
        grads = [torch.empty_like(flat[j]) if (flat[j] is not None and needs[j]) else None for j in range(2 * k)]
        a, b, n, mode, scale = _term_arrays(ctx.modes, ctx.scales, flat)
        arr_p = C.c_void_p * k
        ga = arr_p(*[_p(grads[2 * i]) for i in range(k)])
        gb = arr_p(*[_p(grads[2 * i + 1]) for i in range(k)])
        go = _c(go.float())
        with torch.cuda.device(dev):
            check(lib.crnerf_pair_loss_backward(k, a, b, n, mode, scale, _p(go), ga, gb, _stream(dev)))
        return (None, None, *grads)


def pair_losses(terms: Sequence) -> torch.Tensor:
    """terms: [(mode, scale, a, b_or_None), ...] (at most 4) -> (len(terms),) tensor."""
    modes = [int(t[0]) for t in terms]
    scales = [float(t[1]) for t in terms]
    tensors = []
    for t in terms:
        tensors += [t[2], t[3]]
    return PairLossFn.apply(modes, scales, *tensors)


class MaskSampleFn(torch.autograd.Function):
    """interpolate(pred, size=hw, mode='bilinear', align_corners=False) -> '(h w) c' rows -> [idx]
    (reference train_mask_grid_sample.py:172-175), evaluated only at the sampled pixels."""

    @staticmethod
    def forward(ctx, pred, hw, idx):
        lib = _lib.load()
        _need(pred, "pred", 4)
        if pred.shape[0] != 1:
            raise ValueError("pred must be (1, C, h, w)")
        H, W = int(hw[0]), int(hw[1])
        _, ch, h, w = pred.shape
        pred_c = _c(pred)
        dev = pred.device
        if idx is not None:
            if idx.dtype != torch.int64 or idx.device != dev:
                raise TypeError("idx must be an int64 tensor on pred's device")
            idx = _c(idx.reshape(-1))
            n = idx.numel()
        else:
            n = H * W
        out = torch.empty((n, ch), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.crnerf_mask_sample_forward(_p(pred_c), ch, h, w, H, W, _p(idx), n, _p(out), _stream(dev)))
        ctx.save_for_backward(idx)
        ctx.geom = (ch, h, w, H, W, n)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        lib = _lib.load()
        (idx,) = ctx.saved_tensors
        ch, h, w, H, W, n = ctx.geom
        g_out = _c(g_out.float())
        dev = g_out.device
        g_pred = torch.empty((1, ch, h, w), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.crnerf_mask_sample_backward(_p(g_out), ch, h, w, H, W, _p(idx), n, _p(g_pred), _stream(dev)))
        return g_pred, None, None


def mask_sample(pred: torch.Tensor, hw, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    return MaskSampleFn.apply(pred, hw, idx)
