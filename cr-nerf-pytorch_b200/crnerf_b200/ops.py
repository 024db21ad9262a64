"""Tensor-level wrappers over the C ABI.

PyTorch is used for device memory and streams only: every function validates
its tensors, allocates outputs with ``torch.empty`` on the input's device and
enqueues the library's kernels on ``torch.cuda.current_stream()``.  CPU tensors
are rejected - there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import CrnerfError, check

OPERAND_FP16 = 0
OPERAND_BF16 = 1
OPERAND_FP16X3 = 2      # split precision: hi + lo fp16 operands, three MMAs per product (inference)
_OPERANDS = {"fp16": 0, "bf16": 1, "fp16x3": 2, 0: 0, 1: 1, 2: 2}

# order of the 12 NeRF_sigma layers in crnerf_mlp_weights (reference state_dict prefixes)
MLP_LAYER_KEYS = tuple([f"xyz_encoding_{i}.0" for i in range(1, 9)] +
                       ["xyz_encoding_final", "dir_encoding.0", "static_rgb.0", "static_sigma.0"])


def operand_id(op) -> int:
    try:
        return _OPERANDS[op]
    except KeyError:
        raise ValueError(f"unknown operand format {op!r} (use 'fp16', 'bf16' or 'fp16x3')") from None


def _need(t: torch.Tensor, name: str, dims: Optional[int] = None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise CrnerfError(f"{name} is on {t.device}: crnerf_b200 runs on CUDA (sm_100) only and has "
                          "no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if dims is not None and t.dim() != dims:
        raise ValueError(f"{name} must have {dims} dims, got shape {tuple(t.shape)}")
    return t


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(_lib.load().crnerf_launch_count())


def device_ok() -> bool:
    return bool(_lib.load().crnerf_device_ok())


# --------------------------------------------------------------------------
class PackedMLP:
    """Tensor-core-ready image of one NeRF_sigma's weights (see csrc/nerf_layout.h)."""

    def __init__(self, buf: torch.Tensor, operand: int, e_xyz: int, e_dir: int, status=None):
        self.buf, self.operand, self.e_xyz, self.e_dir = buf, operand, e_xyz, e_dir
        self.status = status     # device int32: 1 if a weight left the operand format's range
        self._host = self._event = None
        # activation-overflow flag of the fp16 formats: one int32 in pinned host memory that the
        # render kernels write (zero-copy) when an operand saturates; read without any sync
        self.overflow = None

    _OVERFLOW_MSG = ("an activation or input reached the fp16 operand limit (|x| >= 65504) in a render pass "
                     "with these weights and was clamped; the results of that pass are not within tolerance. "
                     "Use operand='bf16' (args.crnerf_operand = 'bf16'), which has fp32 range")

    def overflow_ptr(self):
        if self.operand == OPERAND_BF16:
            return None
        if self.overflow is None:
            self.overflow = torch.zeros(1, dtype=torch.int32).pin_memory()
        return self.overflow.data_ptr()

    def check_overflow(self, sync: bool = False):
        """Raise if a render pass that used this image saturated an fp16 operand.  Without ``sync``
        it looks at what has been reported so far (free: a host memory read), so a caller in a
        loop learns of the problem one call later; ``sync=True`` waits for the device first."""
        if self.overflow is None:
            return
        if sync:
            torch.cuda.synchronize(self.buf.device)
        if int(self.overflow[0]) != 0:
            self.overflow[0] = 0
            raise CrnerfError(self._OVERFLOW_MSG)

    _RANGE_MSG = ("a NeRF weight exceeds the fp16 finite range (|w| > 65504) and was clamped by the "
                  "packer; pack with operand='bf16' (args.crnerf_operand = 'bf16')")

    def check_range(self):
        """Raise if the pack kernel saw a weight outside the fp16 range (host sync on first call)."""
        if self.status is not None:
            bad = int(self.status.item()) != 0
            self.status = None
            if bad:
                raise CrnerfError(self._RANGE_MSG)

    def defer_range(self):
        """Start an asynchronous read-back of the verdict (pinned host word + event): the training
        step must not synchronise, so the verdict is looked at by ``poll_range`` one step later."""
        if self.status is not None:
            self._host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._host.copy_(self.status, non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record(torch.cuda.current_stream(self.buf.device))

    def poll_range(self):
        """Non-blocking: raise if the deferred verdict has arrived and is bad."""
        if self._event is not None and self._event.query():
            bad = int(self._host[0]) != 0
            self._event = self._host = self.status = None
            if bad:
                raise CrnerfError(self._RANGE_MSG)

    @property
    def device(self):
        return self.buf.device


def pack_mlp(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], e_xyz: int,
             e_dir: int, operand="fp16", check_range=True) -> PackedMLP:
    """weights/biases: the 12 nn.Linear tensors in ``MLP_LAYER_KEYS`` order.
    ``check_range``: True = verify now (one host sync), "deferred" = leave the verdict on the
    returned object (``PackedMLP.check_range()``), False = do not check."""
    lib = _lib.load()
    op = operand_id(operand)
    if len(weights) != 12 or len(biases) != 12:
        raise ValueError("expected 12 weight and 12 bias tensors")
    dev = weights[0].device
    keep = []
    w = _lib.MlpWeights()
    in_f = [e_xyz] + [256] * 3 + [e_xyz + 256] + [256] * 3 + [256, 256 + e_dir, 128, 256]
    out_f = [256] * 9 + [128, 64, 1]
    for i in range(12):
        wi = _c(_need(weights[i].detach(), f"weight[{i}]", 2))
        bi = _c(_need(biases[i].detach(), f"bias[{i}]", 1))
        if tuple(wi.shape) != (out_f[i], in_f[i]) or tuple(bi.shape) != (out_f[i],):
            raise ValueError(
                f"layer {MLP_LAYER_KEYS[i]}: got weight {tuple(wi.shape)} bias {tuple(bi.shape)}, the "
                f"fused kernel is specialised for D=8, W=256, skips=[4], out_dim=64 and expects "
                f"({out_f[i]}, {in_f[i]})")
        if wi.device != dev or bi.device != dev:
            raise ValueError("all weights must live on one device")
        keep += [wi, bi]
        w.weight[i] = wi.data_ptr()
        w.bias[i] = bi.data_ptr()
    w.e_xyz, w.e_dir = e_xyz, e_dir
    nbytes = lib.crnerf_mlp_packed_bytes_op(e_xyz, e_dir, op)
    with torch.cuda.device(dev):
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev) if check_range else None
        check(lib.crnerf_mlp_pack(C.byref(w), op, buf.data_ptr(), nbytes, _p(status), _stream(dev)))
    del keep
    packed = PackedMLP(buf, op, e_xyz, e_dir, status)
    if check_range == "deferred":
        packed.defer_range()     # the caller polls later (training: no host sync per step)
        return packed
    packed.check_range()
    return packed


def render_partial_rows(n_rays: int, n_samples: int) -> int:
    return int(_lib.load().crnerf_render_partial_rows(int(n_rays), int(n_samples)))


def render_pass(packed: PackedMLP, rays: torch.Tensor, z_vals: torch.Tensor,
                noise: Optional[torch.Tensor] = None, view_dir: Optional[torch.Tensor] = None,
                n_freq_xyz: int = 15, n_freq_dir: int = 4, xyz_jitter: Optional[torch.Tensor] = None,
                want_channel_partials: bool = False, overflow_ptr: Optional[int] = None):
    """One fused pass: embed -> MLP -> composite.  Returns (weights, feature, depth), plus the
    (rows, 64) partial channel sums of ``feature`` with ``want_channel_partials`` (their sum over
    rows is ``feature.sum(0)``: the cross-ray block's mean without another pass over the map).
    ``xyz_jitter`` (n_rays*n_samples, 3) is added to the sample positions (args.pertubeCord)."""
    lib = _lib.load()
    rays = _c(_need(rays, "rays", 2))
    z_vals = _c(_need(z_vals, "z_vals", 2))
    n, s = z_vals.shape
    if rays.shape != (n, 8):
        raise ValueError(f"rays must be ({n}, 8), got {tuple(rays.shape)}")
    if 3 + 6 * n_freq_xyz != packed.e_xyz or 3 + 6 * n_freq_dir != packed.e_dir:
        raise ValueError("weights were packed for a different embedding width")
    if noise is not None:
        noise = _c(_need(noise, "noise", 2))
        if noise.shape != (n, s):
            raise ValueError("noise must match z_vals")
    if view_dir is not None:
        view_dir = _c(_need(view_dir, "view_dir", 2))
        if view_dir.shape != (n, 3):
            raise ValueError("view_dir must be (n_rays, 3)")
    if xyz_jitter is not None:
        xyz_jitter = _c(_need(xyz_jitter, "xyz_jitter", 2))
        if xyz_jitter.shape != (n * s, 3):
            raise ValueError(f"xyz_jitter must be ({n * s}, 3)")
    if overflow_ptr is None:         # the image's own flag; a caller that manages the flag passes it
        packed.check_overflow()      # what earlier passes with these weights reported (no sync)
        overflow_ptr = packed.overflow_ptr()
    dev = rays.device
    with torch.cuda.device(dev):
        weights = torch.empty((n, s), dtype=torch.float32, device=dev)
        feature = torch.empty((n, 64), dtype=torch.float32, device=dev)
        depth = torch.empty((n,), dtype=torch.float32, device=dev)
        partials = None
        if want_channel_partials:
            partials = torch.zeros((max(1, render_partial_rows(n, s)), 64), dtype=torch.float32, device=dev)
        if n == 0:
            return (weights, feature, depth, partials) if want_channel_partials else (weights, feature, depth)
        opts = _lib.RenderOpts(_p(xyz_jitter), _p(partials), overflow_ptr or None)
        check(lib.crnerf_render_pass_opts(packed.buf.data_ptr(), packed.operand, rays.data_ptr(),
                                          _p(view_dir), z_vals.data_ptr(), _p(noise), n, s, n_freq_xyz,
                                          n_freq_dir, weights.data_ptr(), feature.data_ptr(),
                                          depth.data_ptr(), C.byref(opts), _stream(dev)))
    return (weights, feature, depth, partials) if want_channel_partials else (weights, feature, depth)


def mlp_forward(packed: PackedMLP, x: torch.Tensor, sigma_only: bool = False) -> torch.Tensor:
    """NeRF_sigma.forward on embedded rows: (B, e_xyz+e_dir) -> (B, 65) (or (B,1))."""
    lib = _lib.load()
    x = _c(_need(x, "x", 2))
    need = packed.e_xyz if sigma_only else packed.e_xyz + packed.e_dir
    if x.shape[1] != need:
        raise ValueError(f"x must have {need} columns, got {x.shape[1]}")
    dev = x.device
    with torch.cuda.device(dev):
        out = torch.empty((x.shape[0], 1 if sigma_only else 65), dtype=torch.float32, device=dev)
        if x.shape[0] == 0:
            return out
        check(lib.crnerf_mlp_forward(packed.buf.data_ptr(), packed.operand, packed.e_xyz,
                                     packed.e_dir, x.data_ptr(), x.shape[0], x.shape[1],
                                     int(sigma_only), out.data_ptr(), _stream(dev)))
    return out


def pos_embed(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    lib = _lib.load()
    x = _c(_need(x, "x", 2))
    if x.shape[1] != 3:
        raise ValueError("PosEmbedding input must be (B, 3)")
    dev = x.device
    with torch.cuda.device(dev):
        out = torch.empty((x.shape[0], 3 + 6 * n_freqs), dtype=torch.float32, device=dev)
        if x.shape[0] == 0:
            return out
        check(lib.crnerf_pos_embed(x.data_ptr(), x.shape[0], n_freqs, out.data_ptr(), _stream(dev)))
    return out


def coarse_z(rays: torch.Tensor, t_steps: torch.Tensor, perturb_rand: Optional[torch.Tensor] = None,
             use_disp: bool = False) -> torch.Tensor:
    lib = _lib.load()
    rays = _c(_need(rays, "rays", 2))
    t_steps = _c(_need(t_steps, "t_steps", 1))
    n, s = rays.shape[0], t_steps.shape[0]
    if perturb_rand is not None:
        perturb_rand = _c(_need(perturb_rand, "perturb_rand", 2))
        if perturb_rand.shape != (n, s):
            raise ValueError("perturb_rand must be (n_rays, n_samples)")
    dev = rays.device
    with torch.cuda.device(dev):
        z = torch.empty((n, s), dtype=torch.float32, device=dev)
        if n == 0:
            return z
        check(lib.crnerf_coarse_z(rays.data_ptr(), t_steps.data_ptr(), _p(perturb_rand), n, s,
                                  int(use_disp), z.data_ptr(), _stream(dev)))
    return z


def _u_args(u: torch.Tensor, n: int, n_imp: int):
    u = _c(_need(u, "u"))
    if u.dim() == 1 and u.shape[0] == n_imp:
        return u, 0
    if u.dim() == 2 and u.shape == (n, n_imp):
        return u, n_imp
    raise ValueError(f"u must be ({n_imp},) or ({n}, {n_imp}), got {tuple(u.shape)}")


def sample_pdf_merge(z_coarse: torch.Tensor, weights_coarse: torch.Tensor, u: torch.Tensor,
                     n_importance: int, eps: float = 1e-5, return_new: bool = False):
    """z_fine = sort(cat(z_coarse, sample_pdf(mid(z_coarse), weights_coarse[:,1:-1], u)))."""
    lib = _lib.load()
    z_coarse = _c(_need(z_coarse, "z_coarse", 2))
    weights_coarse = _c(_need(weights_coarse.detach(), "weights_coarse", 2))
    n, s = z_coarse.shape
    if weights_coarse.shape != (n, s):
        raise ValueError("weights_coarse must match z_coarse")
    u, u_stride = _u_args(u, n, n_importance)
    dev = z_coarse.device
    with torch.cuda.device(dev):
        z_fine = torch.empty((n, s + n_importance), dtype=torch.float32, device=dev)
        z_new = torch.empty((n, n_importance), dtype=torch.float32, device=dev) if return_new else None
        if n == 0:
            return (z_fine, z_new) if return_new else z_fine
        check(lib.crnerf_sample_pdf_merge(z_coarse.data_ptr(), weights_coarse.data_ptr(),
                                          u.data_ptr(), u_stride, n, s, n_importance, eps,
                                          z_fine.data_ptr(), _p(z_new), _stream(dev)))
    return (z_fine, z_new) if return_new else z_fine


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, u: torch.Tensor, n_importance: int,
               eps: float = 1e-5) -> torch.Tensor:
    lib = _lib.load()
    bins = _c(_need(bins, "bins", 2))
    weights = _c(_need(weights.detach(), "weights", 2))
    n, m = weights.shape
    if bins.shape != (n, m + 1):
        raise ValueError("bins must be (n_rays, n_bins+1)")
    u, u_stride = _u_args(u, n, n_importance)
    dev = bins.device
    with torch.cuda.device(dev):
        out = torch.empty((n, n_importance), dtype=torch.float32, device=dev)
        if n == 0:
            return out
        check(lib.crnerf_sample_pdf(bins.data_ptr(), weights.data_ptr(), u.data_ptr(), u_stride, n,
                                    m, n_importance, eps, out.data_ptr(), _stream(dev)))
    return out


# --------------------------------------------------------------------------
class StyleWeightsRef:
    """Pointers to (a subset of) a style_net's parameters, kept alive by the
    tensors it holds.  ``params`` maps the reference's state_dict keys
    (``multi_net.cnet.convs.0.weight`` ... ``decoder.feat_2_rgb_list.0.bias``) to
    tensors; groups that are absent stay NULL (e.g. a bare NeuralRenderer only
    has the decoder keys, a bare MulLayer has no decoder)."""

    _SHAPES = {"convs.0.weight": (128, 64, 1, 1), "convs.0.bias": (128,),
               "convs.2.weight": (64, 128, 1, 1), "convs.2.bias": (64,),
               "convs.4.weight": (32, 64, 1, 1), "convs.4.bias": (32,),
               "fc.weight": (1024, 1024), "fc.bias": (1024,)}
    _TOP = {"multi_net.compress.weight": (32, 64, 1, 1), "multi_net.compress.bias": (32,),
            "multi_net.unzip.weight": (64, 32, 1, 1), "multi_net.unzip.bias": (64,),
            "decoder.feat_2_rgb_list.0.weight": (3, 64, 1, 1),
            "decoder.feat_2_rgb_list.0.bias": (3,)}

    def __init__(self, params: dict):
        self.tensors = {}
        self.device = None
        s = _lib.StyleWeights()

        def g(key, shape):
            t = params.get(key)
            if t is None:
                return None
            t = _c(_need(t.detach(), key))
            if tuple(t.shape) != shape:
                raise ValueError(f"{key}: expected {shape}, got {tuple(t.shape)} "
                                 "(the cross-ray kernels need nerf_out_dim == 64, matrixSize == 32)")
            if self.device is None:
                self.device = t.device
            elif t.device != self.device:
                raise ValueError("style_net parameters must live on one device")
            self.tensors[key] = t
            return t.data_ptr()

        for name, cw in (("cnet", s.cnet), ("snet", s.snet)):
            pre = f"multi_net.{name}."
            for i, j in enumerate((0, 2, 4)):
                cw.conv_w[i] = g(f"{pre}convs.{j}.weight", self._SHAPES[f"convs.{j}.weight"])
                cw.conv_b[i] = g(f"{pre}convs.{j}.bias", self._SHAPES[f"convs.{j}.bias"])
            cw.fc_w = g(f"{pre}fc.weight", self._SHAPES["fc.weight"])
            cw.fc_b = g(f"{pre}fc.bias", self._SHAPES["fc.bias"])
        s.compress_w = g("multi_net.compress.weight", self._TOP["multi_net.compress.weight"])
        s.compress_b = g("multi_net.compress.bias", self._TOP["multi_net.compress.bias"])
        s.unzip_w = g("multi_net.unzip.weight", self._TOP["multi_net.unzip.weight"])
        s.unzip_b = g("multi_net.unzip.bias", self._TOP["multi_net.unzip.bias"])
        s.rgb_w = g("decoder.feat_2_rgb_list.0.weight", self._TOP["decoder.feat_2_rgb_list.0.weight"])
        s.rgb_b = g("decoder.feat_2_rgb_list.0.bias", self._TOP["decoder.feat_2_rgb_list.0.bias"])
        self.has_fusion = all(k in self.tensors for k in
                              ["multi_net.cnet.fc.weight", "multi_net.snet.fc.weight",
                               "multi_net.compress.weight", "multi_net.unzip.weight"])
        self.has_decoder = "decoder.feat_2_rgb_list.0.weight" in self.tensors
        self.struct = s

    @staticmethod
    def version_key(params: dict):
        return tuple((k, t.data_ptr(), t._version) for k, t in sorted(params.items()))


def _feat_strides(t: torch.Tensor, name: str):
    """(1,64,H,W) feature map (any strides that keep H*W flat) -> (n_pixels, pix_stride, ch_stride).

    Accepts both a contiguous NCHW tensor and the transposed view the reference's
    callers build from the renderer's (N,64) rows (train_mask_grid_sample.py:133-134,
    eval.py:291-292), which is read in place."""
    _need(t, name, 4)
    if t.shape[0] != 1 or t.shape[1] != 64:
        raise ValueError(f"{name} must be (1, 64, H, W), got {tuple(t.shape)}")
    _, _, h, w = t.shape
    sb, sc, sh, sw = t.stride()
    if h * w > 0 and (h == 1 or sh == sw * w):
        return t, h * w, sw, sc
    t = t.contiguous()
    return t, h * w, 1, h * w


def style_forward(sw: StyleWeightsRef, content: torch.Tensor, style: Optional[torch.Tensor],
                  want_trans: bool = False, want_fused: bool = False,
                  channel_sums: Optional[torch.Tensor] = None):
    """style_net.forward: content (1,64,H,W), style (1,64,h,w) or None -> rgb (1,3,H,W).
    ``channel_sums`` (rows, 64), optional: partial sums whose rows add up to the content map's
    channel sums (``render_rays_cross_ray(..., channel_sums=True)['chansum_*']``) - saves the pass
    over the map that would otherwise form them."""
    lib = _lib.load()
    content, n, cps, ccs = _feat_strides(content, "content")
    h, w = content.shape[2], content.shape[3]
    dev = content.device
    if dev != sw.device:
        raise ValueError("content and style_net parameters are on different devices")
    if not sw.has_decoder or (style is not None and not sw.has_fusion):
        raise ValueError("style weights are missing the parameters this call needs")
    ns = sps = scs = 0
    if style is not None:
        style, ns, sps, scs = _feat_strides(style, "style")
        if style.device != dev:
            raise ValueError("content and style are on different devices")
    with torch.cuda.device(dev):
        rgb = torch.empty((1, 3, h, w), dtype=torch.float32, device=dev)
        trans = torch.empty((1, 32, 32), dtype=torch.float32, device=dev) if want_trans else None
        fused = torch.empty((1, 64, h, w), dtype=torch.float32, device=dev) if want_fused else None
        scratch = torch.empty(lib.crnerf_style_scratch_floats(n), dtype=torch.float32, device=dev)
        n_parts = 0
        if channel_sums is not None and style is not None:
            channel_sums = _c(_need(channel_sums, "channel_sums", 2))
            if channel_sums.shape[1] != 64 or channel_sums.shape[0] < 1 or channel_sums.device != dev:
                raise ValueError("channel_sums must be (rows >= 1, 64) on the content's device")
            n_parts = int(channel_sums.shape[0])
        check(lib.crnerf_style_forward_sums(C.byref(sw.struct), content.data_ptr(), n, cps, ccs,
                                            _p(style), ns, sps, scs,
                                            channel_sums.data_ptr() if n_parts else None, n_parts,
                                            rgb.data_ptr(), _p(trans), _p(fused), scratch.data_ptr(),
                                            _stream(dev)))
    out = [rgb]
    if want_trans:
        out.append(trans)
    if want_fused:
        out.append(fused)
    return out[0] if len(out) == 1 else tuple(out)


# keys of the 22 style_net parameters in the order of crnerf_style_backward_layout's offsets
STYLE_GRAD_KEYS = tuple(
    [f"multi_net.{net}.{k}" for net in ("cnet", "snet")
     for k in ("convs.0.weight", "convs.2.weight", "convs.4.weight", "convs.0.bias", "convs.2.bias",
               "convs.4.bias", "fc.weight", "fc.bias")] +
    ["multi_net.compress.weight", "multi_net.compress.bias", "multi_net.unzip.weight",
     "multi_net.unzip.bias", "decoder.feat_2_rgb_list.0.weight", "decoder.feat_2_rgb_list.0.bias"])


def style_forward_train(sw: StyleWeightsRef, content: torch.Tensor, style: torch.Tensor,
                        channel_sums: Optional[torch.Tensor] = None):
    """Training forward of style_net: (rgb (1,3,H,W), aux) - aux is what ``style_backward`` keeps
    from the forward (means, Gram vectors, FC outputs, transmatrix)."""
    lib = _lib.load()
    content, n, cps, ccs = _feat_strides(content, "content")
    style, ns, sps, scs = _feat_strides(style, "style")
    h, w = content.shape[2], content.shape[3]
    dev = content.device
    if dev != sw.device or style.device != dev:
        raise ValueError("content, style and style_net parameters must be on one device")
    if not (sw.has_decoder and sw.has_fusion):
        raise ValueError("style weights are missing the parameters this call needs")
    with torch.cuda.device(dev):
        rgb = torch.empty((1, 3, h, w), dtype=torch.float32, device=dev)
        aux = torch.empty(lib.crnerf_style_aux_floats(), dtype=torch.float32, device=dev)
        scratch = torch.empty(lib.crnerf_style_scratch_floats(n), dtype=torch.float32, device=dev)
        n_parts = 0
        if channel_sums is not None:
            channel_sums = _c(_need(channel_sums, "channel_sums", 2))
            n_parts = int(channel_sums.shape[0])
        check(lib.crnerf_style_forward_train(C.byref(sw.struct), content.data_ptr(), n, cps, ccs, style.data_ptr(), ns,
                                             sps, scs, channel_sums.data_ptr() if n_parts else None, n_parts,
                                             rgb.data_ptr(), aux.data_ptr(), scratch.data_ptr(), _stream(dev)))
    return rgb, aux


_style_grad_layout = None


def style_backward(sw: StyleWeightsRef, content: torch.Tensor, style: torch.Tensor, aux: torch.Tensor,
                   g_rgb: torch.Tensor):
    """Backward of style_net.forward: returns (g_content (1,64,H,W), g_style (1,64,h,w), {key: grad})
    with the 22 parameter gradients as views of one flat buffer."""
    global _style_grad_layout
    lib = _lib.load()
    content, n, cps, ccs = _feat_strides(content, "content")
    style, ns, sps, scs = _feat_strides(style, "style")
    dev = content.device
    g_rgb = _c(_need(g_rgb, "g_rgb")).reshape(3, n)
    if _style_grad_layout is None:
        offs = (C.c_int64 * 22)()
        lib.crnerf_style_backward_layout(offs)
        _style_grad_layout = list(offs)
    with torch.cuda.device(dev):
        flat = torch.empty(lib.crnerf_style_backward_grads_floats(), dtype=torch.float32, device=dev)
        g_c = torch.empty((n, 64), dtype=torch.float32, device=dev)
        g_s = torch.empty((ns, 64), dtype=torch.float32, device=dev)
        scratch = torch.empty(lib.crnerf_style_backward_scratch_floats(n, ns), dtype=torch.float32, device=dev)
        check(lib.crnerf_style_backward(C.byref(sw.struct), content.data_ptr(), n, cps, ccs, style.data_ptr(), ns, sps,
                                        scs, aux.data_ptr(), g_rgb.data_ptr(), g_c.data_ptr(), g_s.data_ptr(),
                                        flat.data_ptr(), scratch.data_ptr(), _stream(dev)))
    grads = {}
    for key, off in zip(STYLE_GRAD_KEYS, _style_grad_layout):
        t = sw.tensors[key]
        grads[key] = flat[off:off + t.numel()].view(t.shape)
    _, _, h, w = content.shape
    _, _, sh, sw_ = style.shape
    return g_c.t().reshape(1, 64, h, w), g_s.t().reshape(1, 64, sh, sw_), grads


def cnn_forward(sw: StyleWeightsRef, which: str, x: torch.Tensor) -> torch.Tensor:
    """CNN.forward (linearStyleTransfer.py:28-37): (1,64,H,W) -> (1,1024)."""
    lib = _lib.load()
    x, n, ps, cs = _feat_strides(x, "x")
    dev = x.device
    cw = sw.struct.cnet if which == "cnet" else sw.struct.snet
    if not cw.fc_w:
        raise ValueError("CNN weights missing")
    with torch.cuda.device(dev):
        out = torch.empty((1, 1024), dtype=torch.float32, device=dev)
        scratch = torch.empty(lib.crnerf_style_scratch_floats(n), dtype=torch.float32, device=dev)
        check(lib.crnerf_cnn_forward(C.byref(cw), x.data_ptr(), n, ps, cs, out.data_ptr(),
                                     scratch.data_ptr(), _stream(dev)))
    return out


def debug_set(buf: Optional[torch.Tensor], layer: int = -1):
    """Tests only: dump post-activation values of `layer` for every point into buf (P,256)."""
    check(_lib.load().crnerf_debug_set(_p(buf), layer))


# --------------------------------------------------------------------------
# Sharded form of the cross-ray block (SURVEY.md 8e scheme B): the three phases of
# style_forward as separate calls so a collective can run between them.
def _rows64(t: torch.Tensor, name: str):
    """(n, 64) row-major features (what render_pass returns) -> (tensor, n, pix_stride, ch_stride)."""
    t = _c(_need(t, name, 2))
    if t.shape[1] != 64:
        raise ValueError(f"{name} must be (n_pixels, 64), got {tuple(t.shape)}")
    return t, t.shape[0], 64, 1


def _style_scratch(dev, n):
    return torch.empty(_lib.load().crnerf_style_scratch_floats(n), dtype=torch.float32, device=dev)


def style_stats1(content_rows: torch.Tensor) -> torch.Tensor:
    """Per-channel SUMS over this rank's pixels: (n,64) -> (64,)."""
    lib = _lib.load()
    x, n, ps, cs = _rows64(content_rows, "content_rows")
    dev = x.device
    with torch.cuda.device(dev):
        sums = torch.zeros(64, dtype=torch.float32, device=dev)
        if n == 0:
            return sums
        check(lib.crnerf_style_stats1(x.data_ptr(), n, ps, cs, sums.data_ptr(),
                                      _style_scratch(dev, n).data_ptr(), _stream(dev)))
    return sums


def sum_rows(parts: torch.Tensor) -> torch.Tensor:
    """(rows, len) -> (len,) column sums in a fixed order (the render kernel's per-CTA channel
    sums -> the channel sums of this rank's block of rays)."""
    lib = _lib.load()
    parts = _c(_need(parts, "parts", 2))
    dev = parts.device
    with torch.cuda.device(dev):
        out = torch.zeros(parts.shape[1], dtype=torch.float32, device=dev)
        if parts.shape[0] == 0:
            return out
        check(lib.crnerf_sum_rows(parts.data_ptr(), int(parts.shape[0]), int(parts.shape[1]), out.data_ptr(),
                                  _stream(dev)))
    return out


def style_stats2(sw: StyleWeightsRef, content_rows: torch.Tensor, mean: torch.Tensor) -> torch.Tensor:
    """Un-normalised Gram of cnet.convs(content - mean) over this rank's pixels -> (32,32)."""
    lib = _lib.load()
    x, n, ps, cs = _rows64(content_rows, "content_rows")
    mean = _c(_need(mean, "mean", 1))
    dev = x.device
    with torch.cuda.device(dev):
        gram = torch.zeros((32, 32), dtype=torch.float32, device=dev)
        if n == 0:
            return gram
        check(lib.crnerf_style_stats2(C.byref(sw.struct), x.data_ptr(), n, ps, cs, mean.data_ptr(),
                                      gram.data_ptr(), _style_scratch(dev, n).data_ptr(), _stream(dev)))
    return gram


def style_apply(sw: StyleWeightsRef, content_rows: torch.Tensor, mean: torch.Tensor,
                gram_normalised: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
    """Fused cross-ray map + decoder on this rank's pixels -> rgb (3, n) planar."""
    lib = _lib.load()
    x, n, ps, cs = _rows64(content_rows, "content_rows")
    mean = _c(_need(mean, "mean", 1))
    gram_normalised = _c(_need(gram_normalised, "gram_normalised", 2))
    style, ns, sps, scs = _feat_strides(style, "style")
    dev = x.device
    with torch.cuda.device(dev):
        rgb = torch.empty((3, n), dtype=torch.float32, device=dev)
        if n == 0:
            return rgb
        check(lib.crnerf_style_apply(C.byref(sw.struct), x.data_ptr(), n, ps, cs, mean.data_ptr(),
                                     gram_normalised.data_ptr(), style.data_ptr(), ns, sps, scs,
                                     rgb.data_ptr(), None, _style_scratch(dev, n).data_ptr(),
                                     _stream(dev)))
    return rgb


# --------------------------------------------------------------------------
# Training step: forward that stores the backward's inputs, composite backward
def render_pass_train(packed: PackedMLP, rays: torch.Tensor, z_vals: torch.Tensor,
                      noise: Optional[torch.Tensor] = None, view_dir: Optional[torch.Tensor] = None,
                      n_freq_xyz: int = 15, n_freq_dir: int = 4, xyz_jitter: Optional[torch.Tensor] = None):
    """render_pass + saved activations.  Returns (weights, feature, depth, acts, raw):
    acts is the byte buffer of crnerf_render_pass_train (tiled 16-bit layout, see ``untile_acts``), raw is
    65*n_points fp32: the sigmoid features as (n_points, 64) rows, then the n_points softplus sigmas."""
    lib = _lib.load()
    rays = _c(_need(rays, "rays", 2))
    z_vals = _c(_need(z_vals, "z_vals", 2))
    n, s = z_vals.shape
    if rays.shape != (n, 8):
        raise ValueError(f"rays must be ({n}, 8), got {tuple(rays.shape)}")
    if 3 + 6 * n_freq_xyz != packed.e_xyz or 3 + 6 * n_freq_dir != packed.e_dir:
        raise ValueError("weights were packed for a different embedding width")
    if noise is not None:
        noise = _c(_need(noise, "noise", 2))
    if view_dir is not None:
        view_dir = _c(_need(view_dir, "view_dir", 2))
    if packed.operand not in (OPERAND_FP16, OPERAND_BF16):
        raise CrnerfError("the training forward takes fp16 or bf16 operands (fp16x3 is inference only)")
    dev = rays.device
    with torch.cuda.device(dev):
        weights = torch.empty((n, s), dtype=torch.float32, device=dev)
        feature = torch.empty((n, 64), dtype=torch.float32, device=dev)
        depth = torch.empty((n,), dtype=torch.float32, device=dev)
        acts = torch.empty((int(lib.crnerf_render_acts_bytes(n * s)),), dtype=torch.uint8, device=dev)
        raw = torch.empty((n * s * 65,), dtype=torch.float32, device=dev)   # (n*s, 64) features, then n*s sigmas
        if n == 0:
            return weights, feature, depth, acts, raw
        opts = None
        if xyz_jitter is not None:
            xyz_jitter = _c(_need(xyz_jitter, "xyz_jitter", 2))
            if xyz_jitter.shape != (n * s, 3):
                raise ValueError(f"xyz_jitter must be ({n * s}, 3)")
            opts = C.byref(_lib.RenderOpts(_p(xyz_jitter), None, None))
        check(lib.crnerf_render_pass_train_opts(packed.buf.data_ptr(), packed.operand, rays.data_ptr(),
                                                _p(view_dir), z_vals.data_ptr(), _p(noise), n, s,
                                                n_freq_xyz, n_freq_dir, weights.data_ptr(),
                                                feature.data_ptr(), depth.data_ptr(), acts.data_ptr(),
                                                raw.data_ptr(), opts, _stream(dev)))
    return weights, feature, depth, acts, raw


def untile_acts(acts: torch.Tensor, n_points: int, operand: int = OPERAND_FP16):
    """Tests / debugging: the saved-activation buffer of ``render_pass_train`` (the backward kernels'
    tiled, swizzled operand layout - include/crnerf_b200.h) as plain row-major tensors:
    ``(trunk [9 x (P,256)], dir_out (P,128), emb (P,128))``."""
    dt = torch.float16 if operand == OPERAND_FP16 else torch.bfloat16
    T = (n_points + 127) // 128
    dev = acts.device
    r = torch.arange(128, device=dev)
    src = (torch.arange(8, device=dev)[None, :] ^ (r[:, None] & 7))          # position of logical chunk c in row r

    def slot(byte0, slabs):
        blk = acts[byte0: byte0 + T * slabs * 16384].view(T, slabs, 128, 8, 16)
        idx = src.view(1, 1, 128, 8, 1).expand(T, slabs, 128, 8, 16)
        rows = torch.gather(blk, 3, idx).reshape(T, slabs, 128, 128)          # logical chunk order
        vals = rows.contiguous().view(dt).reshape(T, slabs, 128, 64)
        return vals.permute(0, 2, 1, 3).reshape(T * 128, slabs * 64)[:n_points]

    trunk = [slot(k * T * 4 * 16384, 4) for k in range(9)]
    base = 9 * T * 4 * 16384
    return trunk, slot(base, 2), slot(base + T * 2 * 16384, 2)


def render_backward(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], operand: int, e_xyz: int,
                    e_dir: int, acts: torch.Tensor, raw: torch.Tensor, z_vals: torch.Tensor,
                    noise: Optional[torch.Tensor], g_feature: Optional[torch.Tensor],
                    g_weights: Optional[torch.Tensor], g_depth: Optional[torch.Tensor]):
    """Backward of one render pass on the tensor cores (crnerf_render_backward): returns the 12 weight
    and 12 bias gradients (fp32, the parameters' shapes, ``MLP_LAYER_KEYS`` order)."""
    lib = _lib.load()
    z_vals = _c(_need(z_vals, "z_vals", 2))
    n, s = z_vals.shape
    raw = _c(_need(raw, "raw"))
    if raw.numel() != n * s * 65:
        raise ValueError("raw must be the 65*n_rays*n_samples floats render_pass_train saved")
    opt = lambda t, name, shape: None if t is None else _c(_need(t, name).reshape(shape))
    noise = opt(noise, "noise", (n, s))
    g_feature = opt(g_feature, "g_feature", (n, 64))
    g_weights = opt(g_weights, "g_weights", (n, s))
    g_depth = opt(g_depth, "g_depth", (n,))
    dev = raw.device
    keep = []
    w = _lib.MlpWeights()
    for i in range(12):
        wi, bi = _c(_need(weights[i].detach(), f"weight[{i}]", 2)), _c(_need(biases[i].detach(), f"bias[{i}]", 1))
        keep += [wi, bi]
        w.weight[i], w.bias[i] = wi.data_ptr(), bi.data_ptr()
    w.e_xyz, w.e_dir = e_xyz, e_dir
    with torch.cuda.device(dev):
        # one flat buffer: weight gradients are written in full by the kernels, bias gradients are
        # accumulated - only that tail is zeroed (one fill instead of 24)
        nw = [int(t.numel()) for t in weights]
        nb = [int(t.numel()) for t in biases]
        flat = torch.empty((sum(nw) + sum(nb),), dtype=torch.float32, device=dev)
        flat[sum(nw):].zero_()
        if n == 0:
            flat.zero_()
        gw, gb, o = [], [], 0
        for i in range(12):
            gw.append(flat[o:o + nw[i]].view(weights[i].shape))
            o += nw[i]
        for i in range(12):
            gb.append(flat[o:o + nb[i]].view(biases[i].shape))
            o += nb[i]
        if n == 0:
            return gw, gb
        bw = torch.empty((int(lib.crnerf_render_backward_weights_bytes(e_xyz)),), dtype=torch.uint8, device=dev)
        scratch = torch.empty((int(lib.crnerf_render_backward_scratch_bytes(n * s)),), dtype=torch.uint8, device=dev)
        pw = (C.c_void_p * 12)(*[t.data_ptr() for t in gw])
        pb = (C.c_void_p * 12)(*[t.data_ptr() for t in gb])
        check(lib.crnerf_render_backward(C.byref(w), operand, acts.data_ptr(), raw.data_ptr(), z_vals.data_ptr(),
                                         _p(noise), _p(g_feature), _p(g_weights), _p(g_depth), n, s,
                                         bw.data_ptr(), scratch.data_ptr(), pw, pb, _stream(dev)))
    del keep
    return gw, gb


def composite_backward(raw: torch.Tensor, z_vals: torch.Tensor, noise: Optional[torch.Tensor],
                       g_feature: Optional[torch.Tensor], g_weights: Optional[torch.Tensor],
                       g_depth: Optional[torch.Tensor]):
    """Backward of the alpha composite: -> (d_rgb_pre (P,64), d_sigma_pre (P,))."""
    lib = _lib.load()
    raw = _c(_need(raw, "raw", 2))
    z_vals = _c(_need(z_vals, "z_vals", 2))
    n, s = z_vals.shape
    if raw.shape != (n * s, 65):
        raise ValueError("raw must be (n_rays*n_samples, 65)")
    opt = lambda t, name, shape: None if t is None else _c(_need(t, name).reshape(shape))
    noise = opt(noise, "noise", (n, s))
    g_feature = opt(g_feature, "g_feature", (n, 64))
    g_weights = opt(g_weights, "g_weights", (n, s))
    g_depth = opt(g_depth, "g_depth", (n,))
    dev = raw.device
    with torch.cuda.device(dev):
        d_rgb = torch.empty((n * s, 64), dtype=torch.float32, device=dev)
        d_sig = torch.empty((n * s,), dtype=torch.float32, device=dev)
        if n == 0:
            return d_rgb, d_sig
        check(lib.crnerf_composite_backward(raw.data_ptr(), z_vals.data_ptr(), _p(noise), _p(g_feature),
                                            _p(g_weights), _p(g_depth), n, s, d_rgb.data_ptr(),
                                            d_sig.data_ptr(), _stream(dev)))
    return d_rgb, d_sig


def relu_bias_grad(g: torch.Tensor, act: Optional[torch.Tensor]) -> torch.Tensor:
    """In place g *= (act > 0) (act: saved 16-bit post-ReLU activations or None) and return the
    column sums of the result (= the layer's bias gradient)."""
    lib = _lib.load()
    _need(g, "g", 2)
    if not g.is_contiguous():
        raise ValueError("g must be contiguous (it is modified in place)")
    n, c = g.shape
    if act is not None:
        if act.shape != g.shape or act.element_size() != 2 or not act.is_contiguous() or act.device != g.device:
            raise ValueError("act must be a contiguous 16-bit tensor of g's shape on g's device")
    dev = g.device
    with torch.cuda.device(dev):
        gb = torch.empty((c,), dtype=torch.float32, device=dev)
        scratch = torch.empty((int(lib.crnerf_relu_bias_grad_scratch_floats(c)),), dtype=torch.float32,
                              device=dev)
        check(lib.crnerf_relu_bias_grad(g.data_ptr(), _p(act), n, c, gb.data_ptr(), scratch.data_ptr(),
                                        _stream(dev)))
    return gb


def generate_rays(height: int, width: int, K, c2w, near: float, far: float, device="cuda") -> torch.Tensor:
    """(height*width, 8) rays [o3, d3, near, far] of a pinhole frame, built on the device
    (reference datasets/ray_utils.py:5-52 + the row assembly of the datasets).  ``K`` is the 3x3
    intrinsics (anything indexable as K[r][c]), ``c2w`` the 3x4 camera-to-world matrix."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise CrnerfError("generate_rays builds the rays in GPU memory; there is no CPU fallback")
    intr = (C.c_float * 4)(float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]))
    m = (C.c_float * 12)(*[float(c2w[r][c]) for r in range(3) for c in range(4)])
    with torch.cuda.device(dev):
        rays = torch.empty((height * width, 8), dtype=torch.float32, device=dev)
        check(lib.crnerf_generate_rays(intr, m, float(near), float(far), int(height), int(width),
                                       rays.data_ptr(), _stream(dev)))
    return rays


def grid_patch(all_rays: torch.Tensor, all_rgbs: torch.Tensor, lin_w: torch.Tensor, lin_h: torch.Tensor,
               img_w: float, img_h: float, image_offset: float, scale: float, h_offset: float, w_offset: float,
               status: Optional[torch.Tensor] = None):
    """One grid-sampled training patch gathered from the device-resident ray cache (reference
    datasets/phototourism_mask_grid_sample.py:241-275).  Returns (rays (g*g,8), ts (g*g) int64,
    rgbs (g*g,3), rgb_idx (g*g) int64, uv_sample (g*g,2)); bit-exact with the reference, whose image
    sizes and cache offset are fp32 values (``image_offset`` = the fp32 sum of w*h before the image)."""
    lib = _lib.load()
    _need(all_rays, "all_rays")
    _need(all_rgbs, "all_rgbs")
    if all_rays.dim() != 2 or all_rays.shape[1] != 9 or all_rgbs.shape != (all_rays.shape[0], 3):
        raise ValueError("all_rays must be (M,9) and all_rgbs (M,3)")
    if not (all_rays.is_contiguous() and all_rgbs.is_contiguous()):
        raise ValueError("the ray cache must be contiguous (it is gathered in place)")
    dev = all_rays.device
    g = int(lin_w.numel())
    if int(lin_h.numel()) != g:
        raise ValueError("lin_w and lin_h must have the same length")
    lw = lin_w.to(dev, torch.float32).contiguous()
    lh = lin_h.to(dev, torch.float32).contiguous()
    with torch.cuda.device(dev):
        rays = torch.empty((g * g, 8), dtype=torch.float32, device=dev)
        ts = torch.empty((g * g,), dtype=torch.int64, device=dev)
        rgbs = torch.empty((g * g, 3), dtype=torch.float32, device=dev)
        idx = torch.empty((g * g,), dtype=torch.int64, device=dev)
        uv = torch.empty((g * g, 2), dtype=torch.float32, device=dev)
        check(lib.crnerf_grid_patch(lw.data_ptr(), lh.data_ptr(), g, float(img_w), float(img_h), float(scale),
                                    float(h_offset), float(w_offset), all_rays.data_ptr(), all_rgbs.data_ptr(),
                                    int(all_rays.shape[0]), float(image_offset), rays.data_ptr(), ts.data_ptr(),
                                    rgbs.data_ptr(), idx.data_ptr(), uv.data_ptr(),
                                    status.data_ptr() if status is not None else None, _stream(dev)))
    return rays, ts, rgbs, idx, uv


def rgb_to_u8(rgb: torch.Tensor) -> torch.Tensor:
    """(1,3,H,W) or (3,n) fp32 rgb -> (H,W,3) / (n,3) uint8 = uint8(clip(x,0,1)*255), reference eval.py:295-297."""
    lib = _lib.load()
    _need(rgb, "rgb")
    shape = None
    if rgb.dim() == 4:
        if rgb.shape[0] != 1 or rgb.shape[1] != 3:
            raise ValueError("rgb must be (1,3,H,W) or (3,n)")
        shape = (rgb.shape[2], rgb.shape[3], 3)
        flat = _c(rgb).reshape(3, -1)
    elif rgb.dim() == 2 and rgb.shape[0] == 3:
        flat = _c(rgb)
    else:
        raise ValueError("rgb must be (1,3,H,W) or (3,n)")
    n = flat.shape[1]
    dev = rgb.device
    with torch.cuda.device(dev):
        out = torch.empty((n, 3), dtype=torch.uint8, device=dev)
        if n:
            check(lib.crnerf_rgb_to_u8(flat.data_ptr(), n, out.data_ptr(), _stream(dev)))
    return out.reshape(shape) if shape is not None else out


# ---- style/content encoder (csrc/encoder.cu) ------------------------------------------------------
class PackedEncoder:
    """Device-side weight image of an ``encoder_sameoutputsize`` (crnerf_encoder_pack)."""

    def __init__(self, buf: torch.Tensor):
        self.buf = buf


def pack_encoder(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]) -> PackedEncoder:
    """weights / biases: conv1..conv7 of the module, as in its state_dict (reference
    models/linearStyleTransfer.py:213-246)."""
    lib = _lib.load()
    shapes = [(3, 3, 1, 1), (64, 3, 3, 3), (64, 64, 3, 3), (128, 64, 3, 3), (128, 128, 3, 3), (128, 128, 3, 3),
              (64, 128, 1, 1)]
    if len(weights) != 7 or len(biases) != 7:
        raise ValueError("expected conv1..conv7")
    w = _lib.EncoderWeights()
    keep = []
    for i, (wt, bt) in enumerate(zip(weights, biases)):
        _need(wt, f"conv{i + 1}.weight")
        _need(bt, f"conv{i + 1}.bias")
        if tuple(wt.shape) != shapes[i] or bt.numel() != shapes[i][0]:
            raise ValueError(f"conv{i + 1}: weight {tuple(wt.shape)} / bias {tuple(bt.shape)}; expected {shapes[i]} "
                             "(out_channel must be 64)")
        wc, bc = _c(wt.detach()), _c(bt.detach())
        keep += [wc, bc]
        w.weight[i], w.bias[i] = wc.data_ptr(), bc.data_ptr()
    dev = keep[0].device
    n = int(lib.crnerf_encoder_packed_bytes())
    buf = torch.empty(n, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.crnerf_encoder_pack(C.byref(w), buf.data_ptr(), n, _stream(dev)))
    return PackedEncoder(buf)


def encoder_forward(packed: PackedEncoder, img: torch.Tensor) -> torch.Tensor:
    """img (1,3,H,W) -> (1,64,32,32): encoder_sameoutputsize.forward (reference
    models/linearStyleTransfer.py:250-276), inference only."""
    lib = _lib.load()
    _need(img, "img", 4)
    if img.shape[0] != 1 or img.shape[1] != 3:
        raise ValueError(f"img must be (1,3,H,W), got {tuple(img.shape)}")
    h, w = int(img.shape[2]), int(img.shape[3])
    if h < 8 or w < 8 or h > 8192 or w > 8192:
        raise ValueError(f"image {h}x{w} unsupported (8..8192 per side)")
    x = _c(img)
    dev = x.device
    nbytes = int(lib.crnerf_encoder_scratch_bytes(h, w))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty((1, 64, 32, 32), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.crnerf_encoder_forward(packed.buf.data_ptr(), x.data_ptr(), h, w, out.data_ptr(),
                                         scratch.data_ptr(), nbytes, _stream(dev)))
    return out


ENCODER_SHAPES = [(3, 3, 1, 1), (64, 3, 3, 3), (64, 64, 3, 3), (128, 64, 3, 3), (128, 128, 3, 3), (128, 128, 3, 3),
                  (64, 128, 1, 1)]


def _encoder_img(img):
    _need(img, "img", 4)
    if img.shape[0] != 1 or img.shape[1] != 3:
        raise ValueError(f"img must be (1,3,H,W), got {tuple(img.shape)}")
    h, w = int(img.shape[2]), int(img.shape[3])
    if h < 8 or w < 8 or h > 8192 or w > 8192:
        raise ValueError(f"image {h}x{w} unsupported (8..8192 per side)")
    return _c(img.detach()), h, w


def encoder_forward_train(packed: PackedEncoder, img: torch.Tensor):
    """Forward of the training step: ``(out (1,64,32,32), tape)`` - the kernels of ``encoder_forward``
    with every layer's activation planes kept for ``encoder_backward`` (crnerf_encoder_forward_train)."""
    lib = _lib.load()
    x, h, w = _encoder_img(img)
    dev = x.device
    nbytes = int(lib.crnerf_encoder_tape_bytes(h, w))
    tape = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty((1, 64, 32, 32), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.crnerf_encoder_forward_train(packed.buf.data_ptr(), x.data_ptr(), h, w, out.data_ptr(),
                                               tape.data_ptr(), nbytes, _stream(dev)))
    return out, tape


def encoder_backward(packed: PackedEncoder, img: torch.Tensor, out: torch.Tensor, tape: torch.Tensor,
                     grad_out: torch.Tensor, want_img_grad: bool = False):
    """Gradients of the 7 weights and 7 biases (state_dict layouts, conv1..conv7) and, optionally, of
    ``img``: crnerf_encoder_backward (tensor-core input- and weight-gradient convolutions, deterministic)."""
    lib = _lib.load()
    x, h, w = _encoder_img(img)
    dev = x.device
    g = _c(grad_out.detach().to(torch.float32))
    if g.numel() != 64 * 32 * 32:
        raise ValueError(f"grad_out must be (1,64,32,32), got {tuple(grad_out.shape)}")
    gw = [torch.empty(sh, dtype=torch.float32, device=dev) for sh in ENCODER_SHAPES]
    gb = [torch.empty(sh[0], dtype=torch.float32, device=dev) for sh in ENCODER_SHAPES]
    grads = _lib.EncoderWeights()
    for i in range(7):
        grads.weight[i], grads.bias[i] = gw[i].data_ptr(), gb[i].data_ptr()
    g_img = torch.empty_like(x) if want_img_grad else None
    nbytes = int(lib.crnerf_encoder_backward_scratch_bytes(h, w))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.crnerf_encoder_backward(packed.buf.data_ptr(), x.data_ptr(), h, w, _c(out).data_ptr(), g.data_ptr(),
                                          tape.data_ptr(), C.byref(grads), g_img.data_ptr() if want_img_grad else None,
                                          scratch.data_ptr(), nbytes, _stream(dev)))
    if _ENCODER_DEBUG is not None:
        _ENCODER_DEBUG["scratch"] = scratch
    return gw, gb, g_img


_ENCODER_DEBUG = None      # tools/enc_bwd_debug.py sets a dict to look at the backward's scratch
