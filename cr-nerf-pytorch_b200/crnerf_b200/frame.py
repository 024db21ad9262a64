"""Whole-frame rendering, sharded by rays across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink).  The frame's rays are
cut into ``world`` contiguous row blocks; every rank renders its block with the
fused kernels (``render_rays_cross_ray`` is per-ray independent, reference
models/rendering.py:50-196) exactly as the reference's ``batched_inference`` loop
does for the whole frame (eval.py:29-59).  The cross-ray block that follows
(``style_net``, reference models/linearStyleTransfer.py:58-90,284-291) needs
statistics over ALL rays of the frame, so it runs in the sharded form of
SURVEY.md 8(e):

  scheme "stats" (default)               scheme "gather" (north-star literal)
    local channel sums                      all-gather feature_fine (N,64)
    all-reduce 64 floats  -> mean           every rank runs the unsharded
    local un-normalised Gram                ``style_net`` on the whole frame
    all-reduce 1,024 floats
    local fused 64->3 map + sigmoid
    all-gather rgb (3 floats / ray)

"stats" moves 20x fewer bytes and has no replicated compute; both give the
reference's result up to fp32 summation order.  Everything between the
collectives is a kernel of libcrnerf_b200.so; the collectives are issued on the
current stream right behind them (no host synchronisation).

The collective plumbing is written against a small ``backend`` object (three
methods: ``sums``, ``gram``, ``apply``) so that the world_size-2 ``gloo`` tests can
drive it on CPU with the oracle standing in for the kernels; the product backend
is ``CudaStyleBackend`` and there is no other.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "batched_render", "CudaStyleBackend", "fuse_decode_sharded",
           "render_frame_sharded"]


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of rank ``rank``: ``world`` contiguous blocks of ceil(n/world) rows
    (the last ones may be short or empty), so gathered blocks concatenate in ray order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def batched_render(models, embeddings, rays, N_samples, N_importance, use_disp, chunk=None, **kwargs):
    """The reference's ``batched_inference`` (eval.py:29-59): eval-mode render of ``rays``.

    The reference chunks the rays (``--chunk``) because every chunk materialises
    ``(chunk*N_samples, 256)`` activations; the fused kernels materialise no per-point tensor, so
    with ``chunk=None`` (default) the whole block of rays is ONE call = one coarse launch, one
    inverse-CDF launch and one fine launch, whatever its size, and nothing is concatenated.
    An explicit ``chunk`` reproduces the reference's loop call for call."""
    from models.rendering import render_rays_cross_ray
    n = rays.shape[0]
    if n == 0:
        typ = "fine" if N_importance > 0 else "coarse"
        return {f"feature_{typ}": rays.new_zeros((0, 64)), f"depth_{typ}": rays.new_zeros((0,))}
    if chunk is None or chunk >= n:
        with torch.no_grad():
            return render_rays_cross_ray(models, embeddings, rays, None, N_samples, use_disp, 0, 0,
                                         N_importance, n, False, test_time=True, **kwargs)
    out = {}
    with torch.no_grad():
        for i in range(0, n, chunk):
            res = render_rays_cross_ray(models, embeddings, rays[i:i + chunk], None, N_samples,
                                        use_disp, 0, 0, N_importance, chunk, False, test_time=True,
                                        **kwargs)
            for k, v in res.items():
                out.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in out.items()}


class CudaStyleBackend:
    """The three kernel phases of the cross-ray block for one ``style_net`` (csrc/crossray.cu)."""

    def __init__(self, decoder):
        from . import ops
        self._ops = ops
        self._sw = decoder._style_ref([("", "")])

    def sums(self, feat, parts=None):           # (n,64) -> (64,) channel sums
        if parts is not None:                   # the render kernel's per-CTA partial sums: no pass over feat
            return self._ops.sum_rows(parts)
        return self._ops.style_stats1(feat)

    def gram(self, feat, mean):                 # -> (32,32) un-normalised Gram of cnet.convs(x-mean)
        return self._ops.style_stats2(self._sw, feat, mean)

    def apply(self, feat, mean, gram_n, style):  # -> (3,n) rgb
        return self._ops.style_apply(self._sw, feat, mean, gram_n, style)


def _all_reduce(t, group):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def _gather_blocks(local: torch.Tensor, n_total: int, group, dim: int) -> torch.Tensor:
    """Concatenate every rank's block along ``dim`` (blocks follow ``shard_bounds``)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    per = -(-n_total // world)
    shape = list(local.shape)
    shape[dim] = per
    padded = local.new_zeros(shape)
    padded.narrow(dim, 0, local.shape[dim]).copy_(local)
    # concatenated-along-dim-0 output form (accepted by both NCCL and gloo), viewed as (world, ...)
    buf = local.new_empty([world * shape[0]] + shape[1:])
    dist.all_gather_into_tensor(buf, padded.contiguous(), group=group)
    # (world, ..., per, ...) -> (..., world*per, ...)[:n_total]
    buf = buf.reshape([world] + shape).movedim(0, dim)   # (..., world, per, ...)
    new_shape = list(local.shape)
    new_shape[dim] = world * per
    return buf.reshape(new_shape).narrow(dim, 0, n_total)


def fuse_decode_sharded(backend, feat_local: torch.Tensor, style: torch.Tensor, n_total: int,
                        group=None, sum_parts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Cross-ray fusion + decoder of a frame whose (n_total,64) feature rows are sharded by
    ``shard_bounds``; returns the whole frame's rgb as (3, n_total) on every rank.  ``sum_parts``:
    the render kernel's partial channel sums of ``feat_local`` (rows add up to its column sums)."""
    local = backend.sums(feat_local, sum_parts) if sum_parts is not None else backend.sums(feat_local)
    sums = _all_reduce(local, group)
    mean = sums / float(n_total)
    gram = _all_reduce(backend.gram(feat_local, mean), group)
    rgb_local = backend.apply(feat_local, mean, gram / float(n_total), style)
    return _gather_blocks(rgb_local, n_total, group, dim=1)


def render_frame_sharded(models, embeddings, rays: Optional[torch.Tensor], style: Optional[torch.Tensor],
                         hw: Tuple[int, int], N_samples: int, N_importance: int, chunk: Optional[int] = None,
                         use_disp: bool = False, scheme: str = "stats", group=None, backend=None,
                         camera=None, **kwargs) -> torch.Tensor:
    """Render one H x W frame whose ``rays`` (H*W, 8) are known to every rank; each rank
    renders its own row block and the frame's rgb (1,3,H,W) is returned on every rank.
    With ``rays=None`` and ``camera=(K, c2w, near, far)`` the rays are built in GPU memory
    (``ops.generate_rays``: no per-frame host meshgrid, no 32 B/ray host-to-device copy).

    Equivalent single-GPU code in the reference: eval.py:279-294 (``batched_inference`` then
    ``models['decoder'](feature, a_embedded_from_img)``)."""
    h, w = hw
    if rays is None:
        if camera is None:
            raise ValueError("give either rays or camera=(K, c2w, near, far)")
        from . import ops
        K, c2w, near, far = camera
        rays = ops.generate_rays(h, w, K, c2w, near, far, device=next(models["coarse"].parameters()).device)
    n_total = rays.shape[0]
    if n_total != h * w:
        raise ValueError(f"rays has {n_total} rows, frame is {h}x{w}")
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(n_total, world, rank)
    typ = "fine" if N_importance > 0 else "coarse"
    want_sums = style is not None and "channel_sums" not in kwargs
    if want_sums:          # the fine pass's epilogue also emits per-CTA channel sums of its features
        kwargs = dict(kwargs, channel_sums=True)
    res = batched_render(models, embeddings, rays[lo:hi], N_samples, N_importance, use_disp, chunk,
                         **kwargs)
    feat = res[f"feature_{typ}"]
    parts = res.get(f"chansum_{typ}") if want_sums and hi > lo else None
    decoder = models["decoder"]
    if scheme == "gather" or style is None:
        full = _gather_blocks(feat, n_total, group, dim=0)              # (N,64)
        content = full.t().reshape(1, 64, h, w)                         # view, read in place
        with torch.no_grad():
            if style is None:
                return decoder(content, None, type="content")
            # after a gather the local partial sums no longer describe the whole map
            return decoder(content, style, channel_sums=parts if world == 1 else None)
    if scheme != "stats":
        raise ValueError(f"unknown scheme {scheme!r}")
    if backend is None:
        backend = CudaStyleBackend(decoder)
    rgb = fuse_decode_sharded(backend, feat, style, n_total, group, sum_parts=parts)
    return rgb.reshape(1, 3, h, w)
