"""Grid-sampled training patches from a ray cache that lives in GPU memory.

The reference's training ``Dataset.__getitem__`` (datasets/phototourism_mask_grid_sample.py:240-275)
picks one training image, lays a ``sqrt(batch_size)``-square lattice over it at a random scale and
offset, and gathers the lattice pixels' rays / colours out of the concatenated ``all_rays (M,9)`` /
``all_rgbs (M,3)`` buffers on the CPU; the DataLoader then copies the batch to the GPU every step.
Here the two buffers are uploaded once (a Phototourism scene at downscale 2 is a few GB - 180 GB of
HBM holds it many times over) and a step's batch is ONE kernel (``crnerf_grid_patch``): index
arithmetic + gathers, bit-exact with the reference.

The random draws stay on the host and are the reference's own calls in the reference's order
(``np.random.seed`` / ``np.random.randint`` for the image, three ``torch.Tensor(1).uniform_`` for
scale and offsets), so a seeded run picks the same patches.
"""
from __future__ import annotations

from math import exp, sqrt
from typing import Sequence

import numpy as np
import torch

from . import ops

__all__ = ["GridPatchSampler"]


class GridPatchSampler:
    """``sample(epoch, idx)`` returns the reference's ``sample`` dict (same keys, same values) with
    the gathered tensors on the GPU.

    all_rays (M,9) fp32 ``[o3 d3 near far image_id]``, all_rgbs (M,3): the reference's buffers
    (:204-221), moved to ``device`` once; all_imgs_wh (n_img,2) ``[w,h]`` per image (:165);
    all_imgs: the per-image ``whole_img`` tensors, passed through untouched (:243)."""

    def __init__(self, all_rays: torch.Tensor, all_rgbs: torch.Tensor, all_imgs_wh, all_imgs: Sequence = None,
                 batch_size: int = 1024, scale_anneal: float = -1, min_scale: float = 0.25, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ops.CrnerfError("GridPatchSampler keeps the ray cache in GPU memory; there is no CPU fallback")
        self.all_rays = all_rays.to(dev, torch.float32).contiguous()
        self.all_rgbs = all_rgbs.to(dev, torch.float32).contiguous()
        # fp32, as the reference builds it (torch.Tensor([img_w, img_h]), :197): every size below is
        # a 0-dim float tensor and the arithmetic on it is fp32 - kept, it decides roundings
        self.all_imgs_wh = torch.as_tensor(all_imgs_wh).cpu().float()
        self.all_imgs = all_imgs
        self.batch_size = int(batch_size)
        self.grid = int(sqrt(self.batch_size))
        self.scale_anneal = scale_anneal
        self.min_scale = min_scale
        self.iterations = len(self.all_rays) // self.batch_size            # __len__ (:226-227)
        self._status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._lin = {}                                                      # (w,h) -> device linspaces

    def _linspaces(self, img_w: torch.Tensor, img_h: torch.Tensor):
        key = (float(img_w), float(img_h))
        if key not in self._lin:
            g = self.grid
            self._lin[key] = (torch.linspace(0, 1 - 1 / img_w, g).to(self.all_rays.device),   # :246
                              torch.linspace(0, 1 - 1 / img_h, g).to(self.all_rays.device))   # :247
        return self._lin[key]

    def sample(self, epoch: int, idx: int) -> dict:
        step = epoch * self.iterations + idx
        np.random.seed(step)                                               # :242
        sample_ts = np.random.randint(0, len(self.all_imgs_wh))            # :243
        img_w, img_h = self.all_imgs_wh[sample_ts]                         # :245 (0-dim fp32 tensors)
        if self.scale_anneal > 0:                                          # :248-251
            min_scale_cur = min(max(self.min_scale, 1. * exp(-step * self.scale_anneal)), 0.9)
        else:
            min_scale_cur = self.min_scale
        scale = torch.Tensor(1).uniform_(min_scale_cur, 1.)                # :252
        h_offset = torch.Tensor(1).uniform_(0, (1 - scale.item()) * (1 - 1 / img_h))   # :253
        w_offset = torch.Tensor(1).uniform_(0, (1 - scale.item()) * (1 - 1 / img_w))   # :254
        lin_w, lin_h = self._linspaces(img_w, img_h)
        offset = (self.all_imgs_wh[:sample_ts, 0] * self.all_imgs_wh[:sample_ts, 1]).sum()   # :266, fp32
        rays, ts, rgbs, rgb_idx, uv = ops.grid_patch(self.all_rays, self.all_rgbs, lin_w, lin_h, float(img_w),
                                                     float(img_h), float(offset), scale.item(), h_offset.item(),
                                                     w_offset.item(), status=self._status)
        return {'rays': rays, 'ts': ts, 'rgbs': rgbs,
                'whole_img': self.all_imgs[sample_ts] if self.all_imgs is not None else None,
                'rgb_idx': rgb_idx, 'min_scale_cur': min_scale_cur,
                'img_wh': self.all_imgs_wh[sample_ts], 'uv_sample': uv}

    def check(self) -> None:
        """Host read of the kernel's status word (one sync): raises if any gathered row ever fell
        outside the cache (inconsistent ``all_imgs_wh``)."""
        if int(self._status.item()) != 0:
            raise ops.CrnerfError("grid_patch: a lattice pixel indexed past the ray cache")
