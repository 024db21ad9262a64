"""Synthetic "Phototourism-shaped" camera rays for benchmarks and tools: a pinhole
camera with a 60 degree field of view, a mildly rotated pose a little off the origin, near 0 / far 5
(scenes are rescaled so that the far bound is 5, reference datasets/phototourism_mask_grid_sample.py:139-141).
Plain tensor code, no dataset and no kernels involved; rows are [o3, d3, near, far] as the renderer
expects (reference models/rendering.py:151-153, directions as datasets/ray_utils.py:5-52)."""
from __future__ import annotations

import math

import torch


def synthetic_pose(seed: int = 0) -> torch.Tensor:
    """Deterministic 3x4 camera-to-world matrix looking down -z."""
    g = torch.Generator().manual_seed(seed)
    ang = (torch.rand(3, generator=g) - 0.5) * 0.4
    cx, sx = math.cos(ang[0]), math.sin(ang[0])
    cy, sy = math.cos(ang[1]), math.sin(ang[1])
    cz, sz = math.cos(ang[2]), math.sin(ang[2])
    rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=torch.float32)
    ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float32)
    rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=torch.float32)
    t = (torch.rand(3, 1, generator=g) - 0.5) * torch.tensor([[0.6], [0.2], [0.6]])
    return torch.cat([rz @ ry @ rx, t], dim=1)


def pinhole_rays(h: int, w: int, c2w: torch.Tensor, near: float = 0.0, far: float = 5.0,
                 fov_deg: float = 60.0) -> torch.Tensor:
    """(h*w, 8) fp32 rays on the CPU, row index j*w + i."""
    f = 0.5 * w / math.tan(0.5 * math.radians(fov_deg))
    j, i = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                          indexing="ij")
    dirs = torch.stack([(i - w / 2) / f, -(j - h / 2) / f, -torch.ones_like(i)], dim=-1)
    rays_d = dirs @ c2w[:, :3].T
    rays_d = (rays_d / rays_d.norm(dim=-1, keepdim=True)).reshape(-1, 3)
    rays_o = c2w[:, 3].expand(h, w, 3).reshape(-1, 3)
    nf = torch.tensor([near, far], dtype=torch.float32).expand(rays_d.shape[0], 2)
    return torch.cat([rays_o, rays_d, nf], dim=1).contiguous()
