"""Small cases of the round-2 kernels for compute-sanitizer: the two-stream Gram kernel (content +
style jobs in one launch, odd pixel counts, partial tiles, both layouts), the apply kernel, the
style_net training forward / backward (chain kernels, FC backward, finish), the grid-patch gather.
  compute-sanitizer --tool memcheck python tools/sanitize_case3.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from conftest import build_mirror_models  # noqa: E402
from crnerf_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
models, _ = build_mirror_models(0)
dec = models["decoder"].to(dev)
out = []
with torch.no_grad():
    for (h, w), shw in (((7, 9), (32, 32)), ((32, 32), (5, 13)), ((40, 52), (32, 32)), ((130, 131), (70, 70))):
        rows = torch.rand(h * w, 64, device=dev)
        style = torch.rand(1, 64, *shw, device=dev)
        parts = torch.stack([c.sum(0) for c in rows.chunk(min(37, h * w))])
        out.append(float(dec(rows.t().reshape(1, 64, h, w), style, channel_sums=parts).sum()))      # row layout + sums
        out.append(float(dec(rows.t().reshape(1, 64, h, w).contiguous(), style).sum()))             # planar layout
        out.append(float(dec(rows.t().reshape(1, 64, h, w), None, type="content").sum()))
dec.train()
for (h, w), shw in (((7, 9), (32, 32)), ((32, 32), (32, 32)), ((40, 52), (5, 13))):
    rows = torch.rand(h * w, 64, device=dev, requires_grad=True)
    style = torch.rand(1, 64, *shw, device=dev, requires_grad=True)
    dec.zero_grad(set_to_none=True)
    rgb = dec(rows.t().reshape(1, 64, h, w), style)
    rgb.square().sum().backward()
    out.append(float(rows.grad.sum()) + float(style.grad.sum()))
all_rays = torch.randn(61 * 47 + 40 * 30, 9, device=dev)
all_rgbs = torch.rand(all_rays.shape[0], 3, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
got = ops.grid_patch(all_rays, all_rgbs, torch.linspace(0, 1 - 1 / 61, 32), torch.linspace(0, 1 - 1 / 47, 32), 61.0, 47.0,
                     1200.0, 0.8, 0.1, 0.05, status=status)
torch.cuda.synchronize()
print("ok", len(out), int(status.item()))
