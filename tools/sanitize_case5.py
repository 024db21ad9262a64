"""Small training steps for compute-sanitizer over the kernels the third round-2 pass changed: the training
forward's 32-byte row stores and split feature / sigma save, the composite backward (block per ray: S <= 128
keeps rows in registers, S > 128 reloads them; a sample count that is not a multiple of 32), dgrad / wgrad
with the staged bulk-store epilogue and descending tile order (several tiles per CTA and a ragged last tile),
the cross-ray chain backward on 8-pixel tiles, and crnerf_b200.optim.Adam (68 tensors = two launches).
  compute-sanitizer --tool memcheck python tools/sanitize_case5.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402
from crnerf_b200 import synthetic  # noqa: E402
from crnerf_b200.optim import Adam  # noqa: E402
from bench import build_models  # noqa: E402
from models.nerf import PosEmbedding  # noqa: E402
from models.rendering import render_rays_cross_ray  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
models, margs = build_models()
models = {k: m.to(dev).train() for k, m in models.items()}
emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
params = [p for m in models.values() for p in m.parameters()]
opt = Adam(params, lr=5e-4, weight_decay=1e-4)
losses = []
# (rays side, coarse samples, fine samples): 12x12 rays x 40 -> 45 tiles (ragged); 8x8 x (72 + 72 = 144 > 128: reload path)
for side, ns, ni in ((12, 40, 24), (8, 72, 72), (6, 33, 17)):
    rays = synthetic.pinhole_rays(side, side, synthetic.synthetic_pose(0)).to(dev)
    style = torch.rand(1, 64, 9, 7, device=dev)
    target = torch.rand(side * side, 3, device=dev)
    for _ in range(2):
        res = render_rays_cross_ray(models, emb, rays, None, ns, False, 1.0, 1.0, ni, 32768, False, args=margs)
        loss = 0
        for typ in ("coarse", "fine"):
            feat = res[f"feature_{typ}"].t().reshape(1, 64, side, side)
            rgb = models["decoder"](feat, style).reshape(3, -1).t()
            loss = loss + 0.5 * ((rgb - target) ** 2).mean() + 1e-3 * res[f"depth_{typ}"].mean() + 1e-3 * res[f"weights_{typ}"].square().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss))
torch.cuda.synchronize()
print("ok", len(losses), sum(l != l for l in losses), float(opt.state[params[0]]["step"]))
