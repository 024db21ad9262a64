"""Small end-to-end case for compute-sanitizer: eval render (coarse + resample + fine), style_net decode,
sharded cross-ray phases, and one training forward/backward.
  compute-sanitizer --tool memcheck  python tools/sanitize_case.py
  compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT): sys.path.insert(0, p)
import torch
from crnerf_b200 import synthetic
from bench import build_models
from models.nerf import PosEmbedding
from models.rendering import render_rays_cross_ray
from crnerf_b200.frame import CudaStyleBackend, fuse_decode_sharded
models, margs = build_models(); dev = torch.device("cuda")
models = {k: m.to(dev) for k, m in models.items()}
emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
rays = synthetic.pinhole_rays(12, 16, synthetic.synthetic_pose(0)).to(dev)
style = torch.rand(1, 64, 32, 32, device=dev)
with torch.no_grad():
    res = render_rays_cross_ray(models, emb, rays, None, 40, False, 0, 0, 24, 4096, False, test_time=True, args=margs)
    rgb = models["decoder"](res["feature_fine"].t().reshape(1, 64, 12, 16), style)
    rgb2 = fuse_decode_sharded(CudaStyleBackend(models["decoder"]), res["feature_fine"], style, 192)
for m in models.values(): m.train()
res = render_rays_cross_ray(models, emb, rays, None, 32, False, 1.0, 1.0, 32, 4096, False, args=margs)
(res["feature_fine"].sum() + res["feature_coarse"].sum()).backward()
torch.cuda.synchronize()
print("sanitize case ok", float(rgb.mean()), float(rgb2.mean()), float(models["fine"].xyz_encoding_1[0].weight.grad.abs().sum()))
