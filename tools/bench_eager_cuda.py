"""Library-kernel baseline on the same GPU: the reference's math (oracle restatement,
device agnostic) executed by stock PyTorch CUDA ops, fp32, TF32 off / on, 4096 rays x (64+128), eval mode."""
import os, sys, json, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), os.path.join(ROOT, "oracle"), ROOT): sys.path.insert(0, p)
import torch
from crnerf_b200 import synthetic
import crnerf_oracle as oracle
from bench import build_models
models, _ = build_models(); dev = torch.device("cuda")
pc = {k: v.to(dev) for k, v in models["coarse"].state_dict().items()}
pf = {k: v.to(dev) for k, v in models["fine"].state_dict().items()}
rays = synthetic.pinhole_rays(64, 64, synthetic.synthetic_pose(0)).to(dev)
rng = {"noise_coarse": torch.zeros(4096, 64, device=dev), "noise_fine": torch.zeros(4096, 192, device=dev)}
out = {}
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    def step():
        with torch.no_grad():
            return oracle.render_rays(pc, pf, rays, n_samples=64, n_importance=128, perturb=0, noise_std=0, chunk=8192 * 16, rng=rng)
    for _ in range(3): step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    torch.cuda.synchronize()
    for e0, e1 in ev:
        e0.record(); step(); e1.record()
    torch.cuda.synchronize()
    ms = statistics.median(e0.elapsed_time(e1) for e0, e1 in ev)
    out["tf32" if tf32 else "fp32"] = {"ms_per_batch": ms, "ray_samples_per_s": 4096 * 192 / (ms * 1e-3)}
print(json.dumps({"workload": "stock PyTorch CUDA eager, reference math, 4096 rays x (64+128)", **out}))
