"""Times crnerf_sample_pdf_merge (4096 rays, 64 coarse + 128 importance samples, shared linspace u).
usage: [CRNERF_B200_LIB=path.so] python tools/time_sample.py"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT): sys.path.insert(0, p)
import torch
from crnerf_b200 import ops
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
z = torch.sort(torch.rand(4096, 64, generator=g) * 4 + 0.5, dim=1)[0].to(dev)
w = torch.rand(4096, 64, generator=g).to(dev) ** 4
u = torch.linspace(0, 1, 128, device=dev)
for _ in range(5): ops.sample_pdf_merge(z, w, u, 128)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
torch.cuda.synchronize()
for e0, e1 in ev:
    e0.record(); out = ops.sample_pdf_merge(z, w, u, 128); e1.record()
torch.cuda.synchronize()
ts = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
ok = bool((out[:, 1:] >= out[:, :-1]).all())
print(f"{os.environ.get('CRNERF_B200_LIB', 'default'):28s} sample_pdf_merge median {statistics.median(ts)*1e3:.1f} us  sorted={ok}")
