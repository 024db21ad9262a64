"""Small cases of the encoder's training path for compute-sanitizer: forward with the activation planes
kept (pool pass, planes epilogues of conv3 / conv5), input-gradient convolutions (kOutRaw), gradient
preparation passes (fold, pool routing), the tcgen05 weight-gradient kernels (tail segments, CTAs
without work), conv2 / conv1 / conv7 gradients.  Odd sizes, sizes below one segment, a width past 128.
  compute-sanitizer --tool memcheck python tools/sanitize_case4.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cr-nerf-pytorch_b200"))
import torch  # noqa: E402
from models.linearStyleTransfer import encoder_sameoutputsize  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
enc = encoder_sameoutputsize(64).to(dev)
out = []
for hw in ((8, 8), (9, 11), (37, 51), (20, 270)):
    x = torch.rand(1, 3, *hw, device=dev, requires_grad=True)
    enc.zero_grad(set_to_none=True)
    enc(x).square().sum().backward()
    out.append(float(x.grad.sum()) + sum(float(p.grad.sum()) for p in enc.parameters()))
torch.cuda.synchronize()
print("ok", len(out), sum(o != o for o in out))
