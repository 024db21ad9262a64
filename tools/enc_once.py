"""One 800x800 encoder forward (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cr-nerf-pytorch_b200"))
from models.linearStyleTransfer import encoder_sameoutputsize  # noqa: E402

torch.manual_seed(0)
enc = encoder_sameoutputsize(64).cuda().eval()
x = torch.rand(1, 3, 800, 800, device="cuda")
with torch.no_grad():
    enc(x)
    enc(x)
torch.cuda.synchronize()
