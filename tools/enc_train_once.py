"""Two forward + backward passes of the encoder's training path on a 340x512 photo (for ncu captures;
the second pass is the warm one)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cr-nerf-pytorch_b200"))
from models.linearStyleTransfer import encoder_sameoutputsize  # noqa: E402

hw = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (340, 512)
torch.manual_seed(0)
enc = encoder_sameoutputsize(64).cuda().train()
x = torch.rand(1, 3, *hw, device="cuda")
g = torch.randn(1, 64, 32, 32, device="cuda")
for _ in range(2):
    enc.zero_grad(set_to_none=True)
    enc(x).backward(g)
torch.cuda.synchronize()
