import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cr-nerf-pytorch_b200"))
import torch
from models.nerf import NeRF_sigma
torch.manual_seed(0)
args = types.SimpleNamespace(nerf_out_dim=64)
m = NeRF_sigma('fine', args, in_channels_xyz=93, in_channels_dir=27).cuda()
x = torch.randn(300, 120, device="cuda")
with torch.no_grad():
    y = m(x)
torch.cuda.synchronize()
print("ok", y.shape, float(y.abs().mean()))
