"""Per-tensor relative L2 error of the native encoder backward against float64 autograd (GPU)."""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT]
import torch
from models.linearStyleTransfer import encoder_sameoutputsize

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))

for hw in [(8, 8), (36, 52), (37, 51), (64, 200), (136, 264)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]:
    torch.manual_seed(11)
    enc = encoder_sameoutputsize(64)
    with torch.no_grad():
        for c in enc._convs():
            c.bias.mul_(3.0)
    enc = enc.cuda()
    gen = torch.Generator().manual_seed(hw[0] * 13 + hw[1])
    x = torch.rand(1, 3, *hw, generator=gen).cuda().requires_grad_(True)
    g = torch.randn(1, 64, 32, 32, generator=gen).cuda()
    out = enc(x)
    out.backward(g)
    ref = copy.deepcopy(enc).double()
    xr = x.detach().double().requires_grad_(True)
    o = ref._stack(xr)
    o.backward(g.double())
    print(hw, "fwd max abs", float((out.double() - o).abs().max()),
          " ".join(f"{n.replace('conv','c').replace('.weight','w').replace('.bias','b')}={rel(p.grad, q.grad):.1e}"
                   for (n, p), (_, q) in zip(enc.named_parameters(), ref.named_parameters())), f"x={rel(x.grad, xr.grad):.1e}")
