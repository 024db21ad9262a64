"""Small cases of the round's later kernels for compute-sanitizer: the style/content encoder (odd
sizes, partial tiles, pooled halo), the fused loss forward/backward, the mask lookup, ray generation and
the uint8 output stage.
  compute-sanitizer --tool memcheck python tools/sanitize_case2.py"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402
import losses  # noqa: E402
from crnerf_b200 import loss, ops, synthetic  # noqa: E402
from models.linearStyleTransfer import encoder_sameoutputsize  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
enc = encoder_sameoutputsize(64).to(dev).eval()
sums = []
with torch.no_grad():
    for hw in ((8, 8), (9, 11), (40, 52), (33, 131)):
        sums.append(float(enc(torch.rand(1, 3, *hw, device=dev)).sum()))
hp = types.SimpleNamespace(maskrs_max=5e-2, maskrs_min=6e-3, maskrs_k=1e-3, maskrd=1e-3, weightKL=1e-5,
                           weightRecA=1e-3, weightcontent=1e-4, mse_on_appearance=False)
n = 777
inp = {"rgb_coarse": torch.rand(n, 3, device=dev, requires_grad=True),
       "rgb_fine": torch.rand(n, 3, device=dev, requires_grad=True),
       "a_embedded": torch.randn(1, 64, 8, 8, device=dev, requires_grad=True),
       "a_embedded_random": torch.randn(1, 64, 8, 8, device=dev),
       "a_embedded_random_rec": torch.randn(1, 64, 8, 8, device=dev, requires_grad=True)}
pred = torch.rand(1, 1, 24, 32, device=dev, requires_grad=True)
idx = torch.randint(0, 340 * 512, (n,), device=dev)
inp["out_mask"] = loss.mask_sample(pred, (340, 512), idx)
ret, _ = losses.CRNeRFLoss(hp)(inp, torch.rand(n, 3, device=dev), hp, 10)
sum(ret.values()).backward()
K = [[300.0, 0.0, 26.0], [0.0, 300.0, 20.0], [0.0, 0.0, 1.0]]
rays = ops.generate_rays(40, 52, K, synthetic.synthetic_pose(0).tolist(), 0.0, 5.0, device=dev)
u8 = ops.rgb_to_u8(torch.rand(1, 3, 40, 52, device=dev))
torch.cuda.synchronize()
print("sanitize case2 ok", sums, float(pred.grad.abs().sum()), tuple(rays.shape), int(u8.sum()))
