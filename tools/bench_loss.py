"""Time the loss tail of a training step (1024 rays): the `losses` mirror (csrc/loss.cu) against
the same terms written as stock PyTorch tensor ops on the same GPU.  Prints one JSON line."""
import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cr-nerf-pytorch_b200"))
import losses  # noqa: E402
from crnerf_b200 import ops  # noqa: E402


def eager_terms(inp, t, hp, w):
    m = inp["out_mask"]
    d = inp["a_embedded_random"].detach() - inp["a_embedded_random_rec"]
    return [torch.mean(inp["a_embedded"] ** 2) * hp.weightKL, d.abs().mean() * hp.weightRecA,
            0.5 * ((1 - m.detach()) * (inp["rgb_coarse"] - t) ** 2).mean(),
            ((inp["content_wo_a_embed"] - inp["content_with_a_embed"]) ** 2).mean() * hp.weightcontent,
            torch.mean(m ** 2) * w, torch.mean(1 / ((m - 0.5) ** 2 + 0.02)) * hp.maskrd,
            0.5 * ((1 - m) * (inp["rgb_fine"] - t) ** 2).mean()]


def main():
    dev = torch.device("cuda:0")
    hp = types.SimpleNamespace(maskrs_max=5e-2, maskrs_min=6e-3, maskrs_k=1e-3, maskrd=1e-3, weightKL=1e-5,
                               weightRecA=1e-3, weightcontent=1e-4, mse_on_appearance=False)
    n = 1024
    inp = {"rgb_coarse": torch.rand(n, 3, device=dev), "rgb_fine": torch.rand(n, 3, device=dev),
           "out_mask": torch.rand(n, 1, device=dev)}
    for k in ("a_embedded", "a_embedded_random", "a_embedded_random_rec", "content_wo_a_embed",
              "content_with_a_embed"):
        inp[k] = torch.randn(1, 64, 32, 32, device=dev)
    for v in inp.values():
        v.requires_grad_(True)
    t = torch.rand(n, 3, device=dev)
    crit = losses.CRNeRFLoss(hp)

    def ours():
        ret, _ = crit(inp, t, hp, 100)
        sum(ret.values()).backward()

    def eager():
        sum(eager_terms(inp, t, hp, crit.Annealing.getWeight(100))).backward()

    out = {}
    for name, fn in (("ours", ours), ("eager", eager)):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        t0 = time.perf_counter()
        for _ in range(200):
            fn()
        torch.cuda.synchronize()
        out[name + "_us"] = (time.perf_counter() - t0) / 200 * 1e6
        out[name + "_native_launches"] = (ops.launch_count() - n0) / 200
    print(json.dumps(out))


if __name__ == "__main__":
    main()
