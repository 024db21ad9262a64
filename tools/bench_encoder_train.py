"""Forward + backward of encoder_sameoutputsize under autograd (the training step's enc_a, reference
models/linearStyleTransfer.py:250-276) on one B200: the native path (csrc/encoder.cu forward with the
activation planes kept, csrc/encoder_train.cuh backward) against the same module on library
convolutions in strict fp32 (what matches the reference's results) and with TF32 allowed (torch's
default).  Prints one JSON line; --profile lists the native path's kernels (torch profiler)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cr-nerf-pytorch_b200"))
from models.linearStyleTransfer import encoder_sameoutputsize  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    enc = encoder_sameoutputsize(64).to(dev).train()
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for hw in ((340, 512), (800, 800), (32, 32)):
        x = torch.rand(1, 3, *hw, device=dev)
        g = torch.randn(1, 64, 32, 32, device=dev)
        for name in ("native", "library_fp32", "library_tf32"):
            enc.train_backend = "native" if name == "native" else "library"
            enc.conv_precision = "tf32" if name.endswith("tf32") else "fp32"

            def step():
                enc.zero_grad(set_to_none=True)
                enc(x).backward(g)

            for _ in range(3):
                step()
            torch.cuda.synchronize()
            n = 10
            e0, e1, e2 = ev(), ev(), ev()
            fwd = 0.0
            e0.record()
            for _ in range(n):
                step()
            e1.record()
            torch.cuda.synchronize()
            out[f"{hw[0]}x{hw[1]}_{name}_ms"] = round(e0.elapsed_time(e1) / n, 4)
            if name == "native":
                e0.record()
                for _ in range(n):
                    y = enc(x)
                e1.record()
                torch.cuda.synchronize()
                out[f"{hw[0]}x{hw[1]}_native_forward_ms"] = round(e0.elapsed_time(e1) / n, 4)
        if "--profile" in sys.argv and hw == (340, 512):
            enc.train_backend = "native"
            with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
                enc.zero_grad(set_to_none=True)
                enc(x).backward(g)
                torch.cuda.synchronize()
            for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:30]:
                print(f"{e.device_time_total / 1e3:8.3f} ms  x{e.count:<3d} {e.key[:120]}", file=sys.stderr)
    out["data"] = "synthetic"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
