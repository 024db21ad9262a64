"""Time encoder_sameoutputsize (reference linearStyleTransfer.py:208-276) on one
B200: csrc/encoder.cu (the module's no-grad path) against the same module on library ops (cuDNN)
with TF32 allowed (torch's default) and in strict fp32, plus each path's max error against a
float64 evaluation of the library path.  Prints one JSON line."""
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
    enc = encoder_sameoutputsize(64).to(dev).eval()
    out = {}
    import copy
    enc64 = copy.deepcopy(enc).double()

    def library(x):          # the module's autograd/library path, without recording a graph
        h = enc.relu2(enc.conv2(enc.reflecPad1(enc.conv1(x))))
        h = enc.relu3(enc.conv3(enc.reflecPad3(h)))
        h, _ = enc.maxPool(h)
        h = enc.relu4(enc.conv4(enc.reflecPad4(h)))
        h = enc.relu5(enc.conv5(enc.reflecPad5(h)))
        h, _ = enc.maxPool2(h)
        h = enc.relu6(enc.conv6(enc.reflecPad6(h)))
        return enc.relu7(enc.conv7(enc.adppool(h)))

    for hw in ((800, 800), (340, 512), (32, 32)):
        x = torch.rand(1, 3, *hw, device=dev)
        with torch.no_grad():
            x64 = x.double()
            h = enc64.relu2(enc64.conv2(enc64.reflecPad1(enc64.conv1(x64))))
            h = enc64.relu3(enc64.conv3(enc64.reflecPad3(h)))
            h, _ = enc64.maxPool(h)
            h = enc64.relu4(enc64.conv4(enc64.reflecPad4(h)))
            h = enc64.relu5(enc64.conv5(enc64.reflecPad5(h)))
            h, _ = enc64.maxPool2(h)
            h = enc64.relu6(enc64.conv6(enc64.reflecPad6(h)))
            truth = enc64.relu7(enc64.conv7(enc64.adppool(h)))
        for name in ("ours", "tf32", "fp32"):
            torch.backends.cudnn.allow_tf32 = name == "tf32"
            torch.backends.cuda.matmul.allow_tf32 = name == "tf32"
            fn = enc if name == "ours" else library
            with torch.no_grad():
                for _ in range(5):
                    y = fn(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    fn(x)
                e1.record()
                torch.cuda.synchronize()
            out[f"{hw[0]}x{hw[1]}_{name}_ms"] = round(e0.elapsed_time(e1) / 20, 4)
            out[f"{hw[0]}x{hw[1]}_{name}_max_rel_err"] = float(((y.double() - truth).abs() /
                                                                (truth.abs() + 1e-3)).max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
