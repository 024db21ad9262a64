"""Time encoder_sameoutputsize (reference linearStyleTransfer.py:208-276; SURVEY 8f rank 1, kept as
library ops) on one B200 so its share of a frame is on record.  Prints one JSON line."""
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
    for hw in ((800, 800), (340, 512), (32, 32)):
        x = torch.rand(1, 3, *hw, device=dev)
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad():
                for _ in range(5):
                    enc(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    enc(x)
                e1.record()
                torch.cuda.synchronize()
            out[f"{hw[0]}x{hw[1]}_{'tf32' if tf32 else 'fp32'}_ms"] = e0.elapsed_time(e1) / 20
    print(json.dumps(out))


if __name__ == "__main__":
    main()
