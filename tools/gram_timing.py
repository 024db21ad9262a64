#!/usr/bin/env python
"""Cycle stamps of the Gram kernel's block 0 (needs a library built with -DCRNERF_GRAM_TIMING:
CRNERF_DEFS=-DCRNERF_GRAM_TIMING CRNERF_OUT=tools/_ab/libgram_timing.so bash cr-nerf-pytorch_b200/csrc/build.sh,
then CRNERF_B200_LIB=tools/_ab/libgram_timing.so python tools/gram_timing.py [pixels])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from bench import build_models
from crnerf_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 640000
dev = torch.device("cuda", 0)
models, _ = build_models()
dec = models["decoder"].to(dev)
feat = torch.rand(n, 64, device=dev) * 0.2 + 0.4
content = feat.t().reshape(1, 64, n // 100 if n % 100 == 0 else 1, 100 if n % 100 == 0 else n)
style = torch.rand(1, 64, 32, 32, device=dev)
parts = torch.stack([c.sum(0) for c in feat.chunk(min(148, n))])
ts = torch.zeros(256, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3):
        dec(content, style, channel_sums=parts)
    torch.cuda.synchronize()
    ops.debug_set(ts.view(torch.float32), -1)
    dec(content, style, channel_sums=parts)
    torch.cuda.synchronize()
    ops.debug_set(None, -1)
t = ts.cpu().tolist()
t0 = t[0]
rel = lambda i: (t[i] - t0) if t[i] else None
print("prologue: weights staged", rel(1), "| setup done", rel(2), "| G read", rel(3), "| end", rel(4))
for tile in range(3):
    print(f"tile {tile} (stream 0): issuer L1 {rel(16+8*tile)} L2 {rel(17+8*tile)} L3 {rel(18+8*tile)} G {rel(19+8*tile)}")
    e = [rel(48 + 8 * tile + k) for k in range(8)]
    print(f"   epilogue warp 0: wait D1 {e[0]} -> got {e[1]} | ep1 done {e[2]} | staged next, wait D2 {e[3]} -> got {e[4]} | ep2 done {e[5]} | got D3 {e[6]} | ep3 done {e[7]}")
