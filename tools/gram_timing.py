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
ts = torch.zeros(1024, dtype=torch.int64, device=dev)
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
print("prologue: weights staged", t[1] - t0, "| setup done", t[2] - t0, "| G read", t[3] - t0, "| end", t[4] - t0)
NAMES = {10: "wait D1(0)", 11: "wait D1(1)", 12: "got D1(0)", 13: "got D1(1)", 14: "ep1(0) done", 15: "ep1(1) done",
         16: "staged(0)", 17: "staged(1)", 20: "got D2(0)", 21: "got D2(1)", 22: "ep2(0) done", 23: "ep2(1) done",
         30: "got D3(0)", 31: "got D3(1)", 32: "ep3(0) done", 33: "ep3(1) done",
         40: "issue L2(0)", 41: "issue L3(0)", 42: "issue G(0)", 43: "issue L1(0)",
         44: "issue L2(1)", 45: "issue L3(1)", 46: "issue G(1)", 47: "issue L1(1)"}
ev = []
for who, base, cnt in (("w0 ", 256, 128), ("w15", 384, 128), ("isC", 600, 100), ("isN", 700, 100)):
    for v in t[base:base + cnt]:
        if v:
            ev.append(((v & 0xffffffffffff) - (t0 & 0xffffffffffff), who, NAMES.get(v >> 48, str(v >> 48))))
ev.sort()
limit = int(os.environ.get("TRACE_UNTIL", "60000"))
for c, who, name in ev:
    if c < limit:
        print(f"{c:8d}  {who}  {name}")

for k, name in enumerate(["got D1(0)", "ep1(0) done", "staged(0)", "ep2(0) done"]):
    vals = [(v & 0xffffffffffff) - (t0 & 0xffffffffffff) for v in t[800 + 20 * k: 800 + 20 * k + 16]]
    print(f"round 2, all warps, {name:12s}:", " ".join(str(v) for v in vals))

fine = [(v & 0xffffffffffff) - (t0 & 0xffffffffffff) if v else None for v in t[900:916]]
names = {0: "ep1: wait D1", 1: "got D1", 2: "ld done", 3: "math done", 4: "st + arrive done", 5: "stage_x done", 6: "load_x issued",
         10: "ep3: wait D3", 11: "got D3", 12: "ld done", 13: "G_DONE ok", 14: "Yt stored", 15: "fence + arrive done"}
print("warp 5, round 2, stream 0 (cycles since kernel start; deltas in brackets):")
prev = None
for k in sorted(names):
    if fine[k] is not None:
        print(f"  {names[k]:22s} {fine[k]:8d}" + (f"  [+{fine[k] - prev}]" if prev is not None else ""))
        prev = fine[k]
