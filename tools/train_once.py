#!/usr/bin/env python
"""A few eager training steps (tools/bench_train.py's step) for ncu captures:
ncu --set full -k regex:'render_fused|dgrad_kernel|wgrad_kernel' -s 100 -c 30 python tools/train_once.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), os.path.join(ROOT, "tools"), ROOT):
    sys.path.insert(0, p)
import torch
import bench_train
step, *_ = bench_train.make_step(torch.device("cuda"), 1, 0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    step()
torch.cuda.synchronize()
print("done")
