#!/usr/bin/env python
"""Host-side profile (cProfile) of the eager training step of tools/bench_train.py: where the Python time goes
between the kernel launches (the eager step is host-bound: 3.4 ms against 2.1 ms of GPU work)."""
import cProfile, os, pstats, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), os.path.join(ROOT, "tools"), ROOT):
    sys.path.insert(0, p)
import torch
import bench_train
step, *_ = bench_train.make_step(torch.device("cuda"), 1, 0)
for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
torch.cuda.synchronize()
pr.disable()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
