#!/usr/bin/env python
"""Where the host-side ops of one training step come from (torch.profiler, with stacks):
python tools/prof_train_ops.py  ->  counts of aten ops per step and the top callers of fill/zero."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), os.path.join(ROOT, "tools"), ROOT):
    sys.path.insert(0, p)
import torch
import bench_train
dev = torch.device("cuda")
step, *_ = bench_train.make_step(dev, 1, 0)
for _ in range(3): step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
                            with_stack=True) as prof:
    step(); torch.cuda.synchronize()
ev = prof.events()
cnt = collections.Counter(e.name for e in ev if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::"))
print("aten ops per step:", sum(cnt.values()))
for k, v in cnt.most_common(25): print(f"  {v:4d} {k}")
callers = collections.Counter()
for e in ev:
    if e.name in ("aten::zeros", "aten::zeros_like", "aten::zero_", "aten::fill_") and e.stack:
        fr = [s for s in e.stack if "site-packages" not in s and "profiler" not in s][:2]
        callers[" <- ".join(fr)] += 1
print("fill/zero callers:")
for k, v in callers.most_common(15): print(f"  {v:4d} {k}")
kern = collections.Counter(e.name[:60] for e in ev if e.device_type == torch.autograd.DeviceType.CUDA)
print("kernels per step:", sum(kern.values()))
