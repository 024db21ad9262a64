#!/usr/bin/env python
"""Where the GPU time of one training step goes (torch.profiler CUDA activity, per kernel name):
python tools/prof_train_kernels.py [--eager]  ->  launches and microseconds per step by kernel, busy time vs step time.
The step is tools/bench_train.py's (1024-ray patch x (64+64), decode x2, MSE, backward, Adam), replayed as one CUDA graph."""
import os, sys, collections, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), os.path.join(ROOT, "tools"), ROOT):
    sys.path.insert(0, p)
import torch
import bench_train

ap = argparse.ArgumentParser()
ap.add_argument("--eager", action="store_true")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda")
if a.eager:
    step, *_ = bench_train.make_step(dev, 1, 0)
else:
    from crnerf_b200.graphs import GraphedTrainStep
    g, *_ = bench_train.make_step(dev, 1, 0, capturable=True)
    step = GraphedTrainStep(g.loss_fn, g.opt)
for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step()
e1.record(); torch.cuda.synchronize()
print(f"step (unprofiled): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(a.reps):
        step()
    torch.cuda.synchronize()
ks = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks.sort(key=lambda e: e.time_range.start)
tot = collections.Counter(); cnt = collections.Counter()
for e in ks:
    n = e.name.replace("crnerf::<unnamed>::", "").replace("void ", "")[:90]
    tot[n] += e.time_range.end - e.time_range.start; cnt[n] += 1
busy = sum(tot.values()) / a.reps
span = (ks[-1].time_range.end - ks[0].time_range.start) / a.reps
print(f"kernels per step {len(ks) / a.reps:.0f}, busy {busy:.1f} us / step, span {span:.1f} us / step")
for n, v in tot.most_common(45):
    print(f"{v / a.reps:9.1f} us {cnt[n] / a.reps:6.1f} x  {n}")
# the timeline of one step in launch order (coalesced runs of the same kernel)
per = len(ks) // a.reps
one = ks[per * (a.reps - 1):]
print("timeline of the last step (start offset us, duration us, gap before us):")
t0 = one[0].time_range.start; prev_end = t0
for e in one:
    n = e.name.replace("crnerf::<unnamed>::", "").replace("void ", "")[:70]
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f} {e.time_range.start - prev_end:7.1f}  {n}")
    prev_end = e.time_range.end
