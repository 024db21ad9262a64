"""Times the fused fine-pass kernel alone (4096 rays x 192 samples), CUDA events, L2 flushed between reps.
usage: [CRNERF_B200_LIB=path.so] python tools/time_kernel.py [reps]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT): sys.path.insert(0, p)
import torch
from crnerf_b200 import synthetic
from bench import build_models
from crnerf_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
models, _ = build_models(); dev = torch.device("cuda")
fine = models["fine"].to(dev)
rays = synthetic.pinhole_rays(64, 64, synthetic.synthetic_pose(0)).to(dev)
z = ops.coarse_z(rays, torch.linspace(0, 1, 192, device=dev))
packed = fine.packed()
flush = torch.empty(64 * 1024 * 1024, device=dev)
for _ in range(5): ops.render_pass(packed, rays, z)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
torch.cuda.synchronize()
for e0, e1 in ev:
    flush.fill_(1.0); e0.record(); ops.render_pass(packed, rays, z); e1.record()
torch.cuda.synchronize()
ts = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
print(f"{os.environ.get('CRNERF_B200_LIB', 'default'):40s} median {statistics.median(ts)*1e3:.1f} us  min {ts[0]*1e3:.1f} us  ({4096*192/statistics.median(ts)/1e3:.1f} M ray-samples/s fine pass alone)")
