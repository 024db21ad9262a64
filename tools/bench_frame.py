#!/usr/bin/env python
"""Whole-frame render (BASELINE configs[2]/[3]): rays sharded across the ranks of one box,
cross-ray fusion + decoder in the sharded form, one rgb all-gather at the end.

  python tools/bench_frame.py [--hw 256 320] [--ns 64 --ni 128] [--reps 3]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_frame.py ...

Prints one JSON line per run on rank 0: frame time (CUDA events, max over ranks), ray-samples/s,
and the stand-alone timing of the cross-ray block against its HBM roofline
(algorithmic bytes = 3 reads of the (H*W,64) fp32 feature map + 12 B/pixel of rgb + 8.6 MB of FC weights)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from crnerf_b200 import synthetic
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", type=int, nargs=2, default=[256, 320])
    ap.add_argument("--ns", type=int, default=64)
    ap.add_argument("--ni", type=int, default=128)
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--scheme", default="stats")
    a = ap.parse_args()
    from bench import build_models
    from models.nerf import PosEmbedding
    from crnerf_b200.frame import render_frame_sharded, CudaStyleBackend, fuse_decode_sharded
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    models, margs = build_models()
    models = {k: m.to(dev) for k, m in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    h, w = a.hw
    rays = synthetic.pinhole_rays(h, w, synthetic.synthetic_pose(0)).to(dev)
    style = torch.rand(1, 64, 32, 32, generator=torch.Generator().manual_seed(1)).to(dev)

    def frame():
        return render_frame_sharded(models, emb, rays, style, (h, w), a.ns, a.ni, chunk=a.chunk,
                                    scheme=a.scheme, args=margs)
    def sync():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    frame(); sync()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync(); e0.record(); rgb = frame(); e1.record(); sync()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([min(ts)], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # cross-ray block alone on this rank's share (features resident)
    n = h * w
    feat = torch.rand(-(-n // world), 64, device=dev)
    be = CudaStyleBackend(models["decoder"])
    fuse_decode_sharded(be, feat, style, feat.shape[0] * world); sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fuse_decode_sharded(be, feat, style, feat.shape[0] * world)
    e1.record(); sync()
    cr_ms = e0.elapsed_time(e1) / 5
    # ray generation alone (crnerf_generate_rays): 32 B written per ray, nothing read
    import math
    from crnerf_b200 import ops
    fl = 0.5 * w / math.tan(math.radians(30.0))
    Kc = [[fl, 0.0, w / 2], [0.0, fl, h / 2], [0.0, 0.0, 1.0]]
    pose = synthetic.synthetic_pose(0).tolist()
    ops.generate_rays(h, w, Kc, pose, 0.0, 5.0, device=dev); sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.generate_rays(h, w, Kc, pose, 0.0, 5.0, device=dev)
    e1.record(); sync()
    rg_ms = e0.elapsed_time(e1) / 10
    # the whole eval loop of one image (reference eval.py:275-297): enc_a(photo) -> rays built on
    # the GPU -> render -> cross-ray fusion + decode -> uint8 -> pinned host buffer
    from models.linearStyleTransfer import encoder_sameoutputsize
    torch.manual_seed(1)
    enc_a = encoder_sameoutputsize(64).to(dev).eval()
    photo = torch.rand(1, 3, h, w, device=dev)
    host_u8 = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()

    def eval_image():
        with torch.no_grad():
            a_emb = enc_a(photo)
        img = render_frame_sharded(models, emb, None, a_emb, (h, w), a.ns, a.ni, chunk=a.chunk, scheme=a.scheme,
                                   args=margs, camera=(Kc, pose, 0.0, 5.0))
        host_u8.copy_(ops.rgb_to_u8(img), non_blocking=True)
    eval_image(); sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eval_image(); e1.record(); sync()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    eval_ms = float(t.item())
    e0.record()
    with torch.no_grad():
        for _ in range(5): enc_a(photo)
    e1.record(); sync()
    enc_ms = e0.elapsed_time(e1) / 5
    if rank == 0:
        bytes_alg = feat.shape[0] * (3 * 256 + 12) + 8.6e6
        peak = 6555.5
        try: peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception: pass
        print(json.dumps({"workload": f"{h}x{w} frame, {a.ns}+{a.ni} samples, style_net fusion+decoder, scheme {a.scheme}",
                          "n_gpus": world, "frame_ms": ms, "ray_samples_per_s": n * (a.ns + a.ni) / (ms * 1e-3),
                          "crossray_ms_per_rank": cr_ms, "crossray_alg_bytes_per_rank": bytes_alg,
                          "crossray_gbs": bytes_alg / (cr_ms * 1e-3) / 1e9, "hbm_peak_gbs": peak,
                          "crossray_frac": bytes_alg / (cr_ms * 1e-3) / 1e9 / peak,
                          "raygen_ms": rg_ms, "raygen_gbs": n * 32 / (rg_ms * 1e-3) / 1e9,
                          "raygen_frac": n * 32 / (rg_ms * 1e-3) / 1e9 / peak,
                          "eval_image_ms": eval_ms, "encoder_ms": enc_ms, "u8_checksum": int(host_u8.sum()),
                          "rgb_mean": float(rgb.mean())}), flush=True)
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
