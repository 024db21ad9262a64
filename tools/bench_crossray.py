#!/usr/bin/env python
"""Cross-ray fusion + decoder (style_net.forward) alone, against its HBM roofline.

  python tools/bench_crossray.py [--sizes 800x800 320x256 32x32] [--reps 20]

Per size, one JSON line: ms per call (CUDA events over `reps` back-to-back calls; the map is larger
than L2 at 800x800, smaller maps are re-read from L2 - stated per line) for
  "sums"   : channel sums supplied as partial sums (what the render kernel's epilogue emits):
             2 reads of the map - SURVEY.md 8(d)'s algorithmic bytes H*W*(512+12) B + 8.6 MB
  "nosums" : plain style_net.forward(content, style): one more read to form the sums
  "sharded": the three-phase form the multi-GPU frame path uses (world 1: no collectives)
and frac = algorithmic bytes / time / measured HBM peak (MEASURED_PEAKS.json)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT):
    sys.path.insert(0, p)
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", nargs="+", default=["800x800", "320x256", "32x32"])
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    from bench import build_models
    from crnerf_b200.frame import CudaStyleBackend, fuse_decode_sharded
    from crnerf_b200 import ops
    dev = torch.device("cuda", 0)
    models, margs = build_models()
    dec = models["decoder"].to(dev)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    style = torch.rand(1, 64, 32, 32, generator=torch.Generator().manual_seed(1)).to(dev)
    for size in a.sizes:
        h, w = (int(v) for v in size.split("x"))
        n = h * w
        feat = torch.rand(n, 64, device=dev) * 0.2 + 0.4
        content = feat.t().reshape(1, 64, h, w)
        parts = torch.stack([c.sum(0) for c in feat.chunk(min(148, n))])    # stand-in for the render epilogue's rows
        be = CudaStyleBackend(dec)

        def timed(fn):
            with torch.no_grad():
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n0 = ops.launch_count()
                e0.record()
                for _ in range(a.reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
            return e0.elapsed_time(e1) / a.reps, (ops.launch_count() - n0) / a.reps

        ms_s, l_s = timed(lambda: dec(content, style, channel_sums=parts))
        ms_n, l_n = timed(lambda: dec(content, style))
        ms_h, l_h = timed(lambda: fuse_decode_sharded(be, feat, style, n, sum_parts=parts))
        with torch.no_grad():
            same = float((dec(content, style, channel_sums=parts) - dec(content, style)).abs().max())
        alg = n * (512 + 12) + 8.6e6
        print(json.dumps({"size": size, "pixels": n, "map_MB": n * 256 / 1e6,
                          "ms_sums": ms_s, "launches_sums": l_s, "ms_nosums": ms_n, "launches_nosums": l_n,
                          "ms_sharded": ms_h, "launches_sharded": l_h,
                          "algorithmic_MB": alg / 1e6, "achieved_GBs": alg / ms_s / 1e6, "hbm_peak_GBs": peak,
                          "frac": alg / ms_s / 1e6 / peak, "max_abs_diff_sums_vs_nosums": same,
                          "l2": "map > L2 (126 MB): streamed from HBM" if n * 256 > 126e6 else "map fits L2: re-reads hit L2"}))


if __name__ == "__main__":
    main()
