#!/usr/bin/env python
"""Measured parity of the CUDA path against the reference's golden outputs, per output.

  python tools/parity_report.py > gpurun_out/parity_report.json        (needs a B200)

For every render golden (default-init, "peaky" and the trained-like set) and every output it
prints the max relative error over the entries that matter (|ref| above a floor), the max
absolute error, and - for ``weights`` / ``depth`` - the error in units of the conditioning bound
used by the tests: the composite multiplies alpha by the transmittance T = exp(-tau), so a relative
error eps in the network's sigma (the 1e-4 north-star bar applies there) becomes up to
(1 + tau)*eps in the weight of a sample at optical depth tau.  tests/test_gpu_parity.py asserts
the same quantities; this tool records the measured values (profiles/r02_parity_report.json).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402

from conftest import build_mirror_models, load_golden  # noqa: E402
from parity_bounds import composite_bounds, rel_err  # noqa: E402


def stagewise(models, g_rays, passes, noise_std):
    from crnerf_b200 import ops
    out = {}
    for typ, z, noise, ref in passes:
        m = models[typ].cuda()
        with torch.no_grad():
            packed = m.packed()
        nz = None if (noise is None or noise_std == 0) else noise.cuda()
        w, f, d = ops.render_pass(packed, g_rays.cuda(), z.contiguous().cuda(), nz)
        bw, bd = composite_bounds(ref[f"weights_{typ}"], z, rtol=1e-4)
        out[typ] = {
            "feature_max_rel": rel_err(f, ref[f"feature_{typ}"], floor=1e-3),
            "feature_max_abs": float((f.cpu() - ref[f"feature_{typ}"]).abs().max()),
            "weights_max_rel_above_1e-3": rel_err(w, ref[f"weights_{typ}"], floor=1e-3),
            "weights_max_abs": float((w.cpu() - ref[f"weights_{typ}"]).abs().max()),
            "weights_err_over_bound": float(((w.cpu() - ref[f"weights_{typ}"]).abs() / bw).max()),
            "depth_max_rel": rel_err(d, ref[f"depth_{typ}"], floor=1e-3),
            "depth_max_abs": float((d.cpu() - ref[f"depth_{typ}"]).abs().max()),
            "depth_err_over_bound": float(((d.cpu() - ref[f"depth_{typ}"]).abs() / bd).max()),
            "max_optical_depth_at_visible_samples": float(
                (-torch.log(torch.clamp(1 - torch.cumsum(ref[f"weights_{typ}"], 1), min=1e-30)))[ref[f"weights_{typ}"] > 1e-3].max())
            if (ref[f"weights_{typ}"] > 1e-3).any() else 0.0,
        }
    return out


def main():
    report = {}
    for name in ("render_c64_eval", "render_64p128_eval", "render_64p128_eval_peaky", "render_64p64_train",
                 "render_32p24_train_peaky", "render_48p48_disp"):
        g = load_golden(name)
        models, _ = build_mirror_models(g["seed"], g["peaky"])
        passes = [("coarse", g["z_coarse"], g["rng"].get("noise_coarse"), g["ref"])]
        if g["n_importance"] > 0:
            passes.append(("fine", g["z_fine"], g["rng"].get("noise_fine"), g["ref"]))
        report[name] = stagewise(models, g["rays"], passes, g["noise_std"])
    path = os.path.join(ROOT, "tests", "golden", "trained.pt")
    if os.path.exists(path):
        t = torch.load(path, map_location="cpu", weights_only=False)
        models, _ = build_mirror_models(0)
        models["coarse"].load_state_dict(t["coarse"])
        models["fine"].load_state_dict(t["fine"])
        for cname in ("eval_64p128", "train_64p64"):
            c = t["cases"][cname]
            passes = [("coarse", c["z_coarse"], c["rng"].get("noise_coarse"), c["ref"]),
                      ("fine", c["z_fine"], c["rng"].get("noise_fine"), c["ref"])]
            report["trained_" + cname] = stagewise(models, c["rays"], passes, 1.0 if c["train"] else 0)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
