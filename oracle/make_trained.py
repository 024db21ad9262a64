"""TEST INFRASTRUCTURE - a "trained-like" weight set, made with the UNMODIFIED reference modules,
and golden vectors on it (default-init weights leave features in [0.44, 0.56] and two
nearly constant images in the PSNR test).

Pretrained CR-NeRF checkpoints are Google-Drive links (reference README.md:138-142) and cannot be
fetched here, so this script trains the reference's own modules on CPU with Adam on a seeded
synthetic target for a few hundred steps:

  * ``NeRF_sigma`` coarse and fine (reference models/nerf.py:115-182, through ``PosEmbedding``):
    supervised per point on an analytic scene sampled along the synthetic camera's rays - a dense
    sphere and a ground slab in empty space (sigma* in {0, 12, 25}: sharp surfaces, strongly peaked
    ray weights) with a 64-channel sinusoidal feature field that also depends on the view direction.
  * ``style_net`` (reference models/linearStyleTransfer.py:278-291) on 32x32 feature patches
    rendered by the reference's ``render_rays_cross_ray`` from the trained NeRFs: every parameter
    except the two 1024x1024 ``fc`` layers (kept at their seeded default init so that the fixture
    stays small) learns to paint the patch's structure in the colour carried by the style feature.

Then the oracle is pinned on the trained weights (``torch.equal`` against the reference, as in
make_golden.py) and inputs + reference outputs + the trained tensors are written to
``tests/golden/trained.pt``.  Build container only (needs /root/reference):

    python oracle/make_trained.py
"""
from __future__ import annotations

import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import crnerf_oracle as oracle  # noqa: E402
import make_golden as mg  # noqa: E402

OUT = os.path.join(mg.GOLD, "trained.pt")
FROZEN = ("multi_net.snet.fc.", "multi_net.cnet.fc.")      # regenerated from the seed by the tests


def scene(x, d, g_feat):
    """Analytic target: sigma* (N,), features* (N,64) in (0.05, 0.95)."""
    c = torch.tensor([0.15, -0.05, -2.6])
    in_sphere = ((x - c).norm(dim=-1) < 0.85)
    in_slab = (x[:, 1] < -0.95) & (x[:, 2] < -0.8) & (x[:, 2] > -4.6) & (x[:, 0].abs() < 2.2)
    sigma = torch.where(in_sphere, torch.tensor(25.0), torch.where(in_slab, torch.tensor(12.0), torch.tensor(0.0)))
    A, phase, B = g_feat
    f = 0.5 + 0.4 * torch.sin(x @ A + phase) + 0.05 * torch.cos(d @ B)
    return sigma, f.clamp(0.05, 0.95)


def feature_field(seed=77):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(3, 64, generator=g) * 2.2
    phase = torch.rand(64, generator=g) * 6.28318
    B = torch.randn(3, 64, generator=g) * 1.5
    return A, phase, B


def train_nerf(model, nerf, steps, seed, field, log):
    e_x, e_d = nerf.PosEmbedding(14, 15), nerf.PosEmbedding(3, 4)
    g = torch.Generator().manual_seed(seed)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.train()
    pool = torch.cat([oracle.pinhole_rays(48, 64, oracle.synthetic_pose(s)) for s in range(6)])
    for it in range(steps):
        idx = torch.randint(0, pool.shape[0], (8192,), generator=g)
        r = pool[idx]
        # half of the samples uniform in depth, half concentrated where the surfaces are
        z = torch.rand(8192, generator=g) * 5.0
        x = r[:, 0:3] + r[:, 3:6] * z[:, None]
        d = r[:, 3:6]
        sig_t, f_t = scene(x, d, field)
        out = model(torch.cat([e_x(x), e_d(d)], 1))
        f, sig = out[:, :64], out[:, 64]
        occ = (sig_t > 0).float()
        loss = ((torch.log1p(sig) - torch.log1p(sig_t)) ** 2).mean() + \
               (((f - f_t) ** 2).mean(1) * (0.15 + occ)).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if it % 100 == 0 or it == steps - 1:
            log(f"    step {it:4d} loss {float(loss):.4f}")
    model.eval()


def style_bank(n, seed):
    """Style features with a colour: channels carry offsets that the decoder learns to read."""
    g = torch.Generator().manual_seed(seed)
    styles, tints = [], []
    for _ in range(n):
        tint = torch.rand(3, generator=g) * 0.8 + 0.1
        s = torch.rand(1, 64, 32, 32, generator=g) * 0.4
        s[:, 0:21] += tint[0] * 0.6
        s[:, 21:42] += tint[1] * 0.6
        s[:, 42:64] += tint[2] * 0.6
        styles.append(s)
        tints.append(tint)
    return styles, tints


def train_decoder(dec, patches, steps, seed, log):
    styles, tints = style_bank(8, seed)
    params = [p for n, p in dec.named_parameters() if not n.startswith(FROZEN)]
    for n, p in dec.named_parameters():
        p.requires_grad_(not n.startswith(FROZEN))
    opt = torch.optim.Adam(params, lr=2e-3)
    g = torch.Generator().manual_seed(seed + 1)
    dec.train()
    for it in range(steps):
        pi = int(torch.randint(0, len(patches), (1,), generator=g))
        si = int(torch.randint(0, len(styles), (1,), generator=g))
        content = patches[pi]
        lum = content[:, :8].mean(1, keepdim=True)
        lum = (lum - lum.min()) / (lum.max() - lum.min() + 1e-6)
        target = (0.08 + 0.9 * lum * tints[si].reshape(1, 3, 1, 1) +
                  0.25 * (content[:, 8:11] - 0.5)).clamp(0.02, 0.98)
        rgb = dec(content, styles[si])
        loss = ((rgb - target) ** 2).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if it % 100 == 0 or it == steps - 1:
            log(f"    step {it:4d} loss {float(loss):.5f}")
    dec.eval()
    for p in dec.parameters():
        p.requires_grad_(True)


def main():
    if not os.path.isdir(mg.REF):
        raise SystemExit(f"reference not found at {mg.REF}")
    torch.set_num_threads(8)
    log = lambda s: print(s, flush=True)
    rendering, nerf, lst = mg.import_reference()
    models, args = mg.build_reference_models(nerf, lst, seed=0)
    field = feature_field()
    t0 = time.time()
    log("training NeRF_sigma coarse (reference module, Adam, point-supervised)")
    train_nerf(models["coarse"], nerf, 500, 11, field, log)
    log("training NeRF_sigma fine")
    train_nerf(models["fine"], nerf, 700, 12, field, log)
    emb = {"xyz": nerf.PosEmbedding(14, 15), "dir": nerf.PosEmbedding(3, 4)}

    def render(rays, ns, ni, train=False, seed=1234):
        torch.manual_seed(seed)
        with torch.no_grad():
            return rendering.render_rays_cross_ray(models, emb, rays, None, ns, False, 1.0 if train else 0,
                                                   1.0 if train else 0, ni, 8192, False, test_time=not train,
                                                   args=args)

    log("rendering feature patches for the decoder")
    patches = []
    for s in range(6):
        rays = oracle.pinhole_rays(32, 32, oracle.synthetic_pose(20 + s))
        patches.append(render(rays, 48, 48)["feature_fine"].t().reshape(1, 64, 32, 32).contiguous())
    log("training style_net (all parameters except the two fc layers)")
    train_decoder(models["decoder"], patches, 400, 31, log)
    log(f"training took {time.time() - t0:.0f} s")

    pc, pf, pd = mg.sd(models["coarse"]), mg.sd(models["fine"]), mg.sd(models["decoder"])
    out = {"kind": "trained", "seed": 0,
           "coarse": pc, "fine": pf,
           "decoder": {k: v for k, v in pd.items() if not k.startswith(FROZEN)},
           "checksum_decoder_frozen": {k: (float(v.double().sum()), float(v.double().abs().sum()))
                                       for k, v in pd.items() if k.startswith(FROZEN)},
           "cases": {}}

    # ---- pin the oracle on the trained weights and record golden vectors
    def pin_render(name, rays, ns, ni, train):
        ref = render(rays, ns, ni, train)
        torch.manual_seed(1234)
        rec = {}
        with torch.no_grad():
            mine = oracle.render_rays(pc, pf, rays, n_samples=ns, n_importance=ni, perturb=1.0 if train else 0,
                                      noise_std=1.0 if train else 0, chunk=8192, record=rec)
        for k, v in mine.items():
            mg.assert_equal(f"{name}:{k}", v, ref[k])
        case = {"rays": rays, "n_samples": ns, "n_importance": ni, "train": train,
                "z_coarse": rec["z_coarse"], "z_fine": rec.get("z_fine"),
                "rng": {k: v for k, v in rec.items() if k in ("perturb_rand", "noise_coarse", "noise_fine")},
                "ref": {k: v.clone() for k, v in ref.items() if k != "feature_fine_random"}}
        if train:
            torch.manual_seed(1234)
            _ = torch.rand_like(rec["z_coarse"])
            _ = torch.randn(rays.shape[0], ns)
            case["rng"]["u"] = torch.rand(rays.shape[0], ni)
        out["cases"][name] = case
        w = ref["weights_fine"]
        log(f"  {name}: pinned; max ray weight {float(w.max()):.3f}, mean of per-ray max {float(w.max(1)[0].mean()):.3f}, "
            f"feature range [{float(ref['feature_fine'].min()):.3f}, {float(ref['feature_fine'].max()):.3f}]")

    pin_render("eval_64p128", mg.make_rays(96, 41, hw=(48, 64)), 64, 128, False)
    pin_render("train_64p64", mg.make_rays(64, 42, hw=(48, 64)), 64, 64, True)

    # MLP rows (NeRF_sigma.forward) incl. the per-layer activation magnitudes the fp16 path must hold
    g = torch.Generator().manual_seed(43)
    rays = mg.make_rays(96, 44, hw=(48, 64))
    z = torch.rand(96, generator=g) * 5.0
    x = rays[:, 0:3] + rays[:, 3:6] * z[:, None]
    with torch.no_grad():
        inp = torch.cat([emb["xyz"](x), emb["dir"](rays[:, 3:6])], 1)
        ref_out = models["fine"](inp)
        mg.assert_equal("trained:mlp", oracle.nerf_sigma_forward(pf, inp), ref_out)
        h, amax = inp[:, :93], []
        for i in range(8):
            layer = getattr(models["fine"], f"xyz_encoding_{i + 1}")
            h = layer(torch.cat([inp[:, :93], h], 1) if i == 4 else h)
            amax.append(float(h.abs().max()))
    out["cases"]["mlp"] = {"x": inp, "ref": ref_out, "act_absmax": amax}
    log(f"  mlp: pinned; per-layer |activation| max {['%.1f' % a for a in amax]}, "
        f"|w| max {max(float(v.abs().max()) for v in pf.values()):.2f}")

    # decoded 48x64 frame with two styles (the PSNR test's A and T)
    styles, _ = style_bank(8, 31)
    rays = oracle.pinhole_rays(48, 64, oracle.synthetic_pose(50))
    ref = render(rays, 64, 128)
    feat = ref["feature_fine"].t().reshape(1, 64, 48, 64)
    with torch.no_grad():
        rgb_a = models["decoder"](feat, styles[0])
        rgb_t = models["decoder"](feat, styles[3])
        mg.assert_equal("trained:style", oracle.style_net_forward(pd, feat, styles[0]), rgb_a)
        mine = oracle.render_rays(pc, pf, rays, n_samples=64, n_importance=128, perturb=0, noise_std=0, chunk=8192)
        mg.assert_equal("trained:frame", mine["feature_fine"], ref["feature_fine"])
    out["cases"]["frame"] = {"rays": rays, "hw": (48, 64), "style_a": styles[0], "style_t": styles[3],
                             "feature_fine": ref["feature_fine"].clone(), "rgb_a": rgb_a, "rgb_t": rgb_t}
    half = lambda t: t[..., 32:]
    log(f"  frame: pinned; PSNR(A, T) right half {oracle.psnr(half(rgb_a), half(rgb_t)):.2f} dB, "
        f"rgb range [{float(rgb_a.min()):.3f}, {float(rgb_a.max()):.3f}]")
    torch.save(out, OUT)
    log(f"wrote {OUT} ({os.path.getsize(OUT) / 1e6:.1f} MB)")


if __name__ == "__main__":
    main()
