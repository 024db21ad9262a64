"""Pin the oracle against the UNMODIFIED reference and write golden fixtures.

Runs only in the build container (needs ``/root/reference``).  For every case it
  1. builds the reference modules with a fixed seed (default PyTorch init, the
     construction order of ``train_mask_grid_sample.py:38-61``: coarse, decoder, fine),
  2. runs the reference's own ``render_rays_cross_ray`` / ``sample_pdf`` /
     ``PosEmbedding`` / ``NeRF_sigma`` / ``style_net`` on CPU,
  3. runs ``oracle/crnerf_oracle.py`` on the same inputs and asserts the outputs
     are bit-identical (``torch.equal``) - that is the pin,
  4. saves inputs + reference outputs (+ the random draws of train-mode cases)
     to ``tests/golden/<case>.pt``.

Weights are NOT stored (2 x 2.5 MB): they are regenerated from the seed by the
product's own module mirror, and each fixture carries a checksum of every
parameter so a drift in init order or RNG is caught.

    python oracle/make_golden.py            # writes tests/golden/*.pt
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CRNERF_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, HERE)
import crnerf_oracle as oracle  # noqa: E402


def import_reference():
    """Import the reference's model files unmodified.  ``kornia`` is not
    installed; ``nerf_decoder_stylenerf.py:104`` only needs the *name*
    ``kornia.filters.filter2d`` at import time (``Blur`` is never constructed at
    n_blocks == 0), so an empty stub module satisfies it."""
    if "kornia" not in sys.modules:
        k = types.ModuleType("kornia")
        kf = types.ModuleType("kornia.filters")
        kf.filter2d = lambda *a, **k_: (_ for _ in ()).throw(RuntimeError("kornia stub"))
        k.filters = kf
        sys.modules["kornia"] = k
        sys.modules["kornia.filters"] = kf
    sys.path.insert(0, REF)
    import importlib
    rendering = importlib.import_module("models.rendering")
    nerf = importlib.import_module("models.nerf")
    lst = importlib.import_module("models.linearStyleTransfer")
    sys.path.remove(REF)
    return rendering, nerf, lst


def ref_args():
    return types.SimpleNamespace(nerf_out_dim=64, pertubeCord=False, img_wh=[32, 32],
                                 N_emb_xyz=15, N_emb_dir=4)


def build_reference_models(nerf, lst, seed=0, peaky=False):
    torch.manual_seed(seed)
    args = ref_args()
    coarse = nerf.NeRF_sigma("coarse", args, in_channels_xyz=93, in_channels_dir=27)
    decoder = lst.style_net(args)
    fine = nerf.NeRF_sigma("fine", args, in_channels_xyz=93, in_channels_dir=27,
                           encode_appearance=True, in_channels_a=48, encode_random=True)
    if peaky:
        sharpen(coarse, seed + 100)
        sharpen(fine, seed + 101)
    return {"coarse": coarse.eval(), "fine": fine.eval(), "decoder": decoder.eval()}, args


def sharpen(model, seed):
    """'Peaky' variant: scale the sigma head so ray weights
    concentrate and sample_pdf's search / empty-bin branches are exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        w = model.static_sigma[0].weight
        w.mul_(30.0)
        w.add_(0.05 * torch.randn(w.shape, generator=g))
        model.static_sigma[0].bias.sub_(2.0)


def checksums(module):
    return {k: (float(v.double().sum()), float(v.double().abs().sum()))
            for k, v in module.state_dict().items()}


def sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def make_rays(n, seed, hw=(24, 32), near_far=(0.0, 5.0), per_ray_nf=False):
    h, w = hw
    rays = oracle.pinhole_rays(h, w, oracle.synthetic_pose(seed), *near_far)
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(rays.shape[0], generator=g)[:n].sort()[0]
    rays = rays[idx].clone()
    if per_ray_nf:
        rays[:, 6] = 0.05 + 0.45 * torch.rand(n, generator=g)
        rays[:, 7] = 2.0 + 3.0 * torch.rand(n, generator=g)
    return rays.contiguous()


def assert_equal(name, a, b):
    if not torch.equal(a, b):
        d = (a.double() - b.double()).abs().max().item()
        raise SystemExit(f"ORACLE != REFERENCE for {name}: max abs diff {d:g}")


def case_render(rendering, nerf, lst, name, n_rays, ns, ni, train, peaky=False, seed=0,
                per_ray_nf=False, use_disp=False):
    models, args = build_reference_models(nerf, lst, seed=seed, peaky=peaky)
    emb = {"xyz": nerf.PosEmbedding(14, 15), "dir": nerf.PosEmbedding(3, 4)}
    rays = make_rays(n_rays, seed + 7, per_ray_nf=per_ray_nf or use_disp)
    perturb, noise_std = (1.0, 1.0) if train else (0, 0)
    with torch.no_grad():
        torch.manual_seed(1234)
        ref = rendering.render_rays_cross_ray(models, emb, rays, None, ns, use_disp, perturb,
                                              noise_std, ni, 8192, False, test_time=not train,
                                              args=args)
        torch.manual_seed(1234)
        rec = {}
        mine = oracle.render_rays(sd(models["coarse"]), sd(models["fine"]), rays, n_samples=ns,
                                  n_importance=ni, use_disp=use_disp, perturb=perturb,
                                  noise_std=noise_std, chunk=8192, record=rec)
    for k, v in mine.items():
        assert_equal(f"{name}:{k}", v, ref[k])
    if ni > 0:
        assert ref["feature_fine_random"] is ref["feature_fine"]
    out = {
        "kind": "render", "seed": seed, "peaky": peaky, "n_samples": ns, "n_importance": ni,
        "perturb": perturb, "noise_std": noise_std, "use_disp": use_disp, "rays": rays,
        "rng": {k: v for k, v in rec.items() if k in ("perturb_rand", "noise_coarse", "noise_fine")},
        "z_coarse": rec["z_coarse"], "z_fine": rec.get("z_fine"),
        "ref": {k: v.clone() for k, v in ref.items() if k != "feature_fine_random"},
        "checksum_coarse": checksums(models["coarse"]), "checksum_fine": checksums(models["fine"]),
    }
    if train and ni > 0:
        # the uniform draws of sample_pdf: replay the generator to capture them
        torch.manual_seed(1234)
        z0 = torch.rand_like(rec["z_coarse"])
        _ = torch.randn(n_rays, ns)
        out["rng"]["u"] = torch.rand(n_rays, ni)
        assert torch.equal(perturb * z0, rec["perturb_rand"])
        with torch.no_grad():
            again = oracle.render_rays(sd(models["coarse"]), sd(models["fine"]), rays, n_samples=ns,
                                       n_importance=ni, perturb=perturb, noise_std=noise_std,
                                       chunk=8192, rng=out["rng"])
        for k, v in again.items():
            assert_equal(f"{name}:replay:{k}", v, ref[k])
    torch.save(out, os.path.join(GOLD, f"{name}.pt"))
    print(f"  {name}: pinned ({', '.join(ref.keys())})")


def case_posenc_mlp(nerf, lst):
    models, args = build_reference_models(nerf, lst, seed=0)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(96, 3, generator=g) - 0.5) * 8.0
    d = torch.nn.functional.normalize(torch.randn(96, 3, generator=g), dim=-1)
    e_x, e_d = nerf.PosEmbedding(14, 15), nerf.PosEmbedding(3, 4)
    with torch.no_grad():
        ex, ed = e_x(x), e_d(d)
        assert_equal("posenc:xyz", oracle.pos_embed(x, 15), ex)
        assert_equal("posenc:dir", oracle.pos_embed(d, 4), ed)
        inp = torch.cat([ex, ed], 1)
        ref_out = models["fine"](inp)
        ref_sig = models["fine"](ex, sigma_only=True)
        assert_equal("mlp", oracle.nerf_sigma_forward(sd(models["fine"]), inp), ref_out)
        assert_equal("mlp:sigma_only",
                     oracle.nerf_sigma_forward(sd(models["fine"]), ex, sigma_only=True), ref_sig)
    torch.save({"kind": "posenc_mlp", "seed": 0, "xyz": x, "dir": d, "emb_xyz": ex, "emb_dir": ed,
                "mlp_out": ref_out, "sigma_only": ref_sig,
                "checksum_fine": checksums(models["fine"])},
               os.path.join(GOLD, "posenc_mlp.pt"))
    print("  posenc_mlp: pinned")


def case_sample_pdf(rendering):
    g = torch.Generator().manual_seed(11)
    n, m = 48, 62
    bins = torch.sort(torch.rand(n, m + 1, generator=g) * 5.0, dim=-1)[0]
    w = torch.rand(n, m, generator=g) ** 4
    w[:8] = 0.0                       # all-empty rays: pdf is uniform through eps
    w[8:16, 10:40] = 0.0              # empty interior bins -> denom < eps branch
    w[16:24] = 0.0
    w[16:24, 30] = 1.0                # a single spike
    out = {"kind": "sample_pdf", "bins": bins, "weights": w, "cases": []}
    for ni, det in ((128, True), (64, True), (33, True), (128, False), (64, False)):
        torch.manual_seed(99)
        ref = rendering.sample_pdf(bins, w, ni, det=det)
        torch.manual_seed(99)
        mine = oracle.sample_pdf(bins, w, ni, det=det)
        assert_equal(f"sample_pdf:{ni}:{det}", mine, ref)
        u = None
        if not det:
            torch.manual_seed(99)
            u = torch.rand(n, ni)
            assert_equal("sample_pdf:u", oracle.sample_pdf(bins, w, ni, det=False, u=u), ref)
        out["cases"].append({"n_importance": ni, "det": det, "u": u, "ref": ref})
    torch.save(out, os.path.join(GOLD, "sample_pdf.pt"))
    print("  sample_pdf: pinned (5 cases)")


def case_style(nerf, lst):
    models, args = build_reference_models(nerf, lst, seed=0)
    dec = models["decoder"]
    g = torch.Generator().manual_seed(21)
    out = {"kind": "style", "seed": 0, "cases": [], "checksum_decoder": checksums(dec)}
    for (h, w) in ((32, 32), (24, 40), (7, 9)):
        feat = torch.rand(h * w, 64, generator=g) * 0.3 + 0.35      # (N,64) as the renderer emits
        content = feat.t().reshape(1, 64, h, w)                      # caller's rearrange: a view
        style = torch.rand(1, 64, 32, 32, generator=g)
        with torch.no_grad():
            ref = dec(content, style)
            ref_c = dec(content, None, type="content")
            fused, trans = dec.multi_net(content, style)
            p = sd(dec)
            assert_equal("style:fused", oracle.style_net_forward(p, content, style), ref)
            assert_equal("style:content", oracle.style_net_forward(p, content, None, type="content"),
                         ref_c)
            f2, t2 = oracle.mul_layer_forward(p, content, style)
            assert_equal("style:mullayer", f2, fused)
            assert_equal("style:trans", t2, trans)
        out["cases"].append({"h": h, "w": w, "feature": feat, "style": style, "rgb": ref,
                             "rgb_content": ref_c, "fused": fused, "trans": trans})
    torch.save(out, os.path.join(GOLD, "style.pt"))
    print("  style_net: pinned (3 sizes)")


def case_loss():
    """CRNeRFLoss (reference losses.py) and the mask lookup (train_mask_grid_sample.py:171-175,
    restated inline from torch calls the script makes) on seeded inputs."""
    sys.path.insert(0, REF)
    import importlib
    ref_losses = importlib.import_module("losses")
    sys.path.remove(REF)
    out = {"kind": "loss", "cases": []}
    for ci, (n, with_mask, with_fine, with_a, with_c, mse_a, step) in enumerate([
            (1024, True, True, True, True, False, 0), (1024, True, True, True, False, True, 1500),
            (777, False, True, False, False, False, 10), (1024, True, False, True, False, False, 3),
            (4096, True, True, False, True, False, 200000)]):
        g = torch.Generator().manual_seed(100 + ci)
        hp = types.SimpleNamespace(maskrs_max=5e-2, maskrs_min=6e-3, maskrs_k=1e-3, maskrd=0.0 if ci != 1 else 1e-3,
                                   weightKL=1e-5, weightRecA=1e-3, weightcontent=1e-4, mse_on_appearance=mse_a)
        inputs = {"rgb_coarse": torch.rand(n, 3, generator=g)}
        targets = torch.rand(n, 3, generator=g)
        if with_fine:
            inputs["rgb_fine"] = torch.rand(n, 3, generator=g)
        if with_mask:
            inputs["out_mask"] = torch.rand(n, 1, generator=g)
        if with_a:
            inputs["a_embedded"] = torch.randn(1, 64, 8, 8, generator=g)
            inputs["a_embedded_random"] = torch.randn(1, 64, 8, 8, generator=g)
            inputs["a_embedded_random_rec"] = torch.randn(1, 64, 8, 8, generator=g)
        if with_c:
            inputs["content_wo_a_embed"] = torch.randn(1, 64, 8, 8, generator=g)
            inputs["content_with_a_embed"] = torch.randn(1, 64, 8, 8, generator=g)
        leaf = {k: v.clone().requires_grad_(True) for k, v in inputs.items()}
        crit = ref_losses.loss_dict["crnerf"](hp, coef=1)
        ref, w = crit(leaf, targets, hp, step)
        total = sum(l for l in ref.values())
        total.backward()
        mine, w2 = oracle.crnerf_loss(inputs, targets, hp, step, coef=1)
        assert list(mine) == list(ref) and w == w2
        for k in ref:
            assert_equal(f"loss:{k}", mine[k], ref[k].detach())
        out["cases"].append({"inputs": inputs, "targets": targets, "hp": vars(hp), "step": step,
                             "ref": {k: v.detach() for k, v in ref.items()}, "weight": w,
                             "grads": {k: v.grad for k, v in leaf.items() if v.grad is not None}})
    # mask lookup: the three torch calls of train_mask_grid_sample.py:172-175
    from einops import rearrange
    out["mask"] = []
    for ci, (c, h, w, H, W, n) in enumerate([(1, 24, 32, 340, 512, 1024), (1, 33, 17, 64, 48, None),
                                             (2, 16, 16, 16, 16, 100), (1, 50, 70, 37, 45, 500)]):
        g = torch.Generator().manual_seed(200 + ci)
        pred = torch.rand(1, c, h, w, generator=g)
        idx = None if n is None else torch.randint(0, H * W, (n,), generator=g)
        up = torch.nn.functional.interpolate(pred, size=(H, W), mode='bilinear', align_corners=False)
        rows = rearrange(up, '1 n h w -> (h w) n')
        ref = rows if idx is None else rows[idx]
        assert_equal("mask_sample", oracle.mask_sample(pred, (H, W), idx), ref)
        out["mask"].append({"pred": pred, "hw": (H, W), "idx": idx, "ref": ref.clone()})
    torch.save(out, os.path.join(GOLD, "loss.pt"))
    print("  CRNeRFLoss + mask lookup: pinned (5 + 4 cases)")


def case_encoder(lst):
    """encoder_sameoutputsize (reference linearStyleTransfer.py:208-276), default init seed 11."""
    torch.manual_seed(11)
    enc = lst.encoder_sameoutputsize(out_channel=64).eval()
    p = sd(enc)
    out = {"kind": "encoder", "seed": 11, "checksum": checksums(enc), "cases": []}
    for i, (h, w) in enumerate([(32, 32), (40, 52), (97, 131)]):
        x = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(300 + i))
        with torch.no_grad():
            ref = enc(x)
            assert_equal(f"encoder:{h}x{w}", oracle.encoder_forward(p, x), ref)
        out["cases"].append({"x": x, "ref": ref})
    torch.save(out, os.path.join(GOLD, "encoder.pt"))
    print("  encoder_sameoutputsize: pinned (3 sizes)")


def case_grid_patch():
    """The training ``__getitem__`` of datasets/phototourism_mask_grid_sample.py:240-275.  The
    dataset module itself cannot be imported here (kornia, torchvision, COLMAP readers), so the
    method's source is cut out of the UNMODIFIED file with ``ast`` and executed as it stands against
    a stand-in ``self`` that carries the attributes it reads."""
    import ast
    import math
    import numpy as np
    path = os.path.join(REF, "datasets", "phototourism_mask_grid_sample.py")
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PhototourismDataset")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__getitem__")
    mod = ast.Module(body=[fn], type_ignores=[])
    gv = types.SimpleNamespace(current_epoch=0)
    ns = {"torch": torch, "np": np, "sqrt": math.sqrt, "exp": math.exp, "global_val": gv}
    exec(compile(mod, path, "exec"), ns)
    getitem = ns["__getitem__"]

    g = torch.Generator().manual_seed(77)
    wh = torch.Tensor([[40, 30], [37, 53], [64, 48], [51, 29]])            # (w, h) per image, fp32 as :197
    n_rows = int((wh[:, 0] * wh[:, 1]).sum())
    all_rays = torch.randn(n_rows, 9, generator=g)
    all_rays[:, 8] = torch.repeat_interleave(torch.arange(4, dtype=torch.float32) + 10, (wh[:, 0] * wh[:, 1]).long())
    all_rgbs = torch.rand(n_rows, 3, generator=g)
    imgs = [torch.rand(3, int(h), int(w), generator=g) for w, h in wh]
    out = {"kind": "grid_patch", "all_rays": all_rays, "all_rgbs": all_rgbs, "all_imgs_wh": wh, "cases": []}
    for ci, (batch, anneal, min_scale, epoch, idx) in enumerate([(1024, -1, 0.25, 0, 0), (1024, -1, 0.5, 3, 17),
                                                                 (256, 0.0025, 0.25, 1, 5), (64, -1, 0.9, 2, 2),
                                                                 (1024, -1, 0.25, 7, 123)]):
        me = types.SimpleNamespace(split="train", all_imgs=imgs, all_imgs_wh=wh, batch_size=batch,
                                   scale_anneal=anneal, min_scale=min_scale, all_rays=all_rays,
                                   all_rgbs=all_rgbs, iterations=n_rows // batch)
        gv.current_epoch = epoch
        torch.manual_seed(500 + ci)
        ref = getitem(me, idx)
        # replay the host draws (same calls, same order) and pin the oracle's arithmetic
        torch.manual_seed(500 + ci)
        step = epoch * me.iterations + idx
        np.random.seed(step)
        ts = np.random.randint(0, len(imgs))
        img_w, img_h = wh[ts]
        msc = min(max(min_scale, 1. * math.exp(-step * anneal)), 0.9) if anneal > 0 else min_scale
        scale = torch.Tensor(1).uniform_(msc, 1.)
        h_off = torch.Tensor(1).uniform_(0, (1 - scale.item()) * (1 - 1 / img_h))
        w_off = torch.Tensor(1).uniform_(0, (1 - scale.item()) * (1 - 1 / img_w))
        mine = oracle.grid_patch(all_rays, all_rgbs, wh, ts, batch, scale, h_off, w_off)
        for k in mine:
            assert_equal(f"grid_patch:{k}", mine[k], ref[k])
        assert ref["min_scale_cur"] == msc and torch.equal(ref["img_wh"], wh[ts]) and ref["whole_img"] is imgs[ts]
        out["cases"].append({"batch_size": batch, "scale_anneal": anneal, "min_scale": min_scale, "epoch": epoch,
                             "idx": idx, "torch_seed": 500 + ci, "sample_ts": int(ts), "scale": scale,
                             "h_offset": h_off, "w_offset": w_off, "min_scale_cur": msc,
                             "ref": {k: ref[k].clone() for k in mine}})
    torch.save(out, os.path.join(GOLD, "grid_patch.pt"))
    print("  grid-sampled training patch: pinned (5 cases)")


def ruffle_cgnet(net, g):
    """Make BatchNorm / PReLU parameters and running statistics non-trivial, as after some training
    (same function in tests/conftest.py)."""
    with torch.no_grad():
        for n_, prm in net.named_parameters():
            if ".bn" in n_ or "b1." in n_ or ".act" in n_ or "bn_prelu" in n_:
                prm.add_(0.1 * torch.randn(prm.shape, generator=g))
        for n_, buf in net.named_buffers():
            if n_.endswith("running_mean"):
                buf.copy_(0.2 * torch.randn(buf.shape, generator=g))
            elif n_.endswith("running_var"):
                buf.copy_(0.5 + torch.rand(buf.shape, generator=g))


def case_cgnet():
    """Context_Guided_Network as the training script builds it (train_mask_grid_sample.py:114):
    state_dict keys / shapes, seeded init, eval- and train-mode forwards, and the gradients of the
    caller's tail (resize -> rows -> [rgb_idx], :171-175).  lightweight_seg.py imports torch only."""
    import importlib.util
    from einops import rearrange
    spec = importlib.util.spec_from_file_location("_ref_lightweight_seg", os.path.join(REF, "models", "lightweight_seg.py"))
    ref_mod = importlib.util.module_from_spec(spec)
    import warnings
    spec.loader.exec_module(ref_mod)
    torch.manual_seed(21)
    net = ref_mod.Context_Guided_Network(classes=1, M=2, N=2, input_channel=3)
    after_init = torch.rand(1)                      # the generator position after construction
    g = torch.Generator().manual_seed(22)
    ruffle_cgnet(net, g)
    # weights are not stored (1 MB): the product's mirror regenerates them from the seed + ruffle;
    # key order, shapes and a checksum of every entry pin that
    out = {"kind": "cgnet", "seed": 21, "after_init": after_init, "checksum": checksums(net),
           "shapes": {k: tuple(v.shape) for k, v in net.state_dict().items()}, "cases": []}
    for ci, (h, w, H, W, n) in enumerate([(48, 64, 48, 64, 256), (70, 52, 35, 26, 100), (33, 47, 66, 94, None)]):
        x = torch.rand(1, 3, h, w, generator=g)
        idx = None if n is None else torch.randint(0, H * W, (n,), generator=g)
        rec = {"x": x, "hw": (H, W), "idx": idx}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            net.eval()
            with torch.no_grad():
                rec["eval"] = net(x)
                assert_equal("cgnet:eval", oracle.cgnet_forward(sd(net), x, 2, 2, train=False), rec["eval"])
            net.train()
            state_before = {k: v.clone() for k, v in net.state_dict().items()}
            net.zero_grad()
            y = net(x)
            assert_equal("cgnet:train", oracle.cgnet_forward(state_before, x, 2, 2, train=True), y.detach())
            rows = rearrange(torch.nn.functional.interpolate(y, size=(H, W), mode='bilinear', align_corners=False),
                             '1 n h w -> (h w) n')
            rows = rows if idx is None else rows[idx]
            gy = torch.rand(rows.shape, generator=g)
            (rows * gy).sum().backward()
        grads = {k: v.grad.clone() for k, v in net.named_parameters()}
        rec.update({"train": y.detach().clone(), "rows": rows.detach().clone(), "g_rows": gy,
                    "grads": grads if ci == 0 else None,
                    "grad_sums": {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in grads.items()},
                    "running_after": {k: v.clone() for k, v in net.named_buffers() if "running" in k}})
        net.load_state_dict(state_before)          # every case starts from the same running statistics
        out["cases"].append(rec)
    torch.save(out, os.path.join(GOLD, "cgnet.pt"))
    print("  Context_Guided_Network: pinned (3 sizes, eval + train + gradients)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", choices=["loss", "encoder", "grid_patch", "cgnet"], help="regenerate a single fixture")
    opt = ap.parse_args()
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}; golden vectors can only be made in the "
                         "build container")
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    print("pinning oracle against", REF)
    if opt.only == "loss":
        case_loss()
        return
    if opt.only == "grid_patch":
        case_grid_patch()
        return
    if opt.only == "cgnet":
        case_cgnet()
        return
    rendering, nerf, lst = import_reference()
    if opt.only == "encoder":
        case_encoder(lst)
        return
    case_encoder(lst)
    case_posenc_mlp(nerf, lst)
    case_sample_pdf(rendering)
    case_style(nerf, lst)
    case_loss()
    case_grid_patch()
    case_cgnet()
    # config[0]-shaped (coarse only), eval and train mode, fine pass, peaky weights
    case_render(rendering, nerf, lst, "render_c64_eval", 64, 64, 0, train=False)
    case_render(rendering, nerf, lst, "render_64p128_eval", 96, 64, 128, train=False)
    case_render(rendering, nerf, lst, "render_64p128_eval_peaky", 96, 64, 128, train=False,
                peaky=True, per_ray_nf=True)
    case_render(rendering, nerf, lst, "render_64p64_train", 64, 64, 64, train=True, seed=3)
    case_render(rendering, nerf, lst, "render_32p24_train_peaky", 40, 32, 24, train=True,
                peaky=True, seed=4, per_ray_nf=True)
    case_render(rendering, nerf, lst, "render_48p48_disp", 32, 48, 48, train=False, seed=5,
                use_disp=True)
    print("done ->", GOLD)


if __name__ == "__main__":
    main()
