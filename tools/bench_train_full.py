#!/usr/bin/env python
"""The reference's WHOLE training step at realistic sizes, on the mirror: grid-sampled 32x32 patch
from a GPU-resident ray cache -> enc_a(photo) -> mask network -> render (64+64, perturb = noise = 1)
-> decode coarse / fine / fine_random -> enc_a(rgb_fine_random) -> CRNeRFLoss -> backward -> Adam
(train_mask_grid_sample.py:150-226, 268-337 with encode_a, encode_random, use_mask; encode_c off as
in command/train.sh).  Prints ms/step and where the time goes (CUDA events around the phases).

  python tools/bench_train_full.py [--photo 340 512] [--steps 10]"""
import argparse, json, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from einops import rearrange


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--photo", type=int, nargs=2, default=[340, 512])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--graph", action="store_true", help="also time the step captured as ONE CUDA graph (GraphedTrainStep)")
    ap.add_argument("--profile", action="store_true", help="print the top CUDA kernels of one step (torch profiler)")
    ap.add_argument("--conv", default="fp32", choices=["fp32", "tf32"], help="precision of the library convolutions")
    ap.add_argument("--encoder", default="native", choices=["native", "library"],
                    help="enc_a under autograd: csrc/encoder.cu + encoder_train.cuh, or library convolutions")
    a = ap.parse_args()
    import _training_step_driver as drv
    from models.nerf import NeRF_sigma, PosEmbedding
    from models.linearStyleTransfer import style_net, encoder_sameoutputsize
    from models.lightweight_seg import Context_Guided_Network
    from models.rendering import render_rays_cross_ray
    from losses import loss_dict
    from crnerf_b200.sampling import GridPatchSampler
    from crnerf_b200 import loss as crloss
    dev = torch.device("cuda", 0)
    hp = drv.hparams()
    hp.N_samples, hp.N_importance, hp.perturb, hp.noise_std, hp.batch_size, hp.img_wh = 64, 64, 1.0, 1.0, 1024, [32, 32]
    torch.manual_seed(0)
    enc_a = encoder_sameoutputsize(64)
    coarse = NeRF_sigma('coarse', hp, in_channels_xyz=93, in_channels_dir=27)
    decoder = style_net(hp)
    fine = NeRF_sigma('fine', hp, in_channels_xyz=93, in_channels_dir=27, encode_appearance=True, in_channels_a=48,
                      encode_random=True)
    mask_net = Context_Guided_Network(classes=1, M=2, N=2, input_channel=3)
    mods = [enc_a, coarse, decoder, fine, mask_net]
    enc_a.conv_precision = mask_net.conv_precision = a.conv
    enc_a.train_backend = a.encoder
    for m in mods:
        m.to(dev).train()
    models = {"coarse": coarse, "decoder": decoder, "fine": fine}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    H, W = a.photo
    n_img = 4
    g = torch.Generator().manual_seed(1)
    rows = n_img * H * W
    rays = torch.zeros(rows, 9)
    rays[:, 2] = 4.0
    d = torch.randn(rows, 3, generator=g) * 0.25 + torch.tensor([0.0, 0.0, -1.0])
    rays[:, 3:6] = d / d.norm(dim=1, keepdim=True)
    rays[:, 6], rays[:, 7] = 0.5, 5.0
    rays[:, 8] = torch.repeat_interleave(torch.arange(n_img, dtype=torch.float32), H * W)
    imgs = [(torch.rand(3, H, W, generator=g) * 2 - 1).to(dev) for _ in range(n_img)]
    sampler = GridPatchSampler(rays, torch.rand(rows, 3, generator=g), torch.Tensor([[W, H]] * n_img), imgs,
                               batch_size=1024, device=dev)
    crit = loss_dict["crnerf"](hp, coef=1)
    params = [p for m in mods for p in m.parameters()]
    from crnerf_b200.optim import Adam   # torch.optim.Adam's update as one launch per 48 tensors (csrc/optim.cu)
    opt = Adam(params, lr=5e-4)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    acc = {}

    def step(i, timed):
        marks = [("start", ev())]
        marks[0][1].record()
        def mark(name):
            e = ev(); e.record(); marks.append((name, e))
        s = sampler.sample(0, i)
        mark("sample")
        whole = (s["whole_img"].unsqueeze(0) + 1) / 2
        a_emb = enc_a(whole)
        mark("enc_a(photo)")
        pred_mask = mask_net.mask_rows(whole, (H, W), s["rgb_idx"])
        mark("mask network")
        res = render_rays_cross_ray(models, emb, s["rays"], s["ts"], 64, False, 1.0, 1.0, 64, 32768, False, args=hp)
        mark("render fwd")
        out = dict(res)
        for typ, key in (("coarse", "feature_coarse"), ("fine", "feature_fine"), ("fine_random", "feature_fine")):
            feat = rearrange(res[key], '(h w) c -> 1 c h w', h=32, w=32)
            img = decoder(feat, a_emb)
            out[f"rgb_{typ}"] = rearrange(img, '1 c h w -> (h w) c')
            if typ == "fine_random":
                out["a_embedded_random_rec"] = enc_a(img)
        out["out_mask"], out["a_embedded"], out["a_embedded_random"] = pred_mask, a_emb, a_emb
        mark("decode x3 + enc_a(patch)")
        ld, _ = crit(out, s["rgbs"], hp, 10 + i)
        loss = sum(ld.values())
        mark("loss")
        opt.zero_grad(set_to_none=True)
        loss.backward()
        mark("backward")
        opt.step()
        mark("adam")
        if timed:
            torch.cuda.synchronize()
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
        return loss

    import models.lightweight_seg as seg
    for _ in range(3):
        step(0, False)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(a.steps):
        step(i, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    for i in range(a.steps):
        step(i, True)
    graph_ms = None
    if a.graph:
        # everything after the (host-driven) patch sampler as one CUDA graph: the sampled batch is copied into
        # static tensors, the graph holds enc_a, mask network, render, decode x3, loss, backward and Adam
        from crnerf_b200.graphs import GraphedTrainStep
        opt_g = Adam(params, lr=5e-4)
        s0 = sampler.sample(0, 0)
        static = {k: s0[k].clone() for k in ("rays", "ts", "rgbs", "rgb_idx")}
        static["whole"] = ((s0["whole_img"].unsqueeze(0) + 1) / 2).clone()

        def step_fn():
            a_emb = enc_a(static["whole"])
            pred_mask = mask_net.mask_rows(static["whole"], (H, W), static["rgb_idx"])
            res = render_rays_cross_ray(models, emb, static["rays"], static["ts"], 64, False, 1.0, 1.0, 64, 32768, False,
                                        args=hp)
            out = dict(res)
            for typ, key in (("coarse", "feature_coarse"), ("fine", "feature_fine"), ("fine_random", "feature_fine")):
                img = decoder(rearrange(res[key], '(h w) c -> 1 c h w', h=32, w=32), a_emb)
                out[f"rgb_{typ}"] = rearrange(img, '1 c h w -> (h w) c')
                if typ == "fine_random":
                    out["a_embedded_random_rec"] = enc_a(img)
            out["out_mask"], out["a_embedded"], out["a_embedded_random"] = pred_mask, a_emb, a_emb
            ld, _ = crit(out, static["rgbs"], hp, 10)
            return sum(ld.values())

        gs = GraphedTrainStep(step_fn, opt_g, warmup=3)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for i in range(a.steps):
            s = sampler.sample(0, i)
            for k in ("rays", "ts", "rgbs", "rgb_idx"):
                static[k].copy_(s[k])
            static["whole"].copy_((s["whole_img"].unsqueeze(0) + 1) / 2)
            gs()
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / a.steps
    if a.profile:
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            step(0, False)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:22]
        for e in rows:
            print(f"{e.device_time_total / 1e3:8.3f} ms  x{e.count:<4d} {e.key[:110]}", file=sys.stderr)
    print(json.dumps({"workload": f"whole training step, photo {H}x{W}, 1024-ray patch x (64+64), perturb=noise=1, "
                                  "enc_a + mask network + render + decode x3 + CRNeRFLoss + Adam",
                      "library_conv_precision": a.conv, "encoder": a.encoder, "ms_per_step": ms, "ms_per_step_graphed": graph_ms, "phases_ms": {k: v / a.steps for k, v in acc.items()}}))


if __name__ == "__main__":
    main()
