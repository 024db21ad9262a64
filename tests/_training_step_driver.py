"""One whole training step of the reference's ``NeRFSystem`` - ``__getitem__`` (the grid-sampled
batch), ``forward``, ``decode``, ``training_step`` - executed VERBATIM from the reference's sources
(cut out with ``ast``: the script itself cannot be imported without pytorch_lightning / kornia /
wandb) against a stand-in ``self`` whose models come from whichever ``models`` / ``losses`` packages
are first on ``sys.path``:

  * impl = "reference": the reference's own files, CPU (run as a subprocess with the extracted
    reference directory first on PYTHONPATH) - the yardstick;
  * impl = "mirror": this repo's mirror + ``GridPatchSampler`` on the GPU.

Returns the loss dict, PSNR, the rendered patch and the gradients of a few representative
parameters.  Used by tests/test_training_step_replay.py."""
import ast
import os
import random
import sys
import textwrap
import types
from collections import defaultdict
from math import exp, sqrt

import numpy as np
import torch
from torch import nn


def _cut(src, name, cls):
    body = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    lines = src.splitlines()
    first = min([fn.lineno] + [d.lineno for d in fn.decorator_list])
    return textwrap.dedent("\n".join(lines[first - 1:fn.end_lineno]))


def hparams():
    return types.SimpleNamespace(
        N_emb_xyz=15, N_emb_dir=4, nerf_out_dim=64, N_a=48, N_vocab=8, img_wh=[16, 16], pertubeCord=False,
        encode_a=True, encode_random=True, encode_c=False, use_mask=True, decoder_num_res_blocks=2,
        N_samples=16, N_importance=16, use_disp=False, perturb=0, noise_std=0, chunk=32768, batch_size=256,
        maskrs_max=5e-2, maskrs_min=6e-3, maskrs_k=1e-3, maskrd=1e-3, weightKL=1e-5, weightRecA=1e-3,
        weightcontent=1e-4, mse_on_appearance=False)


def synthetic_scene(seed=3):
    """Ray cache of three small 'photos' (the reference's all_rays / all_rgbs / all_imgs / all_imgs_wh)."""
    g = torch.Generator().manual_seed(seed)
    wh = torch.Tensor([[64, 48], [56, 40], [48, 64]])
    rows = int((wh[:, 0] * wh[:, 1]).sum())
    rays = torch.zeros(rows, 9)
    rays[:, 0:3] = torch.tensor([0.0, 0.0, 4.0]) + 0.05 * torch.randn(rows, 3, generator=g)
    d = torch.randn(rows, 3, generator=g) * 0.25 + torch.tensor([0.0, 0.0, -1.0])
    rays[:, 3:6] = d / d.norm(dim=1, keepdim=True)
    rays[:, 6] = 0.5
    rays[:, 7] = 5.0
    rays[:, 8] = torch.repeat_interleave(torch.arange(3, dtype=torch.float32), (wh[:, 0] * wh[:, 1]).long())
    rgbs = torch.rand(rows, 3, generator=g)
    imgs = [torch.rand(3, int(h), int(w), generator=g) * 2 - 1 for w, h in wh]      # whole_img in [-1, 1]
    return rays, rgbs, imgs, wh


def run(impl, ref_dir, device):
    hp = hparams()
    src_train = open(os.path.join(ref_dir, "train_mask_grid_sample.py"), encoding="utf-8").read()
    src_data = open(os.path.join(ref_dir, "datasets", "phototourism_mask_grid_sample.py"), encoding="utf-8").read()
    from einops import rearrange
    from models.nerf import NeRF_sigma, PosEmbedding
    from models.linearStyleTransfer import style_net, encoder_sameoutputsize
    from models.lightweight_seg import Context_Guided_Network
    from models.rendering import render_rays_cross_ray
    from losses import loss_dict
    dev = torch.device(device)

    # ---- the modules, in NeRFSystem.__init__'s order (train_mask_grid_sample.py:76-116)
    torch.manual_seed(0)
    enc_a = encoder_sameoutputsize(out_channel=hp.nerf_out_dim)
    coarse = NeRF_sigma(typ='coarse', args=hp, in_channels_xyz=6 * hp.N_emb_xyz + 3, in_channels_dir=6 * hp.N_emb_dir + 3)
    decoder = style_net(args=hp, residual_blocks=hp.decoder_num_res_blocks)
    fine = NeRF_sigma('fine', args=hp, in_channels_xyz=6 * hp.N_emb_xyz + 3, in_channels_dir=6 * hp.N_emb_dir + 3,
                      encode_appearance=hp.encode_a, in_channels_a=hp.N_a, encode_random=hp.encode_random)
    mask_net = Context_Guided_Network(classes=1, M=2, N=2, input_channel=3)
    mods = {"enc_a": enc_a, "coarse": coarse, "decoder": decoder, "fine": fine, "implicit_mask": mask_net}
    for m in mods.values():
        m.to(dev).train()

    # ---- the batch: the reference's own __getitem__ on CPU, the GPU sampler for the mirror
    all_rays, all_rgbs, imgs, wh = synthetic_scene()
    epoch, idx = 1, 3
    torch.manual_seed(77)
    if impl == "reference":
        gv = types.SimpleNamespace(current_epoch=epoch)
        ns = {"torch": torch, "np": np, "sqrt": sqrt, "exp": exp, "global_val": gv}
        exec(compile(_cut(src_data, "__getitem__", "PhototourismDataset"), "dataset", "exec"), ns)
        ds = types.SimpleNamespace(split="train", all_imgs=imgs, all_imgs_wh=wh, batch_size=hp.batch_size, scale_anneal=-1,
                                   min_scale=0.25, all_rays=all_rays, all_rgbs=all_rgbs,
                                   iterations=len(all_rays) // hp.batch_size)
        sample = ns["__getitem__"](ds, idx)
    else:
        from crnerf_b200.sampling import GridPatchSampler
        sampler = GridPatchSampler(all_rays, all_rgbs, wh, imgs, batch_size=hp.batch_size, scale_anneal=-1,
                                   min_scale=0.25, device=dev)
        sample = sampler.sample(epoch, idx)
        sampler.check()
    # what DataLoader(batch_size=1) + Lightning's device transfer hand to training_step
    batch = {k: (v.unsqueeze(0).to(dev) if torch.is_tensor(v) else torch.tensor([v])) for k, v in sample.items()}

    # ---- NeRFSystem's methods, verbatim
    ns = {"torch": torch, "nn": nn, "rearrange": rearrange, "defaultdict": defaultdict, "random": random, "sqrt": sqrt,
          "render_rays_cross_ray": render_rays_cross_ray, "hparams_": hp,
          "psnr": lambda a, b: -10 * torch.log10(torch.mean((a - b) ** 2)),
          "get_learning_rate": lambda opt: opt.param_groups[0]["lr"], "wandb": None}
    for name in ("forward", "decode", "training_step"):
        exec(compile(_cut(src_train, name, "NeRFSystem"), "train_mask_grid_sample.py", "exec"), ns)

    class System:
        forward = ns["forward"]
        decode = ns["decode"]
        training_step = ns["training_step"]
        __call__ = ns["forward"]

        def log(self, *a, **k):
            self.logged[a[0]] = a[1]

    sys_ = System()
    sys_.hparams_ = hp
    sys_.models = {"coarse": coarse, "decoder": decoder, "fine": fine}
    sys_.embeddings = {"xyz": PosEmbedding(hp.N_emb_xyz - 1, hp.N_emb_xyz), "dir": PosEmbedding(hp.N_emb_dir - 1, hp.N_emb_dir)}
    sys_.enc_a = enc_a
    sys_.implicit_mask = mask_net
    sys_.embedding_a_list = [None] * hp.N_vocab
    sys_.train_dataset = types.SimpleNamespace(white_back=False)
    sys_.loss = loss_dict["crnerf"](hp, coef=1)
    params = [p for m in mods.values() for p in m.parameters()]
    sys_.optimizer = torch.optim.Adam(params, lr=5e-4)
    sys_.global_step = 10
    sys_.logged = {}

    loss = sys_.training_step(batch, 0)
    sys_.optimizer.zero_grad(set_to_none=True)
    loss.backward()
    sys_.optimizer.step()
    if dev.type == "cuda":
        torch.cuda.synchronize()

    def grad(m, key):
        return dict(m.named_parameters())[key].grad.detach().float().cpu()

    out = {"loss": float(loss.detach()),
           "logged": {k: (float(v.detach()) if torch.is_tensor(v) else float(v)) for k, v in sys_.logged.items()},
           "batch": {k: v.detach().cpu() for k, v in batch.items() if torch.is_tensor(v)},
           "grads": {
               "fine.xyz_encoding_8.0.weight": grad(fine, "xyz_encoding_8.0.weight"),
               "fine.static_rgb.0.weight": grad(fine, "static_rgb.0.weight"),
               "coarse.xyz_encoding_1.0.weight": grad(coarse, "xyz_encoding_1.0.weight"),
               "decoder.multi_net.cnet.convs.0.weight": grad(decoder, "multi_net.cnet.convs.0.weight"),
               "decoder.multi_net.unzip.weight": grad(decoder, "multi_net.unzip.weight"),
               "decoder.decoder.feat_2_rgb_list.0.weight": grad(decoder, "decoder.feat_2_rgb_list.0.weight"),
               "enc_a.conv2.weight": grad(enc_a, "conv2.weight"),
               "enc_a.conv7.weight": grad(enc_a, "conv7.weight"),
               "implicit_mask.level1_0.conv.weight": grad(mask_net, "level1_0.conv.weight"),
               "implicit_mask.classifier.0.conv.weight": grad(mask_net, "classifier.0.conv.weight")},
           "embedding_a_slot": int(next(i for i, v in enumerate(sys_.embedding_a_list) if v is not None))}
    return out


if __name__ == "__main__":
    # reference mode: python _training_step_driver.py <ref_dir> <out.pt>   (PYTHONPATH = ref_dir)
    ref, out_path = sys.argv[1], sys.argv[2]
    if "kornia" not in sys.modules:     # nerf_decoder_stylenerf.py:104 needs the NAME kornia.filters.filter2d only
        k, kf = types.ModuleType("kornia"), types.ModuleType("kornia.filters")
        kf.filter2d = lambda *a, **k_: (_ for _ in ()).throw(RuntimeError("kornia stub"))
        k.filters = kf
        sys.modules["kornia"], sys.modules["kornia.filters"] = k, kf
    sys.path.insert(0, ref)
    torch.set_num_threads(8)
    torch.save(run("reference", ref, "cpu"), out_path)
