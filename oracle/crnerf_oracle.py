"""CPU oracle for the CR-NeRF volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  This module restates, in plain functional PyTorch on
the CPU, the algorithm of the reference's rendering path so that the CUDA
product can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it; the
product package never does (it fails loudly when its CUDA library is missing).

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference (``/root/reference/models/{rendering,nerf,linearStyleTransfer,
nerf_decoder_stylenerf}.py``) in the build container, runs both on the same
seeded inputs and asserts bit-exact equality on CPU before writing the fixtures
under ``tests/golden/``.  The reference ships no tests or golden vectors of its
own, so those fixtures are the pin.

All functions are dtype-generic: run them on float64 inputs to get a
high-precision "truth", or pass ``operand_dtype=torch.float16`` to the MLP to
emulate tensor-core operand rounding (fp32 accumulate) when choosing tolerances.

Weights are passed as flat ``dict[str, Tensor]`` using the reference's
``state_dict`` key names (e.g. ``xyz_encoding_1.0.weight``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------
# a1  PosEmbedding.forward            (reference models/nerf.py:5-30)
# --------------------------------------------------------------------------
def pos_embed(x: Tensor, n_freqs: int) -> Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] -> (B, 3+6L).

    The reference builds ``freqs = 2**linspace(0, L-1, L)`` (models/nerf.py:12-13,
    callers pass ``PosEmbedding(L-1, L)``) which are exact powers of two, and
    evaluates ``func(freq*x)`` per band (models/nerf.py:26-28).
    """
    freqs = 2.0 ** torch.linspace(0, n_freqs - 1, n_freqs)
    parts = [x]
    for f in freqs:
        fx = f.to(x.dtype) * x
        parts.append(torch.sin(fx))
        parts.append(torch.cos(fx))
    return torch.cat(parts, dim=-1)


# --------------------------------------------------------------------------
# a2  NeRF_sigma.forward              (reference models/nerf.py:157-182)
# --------------------------------------------------------------------------
def _dense(x: Tensor, w: Tensor, b: Tensor, operand_dtype) -> Tensor:
    if operand_dtype is None:
        return F.linear(x, w, b)
    # tensor-core emulation: operands rounded, products/accumulation in fp32
    xr = x.to(operand_dtype).to(torch.float32)
    wr = w.to(operand_dtype).to(torch.float32)
    return xr @ wr.t() + b


def nerf_sigma_forward(p: Params, x: Tensor, *, e_xyz: int = 93, e_dir: int = 27,
                       depth: int = 8, skips=(4,), sigma_only: bool = False,
                       operand_dtype=None) -> Tensor:
    """(B, e_xyz+e_dir) -> (B, out_dim+1) = [sigmoid features | softplus sigma].

    8 x (Linear+ReLU) with the embedded xyz re-concatenated *in front of* the
    hidden state at layer index 4 (models/nerf.py:167-170); sigma =
    softplus(Linear(256->1)) of the trunk output (models/nerf.py:172);
    ``xyz_encoding_final`` then [final | dir] -> Linear+ReLU -> Linear+Sigmoid
    (models/nerf.py:176-180); output is cat([features, sigma]) (:181).
    """
    if sigma_only:
        xyz = x
        dirs = None
    else:
        xyz, dirs = x[:, :e_xyz], x[:, e_xyz:e_xyz + e_dir]
    h = xyz
    for i in range(depth):
        if i in skips:
            h = torch.cat([xyz, h], dim=1)
        h = torch.relu(_dense(h, p[f"xyz_encoding_{i+1}.0.weight"],
                              p[f"xyz_encoding_{i+1}.0.bias"], operand_dtype))
    # sigma head is kept in full precision in every mode (the CUDA path evaluates
    # it as an fp32 dot product in the layer-8 epilogue)
    sigma = F.softplus(F.linear(h, p["static_sigma.0.weight"], p["static_sigma.0.bias"]))
    if sigma_only:
        return sigma
    fin = _dense(h, p["xyz_encoding_final.weight"], p["xyz_encoding_final.bias"], operand_dtype)
    d = torch.relu(_dense(torch.cat([fin, dirs], dim=1), p["dir_encoding.0.weight"],
                          p["dir_encoding.0.bias"], operand_dtype))
    feat = torch.sigmoid(_dense(d, p["static_rgb.0.weight"], p["static_rgb.0.bias"],
                                operand_dtype))
    return torch.cat([feat, sigma], dim=-1)


# --------------------------------------------------------------------------
# a4  sample_pdf                      (reference models/rendering.py:7-46)
# --------------------------------------------------------------------------
def sample_pdf(bins: Tensor, weights: Tensor, n_importance: int, det: bool = False,
               eps: float = 1e-5, u: Optional[Tensor] = None) -> Tensor:
    """Inverse-CDF sampling of ``n_importance`` depths per ray.

    bins (N, M+1), weights (N, M) -> (N, n_importance).  ``u`` may be supplied to
    pin the uniform draws; otherwise they come from ``torch.rand`` exactly where
    the reference draws them (rendering.py:30) so a shared seed gives shared
    samples.
    """
    n_rays, m = weights.shape
    w = weights + eps                                           # :20
    pdf = w / w.sum(dim=1, keepdim=True)                        # :21
    cdf = torch.cumsum(pdf, dim=-1)                             # :22
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], dim=-1)  # :23
    if u is None:
        if det:
            u = torch.linspace(0, 1, n_importance, device=bins.device)   # :27
            u = u.expand(n_rays, n_importance)
        else:
            u = torch.rand(n_rays, n_importance, device=bins.device)     # :30
    u = u.contiguous().to(cdf.dtype)
    idx = torch.searchsorted(cdf, u, right=True)                # :33
    lo = (idx - 1).clamp_min(0)                                 # :34
    hi = idx.clamp_max(m)                                       # :35
    cdf_lo, cdf_hi = cdf.gather(1, lo), cdf.gather(1, hi)       # :38
    bin_lo, bin_hi = bins.gather(1, lo), bins.gather(1, hi)     # :39
    denom = cdf_hi - cdf_lo                                     # :41
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)  # :42
    return bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)    # :45


# --------------------------------------------------------------------------
# a3  inference closure               (reference models/rendering.py:82-145)
# --------------------------------------------------------------------------
def composite(out: Tensor, z_vals: Tensor, noise: Tensor, out_dim: int = 64):
    """(N,S,out_dim+1) MLP outputs + depths -> weights (N,S), feature (N,out_dim), depth (N,).

    deltas with the last one 1e2 (rendering.py:121-123); alpha = 1-exp(-delta *
    relu(sigma+noise)) (:126); transmittance = exclusive cumprod of (1-alpha)
    (:128-130); weights = alpha*T (:132); feature = sum_s w*f (:136-137);
    depth = sum_s w*z (:143).
    """
    feats = out[..., :out_dim]
    sigmas = out[..., out_dim]
    deltas = z_vals[:, 1:] - z_vals[:, :-1]
    deltas = torch.cat([deltas, 1e2 * torch.ones_like(deltas[:, :1])], dim=-1)
    alphas = 1 - torch.exp(-deltas * torch.relu(sigmas + noise))
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas], dim=-1)
    trans = torch.cumprod(shifted[:, :-1], dim=-1)
    weights = alphas * trans
    feature = (weights.unsqueeze(-1) * feats).sum(dim=1)
    depth = (weights * z_vals).sum(dim=1)
    return weights, feature, depth


def _infer(p: Params, rays_o, rays_d, dir_emb, z_vals, noise, n_freq_xyz, chunk,
           out_dim, operand_dtype):
    n, s = z_vals.shape
    xyz = (rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]).reshape(-1, 3)  # :178,:101
    dir_rep = dir_emb[:, None, :].expand(n, s, dir_emb.shape[-1]).reshape(n * s, -1)     # :108
    outs = []
    for i in range(0, n * s, chunk):                                                      # :110-114
        x = torch.cat([pos_embed(xyz[i:i + chunk], n_freq_xyz), dir_rep[i:i + chunk]], dim=1)
        outs.append(nerf_sigma_forward(p, x, e_xyz=3 + 6 * n_freq_xyz,
                                       e_dir=dir_emb.shape[-1], operand_dtype=operand_dtype))
    out = torch.cat(outs, dim=0).reshape(n, s, out_dim + 1)
    return composite(out, z_vals, noise, out_dim)


# --------------------------------------------------------------------------
# a5  render_rays_cross_ray           (reference models/rendering.py:50-196)
# --------------------------------------------------------------------------
def coarse_z_vals(near: Tensor, far: Tensor, n_samples: int, use_disp: bool = False) -> Tensor:
    """Per-ray depths before jitter (rendering.py:161-167).  near/far are (N,1)."""
    t = torch.linspace(0, 1, n_samples, device=near.device).to(near.dtype)
    if not use_disp:
        z = near * (1 - t) + far * t
    else:
        z = 1 / (1 / near * (1 - t) + 1 / far * t)
    return z.expand(near.shape[0], n_samples)


def jitter_z_vals(z: Tensor, perturb_rand: Tensor) -> Tensor:
    """Stratified jitter (rendering.py:169-176); perturb_rand = perturb*U[0,1)."""
    mid = 0.5 * (z[:, :-1] + z[:, 1:])
    upper = torch.cat([mid, z[:, -1:]], dim=-1)
    lower = torch.cat([z[:, :1], mid], dim=-1)
    return lower + (upper - lower) * perturb_rand


def render_rays(p_coarse: Params, p_fine: Optional[Params], rays: Tensor, *,
                n_samples: int = 64, n_importance: int = 0, use_disp: bool = False,
                perturb: float = 0.0, noise_std: float = 1.0, chunk: int = 32768,
                n_freq_xyz: int = 15, n_freq_dir: int = 4, out_dim: int = 64,
                view_dir: Optional[Tensor] = None, rng: Optional[dict] = None,
                operand_dtype=None, record: Optional[dict] = None) -> Dict[str, Tensor]:
    """Restatement of ``render_rays_cross_ray``.  rays (N,8) = [o3, d3, near, far].

    Random numbers are drawn with the same torch calls in the same order as the
    reference (``rand_like(z)`` :175, ``randn_like(sigma)`` coarse :125,
    ``rand(N,Ni)`` :30, ``randn_like`` fine :125) unless ``rng`` supplies them
    (keys ``perturb_rand``, ``noise_coarse``, ``u``, ``noise_fine``).  When
    ``record`` is a dict the draws and the z values are stored in it so the same
    numbers can be fed to the CUDA path.
    """
    rng = rng or {}
    n = rays.shape[0]
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    dir_emb = pos_embed(rays_d if view_dir is None else view_dir, n_freq_dir)     # :155
    z = coarse_z_vals(near, far, n_samples, use_disp)
    if perturb > 0:
        pr = rng.get("perturb_rand")
        if pr is None:
            pr = perturb * torch.rand_like(z)                                    # :175
        z = jitter_z_vals(z, pr)
        if record is not None:
            record["perturb_rand"] = pr
    nz = rng.get("noise_coarse")
    if nz is None:
        nz = torch.randn(n, n_samples, dtype=z.dtype) * noise_std                # :125
    if record is not None:
        record["noise_coarse"] = nz
        record["z_coarse"] = z
    res: Dict[str, Tensor] = {}
    w, f, d = _infer(p_coarse, rays_o, rays_d, dir_emb, z, nz, n_freq_xyz, chunk, out_dim,
                     operand_dtype)
    res["weights_coarse"], res["feature_coarse"], res["depth_coarse"] = w, f, d
    if n_importance > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])                                        # :183
        u = rng.get("u")
        z_new = sample_pdf(mid, w[:, 1:-1].detach(), n_importance,
                           det=(perturb == 0), u=u)                              # :184-185
        z_f = torch.sort(torch.cat([z, z_new], dim=-1), dim=-1)[0]               # :187
        nzf = rng.get("noise_fine")
        if nzf is None:
            nzf = torch.randn(n, n_samples + n_importance, dtype=z.dtype) * noise_std
        if record is not None:
            record["noise_fine"] = nzf
            record["z_fine"] = z_f
        w, f, d = _infer(p_fine, rays_o, rays_d, dir_emb, z_f, nzf, n_freq_xyz, chunk,
                         out_dim, operand_dtype)
        res["weights_fine"], res["feature_fine"], res["depth_fine"] = w, f, d
    return res


# --------------------------------------------------------------------------
# a6  CNN.forward                     (reference models/linearStyleTransfer.py:28-37)
# --------------------------------------------------------------------------
def cnn_forward(p: Params, prefix: str, x: Tensor) -> Tensor:
    """(1,64,H,W) -> (1,1024): three 1x1 convs (LeakyReLU 0.2 between), Gram/(h*w), fc."""
    h = x
    for j, last in ((0, False), (2, False), (4, True)):
        h = F.conv2d(h, p[f"{prefix}.convs.{j}.weight"], p[f"{prefix}.convs.{j}.bias"])
        if not last:
            h = F.leaky_relu(h, 0.2)
    b, c, hh, ww = h.shape
    y = h.reshape(b, c, hh * ww)
    gram = torch.bmm(y, y.transpose(1, 2)) / (hh * ww)
    return F.linear(gram.reshape(b, -1), p[f"{prefix}.fc.weight"], p[f"{prefix}.fc.bias"])


# --------------------------------------------------------------------------
# a7  MulLayer.forward                (reference models/linearStyleTransfer.py:58-90)
# --------------------------------------------------------------------------
def mul_layer_forward(p: Params, content: Tensor, style: Tensor, prefix: str = "multi_net"):
    """content (1,64,H,W), style (1,64,h,w) -> fused (1,64,H,W), transmatrix (1,32,32)."""
    b, c, hh, ww = content.shape
    c_mean = content.reshape(b, c, -1).mean(dim=2, keepdim=True).unsqueeze(3)    # :61-64
    cf = content - c_mean                                                        # :65
    sb, sc, _, _ = style.shape
    s_mean = style.reshape(sb, sc, -1).mean(dim=2, keepdim=True).unsqueeze(3)    # :68-70
    sf = style - s_mean                                                          # :73
    comp = F.conv2d(cf, p[f"{prefix}.compress.weight"], p[f"{prefix}.compress.bias"])  # :76
    comp = comp.reshape(b, comp.shape[1], -1)
    c_mat = cnn_forward(p, f"{prefix}.cnet", cf).reshape(b, 32, 32)               # :81,:85
    s_mat = cnn_forward(p, f"{prefix}.snet", sf).reshape(sb, 32, 32)              # :82,:84
    trans = torch.bmm(s_mat, c_mat)                                              # :86
    y = torch.bmm(trans, comp).reshape(b, 32, hh, ww)                            # :87
    out = F.conv2d(y, p[f"{prefix}.unzip.weight"], p[f"{prefix}.unzip.bias"])     # :88
    return out + s_mean, trans                                                   # :89


# --------------------------------------------------------------------------
# a9  NeuralRenderer.forward at n_blocks == 0
#                                     (reference models/nerf_decoder_stylenerf.py:279-291)
# --------------------------------------------------------------------------
def neural_renderer_forward(p: Params, x: Tensor, prefix: str = "decoder") -> Tensor:
    """sigmoid(conv1x1 64->3).  With featmap_size == img_size the upsampling loop
    has zero iterations (nerf_decoder_stylenerf.py:239, :281)."""
    rgb = F.conv2d(x, p[f"{prefix}.feat_2_rgb_list.0.weight"], p[f"{prefix}.feat_2_rgb_list.0.bias"])
    return torch.sigmoid(rgb)


# --------------------------------------------------------------------------
# a8  style_net.forward               (reference models/linearStyleTransfer.py:284-291)
# --------------------------------------------------------------------------
def style_net_forward(p: Params, content: Tensor, style: Optional[Tensor], type=None) -> Tensor:
    if style is None and type == "content":
        return neural_renderer_forward(p, content)
    fused, _ = mul_layer_forward(p, content, style)
    return neural_renderer_forward(p, fused)


# --------------------------------------------------------------------------
# synthetic "Phototourism-shaped" inputs
# --------------------------------------------------------------------------
BRANDENBURG_POSE = None  # filled lazily; numbers restated from the reference's test-pose script


def pinhole_rays(h: int, w: int, c2w: Tensor, near: float = 0.0, far: float = 5.0,
                 fov_deg: float = 60.0) -> Tensor:
    """(h*w, 8) rays for a pinhole camera; follows datasets/ray_utils.py:5-52
    (directions ((i-cx)/f, -(j-cy)/f, -1), rotated by c2w[:, :3], normalised,
    origin = c2w[:, 3])."""
    f = 0.5 * w / math.tan(0.5 * math.radians(fov_deg))
    j, i = torch.meshgrid(torch.arange(h, dtype=torch.float32),
                          torch.arange(w, dtype=torch.float32), indexing="ij")
    dirs = torch.stack([(i - w / 2) / f, -(j - h / 2) / f, -torch.ones_like(i)], dim=-1)
    rays_d = dirs @ c2w[:, :3].T
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    rays_o = c2w[:, 3].expand(rays_d.shape)
    rays_d = rays_d.reshape(-1, 3)
    rays_o = rays_o.reshape(-1, 3)
    nf = torch.tensor([near, far], dtype=torch.float32).expand(rays_d.shape[0], 2)
    return torch.cat([rays_o, rays_d, nf], dim=1).contiguous()


def synthetic_pose(seed: int = 0) -> Tensor:
    """A deterministic 3x4 camera-to-world pose looking down -z with a mild
    rotation, camera a little off the origin (scene units as Phototourism after
    the far=5 rescale, datasets/phototourism_mask_grid_sample.py:139-141)."""
    g = torch.Generator().manual_seed(seed)
    ang = (torch.rand(3, generator=g) - 0.5) * 0.4
    cx, sx = math.cos(ang[0]), math.sin(ang[0])
    cy, sy = math.cos(ang[1]), math.sin(ang[1])
    cz, sz = math.cos(ang[2]), math.sin(ang[2])
    rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=torch.float32)
    ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float32)
    rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=torch.float32)
    t = (torch.rand(3, 1, generator=g) - 0.5) * torch.tensor([[0.6], [0.2], [0.6]])
    return torch.cat([rz @ ry @ rx, t], dim=1)


def psnr(a: Tensor, b: Tensor) -> float:
    """-10 log10(mse) as metrics.py:4-13."""
    return float(-10.0 * torch.log10(torch.mean((a.double() - b.double()) ** 2)))


# --------------------------------------------------------------------------
# f3  loss + mask tail of the training step
#     CRNeRFLoss.forward (reference losses.py:50-89), ExponentialAnnealingWeight
#     (losses.py:30-39), the mask lookup of NeRFSystem.forward
#     (train_mask_grid_sample.py:171-175)
# --------------------------------------------------------------------------
def annealing_weight(hp, global_step: int) -> float:
    """ExponentialAnnealingWeight.getWeight (losses.py:38-39)."""
    return max(hp.maskrs_min, hp.maskrs_max * math.exp(-global_step * hp.maskrs_k))


def crnerf_loss(inputs: Dict[str, Tensor], targets: Tensor, hp, global_step: int, coef: float = 1.0):
    """Returns (dict of loss terms in the reference's key order, annealing weight)."""
    ret = {}
    size_delta = annealing_weight(hp, global_step)
    if "a_embedded" in inputs:
        ret["kl_a"] = torch.mean(inputs["a_embedded"] ** 2) * hp.weightKL                       # :53, :91-94
        if "a_embedded_random_rec" in inputs:
            d = inputs["a_embedded_random"].detach() - inputs["a_embedded_random_rec"]
            ret["rec_a_random"] = ((d ** 2).mean() if hp.mse_on_appearance else d.abs().mean()) * hp.weightRecA
    if "out_mask" in inputs:
        mask = inputs["out_mask"]
        ret["c_l"] = 0.5 * ((1 - mask.detach()) * (inputs["rgb_coarse"] - targets) ** 2).mean()  # :63-64
    else:
        ret["c_l"] = 0.5 * ((inputs["rgb_coarse"] - targets) ** 2).mean()
    if "content_wo_a_embed" in inputs and "content_with_a_embed" in inputs:
        ret["content_constraint"] = ((inputs["content_wo_a_embed"] - inputs["content_with_a_embed"]) ** 2
                                     ).mean() * hp.weightcontent                                  # :67-68
    if "rgb_fine" in inputs:
        if "out_mask" in inputs:
            ret["r_ms"] = torch.mean(mask ** 2) * size_delta                                      # :80-84
            ret["r_md"] = torch.mean(1 / ((mask - 0.5) ** 2 + 0.02)) * hp.maskrd                  # :86-87
            ret["f_l"] = 0.5 * ((1 - mask) * (inputs["rgb_fine"] - targets) ** 2).mean()          # :72
        else:
            ret["f_l"] = 0.5 * ((inputs["rgb_fine"] - targets) ** 2).mean()
    return {k: coef * v for k, v in ret.items()}, size_delta


def mask_sample(pred: Tensor, hw, idx: Optional[Tensor]) -> Tensor:
    """pred (1,C,h,w) -> bilinear (align_corners=False) at size hw -> '(h w) c' rows -> [idx]
    (train_mask_grid_sample.py:172-175)."""
    up = F.interpolate(pred, size=tuple(hw), mode="bilinear", align_corners=False)
    rows = up[0].permute(1, 2, 0).reshape(-1, pred.shape[1])
    return rows if idx is None else rows[idx]


# --------------------------------------------------------------------------
# f1  encoder_sameoutputsize.forward  (reference models/linearStyleTransfer.py:250-276)
# --------------------------------------------------------------------------
def encoder_forward(p: Params, x: Tensor) -> Tensor:
    """x (1,3,H,W) -> (1,out,32,32); p holds conv1..conv7 .weight/.bias under the module's keys."""
    def conv(name, t):
        return F.conv2d(t, p[f"{name}.weight"], p[f"{name}.bias"])

    def pad(t):
        return F.pad(t, (1, 1, 1, 1), mode="reflect")

    act = lambda t: F.leaky_relu(t, 0.2)
    h = act(conv("conv2", pad(conv("conv1", x))))
    h = act(conv("conv3", pad(h)))
    h = F.max_pool2d(h, 2, 2)
    h = act(conv("conv4", pad(h)))
    h = act(conv("conv5", pad(h)))
    h = F.max_pool2d(h, 2, 2)
    h = act(conv("conv6", pad(h)))
    h = F.adaptive_avg_pool2d(h, 32)
    return act(conv("conv7", h))


# --------------------------------------------------------------------------
# f2  grid-sampled training patch  (reference datasets/phototourism_mask_grid_sample.py:240-275)
# --------------------------------------------------------------------------
def grid_patch(all_rays: Tensor, all_rgbs: Tensor, all_imgs_wh: Tensor, sample_ts: int, batch_size: int,
               scale: Tensor, h_offset: Tensor, w_offset: Tensor) -> Dict[str, Tensor]:
    """The lattice arithmetic and gathers of the training ``__getitem__`` for image ``sample_ts``,
    given the three uniform draws (scale, h_offset, w_offset: (1,) fp32 tensors, :252-254).
    ``all_imgs_wh`` is the reference's fp32 (n_img, 2) tensor (:197): sizes stay 0-dim fp32 tensors,
    so ``1 - 1/img_w``, ``h_sb * img_h`` and the cache offset are fp32 arithmetic, as there."""
    img_w, img_h = all_imgs_wh[sample_ts]
    g = int(math.sqrt(batch_size))
    w_samples, h_samples = torch.meshgrid([torch.linspace(0, 1 - 1 / img_w, g),
                                           torch.linspace(0, 1 - 1 / img_h, g)], indexing="ij")   # :246-247
    h_sb = h_samples * scale + h_offset                                                            # :255
    w_sb = w_samples * scale + w_offset
    h = (h_sb * img_h).floor()                                                                     # :257
    w = (w_sb * img_w).floor()
    idx = (w + h * img_w).permute(1, 0).contiguous().view(-1).long()                               # :260
    uv = torch.cat((h_sb.permute(1, 0).contiguous().view(-1, 1),
                    w_sb.permute(1, 0).contiguous().view(-1, 1)), -1)                              # :262
    rows = (idx + (all_imgs_wh[:sample_ts, 0] * all_imgs_wh[:sample_ts, 1]).sum()).long()          # :264
    return {"rays": all_rays[rows, :8], "ts": all_rays[rows, 8].long(), "rgbs": all_rgbs[rows],
            "rgb_idx": idx, "uv_sample": uv}


# --------------------------------------------------------------------------
# f3  Context_Guided_Network.forward  (reference models/lightweight_seg.py:271-368, built as
#     Context_Guided_Network(classes=1, M=2, N=2, input_channel=3), train_mask_grid_sample.py:114)
# --------------------------------------------------------------------------
def cgnet_forward(p: Params, x: Tensor, m_blocks: int = 2, n_blocks: int = 2, train: bool = False) -> Tensor:
    """x (B,C,H,W) -> sigmoid mask (B,classes,H,W).  ``p`` is the module's state_dict.  ``train``
    selects BatchNorm batch statistics (as ``module.train()``; running stats are not updated here)."""
    def bn_prelu(name_bn, name_act, t):
        t = F.batch_norm(t, p[f"{name_bn}.running_mean"], p[f"{name_bn}.running_var"], p[f"{name_bn}.weight"],
                         p[f"{name_bn}.bias"], training=train, momentum=0.0, eps=1e-3)       # :25,48
        return F.prelu(t, p[f"{name_act}.weight"])

    def conv(name, t, stride=1, dil=1, groups=1):
        w = p[f"{name}.weight"]
        pad = ((w.shape[-1] - 1) // 2) * dil
        return F.conv2d(t, w, None, stride, pad, dil, groups)

    def cbp(name, t, stride=1):                                                               # ConvBNPReLU :28-37
        return bn_prelu(f"{name}.bn", f"{name}.act", conv(f"{name}.conv", t, stride))

    def glo(name, t):                                                                         # FGlo :183-188
        y = t.mean((2, 3))
        y = torch.sigmoid(F.linear(F.relu(F.linear(y, p[f"{name}.fc.0.weight"], p[f"{name}.fc.0.bias"])),
                                   p[f"{name}.fc.2.weight"], p[f"{name}.fc.2.bias"]))
        return t * y[:, :, None, None]

    def down(name, t, dil):                                                                   # :211-224
        o = cbp(f"{name}.conv1x1", t, 2)
        c = o.shape[1]
        j = torch.cat([conv(f"{name}.F_loc.conv", o, groups=c), conv(f"{name}.F_sur.conv", o, dil=dil, groups=c)], 1)
        j = bn_prelu(f"{name}.bn", f"{name}.act", j)
        return glo(f"{name}.F_glo", conv(f"{name}.reduce.conv", j))

    def block(name, t, dil):                                                                  # :244-257
        o = cbp(f"{name}.conv1x1", t)
        c = o.shape[1]
        j = torch.cat([conv(f"{name}.F_loc.conv", o, groups=c), conv(f"{name}.F_sur.conv", o, dil=dil, groups=c)], 1)
        return t + glo(f"{name}.F_glo", bn_prelu(f"{name}.bn_prelu.bn", f"{name}.bn_prelu.act", j))

    pool = lambda t: F.avg_pool2d(t, 3, 2, 1)                                                 # InputInjection :263-268
    o0 = cbp("level1_2", cbp("level1_1", cbp("level1_0", x, 2)))                              # :327-329
    inp1 = pool(x)
    inp2 = pool(pool(x))
    o1_0 = down("level2_0", bn_prelu("b1.bn", "b1.act", torch.cat([o0, inp1], 1)), 2)          # :334-335
    o1 = o1_0
    for i in range(m_blocks - 1):
        o1 = block(f"level2.{i}", o1, 2)
    o2_0 = down("level3_0", bn_prelu("bn_prelu_2.bn", "bn_prelu_2.act", torch.cat([o1, o1_0, inp2], 1)), 4)
    o2 = o2_0
    for i in range(n_blocks - 1):
        o2 = block(f"level3.{i}", o2, 4)
    cat = bn_prelu("bn_prelu_3.bn", "bn_prelu_3.act", torch.cat([o2_0, o2], 1))               # :356
    logits = conv("classifier.0.conv", cat)
    up = F.interpolate(logits, x.shape[2:], mode="bilinear", align_corners=False)             # :365
    return torch.sigmoid(up)
