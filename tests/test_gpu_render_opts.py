"""crnerf_render_pass_opts: the split-precision operand format at default init, the fp16
saturation report, the args.pertubeCord jitter and the channel-sum partials."""
import pytest
import torch

import crnerf_oracle as oracle
from conftest import build_mirror_models, load_golden, state

pytestmark = pytest.mark.gpu
REF = dict(rtol=1e-4, atol=2e-6)


def _setup(peaky=False):
    from models.nerf import PosEmbedding
    models, args = build_mirror_models(0, peaky)
    cpu = {k: state(m) for k, m in models.items()}
    models = {k: m.cuda() for k, m in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    return models, cpu, args, emb


@pytest.mark.parametrize("name", ["render_64p128_eval", "render_64p128_eval_peaky", "render_64p64_train"])
def test_fp16x3_stagewise_matches_golden_tighter_than_fp16(name):
    """The split format must beat plain fp16 by an order of magnitude on the same goldens."""
    from crnerf_b200 import ops
    g = load_golden(name)
    models, _ = build_mirror_models(g["seed"], g["peaky"])
    for typ, z, noise in (("coarse", g["z_coarse"], g["rng"].get("noise_coarse")),
                          ("fine", g["z_fine"], g["rng"].get("noise_fine"))):
        m = models[typ].cuda()
        m.operand = "fp16x3"
        with torch.no_grad():
            packed = m.packed()
        nz = None if (noise is None or g["noise_std"] == 0) else noise.cuda()
        w, f, d = ops.render_pass(packed, g["rays"].cuda(), z.contiguous().cuda(), nz)
        ref = g["ref"]
        assert torch.allclose(f.cpu(), ref[f"feature_{typ}"], rtol=5e-6, atol=2e-7), \
            float((f.cpu() - ref[f"feature_{typ}"]).abs().max())
        assert torch.allclose(w.cpu(), ref[f"weights_{typ}"], rtol=2e-5, atol=5e-7)
        assert torch.allclose(d.cpu(), ref[f"depth_{typ}"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("n_rays,ns,ni", [(1, 16, 0), (77, 40, 24), (300, 64, 128), (513, 100, 60)])
def test_fp16x3_end_to_end_ragged(n_rays, ns, ni):
    from models.rendering import render_rays_cross_ray
    models, cpu, args, emb = _setup(peaky=True)
    models["coarse"].operand = models["fine"].operand = "fp16x3"
    rays = oracle.pinhole_rays(19, 27, oracle.synthetic_pose(4))[:n_rays].contiguous()
    with torch.no_grad():
        res = render_rays_cross_ray(models, emb, rays.cuda(), None, ns, False, 0, 0, ni, 32768, False,
                                    test_time=True, args=args)
        want = oracle.render_rays(cpu["coarse"], cpu["fine"] if ni else None, rays, n_samples=ns, n_importance=ni,
                                  perturb=0, noise_std=0, chunk=8192)
    typ = "fine" if ni else "coarse"
    assert torch.allclose(res[f"feature_{typ}"].cpu(), want[f"feature_{typ}"], **REF)
    assert torch.allclose(res["feature_coarse"].cpu(), want["feature_coarse"], rtol=1e-5, atol=1e-6)


def test_fp16_saturation_is_reported_not_silent():
    """Activations beyond 65504: the fp16 formats clamp (cvt.satfinite) and SAY SO."""
    from crnerf_b200 import CrnerfError, ops
    from models.rendering import render_rays_cross_ray
    models, cpu, args, emb = _setup()
    rays = oracle.pinhole_rays(8, 8, oracle.synthetic_pose(1)).cuda()
    fine = models["fine"]
    with torch.no_grad():
        fine.xyz_encoding_3[0].bias[7] = 2.0e5        # layer-3 unit 7 -> ~2e5 after ReLU: beyond fp16
    render = lambda: render_rays_cross_ray(models, emb, rays, None, 32, False, 0, 0, 32, 32768, False,
                                           test_time=True, args=args)
    for operand in ("fp16", "fp16x3"):
        fine.operand = operand
        with torch.no_grad():
            render()                                   # the pass that saturates
            torch.cuda.synchronize()
            with pytest.raises(CrnerfError, match="65504"):
                render()                               # reported at the next call, without any sync
            packed = fine.packed()
            ops.render_pass(packed, rays, torch.rand(64, 32, device="cuda").sort(1)[0])
            with pytest.raises(CrnerfError, match="65504"):
                packed.check_overflow(sync=True)       # or on demand
    fine.operand = "bf16"                              # fp32 range: nothing to report
    with torch.no_grad():
        res = render()
        res = render()
    assert torch.isfinite(res["feature_fine"]).all()
    fine.packed().check_overflow(sync=True)
    # the coarse model never saturated
    models["coarse"].packed().check_overflow(sync=True)


def test_pertube_cord_matches_reference_math():
    """args.pertubeCord (rendering.py:102-104): xyz += 1e-5*rand, drawn before the noise tensor."""
    from models.rendering import render_rays_cross_ray
    models, cpu, args, emb = _setup()
    args.pertubeCord = True
    rays = oracle.pinhole_rays(6, 7, oracle.synthetic_pose(2))
    n, ns = rays.shape[0], 32
    torch.manual_seed(7)
    with torch.no_grad():
        res = render_rays_cross_ray(models, emb, rays.cuda(), None, ns, False, 0, 0, 0, 32768, False,
                                    test_time=True, args=args)
    # replay the generator: the jitter is the first draw of the pass (perturb == 0)
    torch.manual_seed(7)
    jit = 0.00001 * torch.rand(n * ns, 3, device="cuda")
    z = oracle.coarse_z_vals(rays[:, 6:7], rays[:, 7:8], ns)
    xyz = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]).reshape(-1, 3) + jit.cpu()
    with torch.no_grad():
        x = torch.cat([oracle.pos_embed(xyz, 15), oracle.pos_embed(rays[:, 3:6], 4).repeat_interleave(ns, 0)], 1)
        out = oracle.nerf_sigma_forward(cpu["coarse"], x).reshape(n, ns, 65)
        w, f, d = oracle.composite(out, z, torch.zeros(n, ns))
    assert torch.allclose(res["feature_coarse"].cpu(), f, **REF)
    # and it is not a no-op: the top band sees 2^14 * 1e-5
    args.pertubeCord = False
    with torch.no_grad():
        plain = render_rays_cross_ray(models, emb, rays.cuda(), None, ns, False, 0, 0, 0, 32768, False,
                                      test_time=True, args=args)
    assert (plain["feature_coarse"] - res["feature_coarse"]).abs().max() > 1e-6


@pytest.mark.parametrize("n_rays,ns,ni", [(4096, 64, 128), (333, 40, 24), (1, 16, 16)])
def test_channel_partials_sum_to_feature_sums(n_rays, ns, ni):
    from models.rendering import render_rays_cross_ray
    models, cpu, args, emb = _setup()
    rays = oracle.pinhole_rays(64, 64, oracle.synthetic_pose(0))[:n_rays].contiguous().cuda()
    with torch.no_grad():
        res = render_rays_cross_ray(models, emb, rays, None, ns, False, 0, 0, ni, 32768, False,
                                    test_time=True, args=args, channel_sums=True)
    for typ in ("coarse", "fine"):
        part = res[f"chansum_{typ}"]
        assert part.shape[1] == 64 and part.shape[0] >= 2
        want = res[f"feature_{typ}"].double().sum(0)
        got = part.double().sum(0)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), float((got - want).abs().max())


@pytest.mark.parametrize("hw", [(16, 16), (40, 52)])
def test_style_net_with_render_channel_sums_equals_plain_call(hw):
    """render_rays_cross_ray(..., channel_sums=True) -> style_net.forward(..., channel_sums=...): the
    cross-ray block takes the channel means from the render epilogue's partial sums instead of a
    pass over the feature map; the rgb must agree with the plain call to fp32 rounding
    of the mean (different, fixed, summation orders)."""
    from models.nerf import PosEmbedding
    from models.rendering import render_rays_cross_ray
    h, w = hw
    models, args = build_mirror_models(0)
    models = {k: m.cuda() for k, m in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    rays = oracle.pinhole_rays(h, w, oracle.synthetic_pose(1)).cuda()
    style = torch.rand(1, 64, 32, 32, generator=torch.Generator().manual_seed(2)).cuda()
    with torch.no_grad():
        res = render_rays_cross_ray(models, emb, rays, None, 32, False, 0, 0, 32, 32768, False, test_time=True,
                                    args=args, channel_sums=True)
        content = res["feature_fine"].t().reshape(1, 64, h, w)
        a = models["decoder"](content, style, channel_sums=res["chansum_fine"])
        b = models["decoder"](content, style)
    assert float((a - b).abs().max()) <= 2e-6


def test_pertube_cord_in_the_training_step():
    """args.pertubeCord under autograd (models/rendering.py:102-104 in the training step): the training
    forward takes the same jitter as the inference kernel (same RNG draws in the same order), saves the
    jittered embedding for the weight gradients of layers 1 / 5, and the backward runs."""
    from models.nerf import PosEmbedding
    from models.rendering import render_rays_cross_ray
    models, args = build_mirror_models(0)
    args.pertubeCord = True
    models = {k: m.cuda() for k, m in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    rays = oracle.pinhole_rays(8, 8, oracle.synthetic_pose(1)).cuda()
    torch.manual_seed(7)
    with torch.no_grad():
        want = render_rays_cross_ray(models, emb, rays, None, 16, False, 1.0, 1.0, 16, 32768, False, args=args)
    for m in models.values():
        m.train()
    torch.manual_seed(7)
    got = render_rays_cross_ray(models, emb, rays, None, 16, False, 1.0, 1.0, 16, 32768, False, args=args)
    assert got["feature_fine"].grad_fn is not None
    for k in ("feature_coarse", "feature_fine", "weights_fine", "depth_fine"):
        assert torch.allclose(got[k].detach(), want[k], rtol=1e-5, atol=1e-6), k
    (got["feature_fine"].sum() + got["feature_coarse"].sum()).backward()
    for p in list(models["coarse"].parameters()) + list(models["fine"].parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all()
    args.pertubeCord = False
    torch.manual_seed(7)
    plain = render_rays_cross_ray(models, emb, rays, None, 16, False, 1.0, 1.0, 16, 32768, False, args=args)
    assert not torch.equal(plain["feature_fine"].detach(), got["feature_fine"].detach())   # the jitter did something
