"""Parity on a TRAINED-LIKE weight set (oracle/make_trained.py: the reference's own modules trained
with Adam on a seeded analytic scene; sharp surfaces, features spread over (0, 0.96), activations
up to ~23, decoded images with real contrast).  VERDICT r1 asked for every stage-wise / end-to-end
check and the PSNR check to run in this regime, not only at default init.

CPU: the oracle reproduces the reference's stored outputs bit for bit.
GPU: the CUDA path against the same outputs at the north-star tolerance (tests/parity_bounds.py).
"""
import pytest
import torch

import crnerf_oracle as oracle
from conftest import load_trained, state
from parity_bounds import composite_bounds

REF = dict(rtol=1e-4, atol=2e-6)


def close(a, b, what, rtol, atol):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err, tol = (a - b).abs(), atol + rtol * b.abs()
    if (err > tol).any():
        i = torch.argmax(err / tol)
        raise AssertionError(f"{what}: {int((err > tol).sum())}/{err.numel()} outside tolerance; worst got "
                             f"{a.flatten()[i]:.8g} want {b.flatten()[i]:.8g} (abs {err.flatten()[i]:.3g})")


def within(a, b, bound, what):
    err = (a.detach().cpu().double() - b.double()).abs()
    if (err > bound).any():
        i = torch.argmax(err / bound)
        raise AssertionError(f"{what}: {int((err > bound).sum())}/{err.numel()} outside the conditioning bound; "
                             f"worst err {err.flatten()[i]:.3g} bound {bound.flatten()[i]:.3g}")


@pytest.mark.parametrize("case", ["eval_64p128", "train_64p64"])
def test_oracle_reproduces_reference_on_trained_weights(case):
    t, models, _ = load_trained()
    c = t["cases"][case]
    with torch.no_grad():
        got = oracle.render_rays(state(models["coarse"]), state(models["fine"]), c["rays"],
                                 n_samples=c["n_samples"], n_importance=c["n_importance"],
                                 perturb=1.0 if c["train"] else 0, noise_std=1.0 if c["train"] else 0,
                                 chunk=8192, rng=c["rng"] if c["train"] else None)
    for k, v in c["ref"].items():
        assert torch.equal(got[k], v), k


def test_oracle_mlp_and_style_on_trained_weights():
    t, models, _ = load_trained()
    with torch.no_grad():
        assert torch.equal(oracle.nerf_sigma_forward(state(models["fine"]), t["cases"]["mlp"]["x"]),
                           t["cases"]["mlp"]["ref"])
        f = t["cases"]["frame"]
        h, w = f["hw"]
        rgb = oracle.style_net_forward(state(models["decoder"]), f["feature_fine"].t().reshape(1, 64, h, w),
                                       f["style_a"])
    assert torch.equal(rgb, f["rgb_a"])
    # the regime the fixture is for
    assert max(t["cases"]["mlp"]["act_absmax"]) > 10.0
    assert float(f["rgb_a"].max() - f["rgb_a"].min()) > 0.4


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["eval_64p128", "train_64p64"])
def test_render_pass_stagewise_trained(case):
    from crnerf_b200 import ops
    t, models, _ = load_trained()
    c = t["cases"][case]
    rays = c["rays"].cuda()
    for typ in ("coarse", "fine"):
        z = c[f"z_{typ}"]
        noise = c["rng"].get(f"noise_{typ}") if c["train"] else None
        m = models[typ].cuda()
        with torch.no_grad():
            packed = m.packed()
        w, f, d = ops.render_pass(packed, rays, z.contiguous().cuda(), None if noise is None else noise.cuda())
        close(f, c["ref"][f"feature_{typ}"], f"{case}:{typ} feature", **REF)
        bw, bd = composite_bounds(c["ref"][f"weights_{typ}"], z)
        within(w, c["ref"][f"weights_{typ}"], bw, f"{case}:{typ} weights")
        within(d, c["ref"][f"depth_{typ}"], bd, f"{case}:{typ} depth")


@pytest.mark.gpu
def test_mlp_rows_trained():
    t, models, _ = load_trained()
    c = t["cases"]["mlp"]
    with torch.no_grad():
        out = models["fine"].cuda()(c["x"].cuda())
    close(out[:, :64], c["ref"][:, :64], "features", **REF)
    close(out[:, 64], c["ref"][:, 64], "sigma", rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_end_to_end_and_psnr_trained():
    """render_rays_cross_ray + style_net on the trained set: features within 1e-4 of the reference,
    and the metric's second half - |PSNR(ours, T) - PSNR(ref, T)| <= 0.05 dB on the right half of the
    frame (eval_metric.py:89-93) - in a regime where those PSNRs are realistic (15-30 dB)."""
    from models.nerf import PosEmbedding
    from models.rendering import render_rays_cross_ray
    t, models, args = load_trained()
    f = t["cases"]["frame"]
    h, w = f["hw"]
    models = {k: m.cuda() for k, m in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    with torch.no_grad():
        res = render_rays_cross_ray(models, emb, f["rays"].cuda(), None, 64, False, 0, 0, 128, 32768, False,
                                    test_time=True, args=args)
        rgb = models["decoder"](res["feature_fine"].t().reshape(1, 64, h, w), f["style_a"].cuda()).cpu()
    close(res["feature_fine"], f["feature_fine"], "feature_fine", **REF)
    close(rgb, f["rgb_a"], "rgb", rtol=1e-4, atol=2e-6)
    half = lambda x: x[..., w // 2:]
    p_ours, p_ref = oracle.psnr(half(rgb), half(f["rgb_t"])), oracle.psnr(half(f["rgb_a"]), half(f["rgb_t"]))
    assert 15.0 <= p_ref <= 30.0 and 15.0 <= p_ours <= 30.0, (p_ours, p_ref)
    assert abs(p_ours - p_ref) <= 0.05, (p_ours, p_ref)
    assert oracle.psnr(rgb, f["rgb_a"]) > 80.0
