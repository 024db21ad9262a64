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


def rel_l2(a, b):
    a, b = a.detach().cpu().double(), b.double()
    return float((a - b).norm() / b.norm())


def max_rel(a, b, floor):
    a, b = a.detach().cpu().double(), b.double()
    m = b.abs() > floor
    return float(((a - b).abs()[m] / b.abs()[m]).max())


# What each operand format delivers on trained weights (measured on B200, profiles/r02_parity_report.json):
#   fp16x3 (hi+lo operands, 3 MMAs): the north-star bar - 1e-4 relative.
#   fp16   (11-bit operands): the layers of a trained network cancel large terms, so the operand rounding
#          (2^-12 rms per factor) shows as ~1-2e-3 relative in the features - exactly what a float32 emulation
#          of fp16-rounded operands predicts (tests below compare against that emulation at 2e-5) - i.e.
#          PSNR-level parity (>= 70 dB); TF32 on the reference's own GPU path has the same 11 bits.
@pytest.mark.gpu
@pytest.mark.parametrize("operand", ["fp16x3", "fp16"])
@pytest.mark.parametrize("case", ["eval_64p128", "train_64p64"])
def test_render_pass_stagewise_trained(case, operand):
    from crnerf_b200 import ops
    t, models, _ = load_trained()
    c = t["cases"][case]
    rays = c["rays"].cuda()
    for typ in ("coarse", "fine"):
        z = c[f"z_{typ}"]
        noise = c["rng"].get(f"noise_{typ}") if c["train"] else None
        m = models[typ].cuda()
        m.operand = operand
        with torch.no_grad():
            packed = m.packed()
        w, f, d = ops.render_pass(packed, rays, z.contiguous().cuda(), None if noise is None else noise.cuda())
        ref = c["ref"]
        if operand == "fp16x3":
            close(f, ref[f"feature_{typ}"], f"{case}:{typ} feature", **REF)
            bw, bd = composite_bounds(ref[f"weights_{typ}"], z)
            within(w, ref[f"weights_{typ}"], bw, f"{case}:{typ} weights")
            within(d, ref[f"depth_{typ}"], bd, f"{case}:{typ} depth")
        else:
            assert rel_l2(f, ref[f"feature_{typ}"]) < 1e-3
            assert max_rel(f, ref[f"feature_{typ}"], 1e-2) < 5e-3
            assert float((w.cpu() - ref[f"weights_{typ}"]).abs().max()) < 5e-3
            assert oracle.psnr(f.cpu(), ref[f"feature_{typ}"]) > 70.0
            # ... and it is the operand rounding, nothing else: a float32 emulation with fp16-rounded
            # operands (oracle._dense) is off the reference by the same amount (individual roundings
            # differ - the kernel's embedding is only within 4e-6 of libm's - so the two are compared
            # in norm, not element by element)
            p_cpu = state(models[typ].cpu())
            dir_emb = oracle.pos_embed(c["rays"][:, 3:6], 4)
            we, fe, de = oracle._infer(p_cpu, c["rays"][:, 0:3], c["rays"][:, 3:6], dir_emb, z,
                                       torch.zeros_like(z) if noise is None else noise, 15, 8192, 64, torch.float16)
            e_kernel, e_emul = rel_l2(f, ref[f"feature_{typ}"]), rel_l2(fe, ref[f"feature_{typ}"])
            assert 0.5 * e_emul < e_kernel < 2.0 * e_emul, (e_kernel, e_emul)
        packed.check_overflow(sync=True)


@pytest.mark.gpu
@pytest.mark.parametrize("operand", ["fp16x3", "fp16"])
def test_mlp_rows_trained(operand):
    t, models, _ = load_trained()
    c = t["cases"]["mlp"]
    fine = models["fine"].cuda()
    fine.operand = operand
    with torch.no_grad():
        out = fine(c["x"].cuda())
    if operand == "fp16x3":
        close(out[:, :64], c["ref"][:, :64], "features", **REF)
        close(out[:, 64], c["ref"][:, 64], "sigma", rtol=1e-4, atol=1e-5)
    else:
        assert rel_l2(out, c["ref"]) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("operand", ["fp16x3", "fp16", "bf16"])
def test_end_to_end_and_psnr_trained(operand):
    """render_rays_cross_ray + style_net on the trained set.

    End to end the fine depths are OUR resampling of OUR coarse weights.  On this weight set that
    step is ill-conditioned in the reference itself: re-running the CPU reference with the inverse
    CDF accumulated in float64 instead of float32 moves ``feature_fine`` by up to 1.2e-2 (depths
    shift by 4e-5 and the trained field has content at the 2^14 band), so no implementation -
    including the reference on another device - reproduces the stored end-to-end features to 1e-4.
    What is checked instead:
      * fp16x3: the CPU oracle evaluated at OUR depths reproduces our features to 1e-4 (the
        pipeline is self-consistent), and our depths are a valid resampling (sorted, contain the
        coarse grid, new draws inside the coarse range);
      * every format: the metric's second half - |PSNR(ours, T) - PSNR(ref, T)| <= 0.05 dB on the
        right half of the decoded frame (eval_metric.py:89-93) with both PSNRs in a realistic
        15-30 dB regime (T = the reference frame decoded with another style), and the decoded
        frame itself within 50 dB of the reference's."""
    from models.nerf import PosEmbedding
    from models.rendering import render_rays_cross_ray
    from crnerf_b200 import ops
    t, models, args = load_trained()
    f = t["cases"]["frame"]
    h, w = f["hw"]
    cpu_f = state(models["fine"])
    models = {k: m.cuda() for k, m in models.items()}
    models["coarse"].operand = models["fine"].operand = operand
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    rays = f["rays"].cuda()
    with torch.no_grad():
        res = render_rays_cross_ray(models, emb, rays, None, 64, False, 0, 0, 128, 32768, False,
                                    test_time=True, args=args)
        rgb = models["decoder"](res["feature_fine"].t().reshape(1, 64, h, w), f["style_a"].cuda()).cpu()
    if operand == "fp16x3":
        zc = ops.coarse_z(rays, torch.linspace(0, 1, 64, device="cuda"))
        zf = ops.sample_pdf_merge(zc, res["weights_coarse"], torch.linspace(0, 1, 128, device="cuda"), 128)
        assert (zf[:, 1:] >= zf[:, :-1]).all() and float(zf.min()) >= 0.0 and float(zf.max()) <= 5.0
        sub = slice(0, 768)                                   # a quarter of the frame keeps the CPU pass short
        with torch.no_grad():
            dir_emb = oracle.pos_embed(f["rays"][sub, 3:6], 4)
            zs = zf[sub].cpu()
            _, fe, _ = oracle._infer(cpu_f, f["rays"][sub, 0:3], f["rays"][sub, 3:6], dir_emb, zs,
                                     torch.zeros_like(zs), 15, 8192, 64, None)
        close(res["feature_fine"][sub], fe, "feature_fine vs the oracle at our depths", **REF)
    half = lambda x: x[..., w // 2:]
    p_ours, p_ref = oracle.psnr(half(rgb), half(f["rgb_t"])), oracle.psnr(half(f["rgb_a"]), half(f["rgb_t"]))
    assert 15.0 <= p_ref <= 30.0 and 15.0 <= p_ours <= 30.0, (p_ours, p_ref)
    assert abs(p_ours - p_ref) <= 0.05, (operand, p_ours, p_ref)
    assert oracle.psnr(rgb, f["rgb_a"]) > (50.0 if operand != "bf16" else 40.0)
