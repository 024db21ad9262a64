"""Packed-weight cache: the fused kernel must never render from a stale weight image
(optimizers that update through ``p.data`` do not bump ``p._version``), the CUDA-graph
renderer must not replay against a freed image, and the training step's deferred fp16 range
verdict must actually be looked at."""
import pytest
import torch

from conftest import build_mirror_models

pytestmark = pytest.mark.gpu


def _setup():
    from models.nerf import PosEmbedding
    from crnerf_b200.synthetic import pinhole_rays, synthetic_pose
    models, args = build_mirror_models(0)
    models = {k: m.cuda() for k, m in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    rays = pinhole_rays(16, 16, synthetic_pose(2)).cuda()
    return models, args, emb, rays


def _render(models, args, emb, rays):
    from models.rendering import render_rays_cross_ray
    with torch.no_grad():
        return render_rays_cross_ray(models, emb, rays, None, 32, False, 0, 0, 32, 32768, False,
                                     test_time=True, args=args)["feature_fine"].clone()


class DataSGD(torch.optim.Optimizer):
    """Updates through ``p.data`` like torch_optimizer's RAdam / Ranger (reference
    utils/__init__.py:33-36): leaves ``p._version`` untouched."""

    def __init__(self, params, lr):
        super().__init__(params, dict(lr=lr))

    def step(self, closure=None):
        for grp in self.param_groups:
            for p in grp["params"]:
                if p.grad is not None:
                    p.data.add_(p.grad.data, alpha=-grp["lr"])


def test_optimizer_step_through_data_invalidates_the_image():
    models, args, emb, rays = _setup()
    fine = models["fine"]
    before = _render(models, args, emb, rays)
    w = fine.static_rgb[0].weight
    v0 = w._version
    opt = DataSGD([w], lr=1.0)
    w.grad = torch.full_like(w, 0.05)
    opt.step()
    assert w._version == v0, "this optimizer was supposed to bypass the version counter"
    after = _render(models, args, emb, rays)
    assert (after - before).abs().max() > 1e-3, "rendered from the stale weight image"


def test_manual_data_edit_with_invalidate():
    models, args, emb, rays = _setup()
    fine = models["fine"]
    before = _render(models, args, emb, rays)
    fine.static_rgb[0].bias.data.add_(0.5)        # weight surgery / EMA style edit
    fine.invalidate_packed()
    after = _render(models, args, emb, rays)
    assert (after - before).abs().max() > 1e-2
    sd = {k: v.clone() for k, v in fine.state_dict().items()}
    sd["static_rgb.0.bias"] -= 0.5
    fine.load_state_dict(sd)                      # load_state_dict invalidates by itself
    again = _render(models, args, emb, rays)
    assert torch.allclose(again, before, rtol=0, atol=1e-6)


def test_training_repacks_every_step():
    models, args, emb, rays = _setup()
    from models.rendering import render_rays_cross_ray
    fine = models["fine"].train()
    res = render_rays_cross_ray(models, emb, rays, None, 16, False, 1.0, 1.0, 16, 32768, False, args=args)
    a = res["feature_fine"].detach().clone()
    fine.static_rgb[0].bias.data.add_(0.5)        # no version bump, no invalidate
    torch.manual_seed(0)
    res = render_rays_cross_ray(models, emb, rays, None, 16, False, 1.0, 1.0, 16, 32768, False, args=args)
    assert (res["feature_fine"].detach() - a).abs().max() > 1e-2


def test_graphed_renderer_recaptures_after_weight_change():
    from crnerf_b200.graphs import GraphedRenderer
    models, args, emb, rays = _setup()
    gr = GraphedRenderer(models, emb, rays.shape[0], 32, 32, args=args)
    a = gr(rays)["feature_fine"].clone()
    assert torch.allclose(a, _render(models, args, emb, rays), rtol=0, atol=0)
    with torch.no_grad():
        models["fine"].static_rgb[0].bias.add_(0.25)
    b = gr(rays)["feature_fine"].clone()
    assert gr.captures == 2
    assert torch.equal(b, _render(models, args, emb, rays))
    assert (a - b).abs().max() > 1e-2
    gr(rays)
    assert gr.captures == 2                         # unchanged weights: plain replay


def test_deferred_fp16_range_verdict_is_consumed_in_training():
    from crnerf_b200 import CrnerfError
    models, args, emb, rays = _setup()
    from models.rendering import render_rays_cross_ray
    fine = models["fine"].train()
    with torch.no_grad():
        fine.xyz_encoding_3[0].weight[5, 7] = 1e5   # beyond the fp16 operand range
    step = lambda: render_rays_cross_ray(models, emb, rays, None, 16, False, 1.0, 1.0, 16, 32768, False,
                                         args=args)
    step()                                           # packs (clamped), verdict still in flight
    torch.cuda.synchronize()
    with pytest.raises(CrnerfError, match="fp16"):
        step()                                       # next pack polls the previous verdict
        torch.cuda.synchronize()
        step()
    # inference checks immediately
    fine.eval()
    with pytest.raises(CrnerfError, match="fp16"):
        _render(models, args, emb, rays)
    fine.operand = "bf16"
    fine.invalidate_packed()
    _render(models, args, emb, rays)                 # bf16 operands hold the value
