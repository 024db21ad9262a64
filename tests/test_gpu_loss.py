"""GPU: the loss + mask tail (csrc/loss.cu behind the ``losses`` module mirror) against the
golden vectors the unmodified reference produced (tests/golden/loss.pt) and the oracle.

Tolerances: the terms are fp32 means of up to 12k products; the kernel and torch sum in
different orders, so values agree to rtol 2e-6 and gradients to rtol 1e-5 plus an absolute
2e-6 of the largest gradient entry (the mask gradient is a sum of terms of opposite sign)."""
import types

import pytest
import torch

import crnerf_oracle as oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _cuda_inputs(case, requires_grad=True):
    dev = torch.device("cuda:0")
    inp = {k: v.to(dev).clone().requires_grad_(requires_grad) for k, v in case["inputs"].items()}
    return inp, case["targets"].to(dev), types.SimpleNamespace(**case["hp"])


@pytest.mark.parametrize("ci", range(5))
def test_crnerf_loss_values_and_grads(ci):
    import losses
    case = load_golden("loss")["cases"][ci]
    inp, targets, hp = _cuda_inputs(case)
    crit = losses.loss_dict["crnerf"](hp, coef=1)
    ret, w = crit(inp, targets, hp, case["step"])
    assert list(ret) == list(case["ref"]) and w == case["weight"]
    for k, v in case["ref"].items():
        assert torch.allclose(ret[k].detach().cpu(), v, rtol=2e-6, atol=1e-12), (k, float(ret[k]), float(v))
    sum(l for l in ret.values()).backward()
    for k, gref in case["grads"].items():
        got = inp[k].grad
        assert got is not None, k
        assert torch.allclose(got.cpu(), gref, rtol=1e-5, atol=2e-6 * float(gref.abs().max())), \
            (k, (got.cpu() - gref).abs().max())
    for k in inp:
        if k not in case["grads"]:
            assert inp[k].grad is None, k


def test_crnerf_loss_weighted_terms_and_coef():
    """Upstream gradients other than 1 and coef != 1 (checked against oracle autograd)."""
    import losses
    case = load_golden("loss")["cases"][1]
    inp, targets, hp = _cuda_inputs(case)
    crit = losses.CRNeRFLoss(hp, coef=0.7)
    ret, _ = crit(inp, targets, hp, case["step"])
    wts = {k: 0.5 + i for i, k in enumerate(ret)}
    sum(wts[k] * v for k, v in ret.items()).backward()
    ref_in = {k: v.clone().requires_grad_(True) for k, v in case["inputs"].items()}
    ref, _ = oracle.crnerf_loss(ref_in, case["targets"], hp, case["step"], coef=0.7)
    sum(wts[k] * v for k, v in ref.items()).backward()
    for k in ref:
        assert torch.allclose(ret[k].detach().cpu(), ref[k].detach(), rtol=2e-6, atol=1e-12), k
    for k, v in ref_in.items():
        if v.grad is not None:
            assert torch.allclose(inp[k].grad.cpu(), v.grad, rtol=1e-5, atol=2e-6 * float(v.grad.abs().max())), k


def test_color_loss():
    import losses
    case = load_golden("loss")["cases"][0]
    inp, targets, _ = _cuda_inputs(case, requires_grad=False)
    got = losses.ColorLoss(coef=1)({k: inp[k] for k in ("rgb_coarse", "rgb_fine")}, targets)
    mse = torch.nn.MSELoss()
    want = mse(case["inputs"]["rgb_coarse"], case["targets"]) + mse(case["inputs"]["rgb_fine"], case["targets"])
    assert torch.allclose(got.cpu(), want, rtol=2e-6)


def test_loss_large_batch_is_deterministic():
    """A validation-size batch (multi-block reduction): same bits on every call."""
    from crnerf_b200 import loss
    g = torch.Generator().manual_seed(5)
    n = 640_000
    c, f, t = (torch.rand(n, 3, generator=g) for _ in range(3))
    m = torch.rand(n, 1, generator=g)
    dev = torch.device("cuda:0")
    a = loss.ray_loss(c.to(dev), f.to(dev), t.to(dev), m.to(dev), 1.0, 0.05, 1e-3)
    b = loss.ray_loss(c.to(dev), f.to(dev), t.to(dev), m.to(dev), 1.0, 0.05, 1e-3)
    assert torch.equal(a, b)
    hp = types.SimpleNamespace(maskrs_max=0.05, maskrs_min=0.05, maskrs_k=0.0, maskrd=1e-3, weightKL=0, weightRecA=0,
                               weightcontent=0, mse_on_appearance=False)
    ref, _ = oracle.crnerf_loss({"rgb_coarse": c.double(), "rgb_fine": f.double(), "out_mask": m.double()},
                                t.double(), hp, 0)
    for i, k in enumerate(("c_l", "f_l", "r_ms", "r_md")):
        assert abs(float(a[i]) - float(ref[k])) <= 2e-6 * abs(float(ref[k])), k


@pytest.mark.parametrize("ci", range(4))
def test_mask_sample(ci):
    from crnerf_b200 import loss
    case = load_golden("loss")["mask"][ci]
    dev = torch.device("cuda:0")
    pred = case["pred"].to(dev).requires_grad_(True)
    idx = None if case["idx"] is None else case["idx"].to(dev)
    got = loss.mask_sample(pred, case["hw"], idx)
    assert got.shape == case["ref"].shape
    assert torch.allclose(got.detach().cpu(), case["ref"], rtol=0, atol=1e-6), (got.cpu() - case["ref"]).abs().max()
    g_out = torch.rand(got.shape, generator=torch.Generator().manual_seed(ci))
    got.backward(g_out.to(dev))
    ref_pred = case["pred"].clone().requires_grad_(True)
    oracle.mask_sample(ref_pred, case["hw"], case["idx"]).backward(g_out)
    assert torch.allclose(pred.grad.cpu(), ref_pred.grad, rtol=1e-5, atol=1e-6)


def test_loss_rejects_cpu_tensors():
    from crnerf_b200 import loss
    from crnerf_b200.ops import CrnerfError
    with pytest.raises(CrnerfError):
        loss.ray_loss(torch.rand(8, 3), None, torch.rand(8, 3), None)


def test_loss_accepts_strided_views():
    """rgb maps reach the loss as rearranged views (train_mask_grid_sample.py:220): same values as contiguous."""
    from crnerf_b200 import loss
    g = torch.Generator().manual_seed(9)
    dev = torch.device("cuda:0")
    c3n, f3n, t = torch.rand(3, 500, generator=g).to(dev), torch.rand(3, 500, generator=g).to(dev), \
        torch.rand(500, 3, generator=g).to(dev)
    m = torch.rand(500, 2, generator=g).to(dev)[:, :1]
    a = loss.ray_loss(c3n.t(), f3n.t(), t, m, 1.0, 0.02, 1e-3)
    b = loss.ray_loss(c3n.t().contiguous(), f3n.t().contiguous(), t, m.contiguous(), 1.0, 0.02, 1e-3)
    assert torch.equal(a, b)
