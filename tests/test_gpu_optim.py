"""crnerf_b200.optim.Adam (csrc/optim.cu) against torch.optim.Adam, the optimizer the reference
builds in utils/__init__.py:31-32 (`Adam(parameters, lr, eps=1e-8, weight_decay)`)."""
import copy
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cr-nerf-pytorch_b200"))

pytestmark = pytest.mark.gpu

# parameter shapes: the model's (256x256 layers, biases, the 1024x1024 FC), sizes that are not a multiple of four,
# a single element, more than one 4096-element chunk with a ragged tail
SHAPES = [(256, 256), (256,), (1024, 1024), (3, 64), (1,), (7,), (4099,), (64, 93), (128, 283), (5, 3, 3, 3)]


def _params(n_extra=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = SHAPES + [(17 + i,) for i in range(n_extra)]
    return [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]


def _set_grads(ps, step, seed=0):
    g = torch.Generator().manual_seed(1000 * seed + step)
    for p in ps:
        p.grad = (torch.randn(p.shape, generator=g) * (0.1 + 0.01 * step)).cuda()


@pytest.mark.parametrize("kw", [dict(lr=5e-4, eps=1e-8), dict(lr=1e-2, weight_decay=0.1),
                                dict(lr=3e-3, betas=(0.8, 0.99), maximize=True)])
def test_adam_tracks_torch_adam(kw):
    from crnerf_b200.optim import Adam
    from crnerf_b200 import ops
    ours_p, ref_p = _params(n_extra=50), _params(n_extra=50)      # 60 tensors: two launches per step
    ours, ref = Adam(ours_p, **kw), torch.optim.Adam(ref_p, **kw)
    n0 = ops.launch_count()
    for step in range(25):
        _set_grads(ours_p, step); _set_grads(ref_p, step)
        ours.step(); ref.step()
    assert ops.launch_count() - n0 == 25 * 2
    for a, b in zip(ours_p, ref_p):
        # fp32 element-wise math in the same order; the residue is fma contraction
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), float((a - b).abs().max())
    for a, b in zip(ours_p, ref_p):
        sa, sb = ours.state[a], ref.state[b]
        assert float(sa["step"]) == float(sb["step"]) == 25.0
        # with weight decay the gradient carries the parameters' last-bit differences (1e-7 x weight_decay)
        assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=2e-6, atol=1e-7)
        assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=4e-6, atol=1e-9)


def test_adam_state_dict_moves_between_the_two_optimizers():
    from crnerf_b200.optim import Adam
    kw = dict(lr=1e-3, eps=1e-8)
    a_p, b_p = _params(), _params()
    a, b = Adam(a_p, **kw), torch.optim.Adam(b_p, **kw)
    for step in range(5):
        _set_grads(a_p, step); _set_grads(b_p, step)
        a.step(); b.step()
    # ours -> torch and torch -> ours, then five more steps each
    c_p, d_p = [p.detach().clone().requires_grad_(True) for p in a_p], [p.detach().clone().requires_grad_(True) for p in b_p]
    c, d = torch.optim.Adam(c_p, **kw), Adam(d_p, **kw)
    c.load_state_dict(copy.deepcopy(a.state_dict()))
    d.load_state_dict(copy.deepcopy(b.state_dict()))
    for step in range(5, 10):
        for ps in (a_p, b_p, c_p, d_p):
            _set_grads(ps, step)
        for o in (a, b, c, d):
            o.step()
    for pa, pb, pc, pd in zip(a_p, b_p, c_p, d_p):
        for other in (pb, pc, pd):
            assert torch.allclose(pa, other, rtol=2e-6, atol=2e-7)
    assert float(d.state[d_p[0]]["step"]) == 10.0


def test_adam_replays_inside_a_cuda_graph_and_follows_a_tensor_lr():
    from crnerf_b200.optim import Adam
    lr = torch.tensor(1e-3, device="cuda")
    g_p, e_p = _params(), _params()
    graphed, eager = Adam(g_p, lr=lr), torch.optim.Adam(e_p, lr=1e-3)
    _set_grads(g_p, 0); _set_grads(e_p, 0)
    static_grads = [p.grad for p in g_p]
    graphed.step(); eager.step()                      # allocates the state outside the capture
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        graphed.step()
    # the capture itself applied nothing: replay three times with fresh gradients, the last one at a new lr
    for step in range(1, 4):
        _set_grads(e_p, step)
        for s, p in zip(static_grads, e_p):
            s.copy_(p.grad)
        if step == 3:
            lr.fill_(5e-4)
            eager.param_groups[0]["lr"] = 5e-4
        graph.replay(); eager.step()
    torch.cuda.synchronize()
    assert float(graphed.state[g_p[0]]["step"]) == 4.0
    for a, b in zip(g_p, e_p):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)


def test_adam_errors_are_loud():
    from crnerf_b200.optim import Adam
    with pytest.raises(NotImplementedError):
        Adam(_params(), amsgrad=True)
    cpu = [torch.zeros(4, requires_grad=True)]
    cpu[0].grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no fallback"):
        Adam(cpu).step()
    half = [torch.zeros(4, device="cuda", dtype=torch.float16, requires_grad=True)]
    half[0].grad = torch.ones(4, device="cuda", dtype=torch.float16)
    with pytest.raises(RuntimeError, match="no fallback"):
        Adam(half).step()
    # parameters without gradients are skipped like torch.optim.Adam skips them
    ps = _params()
    ps[0].grad = torch.ones_like(ps[0])
    before = ps[1].detach().clone()
    Adam(ps).step()
    assert torch.equal(ps[1], before)
