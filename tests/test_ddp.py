"""crnerf_b200.ddp: the gradient exchange of the data-parallel training step (one all-reduce per flat
gradient buffer), world_size-2 gloo on CPU; GraphedTrainStep(optimizer=None, parameters=...) on the GPU."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cr-nerf-pytorch_b200"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make(rank):
    """Three 'modules': two whose gradients are views of one flat buffer each (as the backward kernels leave them),
    one parameter whose gradient owns its storage, one without a gradient."""
    g = torch.Generator().manual_seed(10 + rank)
    shapes = [[(4, 3), (3,), (2, 2)], [(5,), (1, 7)]]
    params = []
    for group in shapes:
        flat = torch.randn(sum(int(torch.Size(s).numel()) for s in group), generator=g)
        o = 0
        for s in group:
            p = torch.nn.Parameter(torch.zeros(s))
            n = p.numel()
            p.grad = flat[o:o + n].view(s)
            o += n
            params.append(p)
    own = torch.nn.Parameter(torch.zeros(6))
    own.grad = torch.randn(6, generator=g)
    params += [own, torch.nn.Parameter(torch.zeros(2))]
    return params


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from crnerf_b200.ddp import allreduce_gradients, gradient_buffers
        params = _make(rank)
        assert len(gradient_buffers(params)) == 3
        # the 19-float flat buffer travels alone, the 12-float one and the stand-alone gradient are packed together
        n = allreduce_gradients(params, pack_below_bytes=64)
        torch.save({"n": n, "grads": [None if p.grad is None else p.grad.clone() for p in params]},
                   os.path.join(out_dir, f"g_{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_allreduce_gradients_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want = [None if a.grad is None else (a.grad + b.grad) / 2 for a, b in zip(_make(0), _make(1))]
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"g_{r}.pt"))
        assert got["n"] == 2          # one flat buffer on its own + one packed collective: not one per tensor
        for g, w in zip(got["grads"], want):
            assert (g is None) == (w is None)
            if g is not None:
                assert torch.allclose(g, w, rtol=1e-6, atol=1e-7)


def test_allreduce_gradients_is_a_noop_without_a_group():
    from crnerf_b200.ddp import allreduce_gradients
    params = _make(0)
    before = [None if p.grad is None else p.grad.clone() for p in params]
    assert allreduce_gradients(params) == 0
    for p, b in zip(params, before):
        assert (p.grad is None) == (b is None) and (b is None or torch.equal(p.grad, b))


@pytest.mark.gpu
def test_graphed_forward_backward_without_optimizer_writes_fresh_gradients():
    """The multi-GPU form of GraphedTrainStep: gradients of replay k are those of one backward, not a running sum
    over the warm-up and earlier replays; the library's gradients arrive as views of a few flat buffers."""
    from crnerf_b200.graphs import GraphedTrainStep
    from crnerf_b200.ddp import gradient_buffers
    from crnerf_b200.optim import Adam
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_train
    dev = torch.device("cuda")
    g, *_ = bench_train.make_step(dev, 1, 0, n_rays=256, ns=32, ni=32)
    with pytest.raises(ValueError):
        GraphedTrainStep(g.loss_fn, optimizer=None)
    graphed = GraphedTrainStep(g.loss_fn, optimizer=None, parameters=g.params)
    torch.manual_seed(7)
    graphed()
    first = [p.grad.clone() for p in g.params]
    torch.manual_seed(7)
    graphed()
    torch.cuda.synchronize()
    # same weights (no optimizer step in between); the jitter / noise draws differ between the replays, so the
    # gradients differ a little - but they are of the same size, not twice as large
    for a, p in zip(first, g.params):
        assert torch.isfinite(p.grad).all()
        na, nb = float(a.norm()), float(p.grad.norm())
        assert nb < 1.5 * na + 1e-12 and na < 1.5 * nb + 1e-12
    bufs = gradient_buffers(g.params)
    # each NeRF_sigma's 24 gradients share one flat buffer; the decoder's 22 are fresh tensors (autograd sums the
    # contributions of its two calls out of place)
    assert len(bufs) <= 2 + 22 < len(g.params) and sum(b.numel() for b in bufs) >= sum(p.numel() for p in g.params)
    opt = Adam(g.params, lr=1e-3)
    opt.step()      # views of the flat buffers are what the optimizer consumes
    assert float(opt.state[g.params[0]]["step"]) == 1.0


def test_native_adam_refuses_what_it_cannot_update_without_a_gpu():
    """Host logic of crnerf_b200.optim.Adam: constructor validation like torch.optim.Adam's, and no fallback
    update for parameters the kernel cannot take (raised before any CUDA call)."""
    from crnerf_b200.optim import Adam
    w = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(NotImplementedError):
        Adam([w], amsgrad=True)
    for bad in (dict(lr=-1.0), dict(eps=-1e-8), dict(betas=(1.0, 0.999)), dict(betas=(0.9, -0.1)), dict(weight_decay=-1.0)):
        with pytest.raises(ValueError):
            Adam([w], **bad)
    opt = Adam([w], lr=1e-3)
    assert opt.step() is None                      # no gradients: nothing to do, no library call
    w.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no fallback"):
        opt.step()
    assert opt.state_dict()["param_groups"][0]["lr"] == 1e-3 and opt.state_dict()["state"] == {}
    import copy
    import pickle
    for clone in (copy.deepcopy(opt), pickle.loads(pickle.dumps(opt))):     # caches are not part of the pickled state
        assert clone._tables == {} and clone._group_step == {} and clone.param_groups[0]["lr"] == 1e-3

