"""Sharded whole-frame plumbing (crnerf_b200/frame.py).

CPU: world_size-2 ``gloo`` processes drive the collective orchestration
(``fuse_decode_sharded``: sums -> all-reduce -> Gram -> all-reduce -> apply -> all-gather)
with an oracle-backed stand-in for the three kernel phases, and must reproduce the
unsharded ``style_net`` of the oracle.  GPU: the same function with the CUDA backend on
one rank (world 1) against ``style_net`` itself.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

import crnerf_oracle as oracle
from conftest import build_mirror_models, state


class OracleStyleBackend:
    """The three phases of csrc/crossray.cu restated with oracle pieces (CPU, fp32)."""

    def __init__(self, p):
        self.p = p

    def sums(self, feat, parts=None):            # (n,64) -> (64,); parts: partial sums whose rows add up to it
        return feat.sum(0) if parts is None else parts.sum(0)

    def _convs(self, prefix, x):                 # x (1,64,n,1)
        h = x
        for j, last in ((0, False), (2, False), (4, True)):
            h = F.conv2d(h, self.p[f"{prefix}.convs.{j}.weight"], self.p[f"{prefix}.convs.{j}.bias"])
            if not last:
                h = F.leaky_relu(h, 0.2)
        return h.reshape(32, -1)

    def gram(self, feat, mean):                  # un-normalised Gram of cnet.convs(x - mean)
        if feat.shape[0] == 0:
            return torch.zeros(32, 32)
        y = self._convs("multi_net.cnet", (feat - mean).t().reshape(1, 64, -1, 1))
        return y @ y.t()

    def apply(self, feat, mean, gram_n, style):  # -> (3,n)
        p = self.p
        c_mat = F.linear(gram_n.reshape(1, -1), p["multi_net.cnet.fc.weight"], p["multi_net.cnet.fc.bias"]).reshape(32, 32)
        s_mean = style.reshape(1, 64, -1).mean(2).reshape(64)
        sf = style - s_mean.reshape(1, 64, 1, 1)
        s_mat = oracle.cnn_forward(p, "multi_net.snet", sf).reshape(32, 32)
        trans = s_mat @ c_mat
        cf = (feat - mean).t().reshape(1, 64, -1, 1)
        comp = F.conv2d(cf, p["multi_net.compress.weight"], p["multi_net.compress.bias"]).reshape(32, -1)
        y = (trans @ comp).reshape(1, 32, -1, 1)
        out = F.conv2d(y, p["multi_net.unzip.weight"], p["multi_net.unzip.bias"]) + s_mean.reshape(1, 64, 1, 1)
        return oracle.neural_renderer_forward(p, out).reshape(3, -1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_dir, with_parts=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from crnerf_b200.frame import fuse_decode_sharded, shard_bounds
        torch.manual_seed(0)
        models, _ = build_mirror_models(0)
        p = state(models["decoder"])
        g = torch.Generator().manual_seed(5)
        feat = torch.rand(n_total, 64, generator=g)
        style = torch.rand(1, 64, 32, 32, generator=g)
        lo, hi = shard_bounds(n_total, world, rank)
        local = feat[lo:hi].contiguous()
        # the render epilogue's per-CTA partial channel sums, emulated: rows that add up to the block's sums
        parts = torch.stack([c.sum(0) for c in local.chunk(7)]) if with_parts and hi > lo else None
        rgb = fuse_decode_sharded(OracleStyleBackend(p), local, style, n_total, sum_parts=parts)
        torch.save(rgb, os.path.join(out_dir, f"rgb_{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_in_order():
    from crnerf_b200.frame import shard_bounds
    for n in (0, 1, 7, 640000, 76800):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            per = -(-n // world) if n else 0
            assert all(hi - lo <= per for lo, hi in blocks)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


@pytest.mark.parametrize("n_total,with_parts", [(32 * 24, False), (1001, False), (1001, True)])   # even / ragged split; sums from partials
def test_fuse_decode_sharded_world2_gloo_matches_unsharded(tmp_path, n_total, with_parts):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path), with_parts), nprocs=world, join=True)
    models, _ = build_mirror_models(0)
    p = state(models["decoder"])
    g = torch.Generator().manual_seed(5)
    feat = torch.rand(n_total, 64, generator=g)
    style = torch.rand(1, 64, 32, 32, generator=g)
    want = oracle.style_net_forward(p, feat.t().reshape(1, 64, 1, n_total), style).reshape(3, n_total)
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"rgb_{r}.pt"))
        assert got.shape == (3, n_total)
        # only the fp32 summation order of the two global reductions differs
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), float((got - want).abs().max())
    assert torch.equal(torch.load(os.path.join(tmp_path, "rgb_0.pt")), torch.load(os.path.join(tmp_path, "rgb_1.pt")))


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["stats", "gather"])
def test_render_frame_sharded_single_rank_matches_two_step_path(scheme):
    """world 1: render_frame_sharded == batched render + style_net (the reference's eval.py:279-294 flow)."""
    from crnerf_b200.frame import batched_render, render_frame_sharded
    from models.nerf import PosEmbedding
    models, args = build_mirror_models(0)
    dev = torch.device("cuda")
    models = {k: m.to(dev) for k, m in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    h, w = 24, 40
    rays = oracle.pinhole_rays(h, w, oracle.synthetic_pose(0)).to(dev)
    style = torch.rand(1, 64, 32, 32, generator=torch.Generator().manual_seed(2)).to(dev)
    rgb = render_frame_sharded(models, emb, rays, style, (h, w), 32, 32, chunk=300, scheme=scheme, args=args)
    res = batched_render(models, emb, rays, 32, 32, False, 300, args=args)
    with torch.no_grad():
        want = models["decoder"](res["feature_fine"].t().reshape(1, 64, h, w), style)
    assert rgb.shape == (1, 3, h, w)
    assert torch.allclose(rgb, want, rtol=1e-5, atol=1e-6)
    # and against the CPU oracle end to end
    p = {k: state(m.cpu()) for k, m in models.items()}
    with torch.no_grad():
        ref = oracle.render_rays(p["coarse"], p["fine"], rays.cpu(), n_samples=32, n_importance=32, perturb=0,
                                 noise_std=0, chunk=8192)
        ref_rgb = oracle.style_net_forward(p["decoder"], ref["feature_fine"].t().reshape(1, 64, h, w), style.cpu())
    assert torch.allclose(rgb.cpu(), ref_rgb, rtol=1e-4, atol=2e-6)


@pytest.mark.gpu
def test_render_frame_from_camera_equals_render_from_rays():
    """rays=None + camera=(K, c2w, near, far): rays built on the device by crnerf_generate_rays."""
    import math
    from crnerf_b200.frame import render_frame_sharded
    from models.nerf import PosEmbedding
    models, args = build_mirror_models(0)
    models = {k: m.cuda() for k, m in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    h, w = 20, 28
    c2w = oracle.synthetic_pose(2)
    f = 0.5 * w / math.tan(0.5 * math.radians(60.0))
    K = [[f, 0.0, w / 2], [0.0, f, h / 2], [0.0, 0.0, 1.0]]
    style = torch.rand(1, 64, 32, 32, generator=torch.Generator().manual_seed(2)).cuda()
    a = render_frame_sharded(models, emb, None, style, (h, w), 32, 32, chunk=256, camera=(K, c2w.tolist(), 0.0, 5.0),
                             args=args)
    b = render_frame_sharded(models, emb, oracle.pinhole_rays(h, w, c2w).cuda(), style, (h, w), 32, 32, chunk=256,
                             args=args)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)


def _nccl_worker(rank, world, port, hw, scheme, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from crnerf_b200.frame import render_frame_sharded
        from models.nerf import PosEmbedding
        models, args = build_mirror_models(0)
        models = {k: m.to(dev) for k, m in models.items()}
        emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
        h, w = hw
        rays = oracle.pinhole_rays(h, w, oracle.synthetic_pose(0)).to(dev)
        style = torch.rand(1, 64, 32, 32, generator=torch.Generator().manual_seed(2)).to(dev)
        rgb = render_frame_sharded(models, emb, rays, style, (h, w), 32, 32, scheme=scheme, args=args)
        torch.cuda.synchronize(dev)
        torch.save(rgb.cpu(), os.path.join(out_dir, f"rgb_{scheme}_{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["stats", "gather"])
def test_render_frame_sharded_two_ranks_nccl_matches_one_rank(tmp_path, scheme):
    """The CUDA cross-ray phases under NCCL with world 2 (needs two GPUs; `bench.py --gpus N` runs the same
    comparison at every N as `frame_check`): ragged row split, both schemes, every rank ends up with the
    frame a single GPU renders."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from crnerf_b200.frame import render_frame_sharded
    from models.nerf import PosEmbedding
    hw = (23, 41)          # 943 rays: the two blocks differ in size
    mp.spawn(_nccl_worker, args=(2, _free_port(), hw, scheme, str(tmp_path)), nprocs=2, join=True)
    models, args = build_mirror_models(0)
    models = {k: m.cuda() for k, m in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    rays = oracle.pinhole_rays(*hw, oracle.synthetic_pose(0)).cuda()
    style = torch.rand(1, 64, 32, 32, generator=torch.Generator().manual_seed(2)).cuda()
    want = render_frame_sharded(models, emb, rays, style, hw, 32, 32, scheme=scheme, args=args).cpu()
    for r in range(2):
        got = torch.load(os.path.join(tmp_path, f"rgb_{scheme}_{r}.pt"))
        assert float((got - want).abs().max()) <= 1e-5
