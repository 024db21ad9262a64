"""GPU: the grid-sampled training patch (second half) - ``crnerf_grid_patch`` and the
``GridPatchSampler`` built on it - against the reference's own ``__getitem__`` (golden vectors from
the unmodified datasets/phototourism_mask_grid_sample.py:240-275).  Integer / index work: bit-exact."""
import numpy as np
import pytest
import torch

import crnerf_oracle as oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_sampler_reproduces_the_reference_getitem():
    """Same seeds -> same host draws -> the same patch, every tensor bit-identical."""
    from crnerf_b200.sampling import GridPatchSampler
    g = load_golden("grid_patch")
    imgs = [object() for _ in range(len(g["all_imgs_wh"]))]
    for c in g["cases"]:
        s = GridPatchSampler(g["all_rays"], g["all_rgbs"], g["all_imgs_wh"], imgs, batch_size=c["batch_size"],
                             scale_anneal=c["scale_anneal"], min_scale=c["min_scale"])
        torch.manual_seed(c["torch_seed"])
        out = s.sample(c["epoch"], c["idx"])
        s.check()
        assert list(out) == ['rays', 'ts', 'rgbs', 'whole_img', 'rgb_idx', 'min_scale_cur', 'img_wh', 'uv_sample']
        for k, v in c["ref"].items():
            assert out[k].is_cuda and out[k].dtype == v.dtype and out[k].shape == v.shape, k
            assert torch.equal(out[k].cpu(), v), k
        assert out["min_scale_cur"] == c["min_scale_cur"]
        assert out["whole_img"] is imgs[c["sample_ts"]]
        assert torch.equal(out["img_wh"], g["all_imgs_wh"][c["sample_ts"]])


def test_kernel_matches_oracle_on_many_draws_and_odd_sizes():
    """Direct kernel calls over random image sizes / scales / offsets, lattice sizes 1..37."""
    from crnerf_b200 import ops
    gen = torch.Generator().manual_seed(5)
    for trial in range(40):
        n_img = 3
        wh = torch.randint(17, 200, (n_img, 2), generator=gen).float()
        rows = int((wh[:, 0] * wh[:, 1]).sum())
        all_rays = torch.randn(rows, 9, generator=gen)
        all_rays[:, 8] = torch.randint(0, 1500, (rows,), generator=gen).float()
        all_rgbs = torch.rand(rows, 3, generator=gen)
        ts = int(torch.randint(0, n_img, (1,), generator=gen))
        grid = int(torch.randint(1, 38, (1,), generator=gen))
        img_w, img_h = wh[ts]
        scale = torch.rand(1, generator=gen) * 0.75 + 0.25
        h_off = torch.rand(1, generator=gen) * ((1 - scale.item()) * (1 - 1 / img_h))
        w_off = torch.rand(1, generator=gen) * ((1 - scale.item()) * (1 - 1 / img_w))
        want = oracle.grid_patch(all_rays, all_rgbs, wh, ts, grid * grid, scale, h_off, w_off)
        off = (wh[:ts, 0] * wh[:ts, 1]).sum()
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        got = ops.grid_patch(all_rays.cuda(), all_rgbs.cuda(), torch.linspace(0, 1 - 1 / img_w, grid),
                             torch.linspace(0, 1 - 1 / img_h, grid), float(img_w), float(img_h), float(off),
                             scale.item(), h_off.item(), w_off.item(), status=status)
        assert int(status.item()) == 0
        for k, t in zip(("rays", "ts", "rgbs", "rgb_idx", "uv_sample"), got):
            assert torch.equal(t.cpu(), want[k]), (trial, k)


def test_fp32_cache_offset_rounds_like_the_reference_beyond_2p24_rows():
    """all_imgs_wh is fp32 in the reference, so index + offset is an fp32 sum (:266): with more than
    2^24 rows before the image the gathered row is the ROUNDED one.  Same rows here."""
    from crnerf_b200 import ops
    wh = torch.Tensor([[4099, 4099], [61, 47]])            # 16.8 M rows before image 1 (odd -> rounding)
    rows = int(4099 * 4099 + 61 * 47)
    all_rays = torch.zeros(rows, 9)
    all_rays[:, 0] = torch.arange(rows, dtype=torch.float32)         # row id (as fp32), enough to tell rows apart
    all_rays[:, 8] = (torch.arange(rows) % 7).float()
    all_rgbs = torch.zeros(rows, 3)
    all_rgbs[:, 1] = (torch.arange(rows) % 1024).float()
    scale, h_off, w_off = torch.Tensor([0.8]), torch.Tensor([0.1]), torch.Tensor([0.05])
    want = oracle.grid_patch(all_rays, all_rgbs, wh, 1, 1024, scale, h_off, w_off)
    off = (wh[:1, 0] * wh[:1, 1]).sum()
    exact = want["rgb_idx"] + 4099 * 4099
    assert not torch.equal(all_rays[exact, :8], want["rays"]), "case does not exercise the fp32 rounding"
    got = ops.grid_patch(all_rays.cuda(), all_rgbs.cuda(), torch.linspace(0, 1 - 1 / wh[1, 0], 32),
                         torch.linspace(0, 1 - 1 / wh[1, 1], 32), 61.0, 47.0, float(off), 0.8, 0.1, 0.05)
    for k, t in zip(("rays", "ts", "rgbs", "rgb_idx", "uv_sample"), got):
        assert torch.equal(t.cpu(), want[k]), k


def test_inconsistent_sizes_are_reported_not_clamped():
    from crnerf_b200 import ops
    all_rays, all_rgbs = torch.zeros(100, 9).cuda(), torch.zeros(100, 3).cuda()
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.grid_patch(all_rays, all_rgbs, torch.linspace(0, 0.98, 8), torch.linspace(0, 0.98, 8), 50.0, 50.0, 0.0,
                   1.0, 0.0, 0.0, status=status)
    assert int(status.item()) == 1
    with pytest.raises(Exception):
        ops.grid_patch(all_rays.cpu(), all_rgbs.cpu(), torch.linspace(0, 1, 8), torch.linspace(0, 1, 8), 5.0, 5.0, 0.0,
                       1.0, 0.0, 0.0)
