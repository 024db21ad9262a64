"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU.

Only tests (and smoke / bench's CPU-baseline leg) may import ``oracle/``; the
product package under ``cr-nerf-pytorch_b200/`` never does.
"""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cr-nerf-pytorch_b200")
for p in (os.path.join(ROOT, "oracle"), PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100) device")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), map_location="cpu", weights_only=False)


def make_args(**kw):
    d = dict(nerf_out_dim=64, pertubeCord=False, img_wh=[32, 32], N_emb_xyz=15, N_emb_dir=4)
    d.update(kw)
    return types.SimpleNamespace(**d)


def sharpen(model, seed):
    """Same 'peaky' transform as oracle/make_golden.py::sharpen."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        w = model.static_sigma[0].weight
        w.mul_(30.0)
        w.add_(0.05 * torch.randn(w.shape, generator=g))
        model.static_sigma[0].bias.sub_(2.0)


def build_mirror_models(seed=0, peaky=False):
    """The product's module mirror, built in the reference's order (coarse, decoder,
    fine - train_mask_grid_sample.py:38-61) so seeded default init equals the reference's."""
    from models.nerf import NeRF_sigma
    from models.linearStyleTransfer import style_net
    torch.manual_seed(seed)
    args = make_args()
    coarse = NeRF_sigma('coarse', args, in_channels_xyz=93, in_channels_dir=27)
    decoder = style_net(args)
    fine = NeRF_sigma('fine', args, in_channels_xyz=93, in_channels_dir=27, encode_appearance=True,
                      in_channels_a=48, encode_random=True)
    if peaky:
        sharpen(coarse, seed + 100)
        sharpen(fine, seed + 101)
    return {"coarse": coarse.eval(), "fine": fine.eval(), "decoder": decoder.eval()}, args


def state(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def check_checksums(module, sums):
    sd = module.state_dict()
    assert set(sd.keys()) == set(sums.keys()), (sorted(sd.keys()), sorted(sums.keys()))
    for k, (s, a) in sums.items():
        v = sd[k].double()
        assert float(v.sum()) == s and float(v.abs().sum()) == a, f"init drift in {k}"


def load_trained():
    """The trained-like weight set of oracle/make_trained.py loaded into the product's module
    mirror: NeRF coarse/fine entirely, style_net except its two seeded-default fc layers (whose
    checksums the fixture carries)."""
    t = load_golden("trained")
    models, args = build_mirror_models(t["seed"], False)
    models["coarse"].load_state_dict(t["coarse"], strict=True)
    models["fine"].load_state_dict(t["fine"], strict=True)
    missing, unexpected = models["decoder"].load_state_dict(t["decoder"], strict=False)
    assert not unexpected and all(k.startswith(("multi_net.snet.fc.", "multi_net.cnet.fc.")) for k in missing)
    sd = models["decoder"].state_dict()
    for k, (s, a) in t["checksum_decoder_frozen"].items():
        v = sd[k].double()
        assert float(v.sum()) == s and float(v.abs().sum()) == a, f"init drift in {k}"
    return t, models, args


@pytest.fixture(scope="session")
def mirror_default():
    return build_mirror_models(0, False)


def ruffle_cgnet(net, g):
    """Same transform as oracle/make_golden.py::ruffle_cgnet."""
    with torch.no_grad():
        for n_, prm in net.named_parameters():
            if ".bn" in n_ or "b1." in n_ or ".act" in n_ or "bn_prelu" in n_:
                prm.add_(0.1 * torch.randn(prm.shape, generator=g))
        for n_, buf in net.named_buffers():
            if n_.endswith("running_mean"):
                buf.copy_(0.2 * torch.randn(buf.shape, generator=g))
            elif n_.endswith("running_var"):
                buf.copy_(0.5 + torch.rand(buf.shape, generator=g))


def build_mirror_cgnet(golden):
    """The mirror's Context_Guided_Network rebuilt the way the golden's reference network was
    (seed, construction, ruffle); returns (net, generator positioned where the cases start)."""
    from models.lightweight_seg import Context_Guided_Network
    torch.manual_seed(golden["seed"])
    net = Context_Guided_Network(classes=1, M=2, N=2, input_channel=3)
    assert torch.equal(torch.rand(1), golden["after_init"]), "construction consumed the RNG differently"
    g = torch.Generator().manual_seed(22)
    ruffle_cgnet(net, g)
    return net, g
