"""GPU: csrc/encoder.cu (tcgen05 implicit-GEMM convolutions, fp16 hi/lo split operands) behind
``encoder_sameoutputsize`` against the golden vectors of the unmodified reference and the oracle.

Tolerance: rtol 1e-4 (the north-star fp32 bar) + atol 1e-5 for outputs near zero (the output
is a LeakyReLU of O(0.1) sums)."""
import pytest
import torch

import crnerf_oracle as oracle
from conftest import check_checksums, load_golden, state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _encoder(seed):
    from models.linearStyleTransfer import encoder_sameoutputsize
    torch.manual_seed(seed)
    return encoder_sameoutputsize(out_channel=64).eval()


def test_encoder_golden():
    from crnerf_b200 import ops
    g = load_golden("encoder")
    enc = _encoder(g["seed"])
    check_checksums(enc, g["checksum"])
    enc = enc.to(DEV)
    enc.packed()
    for case in g["cases"]:
        n0 = ops.launch_count()
        with torch.no_grad():
            got = enc(case["x"].to(DEV))
        assert ops.launch_count() - n0 == 7, "the native encoder kernels did not run"
        assert got.shape == case["ref"].shape
        err = (got.cpu() - case["ref"]).abs().max()
        assert torch.allclose(got.cpu(), case["ref"], rtol=1e-4, atol=1e-5), (tuple(case["x"].shape), float(err))


@pytest.mark.parametrize("hw", [(8, 8), (9, 11), (33, 129), (130, 258), (200, 264), (64, 1030)])
def test_encoder_vs_oracle_shapes(hw):
    """Odd sizes (pool floors, partial tiles, rows longer and shorter than a tile), scaled weights."""
    enc = _encoder(3)
    with torch.no_grad():
        for c in enc._convs():
            c.bias.mul_(3.0)      # default-init biases are tiny; make them matter
    x = torch.rand(1, 3, *hw, generator=torch.Generator().manual_seed(hw[0] * 7 + hw[1]))
    with torch.no_grad():
        want = oracle.encoder_forward(state(enc), x)
        got = enc.to(DEV)(x.to(DEV))
    err = (got.cpu() - want).abs().max()
    assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-5), float(err)


def test_encoder_repacks_after_update_and_trains_on_either_path():
    from crnerf_b200 import ops
    enc = _encoder(5).to(DEV)
    x = torch.rand(1, 3, 32, 32, device=DEV)
    with torch.no_grad():
        a = enc(x)
        enc.conv6.weight.mul_(1.5)
        b = enc(x)
    assert not torch.equal(a, b)
    with torch.no_grad():
        want = oracle.encoder_forward({k: v.cpu() for k, v in state(enc).items()}, x.cpu())
    assert torch.allclose(b.cpu(), want, rtol=1e-4, atol=1e-5)
    n0 = ops.launch_count()
    out = enc(x)                      # autograd on: the native training path (tests/test_gpu_encoder_backward.py)
    out.mean().backward()
    assert ops.launch_count() > n0 and enc.conv3.weight.grad is not None
    native = enc.conv3.weight.grad.clone()
    enc.zero_grad()
    enc.train_backend = "library"     # differentiable library ops, kept for A/B checks
    n0 = ops.launch_count()
    enc(x).mean().backward()
    assert ops.launch_count() == n0
    assert torch.allclose(native, enc.conv3.weight.grad, rtol=1e-3, atol=1e-7)


def test_encoder_full_frame_and_strided_input():
    """An 800x800 photo (the frame size of BASELINE configs[3]) against the CPU oracle, fed once as a
    contiguous tensor and once as a channels-last view (the wrapper must not assume NCHW strides)."""
    enc = _encoder(7)
    x = torch.rand(1, 3, 800, 800, generator=torch.Generator().manual_seed(42))
    with torch.no_grad():
        want = oracle.encoder_forward(state(enc), x)
        enc = enc.to(DEV)
        got = enc(x.to(DEV))
        got_cl = enc(x.to(DEV).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
    assert torch.equal(got, got_cl)
    err = (got.cpu() - want).abs().max()
    assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-5), float(err)


def test_no_swizzle_descriptor_probe():
    """tools/nosw_probe (built by __graft_entry__.build): the SWIZZLE_NONE K-major descriptor with a
    row-shifted start address - what makes a 3x3 tap a pointer shift - is bit-exact on this GPU."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "nosw_probe")
    if not os.path.exists(exe):
        pytest.skip("tools/nosw_probe not built")
    for shift, rows in ((0, 130), (1, 130), (2, 130), (131, 390), (262, 390)):
        out = subprocess.run([exe, str(shift), str(rows), "0"], capture_output=True, text=True, timeout=60).stdout
        assert ": ok (0 /" in out, out
