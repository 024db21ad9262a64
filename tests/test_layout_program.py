"""CPU: the weight-chunk PROGRAM the fused kernel executes (csrc/nerf_layout.h,
exported through crnerf_debug_program) is emulated in float64 - same units, same
chunks, same A-operand sources, same padding - and must reproduce the oracle's
NeRF_sigma.forward.  Catches tiling / column-mapping bugs without a GPU."""
import ctypes

import pytest
import torch

import crnerf_oracle as oracle
from conftest import build_mirror_models, state

LAYER_KEYS = [f"xyz_encoding_{i}.0" for i in range(1, 9)] + ["xyz_encoding_final", "dir_encoding.0",
                                                             "static_rgb.0"]


def get_program(e_xyz, e_dir):
    from crnerf_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_int32 * 4096)()
    n = lib.crnerf_debug_program(e_xyz, e_dir, buf, 4096)
    assert n > 0
    nc, nu, image_bytes = buf[0], buf[1], buf[2]
    ck = ["offset", "bytes", "layer", "rows", "row0", "wcol0", "wcols", "a_src", "a_k0", "nk", "kind"]
    uk = ["layer", "half", "n", "chunk0", "nchunks", "first", "last"]
    chunks = [dict(zip(ck, buf[3 + 11 * i: 14 + 11 * i])) for i in range(nc)]
    base = 3 + 11 * nc
    units = [dict(zip(uk, buf[base + 7 * i: base + 7 * i + 7])) for i in range(nu)]
    return chunks, units, image_bytes


def emulate(p, x_xyz, x_dir, e_xyz, e_dir):
    chunks, units, _ = get_program(e_xyz, e_dir)
    n = x_xyz.shape[0]
    emb = torch.zeros(n, 128, dtype=torch.float64)
    emb[:, :e_xyz] = x_xyz
    emb[:, 96:96 + e_dir] = x_dir
    act = torch.zeros(n, 256, dtype=torch.float64)
    acc = torch.zeros(n, 256, dtype=torch.float64)
    sigma = None
    for u in units:
        key = LAYER_KEYS[u["layer"]]
        W, b = p[key + ".weight"].double(), p[key + ".bias"].double()
        d = torch.zeros(n, u["n"], dtype=torch.float64)
        for c in chunks[u["chunk0"]: u["chunk0"] + u["nchunks"]]:
            assert c["layer"] == u["layer"] and c["rows"] == u["n"] and c["row0"] == u["half"] * 128
            assert c["kind"] == 0      # pure weight chunks; biases live in the fp32 side blob
            assert c["nk"] == (c["wcols"] + 15) // 16 and c["bytes"] == c["rows"] * 128
            kk = c["nk"] * 16
            wc = torch.zeros(c["rows"], kk, dtype=torch.float64)
            wc[:, :c["wcols"]] = W[c["row0"]: c["row0"] + c["rows"], c["wcol0"]: c["wcol0"] + c["wcols"]]
            src = emb if c["a_src"] == 0 else act
            a = src[:, c["a_k0"] * 16: c["a_k0"] * 16 + kk]
            d += a @ wc.t()
        # the layer epilogue adds the fp32 bias to the drained accumulator
        acc[:, u["half"] * 128: u["half"] * 128 + u["n"]] = d + b[u["half"] * 128: u["half"] * 128 + u["n"]]
        if u["last"]:
            width = W.shape[0]
            y = acc[:, :width]
            if u["layer"] < 8 or u["layer"] == 9:
                y = torch.relu(y)
            if u["layer"] == 7:
                sigma = torch.nn.functional.softplus(
                    y @ p["static_sigma.0.weight"].double().t() + p["static_sigma.0.bias"].double())
            if u["layer"] == 10:
                return torch.cat([torch.sigmoid(y), sigma], dim=1)
            act = torch.zeros(n, 256, dtype=torch.float64)
            act[:, :width] = y
    raise AssertionError("program has no rgb unit")


def test_program_covers_every_weight_once():
    chunks, units, image_bytes = get_program(93, 27)
    assert image_bytes == sum(c["bytes"] for c in chunks)
    in_f = [93, 256, 256, 256, 349, 256, 256, 256, 256, 283, 128]
    out_f = [256] * 9 + [128, 64]
    cover = [torch.zeros(o, i, dtype=torch.int32) for o, i in zip(out_f, in_f)]
    off = 0
    for c in chunks:
        assert c["offset"] == off and c["offset"] % 16 == 0
        off += c["bytes"]
        cover[c["layer"]][c["row0"]: c["row0"] + c["rows"], c["wcol0"]: c["wcol0"] + c["wcols"]] += 1
    for l, cv in enumerate(cover):
        assert int(cv.min()) == 1 and int(cv.max()) == 1, f"layer {l} not covered exactly once"
    assert [u["layer"] for u in units] == [l for l in range(9) for _ in range(2)] + [9, 10]


@pytest.mark.parametrize("n_freq_xyz,n_freq_dir", [(15, 4), (10, 4), (12, 2)])
def test_program_emulation_matches_oracle(n_freq_xyz, n_freq_dir):
    from models.nerf import NeRF_sigma
    from conftest import make_args
    e_xyz, e_dir = 3 + 6 * n_freq_xyz, 3 + 6 * n_freq_dir
    torch.manual_seed(3)
    m = NeRF_sigma('fine', make_args(), in_channels_xyz=e_xyz, in_channels_dir=e_dir)
    p = state(m)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, e_xyz + e_dir, generator=g, dtype=torch.float64)
    got = emulate(p, x[:, :e_xyz], x[:, e_xyz:], e_xyz, e_dir)
    want = oracle.nerf_sigma_forward({k: v.double() for k, v in p.items()}, x, e_xyz=e_xyz, e_dir=e_dir)
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-13)
