"""Context_Guided_Network mirror (reference models/lightweight_seg.py:271-368,
train_mask_grid_sample.py:114,170-176).  Golden vectors come from the unmodified reference file
(oracle/make_golden.py::case_cgnet).

CPU: state_dict keys / order / shapes and seeded init equal the reference's; the oracle's functional
restatement and the mirror module reproduce the golden forwards bit for bit.
GPU: the mirror on cuDNN (fp32, TF32 off) against the golden forwards and gradients at 1e-4, the
fused resize+gather tail, and the network inside a captured training step."""
import pytest
import torch

import crnerf_oracle as oracle
from conftest import build_mirror_cgnet, check_checksums, load_golden

REL = dict(rtol=1e-4, atol=2e-6)      # north-star bar for floating point


def test_state_dict_and_seeded_init_equal_the_reference():
    g = load_golden("cgnet")
    net, _ = build_mirror_cgnet(g)
    sd = net.state_dict()
    assert list(sd) == list(g["shapes"])                       # same keys in the same order
    assert all(tuple(sd[k].shape) == g["shapes"][k] for k in sd)
    check_checksums(net, g["checksum"])                        # same values: default init + kaiming pass + ruffle


def _same(a, b):
    """Bit-identical on the machine that made the goldens (oracle/make_golden.py asserts torch.equal
    there); oneDNN partitions convolutions by thread count, so another host may differ in the last
    bits - a few ulp are allowed, nothing more."""
    return torch.equal(a, b) or torch.allclose(a, b, rtol=2e-6, atol=2e-7)


def test_oracle_and_mirror_forward_match_reference_on_cpu():
    g = load_golden("cgnet")
    net, _ = build_mirror_cgnet(g)
    state = {k: v.clone() for k, v in net.state_dict().items()}
    for c in g["cases"]:
        with torch.no_grad():
            assert _same(oracle.cgnet_forward(state, c["x"], 2, 2, train=False), c["eval"])
            assert _same(oracle.cgnet_forward(state, c["x"], 2, 2, train=True), c["train"])
            assert _same(net.eval()(c["x"]), c["eval"])
        net.load_state_dict(state)
        net.train()
        y = net(c["x"])
        assert _same(y.detach(), c["train"])
        for k, v in net.named_buffers():                       # BatchNorm running statistics advance as there
            if "running" in k:
                assert _same(v, c["running_after"][k]), k
        net.load_state_dict(state)


@pytest.mark.gpu
def test_mirror_on_gpu_matches_reference_forward_and_gradients():
    g = load_golden("cgnet")
    net, _ = build_mirror_cgnet(g)
    net = net.cuda()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    for c in g["cases"]:
        x = c["x"].cuda()
        with torch.no_grad():
            torch.testing.assert_close(net.eval()(x).cpu(), c["eval"], **REL)
        net.load_state_dict(state)
        net.train()
        net.zero_grad()
        rows = net.mask_rows(x, c["hw"], None if c["idx"] is None else c["idx"].cuda())
        torch.testing.assert_close(rows.detach().cpu(), c["rows"], **REL)
        (rows * c["g_rows"].cuda()).sum().backward()
        for k, p in net.named_parameters():
            s, a = c["grad_sums"][k]
            got = p.grad.double().cpu()
            # gradients pass through batch-statistics BatchNorm (B=1) of a 20-layer stack: compared per
            # tensor in relative L2 / by their sums, the fp32 summation order being cuDNN's
            if c["grads"] is not None:
                ref = c["grads"][k].double()
                assert float((got - ref).norm()) <= 2e-4 * float(ref.norm()) + 1e-7, k
            assert abs(float(got.abs().sum()) - a) <= 2e-4 * a + 1e-6, k
        net.load_state_dict(state)


@pytest.mark.gpu
def test_mask_network_replays_inside_a_captured_training_step():
    """The survey's bar for this row: library convolutions under a CUDA graph.  forward -> fused
    resize+gather tail -> loss -> backward -> Adam, captured once and replayed; replays follow the
    eager trajectory."""
    from crnerf_b200.graphs import GraphedTrainStep
    g = load_golden("cgnet")
    c = g["cases"][0]

    def make():
        net, _ = build_mirror_cgnet(g)
        net = net.cuda().train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)
        return net, opt

    x = c["x"].cuda()
    idx = c["idx"].cuda()
    target = torch.rand(idx.numel(), 1, generator=torch.Generator().manual_seed(3)).cuda()

    def losses(graphed, steps=8):
        net, opt = make()
        step_fn = lambda: ((net.mask_rows(x, c["hw"], idx) - target) ** 2).mean()
        out = []
        if graphed:
            gs = GraphedTrainStep(step_fn, opt, warmup=3)
            for _ in range(steps):
                out.append(float(gs().detach()))
        else:
            for i in range(steps + 3):
                opt.zero_grad(set_to_none=True)
                loss = step_fn()
                loss.backward()
                opt.step()
                if i >= 3:
                    out.append(float(loss.detach()))
        return out

    eager, replay = losses(False), losses(True)
    assert replay[-1] < replay[0]
    # GraphedTrainStep runs 3 eager warm-up steps before capture; the captured step is step 4
    assert all(abs(a - b) <= 5e-3 * abs(a) for a, b in zip(eager, replay)), (eager, replay)
