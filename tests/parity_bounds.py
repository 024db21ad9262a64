"""Tolerance model of the alpha composite, shared by tests/test_gpu_parity.py and
tools/parity_report.py.

North star: outputs within 1e-4 relative of the fp32 reference.  That bar applies directly to
what the network emits per point (features, sigma) and to ``feature_*``.  The composite
(reference models/rendering.py:121-143) then forms

    w_i = alpha_i * T_i,   alpha_i = 1 - exp(-delta_i sigma_i),   T_i = prod_{j<i} (1 - alpha_j)
                                                                      = exp(-tau_i),  tau_i = sum_{j<i} delta_j sigma_j

so a relative perturbation eps of every sigma_j moves ln T_i by up to eps*tau_i and ln alpha_i by
at most eps (x/(e^x - 1) <= 1): |dw_i| <= eps (1 + tau_i) w_i.  A sample that is still visible
behind an optical depth of 5 carries a 6x amplified relative error whatever the arithmetic - the
same holds between torch-CPU and torch-CUDA fp32 runs of the reference itself.  The tests
therefore hold

    |w - w_ref|     <= rtol (1 + tau_i) w_ref + atol_w
    |depth - d_ref| <= sum_i rtol (1 + tau_i) w_ref,i z_i + atol_d

with rtol = 1e-4 (the north-star figure), tau_i taken from the reference weights
(T_i = 1 - sum_{j<i} w_j), atol_w = 5e-6 and atol_d = 2e-5 (absolute floors for entries
near zero; depth is a sum of up to 192 products with z <= 5).
"""
import torch

ATOL_W = 5e-6
ATOL_D = 2e-5


def composite_bounds(w_ref, z, rtol=1e-4, atol_w=ATOL_W, atol_d=ATOL_D):
    """Per-entry tolerance for weights (N,S) and per-ray tolerance for depth (N,)."""
    w = w_ref.double().cpu()
    z = z.double().cpu()
    t_excl = torch.clamp(1.0 - (torch.cumsum(w, 1) - w), min=1e-30)       # T_i from the reference
    tau = -torch.log(t_excl)
    bw = rtol * (1.0 + tau) * w.abs() + atol_w
    bd = (rtol * (1.0 + tau) * w.abs() * z.abs()).sum(1) + atol_d
    return bw, bd


def rel_err(got, ref, floor):
    """max |got-ref|/|ref| over entries with |ref| > floor (0.0 if none)."""
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    m = ref.abs() > floor
    if not m.any():
        return 0.0
    return float(((got - ref).abs()[m] / ref.abs()[m]).max())
