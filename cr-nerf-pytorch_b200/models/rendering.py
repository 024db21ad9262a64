"""``models.rendering`` of the reference (models/rendering.py), backed by sm_100a kernels.

``render_rays_cross_ray`` keeps the reference's positional signature
(rendering.py:50-63) and result dictionary.  The per-chunk Python loop over
``embedding -> cat -> model -> cat`` and the ~200 small launches of the
reference collapse into, per call: one depth-sampling kernel, one fused
embed+MLP+composite kernel per pass, and one inverse-CDF+sort kernel.

Random numbers stay on the host side of the boundary: they are drawn with the
same torch calls, in the same order and on the same device as the reference
(rand_like(z) :175, randn_like(sigma) :125, rand(N,Ni) :30, randn_like :125) and
handed to the kernels, so seeded CUDA runs consume the generator identically.
"""
import torch

from crnerf_b200 import autograd as crnerf_autograd
from crnerf_b200 import ops
from crnerf_b200 import torch_ops  # noqa: F401  (registers torch.ops.crnerf.*)

_K = torch.ops.crnerf          # the library's kernels as torch operators (crnerf_b200/torch_ops.py)

__all__ = ['render_rays_cross_ray']


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5):
    """Inverse-CDF sampling, reference models/rendering.py:7-46.

    bins (N_rays, N_samples_+1), weights (N_rays, N_samples_) -> (N_rays, N_importance).
    """
    N_rays = weights.shape[0]
    if det:
        u = torch.linspace(0, 1, N_importance, device=bins.device)
    else:
        u = torch.rand(N_rays, N_importance, device=bins.device)
    return _K.sample_pdf(bins, weights.detach(), u, N_importance, eps)


def _n_freqs(embedding):
    n = getattr(embedding, 'N_freqs', None)
    return int(n if n is not None else len(embedding.freqs))


def render_rays_cross_ray(models,
                          embeddings,
                          rays,
                          ts,
                          N_samples=64,
                          use_disp=False,
                          perturb=0,
                          noise_std=1,
                          N_importance=0,
                          chunk=1024 * 32,
                          white_back=False,
                          test_time=False,
                          **kwargs):
    """Render rays: coarse pass, importance resampling, fine pass.

    Same contract as reference models/rendering.py:50-196, including autograd with respect to
    the models' parameters (crnerf_b200/autograd.py) when called with gradients enabled.  ``ts``, ``chunk``,
    ``white_back`` and ``test_time`` are accepted and unused, as in the reference
    (point chunking is unnecessary because no per-point tensor is
    materialised).  ``args.pertubeCord`` (:102-104) is honoured in inference and training.  One extension:
    ``channel_sums=True`` adds ``chansum_{typ}``, partial column sums of ``feature_{typ}`` that
    the cross-ray block uses instead of a pass over the feature map.  Returns ``weights_{typ} (N,S)``, ``feature_{typ} (N,64)``,
    ``depth_{typ} (N,)`` for the coarse model and, if ``N_importance > 0``, the
    fine model, plus the ``feature_fine_random`` alias (rendering.py:140-141).
    """
    args = kwargs['args']
    pertube = bool(getattr(args, 'pertubeCord', False))
    if not rays.is_cuda:
        raise ops.CrnerfError(f"rays are on {rays.device}: crnerf_b200 renders on CUDA (sm_100) only "
                              "and has no CPU fallback")
    coarse = models['coarse']
    n_fx, n_fd = _n_freqs(embeddings['xyz']), _n_freqs(embeddings['dir'])
    rays = rays.contiguous().float()
    N_rays = rays.shape[0]
    view_dir = kwargs.get('view_dir', None)
    dev = rays.device

    # depths: z = near*(1-t)+far*t (+ stratified jitter), rendering.py:161-176
    t_steps = torch.linspace(0, 1, N_samples, device=dev)
    perturb_rand = None
    if perturb > 0:
        perturb_rand = perturb * torch.rand(N_rays, N_samples, device=dev)
    z_vals = _K.coarse_z(rays, t_steps, perturb_rand, bool(use_disp))

    results = {}

    want_sums = bool(kwargs.get('channel_sums', False))

    def run(model, z):
        # xyz_ += pertube_ratio * torch.rand(xyz_.size()) (:102-104), drawn before the noise as there
        jitter = 0.00001 * torch.rand(z.numel(), 3, device=dev) if pertube else None
        # the reference always draws the noise tensor, even when noise_std == 0 (:125)
        noise = torch.randn(z.shape, device=dev) * noise_std
        noise = noise if noise_std != 0 else None
        typ = model.typ
        if model.wants_grad():
            # training step: same fused kernel, plus saved activations for the backward
            w, f, d = crnerf_autograd.render_pass(model, rays, z, noise, view_dir, n_fx, n_fd, jitter)
        else:
            pk = model.packed()
            pk.check_overflow()      # fp16 saturation reported by earlier passes (host read, no sync)
            w, f, d, part = _K.render_pass(pk.buf, pk.operand, rays, z, noise, view_dir, n_fx, n_fd, jitter,
                                           pk.overflow_ptr() or 0, want_sums)
            if want_sums:
                # (rows, 64) partial channel sums of feature_{typ}: rows concatenated over chunks
                # still sum to the frame's channel sums (crnerf_b200/frame.py consumes them)
                results[f'chansum_{typ}'] = part
        results[f'weights_{typ}'] = w
        results[f'feature_{typ}'] = f
        results[f'depth_{typ}'] = d
        return w, f

    w_coarse, _ = run(coarse, z_vals)

    if N_importance > 0:
        fine = models['fine']
        if perturb == 0:     # det: u = linspace(0,1,N_importance) shared by all rays (:26-28)
            u = torch.linspace(0, 1, N_importance, device=dev)
        else:                # :30
            u = torch.rand(N_rays, N_importance, device=dev)
        # sample_pdf on the coarse mid-points with weights_coarse[:,1:-1] (detached),
        # then sort(cat(z, z_new)) - one kernel (rendering.py:183-187)
        z_fine = _K.sample_pdf_merge(z_vals, w_coarse.detach(), u, N_importance, 1e-5)
        _, f_fine = run(fine, z_fine)
        if kwargs.get('output_random', True) and fine.encode_random:
            results['feature_fine_random'] = f_fine   # same tensor object, as the reference

    return results
