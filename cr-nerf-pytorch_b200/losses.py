"""Drop-in for the reference's ``losses`` module (``from losses import loss_dict``,
train_mask_grid_sample.py:22): same class names, constructor arguments, result keys and key
order; the arithmetic runs in the kernels of csrc/loss.cu - one forward launch for the
per-ray terms (c_l, f_l, r_ms, r_md), one for the embedding terms (kl_a, rec_a_random,
content_constraint), and one backward launch each - instead of ~40 elementwise/reduction
launches on (1024, 3) tensors.  CUDA fp32 tensors only; no CPU fallback.
"""
import math

from torch import nn

from crnerf_b200 import loss as _loss


class ColorLoss(nn.Module):
    """coef * (mse(rgb_coarse, targets) [+ mse(rgb_fine, targets)])  (reference losses.py:6-17)."""

    def __init__(self, coef=1):
        super().__init__()
        self.coef = coef

    def forward(self, inputs, targets):
        out = _loss.ray_loss(inputs['rgb_coarse'], inputs.get('rgb_fine'), targets, None, coef=1.0)
        # the kernel returns 0.5 * mean(...) per term; 2 * x is exact
        total = out[0] * 2 if 'rgb_fine' not in inputs else (out[0] + out[1]) * 2
        return self.coef * total


class _Schedule:
    """Weight schedules of the mask-size regulariser; ``getWeight(step)`` as in the reference."""

    def __init__(self, hi, lo):
        self.max, self.min = hi, lo


class CosineAnnealingWeight(_Schedule):
    """min + (max - min) (1 + cos(pi t / Tmax)) / 2  (reference losses.py:19-28)."""

    def __init__(self, max, min, Tmax):
        super().__init__(max, min)
        self.Tmax = Tmax

    def getWeight(self, Tcur):
        phase = math.cos(math.pi * Tcur / self.Tmax)
        return self.min + (self.max - self.min) * (1 + phase) / 2


class ExponentialAnnealingWeight(_Schedule):
    """max(min, max * exp(-k t))  (reference losses.py:30-39)."""

    def __init__(self, max, min, k):
        super().__init__(max, min)
        self.k = k

    def getWeight(self, Tcur):
        decayed = self.max * math.exp(-Tcur * self.k)
        return decayed if decayed > self.min else self.min


class CRNeRFLoss(nn.Module):
    """Reference losses.py:42-94.  ``forward(inputs, targets, hparams, global_step)`` returns
    ``(dict of loss terms, annealing weight)`` with the reference's keys in its order:
    kl_a, rec_a_random, c_l, content_constraint, r_ms, r_md, f_l (those that apply)."""

    def __init__(self, hparams, coef=1, lambda_u=0.01):
        super().__init__()
        self.coef = coef
        self.lambda_u = lambda_u
        self.Annealing = ExponentialAnnealingWeight(max=hparams.maskrs_max, min=hparams.maskrs_min,
                                                    k=hparams.maskrs_k)

    def forward(self, inputs, targets, hparams, global_step):
        weight = self.Annealing.getWeight(global_step)
        terms, names = [], []
        if 'a_embedded' in inputs:
            names.append('kl_a')
            terms.append((_loss.MODE_SQUARE, self.coef * hparams.weightKL, inputs['a_embedded'], None))
            if 'a_embedded_random_rec' in inputs:
                names.append('rec_a_random')
                mode = _loss.MODE_SQ_DIFF if hparams.mse_on_appearance else _loss.MODE_ABS_DIFF
                terms.append((mode, self.coef * hparams.weightRecA, inputs['a_embedded_random'].detach(),
                              inputs['a_embedded_random_rec']))
        if 'content_wo_a_embed' in inputs and 'content_with_a_embed' in inputs:
            names.append('content_constraint')
            terms.append((_loss.MODE_SQ_DIFF, self.coef * hparams.weightcontent, inputs['content_wo_a_embed'],
                          inputs['content_with_a_embed']))
        pair = dict(zip(names, _loss.pair_losses(terms).unbind(0))) if terms else {}

        mask = inputs.get('out_mask')
        fine = inputs.get('rgb_fine')
        if fine is None and mask is not None:
            mask = mask.detach()          # only c_l uses it, detached (reference losses.py:64)
        c_l, f_l, r_ms, r_md = _loss.ray_loss(inputs['rgb_coarse'], fine, targets, mask, coef=self.coef,
                                              size_delta=weight, digit_delta=hparams.maskrd).unbind(0)
        ret = {}
        for k in ('kl_a', 'rec_a_random'):
            if k in pair:
                ret[k] = pair[k]
        ret['c_l'] = c_l
        if 'content_constraint' in pair:
            ret['content_constraint'] = pair['content_constraint']
        if fine is not None:
            if mask is not None:
                ret['r_ms'], ret['r_md'] = r_ms, r_md
            ret['f_l'] = f_l
        return ret, weight

    def mask_regularize(self, mask, size_delta, digit_delta):
        """(r_ms, r_md) alone (reference losses.py:77-89)."""
        z = mask.detach().new_zeros((mask.numel(), 3))
        out = _loss.ray_loss(z, None, z, mask, coef=1.0, size_delta=size_delta, digit_delta=digit_delta)
        return out[2], out[3]

    def _l2_regularize(self, mu):
        return _loss.pair_losses([(_loss.MODE_SQUARE, 1.0, mu, None)])[0]


loss_dict = {'color': ColorLoss,
             'crnerf': CRNeRFLoss}
