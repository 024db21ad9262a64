"""Adam for the training step on the library's own kernel (``csrc/optim.cu``).

The reference builds ``torch.optim.Adam(parameters, lr=hparams.lr, eps=1e-8, weight_decay=...)``
(utils/__init__.py:31-32) and steps it once per batch.  In its graph-capturable form the tensor
library's Adam is ~160 launches per step for the 68 parameter tensors of this model (0.45 ms of a
3.5 ms step); :class:`Adam` is the same update as ONE launch per 48 tensors plus the step counter's
increment, capturable as is.

Drop-in for ``torch.optim.Adam`` (``amsgrad=False``): same constructor arguments and defaults, same
``param_groups`` (LR schedulers work: a float ``lr`` is read at every eager step, a tensor ``lr`` also
inside a replayed graph), and the same ``state_dict`` layout (``step`` / ``exp_avg`` / ``exp_avg_sq``
per parameter), so checkpoints move between the two in either direction.  CUDA fp32 parameters only;
anything else raises - there is no fallback update.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .ops import _stream, check


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *,
                 maximize=False, capturable=True):
        if amsgrad:
            raise NotImplementedError("crnerf_b200.optim.Adam: amsgrad is not implemented (the reference never sets it)")
        if isinstance(lr, torch.Tensor):
            if lr.numel() != 1:
                raise ValueError("a tensor lr must hold one element")
        elif not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid betas: {betas}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        # `capturable` is accepted for signature compatibility; the step is always capturable
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                        maximize=maximize, capturable=True)
        super().__init__(params, defaults)
        self._tables = {}      # group index -> (key, ctypes arrays)
        self._group_step = {}  # group index -> shared device step counter
        self._lr_keepalive = {}

    def __setstate__(self, state):
        super().__setstate__(state)       # unpickled / deep-copied: the launch-table caches are rebuilt on demand
        self._tables, self._group_step, self._lr_keepalive = {}, {}, {}

    # ---- state ---------------------------------------------------------------------------
    def _shared_step(self, gi: int, plist) -> torch.Tensor:
        """One device counter per group, shared by the `step` entries of its parameters' state (all
        parameters of a group step together).  Rebuilt when a loaded state_dict brought separate tensors."""
        shared = self._group_step.get(gi)
        if shared is not None and all(self.state[p].get("step") is shared for p in plist):
            return shared
        dev = plist[0].device
        known = [self.state[p]["step"] for p in plist if "step" in self.state[p]]
        if known:
            vals = {float(torch.as_tensor(s).item()) for s in known}
            if len(vals) != 1 or len(known) != len(plist):
                raise RuntimeError("crnerf_b200.optim.Adam: the parameters of one group carry different step counts "
                                   f"({sorted(vals)}, {len(plist) - len(known)} without state); split them into groups")
            value = vals.pop()
        else:
            value = 0.0
        shared = torch.full((), value, dtype=torch.float32, device=dev)
        for p in plist:
            self.state[p]["step"] = shared
        self._group_step[gi] = shared
        return shared

    def state_dict(self):
        """torch.optim.Adam's layout.  The live state shares ONE step counter per group; a consumer that
        increments every parameter's ``step`` (torch.optim.Adam after ``load_state_dict``, which keeps the
        tensors it is given) must receive separate tensors."""
        sd = super().state_dict()
        sd["state"] = {k: ({**v, "step": v["step"].clone()} if isinstance(v.get("step"), torch.Tensor) else v)
                       for k, v in sd["state"].items()}
        return sd

    def _table(self, gi: int, plist):
        st = [self.state[p] for p in plist]
        key = tuple((p.data_ptr(), p.grad.data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr(), p.numel())
                    for p, s in zip(plist, st))
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1]
        n = len(plist)
        arr = lambda vals, ty: (ty * n)(*vals)
        tab = (n, arr([k[0] for k in key], C.c_void_p), arr([k[1] for k in key], C.c_void_p),
               arr([k[2] for k in key], C.c_void_p), arr([k[3] for k in key], C.c_void_p),
               arr([k[4] for k in key], C.c_int64))
        self._tables[gi] = (key, tab)
        return tab

    # ---- update --------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad"):
                raise NotImplementedError("crnerf_b200.optim.Adam: amsgrad is not implemented")
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            dev = plist[0].device
            for p in plist:
                if p.grad.is_sparse:
                    raise RuntimeError("crnerf_b200.optim.Adam does not support sparse gradients")
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.device == dev):
                    raise RuntimeError("crnerf_b200.optim.Adam needs contiguous CUDA fp32 parameters on one device per "
                                       f"group (got {p.dtype} on {p.device}); there is no fallback update")
                if p.grad.dtype != torch.float32 or p.grad.device != dev:
                    raise RuntimeError("crnerf_b200.optim.Adam needs fp32 gradients on the parameters' device")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                s = self.state[p]
                if "exp_avg" not in s:
                    s["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    s["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                elif not (s["exp_avg"].is_cuda and s["exp_avg"].is_contiguous() and s["exp_avg_sq"].is_contiguous()):
                    s["exp_avg"] = s["exp_avg"].to(dev, torch.float32).contiguous()
                    s["exp_avg_sq"] = s["exp_avg_sq"].to(dev, torch.float32).contiguous()
            step_t = self._shared_step(gi, plist)
            n, pp, gp, mp, vp, nn = self._table(gi, plist)
            lr = group["lr"]
            lr_dev: Optional[torch.Tensor] = None
            if isinstance(lr, torch.Tensor):
                lr_dev = lr if (lr.is_cuda and lr.dtype == torch.float32) else lr.to(dev, torch.float32)
                self._lr_keepalive[gi] = lr_dev   # the launch reads it asynchronously
                lr = 0.0
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                check(lib.crnerf_adam_step(n, pp, gp, mp, vp, nn, step_t.data_ptr(),
                                           lr_dev.data_ptr() if lr_dev is not None else None, float(lr), float(b1),
                                           float(b2), float(group["eps"]), float(group["weight_decay"]),
                                           int(bool(group.get("maximize", False))), _stream(dev)))
                step_t.add_(1.0)
        return loss
