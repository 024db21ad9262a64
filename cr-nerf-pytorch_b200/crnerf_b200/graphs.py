"""CUDA-graph replay of the eval-mode render call.

One eval batch is ~12 launches (depth grid, coarse pass, inverse-CDF + sort, fine pass and the
torch helpers that keep the reference's RNG consumption); at 4096 rays x (64+128) the kernels
take ~1.15 ms and the launch gaps between them ~50 us.  ``GraphedRenderer`` captures
``render_rays_cross_ray`` once for a fixed batch shape and replays the graph: same kernels,
same results, one launch.  (``torch.cuda.graph`` captures the library's launches because every
C-ABI call enqueues on ``torch.cuda.current_stream()`` and never synchronises or allocates.)
"""
from __future__ import annotations

import torch

__all__ = ["GraphedRenderer", "GraphedTrainStep"]


class GraphedRenderer:
    """``renderer(rays) -> results`` for a fixed ``(n_rays, N_samples, N_importance)``, eval mode
    (``perturb=0, noise_std=0`` as eval.py:46-47).  The returned tensors are the graph's static
    outputs: they are overwritten by the next call (copy them if they must outlive it)."""

    def __init__(self, models, embeddings, n_rays, N_samples, N_importance, use_disp=False,
                 device=None, **kwargs):
        from models.rendering import render_rays_cross_ray
        dev = torch.device(device) if device is not None else next(models["coarse"].parameters()).device
        if dev.type != "cuda":
            raise ValueError("GraphedRenderer needs CUDA models")
        self.n_rays = n_rays
        self.rays = torch.zeros(n_rays, 8, device=dev)
        self.rays[:, 3:6] = torch.tensor([0.0, 0.0, -1.0], device=dev)
        self.rays[:, 7] = 1.0

        self._models = models
        self._keys = ["coarse"] + (["fine"] if N_importance > 0 else [])
        self._dev = dev

        def run():
            with torch.no_grad():
                return render_rays_cross_ray(models, embeddings, self.rays, None, N_samples, use_disp, 0, 0,
                                             N_importance, n_rays, False, test_time=True, **kwargs)

        self._run = run
        self._capture()

    def _capture(self):
        """(Re)capture.  The graph bakes in the address of each model's packed weight image, so the
        PackedMLP objects it was captured with are kept alive here and compared on every call."""
        dev, run = self._dev, self._run
        # warm up on a side stream (weight packing, allocator) before capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        from . import ops
        with torch.cuda.stream(side):
            run()
            n0 = ops.launch_count()
            run()
            self.kernels_per_replay = ops.launch_count() - n0   # library kernels inside one replay
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            self._packed = [self._models[k].packed() for k in self._keys]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.results = run()
        with torch.no_grad():
            if any(self._models[k].packed() is not p for k, p in zip(self._keys, self._packed)):
                raise RuntimeError("the weight image changed during graph capture")
        self.captures = getattr(self, "captures", 0) + 1

    def __call__(self, rays: torch.Tensor, non_blocking: bool = True):
        if rays.shape != self.rays.shape:
            raise ValueError(f"this graph renders batches of {tuple(self.rays.shape)}, got {tuple(rays.shape)}")
        # weights re-packed since the capture (optimizer step, load_state_dict, invalidate_packed):
        # the old image may be freed - capture again against the new one
        with torch.no_grad():        # inference-mode packed(): cached, keyed on the weight version
            stale = any(self._models[k].packed() is not p for k, p in zip(self._keys, self._packed))
        if stale:
            self._capture()
        for p in self._packed:
            p.check_overflow()       # fp16 saturation reported by earlier replays (host read, no sync)
        self.rays.copy_(rays, non_blocking=non_blocking)     # H2D or D2D into the static input
        self.graph.replay()
        return self.results


class GraphedTrainStep:
    """One whole training step - render under autograd, decode, loss, backward, optimizer - captured
    in a CUDA graph and replayed (the training step of train_mask_grid_sample.py:268-337 issues
    ~340 kernels through ~1,600 tensor-library calls; at 1,024-ray patches the host, not the GPU,
    bounds it).  Static shapes: ``step_fn()`` must read its inputs from fixed tensors (update them
    in place between replays) and return the loss tensor; it must NOT call ``backward`` or the
    optimizer - this class does.  The optimizer must be capturable
    (``torch.optim.Adam(..., capturable=True)``).  Random draws inside ``step_fn`` (stratified
    jitter, density noise) advance with every replay, as torch's graph-safe generator does.

    Multi-GPU: capture forward + backward per rank (``optimizer=None, parameters=<the trained parameters>``),
    replay, exchange the gradients (``crnerf_b200.ddp.allreduce_gradients``: one all-reduce per flat
    gradient buffer) and step the optimizer outside.  ``parameters`` are needed in that form so that their
    gradients are unset before the capture: the captured backward then WRITES static gradient tensors on
    every replay instead of accumulating into whatever the warm-up left."""

    def __init__(self, step_fn, optimizer=None, warmup: int = 3, device=None, parameters=None):
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.optimizer = optimizer
        if optimizer is None and parameters is None:
            raise ValueError("GraphedTrainStep(optimizer=None) needs parameters=: their gradients must be unset "
                             "before the capture, or every replay would accumulate into the warm-up's")
        self.parameters = list(parameters) if parameters is not None else None

        def unset_grads():
            if optimizer is not None:
                optimizer.zero_grad(set_to_none=True)
            if self.parameters is not None:
                for p in self.parameters:
                    p.grad = None

        def whole():
            loss = step_fn()
            loss.backward()
            if optimizer is not None:
                optimizer.step()
            return loss

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):           # eager steps: allocator warm-up, weight-range verdicts
                unset_grads()
                whole()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        unset_grads()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = whole()

    def __call__(self):
        self.graph.replay()
        return self.loss
