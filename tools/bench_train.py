#!/usr/bin/env python
"""Training-step timing (BASELINE configs[4] shapes): 1024-ray (32x32) patch, 64+64 samples,
perturb=1, noise_std=1, render under autograd -> style_net decode (coarse, fine) -> MSE ->
backward -> Adam.  Prints our step next to the same step on stock PyTorch eager ops (the
oracle's restatement of the reference math moved to the same GPU) as the library baseline.

  python tools/bench_train.py [--steps 20]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py   (DDP, 1 patch / rank)
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)
import torch
from crnerf_b200 import synthetic
import torch.distributed as dist


OPTIMIZER = os.environ.get("CRNERF_BENCH_OPTIMIZER", "native")   # "native": crnerf_b200.optim.Adam, "torch": torch.optim.Adam


def make_step(dev, world, rank, n_rays=1024, ns=64, ni=64, operand="fp16", bwd=None, capturable=False):
    """Build the models + optimizer and return (step_fn, models, margs, rays, style, target, side)."""
    from bench import build_models
    from models.nerf import PosEmbedding
    from models.rendering import render_rays_cross_ray
    from crnerf_b200 import autograd as ag
    if bwd is not None:
        ag.BACKWARD_MATMUL = bwd
    models, margs = build_models()
    for k in ("coarse", "fine"):
        models[k].operand = operand
    models = {k: m.to(dev).train() for k, m in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    side = int(n_rays ** 0.5)
    rays = synthetic.pinhole_rays(side, side, synthetic.synthetic_pose(rank)).to(dev)
    style = torch.rand(1, 64, 32, 32, device=dev)
    target = torch.rand(side * side, 3, device=dev)
    params = [p for m in models.values() for p in m.parameters()]
    if OPTIMIZER == "native":
        from crnerf_b200.optim import Adam      # one launch per 48 tensors (csrc/optim.cu), always capturable
        opt = Adam(params, lr=5e-4)
    else:
        opt = torch.optim.Adam(params, lr=5e-4, capturable=capturable)
    flat_n = sum(p.numel() for p in params)

    def allreduce_grads():
        if world > 1:   # one in-place all-reduce per gradient storage (24 for the 68 tensors)
            from crnerf_b200.ddp import allreduce_gradients
            allreduce_gradients(params)

    def loss_fn():
        res = render_rays_cross_ray(models, emb, rays, None, ns, False, 1.0, 1.0, ni, 32768, False, args=margs)
        loss = 0
        for typ in ("coarse", "fine"):
            feat = res[f"feature_{typ}"].t().reshape(1, 64, side, side)
            rgb = models["decoder"](feat, style).reshape(3, -1).t()
            loss = loss + 0.5 * ((rgb - target) ** 2).mean()
        return loss

    def step_ours():
        loss = loss_fn()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        allreduce_grads()
        opt.step()
        return loss

    step_ours.loss_fn, step_ours.opt, step_ours.params = loss_fn, opt, params
    return step_ours, models, margs, rays, style, target, side, flat_n


def measure(dev, world, rank, barrier, reps=10, warm=3, n_rays=1024, ns=64, ni=64):
    """bench.py's training leg: the step in both operand formats (fp16 = the 1e-4 parity mode,
    bf16 = the configuration BASELINE configs[4] names), one patch per rank, gradient all-reduce.
    Device-timed with CUDA events, max over ranks."""
    from crnerf_b200 import ops
    out = {}
    for operand in ("fp16", "bf16"):
        from crnerf_b200 import autograd as ag
        default_bwd = getattr(ag, "DEFAULT_BACKWARD_MATMUL", ag.BACKWARD_MATMUL)
        step, *_rest, flat_n = make_step(dev, world, rank, n_rays, ns, ni, operand,
                                         default_bwd if (operand == "fp16" or default_bwd == "native") else "bf16")
        for _ in range(warm):
            step()
        barrier()
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            loss = step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out[f"train_step_ms_{operand}"] = ms
        out[f"train_ray_samples_per_s_{operand}"] = world * n_rays * (ns + ni) / (ms * 1e-3)
        out[f"train_native_launches_per_step_{operand}"] = (ops.launch_count() - n0) / reps
        out[f"train_loss_{operand}"] = float(loss)
        if world > 1:
            # forward + backward replayed as one CUDA graph per rank, gradient exchange and optimizer outside it
            try:
                from crnerf_b200.graphs import GraphedTrainStep
                from crnerf_b200.ddp import allreduce_gradients
                gstep, *_r = make_step(dev, world, rank, n_rays, ns, ni, operand, None, capturable=True)
                graphed = GraphedTrainStep(gstep.loss_fn, optimizer=None, parameters=gstep.params)

                def dstep():
                    gl = graphed()
                    allreduce_gradients(gstep.params)
                    gstep.opt.step()
                    return gl
                for _ in range(warm):
                    dstep()
                barrier()
                e0.record()
                for _ in range(reps):
                    gl = dstep()
                e1.record()
                barrier()
                t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                out[f"train_step_ms_{operand}_graphed"] = float(t.item())
                out[f"train_ray_samples_per_s_{operand}_graphed"] = world * n_rays * (ns + ni) / (float(t.item()) * 1e-3)
                out[f"train_loss_{operand}_graphed"] = float(gl)
            except Exception as e:   # noqa: BLE001
                out[f"train_graph_error_{operand}"] = f"{type(e).__name__}: {e}"[:300]
        if world == 1:
            # the same step replayed as one CUDA graph (crnerf_b200.graphs.GraphedTrainStep): the eager
            # step is bound by ~1,600 host-side tensor-library calls, not by the GPU
            try:
                from crnerf_b200.graphs import GraphedTrainStep
                gstep, *_r = make_step(dev, world, rank, n_rays, ns, ni, operand, None, capturable=True)
                graphed = GraphedTrainStep(gstep.loss_fn, gstep.opt)
                for _ in range(warm):
                    graphed()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(reps):
                    gl = graphed()
                e1.record()
                torch.cuda.synchronize()
                out[f"train_step_ms_{operand}_graphed"] = e0.elapsed_time(e1) / reps
                out[f"train_loss_{operand}_graphed"] = float(gl)
            except Exception as e:   # noqa: BLE001
                out[f"train_graph_error_{operand}"] = f"{type(e).__name__}: {e}"[:300]
    out["train"] = {"workload": f"{n_rays}-ray (32x32) patch per rank x ({ns}+{ni}) samples, perturb=1, noise_std=1, "
                                "style_net decode of coarse and fine, MSE, backward, Adam",
                    "parallelism": f"data parallel x{world}, {flat_n * 4} B of gradients per step exchanged as one in-place all-reduce per "
                                   "gradient storage (crnerf_b200.ddp: 24 collectives for 68 tensors); _graphed = forward + backward as one CUDA graph per rank, "
                                   "exchange and optimizer outside it",
                    "optimizer": "crnerf_b200.optim.Adam (csrc/optim.cu, torch.optim.Adam's update as one launch per 48 tensors)"
                                 if OPTIMIZER == "native" else "torch.optim.Adam",
                    "reps": reps, "warmup": warm, "timing": "CUDA events over the reps, max over ranks"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--ns", type=int, default=64)
    ap.add_argument("--ni", type=int, default=64)
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--operand", default="fp16", choices=["fp16", "bf16"], help="tensor-core operand format of the MLP")
    ap.add_argument("--bwd", default=None, choices=["native"], help="(kept for old command lines)")
    ap.add_argument("--graph", action="store_true", help="also time the step replayed as one CUDA graph")
    a = ap.parse_args()
    import crnerf_oracle as oracle
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    step_ours, models, margs, rays, style, target, side, _ = make_step(dev, world, rank, a.rays, a.ns, a.ni,
                                                                    a.operand, a.bwd)
    from crnerf_b200 import autograd as ag
    # library baseline: the same math as plain differentiable torch ops on the GPU
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in models["coarse"].state_dict().items()}
    pf = {k: v.detach().clone().requires_grad_(True) for k, v in models["fine"].state_dict().items()}
    pd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k.find("rgb_upsample") < 0) for k, v in models["decoder"].state_dict().items()}
    eparams = [v for d_ in (pc, pf, pd) for v in d_.values() if v.requires_grad]
    eopt = torch.optim.Adam(eparams, lr=5e-4)

    def step_eager():
        n = rays.shape[0]
        rng = {"perturb_rand": torch.rand(n, a.ns, device=dev), "noise_coarse": torch.randn(n, a.ns, device=dev),
               "u": torch.rand(n, a.ni, device=dev), "noise_fine": torch.randn(n, a.ns + a.ni, device=dev)}
        res = oracle.render_rays(pc, pf, rays, n_samples=a.ns, n_importance=a.ni, perturb=1.0, noise_std=1.0,
                                 chunk=1 << 30, rng=rng)
        loss = 0
        for typ in ("coarse", "fine"):
            feat = res[f"feature_{typ}"].t().reshape(1, 64, side, side)
            rgb = oracle.style_net_forward(pd, feat, style).reshape(3, -1).t()
            loss = loss + 0.5 * ((rgb - target) ** 2).mean()
        eopt.zero_grad(set_to_none=True)
        loss.backward()
        eopt.step()
        return loss

    def timeit(fn, steps):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(steps): l = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps, float(l)

    ours_dev, ours_wall, l1 = timeit(step_ours, a.steps)
    graphed_ms = None
    if a.graph and world == 1:
        from crnerf_b200.graphs import GraphedTrainStep
        gstep, *_r = make_step(dev, world, rank, a.rays, a.ns, a.ni, a.operand, None, capturable=True)
        graphed = GraphedTrainStep(gstep.loss_fn, gstep.opt)
        graphed_ms, _, lg = timeit(graphed, a.steps)
    out = {"workload": f"train step, {a.rays} rays x ({a.ns}+{a.ni}), perturb=1 noise=1, style_net decode x2, MSE, Adam",
           "operand": a.operand, "backward_gemm": ag.BACKWARD_MATMUL,
           "n_gpus": world, "ours_ms_device": ours_dev, "ours_ms_wall": ours_wall,
           "ours_ray_samples_per_s": world * a.rays * (a.ns + a.ni) / (ours_wall * 1e-3), "loss": l1,
           "ours_ms_graphed": graphed_ms}
    if not a.no_eager and world == 1:
        try:
            eg_dev, eg_wall, l2 = timeit(step_eager, max(3, a.steps // 4))
            out.update({"torch_eager_fp32_ms_device": eg_dev, "torch_eager_fp32_ms_wall": eg_wall,
                        "speedup_vs_torch_eager": eg_wall / ours_wall})
        except Exception as e:   # noqa: BLE001
            out["torch_eager_error"] = repr(e)[:200]
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
