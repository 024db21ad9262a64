#!/usr/bin/env python
"""bench.py - ray-samples/s of the CR-NeRF volume-rendering hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one eval-mode pass of ``render_rays_cross_ray`` over a batch of
4096 synthetic Phototourism-shaped rays with 64 coarse + 128 fine samples
(coarse pass -> inverse-CDF resampling -> fine pass), the configuration
BASELINE.json's metric is quoted on.  One "ray-sample" = one of the
4096*(64+128) = 786,432 fine-pass sample slots; each step
evaluates the MLP at 4096*(64+192) = 1,048,576 points = 1.29306 TFLOP.

  value : device-timed (CUDA events, max over ranks), inputs resident in HBM, L2
          flushed between timed steps.
  e2e   : same call through the public API with the rays in pinned HOST memory:
          H2D copy + render + D2H read of the result inside the wall-clock region,
          two steps in flight (the host waits for step i-1's result while step i runs).
  roofline : the dominant kernel (fused fine pass) timed alone, algorithmic
          FLOPs / duration against the measured bf16 peak of MEASURED_PEAKS.json.
  cpu_baseline : the oracle (torch-CPU port of the reference path) on the host cores.

  sustained : >= 2 s of back-to-back steps (no L2 flush: the step's working set is the
          1.3 MB weight images + a 131 KB ray batch), clocks sampled during the loop, against
          MEASURED_PEAKS.json's bf16_tflops_sustained.
  frame     : BASELINE configs[3] - ONE 800x800 frame (640,000 rays x (64+128)), rays built on
          the GPU, ray rows sharded across the N ranks, cross-ray statistics combined with two
          small all-reduces, rgb blocks with one all-gather (crnerf_b200/frame.py, scheme
          "stats"); strong scaling; ``frame_check`` = max |rgb| difference between the N-rank
          frame and rank 0 rendering the same frame alone through the unsharded style_net.
  train     : BASELINE configs[4] - one training step of the reference's call pattern
          (1024-ray patch x (64+64), perturb = noise = 1, decode x2, MSE, backward, Adam), one
          patch per rank, DDP gradient all-reduce.

``--impl reference`` times the reference's own CPU implementation of the path: the
UNMODIFIED ``models/rendering.py`` + ``models/nerf.py`` staged by ``build()`` under the
git-ignored ``oracle/_ref/`` (``kind: "reference"``); if they are absent, the pinned port
``oracle/crnerf_oracle.py`` (``kind: "port"``, bit-exact against them - oracle/make_golden.py).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

N_RAYS, NS, NI = 4096, 64, 128
FLOP_PER_POINT = 1233152                      # 2 * 616,576 MAC
RAY_SAMPLES_PER_STEP = N_RAYS * (NS + NI)     # 786,432
POINTS_PER_STEP = N_RAYS * (NS + NS + NI)     # 1,048,576
METRIC = "ray-samples/sec at 4096 rays x (64+128) samples"
UNIT = "ray-samples/s"
WORKLOAD = ("configs[1]: 4096-ray eval batches (slices of a 320x256 synthetic Brandenburg-Gate-shaped "
            "frame), 64 coarse + 128 fine samples, N_emb_xyz=15, N_emb_dir=4, nerf_out_dim=64, "
            "default-init weights seed 0")


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import crnerf_oracle
    return crnerf_oracle


class CpuPath:
    """The CPU arm: the reference's own files when staged (oracle/_ref), else the pinned port.
    bench.py's CPU legs are the only place outside tests/ and smoke() that touches oracle/."""

    def __init__(self, state):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_loader
        self.kind, self._ref, self._oracle, self._state = "port", None, None, state
        if ref_loader.find_reference() is not None:
            try:
                self._ref = ref_loader.ReferenceRenderer(state[0], state[1])
                self.kind = "reference"
            except Exception as e:   # noqa: BLE001  (e.g. einops missing on the box)
                print(f"[bench] staged reference unusable ({type(e).__name__}: {e}); using the port",
                      file=sys.stderr)
        if self._ref is None:
            self._oracle = load_oracle()

    def render(self, rays):
        with torch.no_grad():
            if self._ref is not None:
                return self._ref.render_rays(rays, NS, NI, 8192)
            return self._oracle.render_rays(self._state[0], self._state[1], rays, n_samples=NS,
                                            n_importance=NI, perturb=0, noise_std=0, chunk=8192)


def make_args():
    return types.SimpleNamespace(nerf_out_dim=64, pertubeCord=False, img_wh=[320, 256])


def build_models(seed=0):
    from models.nerf import NeRF_sigma
    from models.linearStyleTransfer import style_net
    torch.manual_seed(seed)
    args = make_args()
    coarse = NeRF_sigma('coarse', args, in_channels_xyz=93, in_channels_dir=27)
    decoder = style_net(args)
    fine = NeRF_sigma('fine', args, in_channels_xyz=93, in_channels_dir=27, encode_appearance=True,
                      in_channels_a=48, encode_random=True)
    return {"coarse": coarse.eval(), "fine": fine.eval(), "decoder": decoder.eval()}, args


def frame_rays():
    """A 320x256 frame = 20 batches of 4096 rays, fov 60 deg, near 0 / far 5."""
    from crnerf_b200.synthetic import pinhole_rays, synthetic_pose
    return pinhole_rays(256, 320, synthetic_pose(0), 0.0, 5.0)


def cpu_state(models):
    sd = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    return sd(models["coarse"]), sd(models["fine"])


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host core
    (only rank 0 runs it, so there is no oversubscription)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_reference_throughput(cpu, rays, reps, warm=1):
    """ray-samples/s of the CPU arm on all host threads; each rep = one 4096-ray batch."""
    use_all_host_threads()
    times = []
    for i in range(warm + reps):
        batch = rays[(i % 20) * N_RAYS:(i % 20 + 1) * N_RAYS]
        t0 = time.perf_counter()
        cpu.render(batch)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return RAY_SAMPLES_PER_STEP / statistics.mean(times), times


def psnr_delta(oracle, models_gpu, state_cpu, emb, margs, dev):
    """'PSNR delta vs ref' half of the metric: one 64x64 frame (4096 rays, 64+128
    samples) rendered + decoded by the CUDA path and by the CPU port of the reference; both are
    scored against a pseudo ground truth T (the reference render decoded with a differently seeded
    style feature) on the right half of the image, as eval_metric.py:89-93 does."""
    from models.rendering import render_rays_cross_ray
    h = w = 64
    from crnerf_b200.synthetic import pinhole_rays, synthetic_pose
    rays = pinhole_rays(h, w, synthetic_pose(0), 0.0, 5.0)
    g = torch.Generator().manual_seed(1)
    style = torch.rand(1, 64, 32, 32, generator=g)
    style_t = torch.rand(1, 64, 32, 32, generator=g)
    pd = {k: v.detach().cpu().clone() for k, v in models_gpu["decoder"].state_dict().items()}
    with torch.no_grad():
        ref = oracle.render_rays(state_cpu[0], state_cpu[1], rays, n_samples=NS, n_importance=NI, perturb=0,
                                 noise_std=0, chunk=8192)
        feat_ref = ref["feature_fine"].t().reshape(1, 64, h, w)
        rgb_ref = oracle.style_net_forward(pd, feat_ref, style)
        rgb_t = oracle.style_net_forward(pd, feat_ref, style_t)
        res = render_rays_cross_ray(models_gpu, emb, rays.to(dev), None, NS, False, 0, 0, NI, 32768, False,
                                    test_time=True, args=margs)
        rgb = models_gpu["decoder"](res["feature_fine"].t().reshape(1, 64, h, w), style.to(dev)).cpu()
    half = lambda t: t[..., w // 2:]
    p_ours, p_ref = oracle.psnr(half(rgb), half(rgb_t)), oracle.psnr(half(rgb_ref), half(rgb_t))
    return {"psnr_ours_vs_ref_db": oracle.psnr(rgb, rgb_ref), "psnr_ours_vs_T_db": p_ours,
            "psnr_ref_vs_T_db": p_ref, "psnr_delta_db": abs(p_ours - p_ref),
            "frame": "64x64, 64+128 samples, style_net decode, right half scored"}


def psnr_trained(emb, margs, dev):
    """The same metric on the TRAINED-like weight set (tests/golden/trained.pt: weights, rays and the
    decoded frames rgb_a / rgb_t produced by the UNMODIFIED reference on CPU, oracle/make_trained.py) -
    the regime the metric is meant for: both PSNRs against T land in 15-30 dB.  Nothing of oracle/
    runs here; the reference side is the stored frame."""
    from models.rendering import render_rays_cross_ray
    path = os.path.join(ROOT, "tests", "golden", "trained.pt")
    if not os.path.isfile(path):
        return None
    t = torch.load(path, map_location="cpu", weights_only=False)
    models, _ = build_models(t["seed"])
    models["coarse"].load_state_dict(t["coarse"], strict=True)
    models["fine"].load_state_dict(t["fine"], strict=True)
    models["decoder"].load_state_dict(t["decoder"], strict=False)     # its two seeded-default fc layers are not stored
    models = {k: m.to(dev).eval() for k, m in models.items()}
    f = t["cases"]["frame"]
    h, w = f["hw"]
    half = lambda x: x[..., w // 2:]
    psnr = lambda a, b: float(-10.0 * torch.log10(torch.mean((a.double() - b.double()) ** 2)))
    out = {"frame": f"{h}x{w}, 64+128 samples, trained-like weights, style_net decode, right half scored; "
                    "reference side = frames stored by the unmodified reference (CPU fp32)",
           "psnr_ref_vs_T_db": psnr(half(f["rgb_a"]), half(f["rgb_t"]))}
    for operand in ("fp16", "fp16x3"):
        models["coarse"].operand = models["fine"].operand = operand
        with torch.no_grad():
            res = render_rays_cross_ray(models, emb, f["rays"].to(dev), None, 64, False, 0, 0, 128, 32768, False,
                                        test_time=True, args=margs)
            rgb = models["decoder"](res["feature_fine"].t().reshape(1, 64, h, w), f["style_a"].to(dev)).cpu()
        p = psnr(half(rgb), half(f["rgb_t"]))
        out[operand] = {"psnr_ours_vs_T_db": p, "psnr_delta_db": abs(p - out["psnr_ref_vs_T_db"]),
                        "psnr_ours_vs_ref_db": psnr(rgb, f["rgb_a"])}
    return out


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML is
    polled from a thread every ~2 ms (the timed region of the default run is ~25 ms, too short for
    `nvidia-smi -lms`); falls back to one `nvidia-smi` query if the NVML binding is missing."""

    def __init__(self, index, period=0.002):
        self.index, self.rows, self._stop, self._thr, self._h = index, [], False, None, None
        self.period = period
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:   # noqa: BLE001
            self._nv = None

    def _poll(self):
        nv = self._nv
        while not self._stop:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)))
            except Exception:   # noqa: BLE001
                break
            time.sleep(self.period)

    def start(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()

    def stop(self):
        if self._nv is None:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=20).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1,
                        "note": "NVML binding unavailable: one nvidia-smi sample after the timed region"}
            except Exception:   # noqa: BLE001
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self._stop = True
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        nv = self._nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r in self.rows))
        sm = [c for c, _ in self.rows]
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM))
        except Exception:   # noqa: BLE001
            mx = None
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(sm)}


def shared_config():
    """`config` is byte-identical in both arms (same workload); per-arm details go to `detail`."""
    return {"workload": WORKLOAD, "rays_per_step": N_RAYS, "n_coarse": NS, "n_fine": NI}


def traffic_record():
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (bench.py cannot run ncu on itself).  profiles/ncu_traffic.json also stores a hash of the
    kernel's sources at capture time: a mismatch with today's sources flags the number as stale
    instead of letting a regression hide behind a constant."""
    import hashlib
    rec_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(rec_path))
    except (OSError, ValueError):
        return None, {"traffic_source": "no committed capture"}
    h = hashlib.sha256()
    for rel in rec.get("sources", []):
        try:
            h.update(open(os.path.join(ROOT, rel), "rb").read())
        except OSError:
            h.update(b"missing")
    fresh = h.hexdigest() == rec.get("sources_sha256")
    return rec.get("dram_bytes_per_launch"), {
        "traffic_unit": "bytes/launch", "traffic_source": rec.get("capture"),
        "traffic_matches_current_sources": fresh}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    models, _ = build_models()
    rays = frame_rays()
    cores = use_all_host_threads()
    cpu = CpuPath(cpu_state(models))      # the reference arm IS the CPU path
    warm = max(1, args.warmup)
    value, times = cpu_reference_throughput(cpu, rays, reps=max(1, args.steps), warm=warm)
    sample = f"{len(times)} x one 4096-ray batch (786,432 ray-samples each) on {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": warm,
        "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(),
        "detail": {"device": "cpu", "chunk": 8192, "host_cpus": os.cpu_count(),
                   "code": ("unmodified reference models/rendering.py + models/nerf.py (oracle/_ref)"
                            if cpu.kind == "reference" else "oracle/crnerf_oracle.py (pinned port)")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": cpu.kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def frame_leg(models_gpu, emb, margs, dev, world, rank, barrier, reps=3, warm=2):
    """BASELINE configs[3]: one 800x800 frame sharded over the ranks (see module docstring)."""
    import math
    import torch.distributed as dist
    from crnerf_b200 import ops
    from crnerf_b200.frame import batched_render, render_frame_sharded
    from crnerf_b200.synthetic import synthetic_pose
    h = w = 800
    fl = 0.5 * w / math.tan(math.radians(30.0))
    K = [[fl, 0.0, w / 2], [0.0, fl, h / 2], [0.0, 0.0, 1.0]]
    pose = synthetic_pose(0).tolist()
    camera = (K, pose, 0.0, 5.0)
    style = torch.rand(1, 64, 32, 32, generator=torch.Generator().manual_seed(1)).to(dev)

    def frame():
        return render_frame_sharded(models_gpu, emb, None, style, (h, w), NS, NI, scheme="stats",
                                    camera=camera, args=margs)

    for _ in range(warm):
        rgb = frame()
    barrier()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        rgb = frame()
        e1.record()
        barrier()
        ms.append(e0.elapsed_time(e1))
    t = torch.tensor(ms, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # per frame: the slowest rank
    frame_ms = float(t.mean().item())
    check = None
    if rank == 0:
        # the same frame on rank 0 alone, through the UNSHARDED public path: batched render of all
        # 640,000 rays, then style_net.forward on the (1,64,H,W) view (reference eval.py:279-294)
        with torch.no_grad():
            rays = ops.generate_rays(h, w, K, pose, 0.0, 5.0, device=dev)
            res = batched_render(models_gpu, emb, rays, NS, NI, False, None, args=margs)
            solo = models_gpu["decoder"](res["feature_fine"].t().reshape(1, 64, h, w), style)
        check = float((solo - rgb).abs().max().item())
        del rays, res, solo
    barrier()
    return {"frame_ms": frame_ms, "frame_ray_samples_per_s": h * w * (NS + NI) / (frame_ms * 1e-3),
            "frame_check": check, "frame_check_tolerance": 1e-5,
            "frame": {"hw": [h, w], "rays": h * w, "samples": [NS, NI], "scheme": "stats", "reps": reps,
                      "warmup": warm, "scaling": "strong", "rays_per_rank": -(-h * w // world),
                      "collectives_per_frame": "all_reduce(64 f32) + all_reduce(1024 f32) + all_gather(rgb 12 B/ray)"
                                               if world > 1 else "none",
                      "timing": "CUDA events around the whole frame call, max over ranks per frame, mean of reps",
                      "check": "max |rgb| diff vs rank 0 rendering the frame alone through style_net.forward"}}


def train_leg(dev, world, rank, barrier, reps=10, warm=3):
    """BASELINE configs[4]: tools/bench_train.py's step, one 1024-ray patch per rank."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_train
        return bench_train.measure(dev, world, rank, barrier, reps=reps, warm=warm)
    except Exception as e:   # noqa: BLE001  (an extra leg must not take the headline down)
        return {"train_error": f"{type(e).__name__}: {e}"}


def encoder_train_leg(dev, reps=10, warm=3):
    """SURVEY 8f-1 under autograd: forward + backward of enc_a on a 340x512 photo (the training step's
    per-step encoder call), native kernels, CUDA events; the same module on library convolutions in fp32
    beside it."""
    try:
        import torch
        from models.linearStyleTransfer import encoder_sameoutputsize
        torch.manual_seed(0)
        enc = encoder_sameoutputsize(64).to(dev).train()
        x = torch.rand(1, 3, 340, 512, device=dev)
        g = torch.randn(1, 64, 32, 32, device=dev)
        out = {}
        for name in ("native", "library"):
            enc.train_backend = name
            for i in range(warm + reps):
                if i == warm:
                    torch.cuda.synchronize(dev)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                enc.zero_grad(set_to_none=True)
                enc(x).backward(g)
            e1.record()
            torch.cuda.synchronize(dev)
            out[name] = e0.elapsed_time(e1) / reps
        return {"encoder_train": {"workload": "encoder_sameoutputsize forward + backward, 340x512 photo",
                                  "ms_native": out["native"], "ms_library_fp32": out["library"]}}
    except Exception as e:   # noqa: BLE001  (an extra leg must not take the headline down)
        return {"encoder_train_error": f"{type(e).__name__}: {e}"}


def run_ours(args):
    import torch.distributed as dist
    from models.nerf import PosEmbedding
    from models.rendering import render_rays_cross_ray
    from crnerf_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    models, margs = build_models()
    rays_cpu = frame_rays()
    state_cpu = cpu_state(models)          # nn.Module.to() moves in place: keep a CPU copy
    models_gpu = {k: v.to(dev) for k, v in models.items()}
    emb = {"xyz": PosEmbedding(14, 15), "dir": PosEmbedding(3, 4)}
    # every rank renders its own 4096-ray batches of the frame (weak scaling)
    n_batches = rays_cpu.shape[0] // N_RAYS
    rays_dev = rays_cpu.to(dev)
    batch = lambda i: rays_dev[((i * world + rank) % n_batches) * N_RAYS:
                               ((i * world + rank) % n_batches + 1) * N_RAYS]
    # the patch all-gather runs on a side stream: step i's gather overlaps step i+1's render
    gather_bufs = [torch.empty(world * N_RAYS, 64, device=dev) for _ in range(2)] if world > 1 else None
    gather_src = [torch.empty(N_RAYS, 64, device=dev) for _ in range(2)] if world > 1 else None
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    # The per-step patch gather runs on a side stream UNDER the next step's render.  Limiting its
    # communicator's CTAs (so that fewer SMs are taken from the persistent render kernel) was measured
    # at 8 GPUs and LOSES: 5.93 G ray-samples/s with NCCL's default against 5.22 / 4.86 / 5.29 G with
    # max_ctas = 2 / 4 / 8 (the gather then outlasts the step).  CRNERF_GATHER_CTAS keeps the experiment.
    gather_pg = None
    gather_ctas = int(os.environ.get("CRNERF_GATHER_CTAS", "0"))      # 0 = NCCL's default channel count
    if world > 1 and gather_ctas > 0:
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = gather_ctas
            opts.config.min_ctas = 1
            gather_pg = dist.new_group(pg_options=opts)
        except Exception:   # noqa: BLE001  (older torch: no per-communicator config)
            gather_pg = None
    main_stream = torch.cuda.current_stream(dev)

    # the public eval API for a fixed batch shape: render_rays_cross_ray captured in a CUDA graph
    # (crnerf_b200/graphs.py; same kernels, one launch); --no-graph times the plain call
    graphed = None
    if not args.no_graph:
        from crnerf_b200.graphs import GraphedRenderer
        graphed = GraphedRenderer(models_gpu, emb, N_RAYS, NS, NI, args=margs)

    def render(rays):
        if graphed is not None:
            return graphed(rays)
        with torch.no_grad():
            return render_rays_cross_ray(models_gpu, emb, rays, None, NS, False, 0, 0, NI, 32768, False,
                                         test_time=True, args=margs)

    deferred = [None]     # slot whose features wait to be gathered (launched with the NEXT step)
    pending = [None]      # completion event of the gather in flight

    def launch_gather():
        s_prev, deferred[0] = deferred[0], None
        mark = torch.cuda.Event()
        mark.record(main_stream)                 # inside the timed window of the step that hosts it
        comm.wait_event(mark)
        with torch.cuda.stream(comm):
            dist.all_gather_into_tensor(gather_bufs[s_prev], gather_src[s_prev], group=gather_pg)
            done = torch.cuda.Event()
            done.record(comm)
        pending[0] = done

    def step(rays, i):
        """One step = render of this rank's batch; with N > 1 also the patch all-gather of the
        PREVIOUS step's features, launched on a side stream at the start of this step (never
        earlier: it must not hide under the untimed L2 flush) and awaited before the step ends."""
        if world > 1 and deferred[0] is not None:
            launch_gather()
        res = render(rays)
        if world > 1:
            s = i & 1
            gather_src[s].copy_(res["feature_fine"])          # the graph's static output is reused
            if pending[0] is not None:
                main_stream.wait_event(pending[0])             # previous gather ends inside this step
                pending[0] = None
            deferred[0] = s
        return res

    def drain():
        """The last step's gather: launched and awaited on its own."""
        if world > 1 and deferred[0] is not None:
            launch_gather()
        if world > 1 and pending[0] is not None:
            main_stream.wait_event(pending[0])
            pending[0] = None

    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_warm = max(3, args.warmup)
    for i in range(n_warm):
        step(batch(i), i)
    drain()
    barrier()

    # ---- value: device-timed steps, inputs resident, L2 flushed between steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps + 1)]
    n0 = ops.launch_count()
    barrier()
    for i in range(args.steps):
        e0, e1 = evs[i]
        flush.fill_(float(i))
        e0.record()
        step(batch(i), i)
        e1.record()
    e0, e1 = evs[args.steps]           # the last step's gather, timed on its own (N > 1)
    e0.record()
    drain()
    e1.record()
    barrier()
    launches = ops.launch_count() - n0
    if graphed is not None:   # replays do not pass through the library's host-side counter
        launches = graphed.kernels_per_replay * args.steps
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * RAY_SAMPLES_PER_STEP * args.steps / (dev_ms * 1e-3)

    # ---- e2e: rays in pinned host memory, result read back, wall clock
    host_rays = [rays_cpu[((i * world + rank) % n_batches) * N_RAYS:
                          ((i * world + rank) % n_batches + 1) * N_RAYS].clone().pin_memory()
                 for i in range(min(args.steps, n_batches))]
    # two steps in flight (a frame loop's natural shape): step i is enqueued - H2D of its rays, the
    # render, D2H of its result - before the host blocks on step i-1's result, so the GPU never
    # waits for the host.  Every step's result is still read back and waited for.
    host_out = [torch.empty(N_RAYS, 64).pin_memory() for _ in range(2)]
    host_depth = [torch.empty(N_RAYS).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    # The read-back runs on its own stream: the render stream only pays a device-to-device copy of the
    # result (1 MB, ~2 us) into a staging slot, and the PCIe transfer of step i overlaps the render of
    # step i+1 (on one stream the ~30 us transfer sits between two renders).
    copy_stream = torch.cuda.Stream(device=dev)
    stage_f = [torch.empty(N_RAYS, 64, device=dev) for _ in range(2)]
    stage_d = [torch.empty(N_RAYS, device=dev) for _ in range(2)]
    staged = [torch.cuda.Event(), torch.cuda.Event()]
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        s = i & 1
        r = host_rays[i % len(host_rays)]
        res = step(r if graphed is not None else r.to(dev, non_blocking=True), i)   # H2D inside either way
        stage_f[s].copy_(res["feature_fine"], non_blocking=True)     # slot s was read back two steps ago (waited below)
        stage_d[s].copy_(res["depth_fine"], non_blocking=True)
        staged[s].record(main_stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(staged[s])
            host_out[s].copy_(stage_f[s], non_blocking=True)
            host_depth[s].copy_(stage_d[s], non_blocking=True)
            done[s].record(copy_stream)
        if i:
            done[s ^ 1].synchronize()    # the caller consumes step i-1's result while step i runs
    drain()
    done[(args.steps - 1) & 1].synchronize()
    barrier()
    wall = time.perf_counter() - t0
    t = torch.tensor([wall], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * RAY_SAMPLES_PER_STEP * args.steps / float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- sustained: >= 2 s of back-to-back steps under the power cap, clocks sampled
    sustained = None
    if not args.no_sustained:
        n_sus = max(200, int(args.sustained_s * 1e3 / max(1e-3, dev_ms / args.steps)))
        sus_sampler = ClockSampler(local, period=0.05)
        barrier()
        if rank == 0:
            sus_sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_sus):
            step(batch(i), i)
        drain()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sus_ms = float(t.item())
        if rank == 0:
            sus_clocks = sus_sampler.stop()
            sus_val = world * RAY_SAMPLES_PER_STEP * n_sus / (sus_ms * 1e-3)
            tf = POINTS_PER_STEP * FLOP_PER_POINT * n_sus / (sus_ms * 1e-3) / 1e12     # per GPU
            peak_s = None
            try:
                peak_s = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
            except (OSError, KeyError, ValueError):
                pass
            sustained = {"value": sus_val, "unit": UNIT, "steps": n_sus, "seconds": sus_ms * 1e-3,
                         "ms_per_step": sus_ms / n_sus, "tflops_per_gpu": tf,
                         "peak": peak_s, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained",
                         "frac": (tf / peak_s) if peak_s else None, "clocks": sus_clocks,
                         "l2": "no flush between steps (working set: 2 x 1.3 MB weight images + 131 KB of rays)",
                         "note": "whole step (coarse + resample + fine) back to back; frac = algorithmic "
                                 "TFLOP/s of the step / sustained cuBLAS bf16 peak"}

    # ---- roofline of the dominant kernel: the fused fine pass, timed alone
    roof = None
    cpu_base = None
    cpu = None
    if rank == 0:
        with torch.no_grad():
            res = render_rays_cross_ray(models_gpu, emb, batch(0), None, NS, False, 0, 0, NI, 32768,
                                        False, test_time=True, args=margs)
            t_steps = torch.linspace(0, 1, NS, device=dev)
            zc = ops.coarse_z(batch(0), t_steps)
            zf = ops.sample_pdf_merge(zc, res["weights_coarse"], torch.linspace(0, 1, NI, device=dev), NI)
            packed = models_gpu["fine"].packed()
            for _ in range(3):
                ops.render_pass(packed, batch(0), zf)
            reps = 20
            kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(reps)]
            torch.cuda.synchronize()
            for e0, e1 in kev:
                flush.fill_(1.0)
                e0.record()
                ops.render_pass(packed, batch(0), zf)
                e1.record()
            torch.cuda.synchronize()
        k_ms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in kev)
        flop = N_RAYS * (NS + NI) * FLOP_PER_POINT
        achieved = flop / (k_ms * 1e-3) / 1e12
        peak, peak_src = 1590.0, "fallback (B200_PROFILING.md)"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(mp["bf16_tflops"]), "measured burst (MEASURED_PEAKS.json bf16_tflops)"
        except (OSError, KeyError, ValueError):
            pass
        traffic, traffic_info = traffic_record()
        roof = {"bound": "tensor", "kernel": "render_fused_kernel<fp16> fine pass, 4096x192 points",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, **traffic_info, "peak_source": peak_src, "kernel_ms": k_ms,
                "flop_per_launch": flop}
        if world == 1 and not args.no_cpu_baseline:
            cpu = CpuPath(state_cpu)     # cpu_baseline / PSNR legs only: the checker, never the thing measured
            cores = use_all_host_threads()
            v, times = cpu_reference_throughput(cpu, rays_cpu, reps=5, warm=1)
            cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": cpu.kind,
                        "sample": f"5 x one 4096-ray batch (786,432 ray-samples each), "
                                  f"{sum(times):.1f} s of CPU work on {cores} threads"}

    psnr = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        psnr = psnr_delta(load_oracle(), models_gpu, state_cpu, emb, margs, dev)
    psnr_tr = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            psnr_tr = psnr_trained(emb, margs, dev)
        except Exception as e:   # noqa: BLE001  (an extra record must not take the headline down)
            psnr_tr = {"error": f"{type(e).__name__}: {e}"}

    extra = {}
    if not args.no_frame:
        extra.update(frame_leg(models_gpu, emb, margs, dev, world, rank, barrier))
    if not args.no_train:
        extra.update(train_leg(dev, world, rank, barrier))
        if rank == 0 and world == 1:
            extra.update(encoder_train_leg(dev))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands, fp32 accumulate (tcgen05 kind::f16); fp32 everywhere else",
            "data": "synthetic",
            "config": shared_config(),
            "detail": {"rays_per_gpu_per_step": N_RAYS,
                       "parallelism": f"rays sharded x{world}" + (
                           ", all_gather(feature_fine) of step i on a side stream under step i+1's render; "
                           "the last gather is timed on its own" if world > 1 else ""),
                       "api": ("crnerf_b200.graphs.GraphedRenderer (render_rays_cross_ray captured in a CUDA graph)"
                               if graphed is not None else "models.rendering.render_rays_cross_ray"),
                       "l2": "256 MiB buffer written between timed steps (outside the event pairs)",
                       "timing": "CUDA events per step on the launch stream, summed, max over ranks",
                       "e2e_steps_in_flight": 2},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": N_RAYS * 8 * 4, "d2h_bytes_per_step": N_RAYS * 65 * 4,
                    "timing": "wall clock, pinned host rays -> H2D -> render -> D2H feature+depth (read-back on "
                              "its own stream behind a device-side staging copy), every step's result waited for"},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu_base, "psnr": psnr, "psnr_trained": psnr_tr, "clocks": clocks,
            "sustained": sustained,
            "points_per_step": POINTS_PER_STEP,
            "tflops_per_step_device": POINTS_PER_STEP * FLOP_PER_POINT / (dev_ms / args.steps * 1e-3) / 1e12,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the plain API call instead of the graphed one")
    ap.add_argument("--no-frame", action="store_true", help="skip the 800x800 sharded-frame leg")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained loop")
    ap.add_argument("--sustained-s", type=float, default=2.5)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
