"""Cycle breakdown of the fused kernel (CTA 0): where the MMA issuer and one epilogue
thread of each tile group spend their time.  Run on a B200: python tools/prof_breakdown.py"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cr-nerf-pytorch_b200"))
import torch
from crnerf_b200 import synthetic
from crnerf_b200 import ops
from models.nerf import NeRF_sigma

torch.manual_seed(0)
args = types.SimpleNamespace(nerf_out_dim=64, pertubeCord=False, img_wh=[64, 64])
fine = NeRF_sigma('fine', args, in_channels_xyz=93, in_channels_dir=27).cuda()
rays = synthetic.pinhole_rays(64, 64, synthetic.synthetic_pose(0)).cuda()
EXPS = [(-2, "normal"), (-3, "EXP1: epilogue skips TMEM traffic + math"),
        (-4, "EXP2: producer skips weight copies"), (-5, "EXP1+EXP2"), (-6, "EXP4: no bias MMAs")]
if len(sys.argv) > 1 and sys.argv[1] == "bias":
    EXPS = [EXPS[0], EXPS[4]]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    EXPS = EXPS[:1]
for code, label in EXPS:
  print("====", label)
  for S in (192,):
    z = ops.coarse_z(rays, torch.linspace(0, 1, S, device="cuda"))
    packed = fine.packed()
    for _ in range(2):
        ops.render_pass(packed, rays, z)
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    ops.debug_set(buf, code)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.render_pass(packed, rays, z)
    e1.record()
    torch.cuda.synchronize()
    ops.debug_set(None, -1)
    print(f"   kernel time {e0.elapsed_time(e1)*1e3:.0f} us")
    c = buf.cpu().tolist()
    life, pairs = c[0], max(c[5], 1)
    print(f"S={S}: issuer lifetime {life} cyc over {pairs} tile pairs = {life/pairs:.0f} cyc/pair "
          f"(MMA-bound floor ~38.5k)")
    for nm, off in (("X", 0), ("Y", 24)):
        lf = max(c[off], 1)
        print(f"   issuer {nm}: lifetime {c[off]}; blocked on " + ", ".join(
            f"{name} {v} ({100*v/lf:.1f}%)" for name, v in zip(["emb_full", "a_full", "d_empty", "ring_full", "pipe_turn"], c[off+1:off+5] + [c[off+6]])) + f"; issue bursts {c[off+7]} ({100*c[off+7]/lf:.1f}%, {c[off+7]/pairs/19:.0f} cyc/unit)")
    for b in (0, 1):
        o = c[8 + 8*b: 8 + 8*b + 8]
        if o[0] == 0: continue
        print(f"   epilogue group {'XY'[b]}: lifetime {o[0]}, blocked on d_full {o[1]} ({100*o[1]/o[0]:.1f}%), "
              f"embedding {o[2]} ({o[2]/pairs:.0f}/tile), sigma+scan {o[3]} ({o[3]/pairs:.0f}/tile), "
              f"feature reduction {o[4]} ({o[4]/pairs:.0f}/tile); layer epilogues {o[5]} "
              f"({o[5]/pairs:.0f}/tile): stage-half {o[6]/pairs/9:.0f} cyc each, flush-half/dir {o[7]/pairs/10:.0f} cyc each")
