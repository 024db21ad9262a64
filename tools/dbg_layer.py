"""Debug helper: dump one layer's activations from the fused kernel and show which columns differ
from the fp16-operand emulation.  usage: python tools/dbg_layer.py [layer]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cr-nerf-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import build_mirror_models, state
from crnerf_b200 import ops
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 0
g = torch.load(os.path.join(ROOT, "tests/golden/posenc_mlp.pt"))
models, _ = build_mirror_models(0)
fine = models["fine"]; p = state(fine)
x = torch.cat([g["emb_xyz"], g["emb_dir"]], 1)
r16 = lambda t: t.to(torch.float16).float()
h = torch.relu(r16(x[:, :93]) @ r16(p["xyz_encoding_1.0.weight"]).t() + p["xyz_encoding_1.0.bias"])
fine = fine.cuda()
buf = torch.zeros(96, 256, device="cuda")
ops.debug_set(buf, layer)
with torch.no_grad():
    fine(x.cuda())
torch.cuda.synchronize()
ops.debug_set(None, -1)
d = (buf.cpu() - h).abs()
print("max err per 16-col group:", [round(float(d[:, c:c+16].max()), 4) for c in range(0, 256, 16)])
print("rows with err:", (d.max(1).values > 1e-3).nonzero().flatten().tolist()[:20])
print("got[0,:8]", buf[0, :8].cpu().tolist()); print("want[0,:8]", h[0, :8].tolist())
print("nobias[0,:8]", torch.relu(r16(x[:, :93]) @ r16(p["xyz_encoding_1.0.weight"]).t())[0, :8].tolist())
