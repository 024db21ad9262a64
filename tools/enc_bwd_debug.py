"""Debug: gradient planes of one stage of the native encoder backward vs float64 autograd (GPU).
usage: CRNERF_ENC_BWD_STOP=<1..4> python tools/enc_bwd_debug.py H W   (1: dZ6, 2: dZ5, 3: dZ4, 4: dZ3)"""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cr-nerf-pytorch_b200"), ROOT]
import torch
import torch.nn.functional as Fn
from models.linearStyleTransfer import encoder_sameoutputsize
from crnerf_b200 import ops

H, W = int(sys.argv[1]), int(sys.argv[2])
stop = int(os.environ["CRNERF_ENC_BWD_STOP"])
torch.manual_seed(11)
enc = encoder_sameoutputsize(64)
with torch.no_grad():
    for c in enc._convs():
        c.bias.mul_(3.0)
enc = enc.cuda()
gen = torch.Generator().manual_seed(H * 13 + W)
x = torch.rand(1, 3, H, W, generator=gen).cuda()
g = torch.randn(1, 64, 32, 32, generator=gen).cuda()
ops._ENCODER_DEBUG = {}
out = enc(x)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_encoder_backward import _tape_planes
tape = out.grad_fn.tape
like3, like5 = _tape_planes(tape, H, W, "a3"), _tape_planes(tape, H, W, "a5")
out.backward(g)
def pool(t, like):
    _, idx = Fn.max_pool2d(like.double(), 2, return_indices=True)
    return t.flatten(2).gather(2, idx.flatten(2)).view(idx.shape)
scratch = ops._ENCODER_DEBUG["scratch"]

ref = copy.deepcopy(enc).double()
pad = lambda t: Fn.pad(t, (1, 1, 1, 1), mode="reflect")
lre = lambda t: Fn.leaky_relu(t, 0.2)
z = {}
h = ref.conv1(x.double())
z[2] = ref.conv2(pad(h)); h = lre(z[2])
z[3] = ref.conv3(pad(h)); h = pool(lre(z[3]), like3)
z[4] = ref.conv4(pad(h)); h = lre(z[4])
z[5] = ref.conv5(pad(h)); h = pool(lre(z[5]), like5)
z[6] = ref.conv6(pad(h)); h = lre(z[6])
o = lre(ref.conv7(Fn.adaptive_avg_pool2d(h, 32)))
for t in z.values():
    t.retain_grad()
o.backward(g.double())

gs = lambda w: ((w + 15) & ~15) + 4
al = lambda b: (b + 255) & ~255
gp = lambda C, h, w: al(2 * C * (h + 4) * gs(w) * 2 + 256)
H2, W2 = H // 2, W // 2
H4, W4 = H2 // 2, W2 // 2
ga = max(gp(128, H4, W4), gp(128, H2, W2), gp(64, H, W))
gb = max(gp(128, H2, W2), gp(64, H, W))
dxb = max(al((h + 2) * (gs(w) - 2) * C * 4) for C, h, w in [(128, H4, W4), (128, H2, W2), (64, H2, W2), (64, H, W)])
part = al(max(148 * 3 * 128 * 128, 296 * 1728) * 4)
small = ga + gb + dxb + part + al(1024 * 128 * 4) + al(1024 * 64 * 4) + al(16 * 4 * 296 * 8 * 4) + al(296 * 12 * 4)
scales = scratch[small:small + 32].view(torch.float32)
maxbits = scratch[small + 32:small + 64].view(torch.int32)
print("scales", scales.tolist(), "max", maxbits.view(torch.float32).tolist())
layer, C, h, w, off = {1: (6, 128, H4, W4, 0), 2: (5, 128, H2, W2, ga), 3: (4, 128, H2, W2, 0), 4: (3, 64, H, W, ga)}[stop]
Wg = gs(w)
n = C * (h + 4) * Wg
pl = scratch[off:off + 4 * n].view(torch.float16)
hi = pl[:n].view(C // 8, h + 4, Wg, 8).float()
lo = pl[n:].view(C // 8, h + 4, Wg, 8).float()
G = (hi + lo).permute(0, 3, 1, 2).reshape(C, h + 4, Wg).double() / float(scales[layer])
want = z[layer].grad[0]
inner = G[:, 2:2 + h, 2:2 + w]
halo = G.clone(); halo[:, 2:2 + h, 2:2 + w] = 0
print(f"dZ{layer}: rel L2 {float((inner - want).norm() / want.norm()):.2e}  max|halo| {float(halo.abs().max()):.2e}  max|G| {float(inner.abs().max()*scales[layer]):.3g}")
err = (inner - want).abs()
idx = torch.nonzero(err > 1e-3 * want.abs().max())
print("bad entries:", idx.shape[0], idx[:12].tolist())
for L, like in ((3, like3), (5, like5)):
    d = (lre(z[L]).detach().float() - like).abs().max()
    flips = ((z[L].detach() > 0) != (like > 0)).sum()
    print(f"a{L}: max abs diff vs float64 {float(d):.2e}, sign flips {int(flips)}")
if stop in (2, 4) and idx.shape[0]:
    c, yy, xx = idx[0].tolist()
    y0, x0 = yy & ~1, xx & ~1
    print("ref lrelu(z) window:", lre(z[layer])[0, c, y0:y0 + 2, x0:x0 + 2].tolist())
    zf = {2: None}
    ref32 = copy.deepcopy(enc).float()
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        h = ref32.conv1(x)
        h = lre(ref32.conv2(pad(h)))
        a3 = lre(ref32.conv3(pad(h))); h = Fn.max_pool2d(a3, 2)
        h = lre(ref32.conv4(pad(h)))
        a5 = lre(ref32.conv5(pad(h)))
    print("fp32 torch window  :", (a5 if stop == 2 else a3)[0, c, y0:y0 + 2, x0:x0 + 2].tolist())
    print("want grad window:", want[c, y0:y0 + 2, x0:x0 + 2].tolist())
    print("got  grad window:", inner[c, y0:y0 + 2, x0:x0 + 2].tolist())
