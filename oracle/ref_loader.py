"""TEST INFRASTRUCTURE - loads the UNMODIFIED reference files of the hot path, when present.

``__graft_entry__.build()`` stages ``models/{rendering,nerf}.py`` (and the other files of the
path) from ``/root/reference`` into ONE archive, the git-ignored ``oracle/_ref/reference_path.zip``
(a build output like a compiled ``.so``: never committed, byte-identical members), so that they
travel to the GPU box with the snapshot (nothing at run time reads ``/root/reference``).  ``bench.py``'s CPU
legs (``--impl reference`` and ``cpu_baseline``) time these real files when they are there
(``kind: "reference"``) and fall back to the pinned port ``oracle/crnerf_oracle.py``
(``kind: "port"``) when they are not.  Never imported by the product package.

The two files needed for BASELINE configs[1] are self-contained (``models/rendering.py:1-2``
imports torch + einops, ``models/nerf.py:1-3`` torch + os), so they are loaded by path under
private module names and cannot collide with the product's ``models`` mirror.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref", "reference_path.zip")
# files of the path staged by build(); the first two are what the CPU legs run
FILES = ("models/rendering.py", "models/nerf.py", "models/linearStyleTransfer.py",
         "models/nerf_decoder_stylenerf.py", "models/__init__.py", "losses.py",
         # off-path modules and the three caller scripts: only the drop-in tests read them
         # (tests/test_dropin.py: the callers' import lines and their batched_inference / decode
         # functions are executed, unmodified, against the mirror)
         "models/esrgan.py", "models/lightweight_seg.py", "models/networks.py", "models/conv_decoder.py",
         "eval.py", "train_mask_grid_sample.py", "appearance_modification_video.py",
         # the training Dataset (its __getitem__ is cut out and executed by tests/test_training_step_replay.py)
         "datasets/phototourism_mask_grid_sample.py")


def find_reference() -> Optional[str]:
    """The staged archive, else a reference checkout (build container only)."""
    if os.path.isfile(STAGED):
        return STAGED
    for base in (os.environ.get("CRNERF_REFERENCE", ""), "/root/reference"):
        if base and os.path.isfile(os.path.join(base, "models", "rendering.py")) \
                and os.path.isfile(os.path.join(base, "models", "nerf.py")):
            return base
    return None


def stage(src: str = "/root/reference") -> bool:
    """Archive the path's files from the reference checkout into oracle/_ref/ (build step)."""
    import zipfile
    if not os.path.isfile(os.path.join(src, "models", "rendering.py")):
        return False
    os.makedirs(os.path.dirname(STAGED), exist_ok=True)
    with zipfile.ZipFile(STAGED, "w", zipfile.ZIP_DEFLATED) as z:
        for rel in FILES:
            if os.path.isfile(os.path.join(src, rel)):
                z.write(os.path.join(src, rel), rel)
    return True


def read_source(base: str, rel: str) -> str:
    if base.endswith(".zip"):
        import zipfile
        with zipfile.ZipFile(base) as z:
            return z.read(rel).decode("utf-8")
    with open(os.path.join(base, rel), encoding="utf-8") as f:
        return f.read()


def _load(base: str, rel: str, name: str):
    """Execute one self-contained reference file as module ``name`` (source unmodified)."""
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(base, rel)
    sys.modules[name] = mod
    exec(compile(read_source(base, rel), mod.__file__, "exec"), mod.__dict__)
    return mod


def extract(dst: str, base: Optional[str] = None) -> Optional[str]:
    """Unpack the staged files into ``dst`` (a scratch directory) and return it, or None."""
    base = base or find_reference()
    if base is None:
        return None
    if not base.endswith(".zip"):
        return base
    import zipfile
    with zipfile.ZipFile(base) as z:
        z.extractall(dst)
    return dst


class ReferenceRenderer:
    """The reference's own ``render_rays_cross_ray`` + ``NeRF_sigma`` + ``PosEmbedding`` on CPU,
    called the way ``eval.py:29-59`` calls them."""

    def __init__(self, state_coarse: dict, state_fine: dict, base: Optional[str] = None):
        base = base or find_reference()
        if base is None:
            raise FileNotFoundError("reference files not staged (oracle/_ref) and /root/reference absent")
        self.base = base
        self.rendering = _load(base, "models/rendering.py", "_crnerf_ref_rendering")
        self.nerf = _load(base, "models/nerf.py", "_crnerf_ref_nerf")
        self.args = types.SimpleNamespace(nerf_out_dim=64, pertubeCord=False, img_wh=[32, 32])
        coarse = self.nerf.NeRF_sigma("coarse", self.args, in_channels_xyz=93, in_channels_dir=27)
        fine = self.nerf.NeRF_sigma("fine", self.args, in_channels_xyz=93, in_channels_dir=27,
                                    encode_appearance=True, in_channels_a=48, encode_random=True)
        coarse.load_state_dict(state_coarse, strict=True)
        fine.load_state_dict(state_fine, strict=True)
        self.models = {"coarse": coarse.eval(), "fine": fine.eval()}
        self.emb = {"xyz": self.nerf.PosEmbedding(14, 15), "dir": self.nerf.PosEmbedding(3, 4)}

    @torch.no_grad()
    def render_rays(self, rays, n_samples=64, n_importance=128, chunk=8192):
        return self.rendering.render_rays_cross_ray(self.models, self.emb, rays, None, n_samples, False,
                                                    0, 0, n_importance, chunk, False, test_time=True,
                                                    args=self.args)
