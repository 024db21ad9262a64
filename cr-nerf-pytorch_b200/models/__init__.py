"""Drop-in mirror of the reference's ``models`` package (same module names,
class/function names, call signatures and ``state_dict`` keys) for the
volume-rendering hot path, backed by the sm_100a kernels in ``crnerf_b200``.

Put ``cr-nerf-pytorch_b200/`` on ``sys.path`` ahead of the reference checkout and
``from models.rendering import *`` / ``from models.nerf import *`` in
``train_mask_grid_sample.py`` / ``eval.py`` resolve here.
"""
import os as _os
import sys as _sys

_pkg_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _pkg_root not in _sys.path:
    _sys.path.insert(0, _pkg_root)
