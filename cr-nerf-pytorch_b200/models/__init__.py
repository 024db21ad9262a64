"""Drop-in mirror of the reference's ``models`` package (same module names,
class/function names, call signatures and ``state_dict`` keys) for the
volume-rendering hot path, backed by the sm_100a kernels in ``crnerf_b200``.

Put ``cr-nerf-pytorch_b200/`` on ``sys.path`` ahead of the reference checkout and
``from models.rendering import *`` / ``from models.nerf import *`` in
``train_mask_grid_sample.py`` / ``eval.py`` resolve here.

Only the hot-path modules are mirrored (``rendering``, ``nerf``,
``linearStyleTransfer``, ``nerf_decoder_stylenerf``).  The callers also import
``models.esrgan``, ``models.lightweight_seg`` and ``models.networks``
(train_mask_grid_sample.py:15,20; eval.py:18; appearance_modification_video.py:13),
which are off the path: this package's ``__path__`` is extended over every other
``models`` directory found later on ``sys.path``, so those names resolve to the
reference's own files, unmodified, while the mirrored names keep resolving here
(the mirror's directory comes first in ``__path__``).
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg_root = _os.path.dirname(_here)
if _pkg_root not in _sys.path:
    _sys.path.insert(0, _pkg_root)


def _extend_over_reference():
    """Append the other ``models`` package directories on ``sys.path`` to ``__path__``."""
    seen = {_os.path.realpath(p) for p in __path__}
    for entry in list(_sys.path):
        cand = _os.path.join(entry or _os.getcwd(), "models")
        real = _os.path.realpath(cand)
        if real in seen or not _os.path.isfile(_os.path.join(cand, "__init__.py")):
            continue
        seen.add(real)
        __path__.append(cand)


_extend_over_reference()
