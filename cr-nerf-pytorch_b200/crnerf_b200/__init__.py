"""crnerf_b200: B200-native (sm_100a) kernels for the CR-NeRF volume-rendering path.

``crnerf_b200.ops`` wraps the C ABI of ``libcrnerf_b200.so`` (see
``include/crnerf_b200.h``); the sibling ``models`` package mirrors the
reference's ``models/`` API on top of it.  Importing this package does not load
the shared library; the first op does, and raises if it is missing.
"""
from . import _lib, ops  # noqa: F401
from . import torch_ops  # noqa: F401  (registers torch.ops.crnerf.*)
from ._lib import CrnerfError  # noqa: F401

__all__ = ["ops", "CrnerfError"]
