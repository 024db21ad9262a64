"""Data-parallel gradient exchange for the training step (BASELINE configs[4]: one 1,024-ray patch per rank).

The reference trains under Lightning's DDP wrapper, which copies every gradient into 25 MB buckets and
all-reduces the buckets; that keeps working on the mirror.  This is the short path for the library's own
backward kernels: they write a module's gradients into ONE flat buffer (``ops.render_backward`` - 24 tensors
of a NeRF_sigma; ``csrc/style_backward.cu`` - the 22 tensors of style_net) and ``param.grad`` are views of
it (they share its storage), so the exchange is one in-place all-reduce per gradient storage, no bucket copies:
one call for each NeRF_sigma's 24 tensors (5.3 MB each), one for each of the decoder's two 4 MB FC matrices and one
for its 20 small tensors packed together (autograd sums the decoder's two calls per step into fresh tensors, so
those do not share a buffer) - 5 collectives instead of 68 for the training step - and it fits between a
graph replay of forward + backward
(``GraphedTrainStep(step_fn, optimizer=None, parameters=...)``) and the optimizer launch.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def gradient_buffers(params: Iterable[torch.nn.Parameter]) -> List[torch.Tensor]:
    """One flat tensor per storage that holds gradients (spanning the first to the last gradient element in it),
    in first-use order.  Gradients are grouped by storage, not by ``_base``: what autograd leaves in
    ``param.grad`` is a detached alias that no longer names the buffer it was cut from."""
    groups = {}
    for p in params:
        g = p.grad
        if g is None:
            continue
        if g.is_sparse:
            raise RuntimeError("sparse gradients are not supported")
        if not g.is_contiguous():
            raise RuntimeError("a gradient is not contiguous")
        st = g.untyped_storage()
        key = (st.data_ptr(), g.dtype, g.device)
        lo, hi = g.storage_offset(), g.storage_offset() + g.numel()
        if key in groups:
            e = groups[key]
            e[1], e[2] = min(e[1], lo), max(e[2], hi)
        else:
            groups[key] = [st, lo, hi, g.dtype, g.device]
    out = []
    for st, lo, hi, dtype, device in groups.values():
        out.append(torch.empty(0, dtype=dtype, device=device).set_(st, lo, (hi - lo,), (1,)))
    return out


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                        average: bool = True, pack_below_bytes: int = 1 << 18) -> int:
    """Mean (or sum) of the gradients over the ranks, in place; returns the number of collectives issued.
    Buffers below ``pack_below_bytes`` travel together: one concatenation, one all-reduce, one multi-tensor copy
    back (a collective costs ~30 us of latency at 8 ranks whatever its size; the decoder's 20 small tensors would
    be 20 of them).  A no-op on one rank or without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)

    def reduce_(b):
        if average and b.is_cuda:
            dist.all_reduce(b, op=dist.ReduceOp.AVG, group=group)
        else:   # gloo has no AVG
            dist.all_reduce(b, group=group)
            if average:
                b.div_(world)

    bufs = gradient_buffers(params)
    small = [b for b in bufs if b.numel() * b.element_size() < pack_below_bytes]
    if len(small) < 2 or len({(b.dtype, b.device) for b in small}) != 1:
        small = []
    packed = {id(b) for b in small}
    n = 0
    for b in bufs:
        if id(b) not in packed:
            reduce_(b)
            n += 1
    if small:
        flat = torch.cat(small)
        reduce_(flat)
        torch._foreach_copy_(small, list(flat.split([b.numel() for b in small])))
        n += 1
    return n
