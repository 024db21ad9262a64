#!/usr/bin/env python
"""Pure-write, pure-read and copy bandwidth of this GPU through the tensor library's own kernels (2 GiB buffers,
CUDA events, best of 5): the ceilings the training forward's activation stores (write-only, 1.0 GB per step) and the
backward GEMMs' operand streams (read-mostly) are compared with in DESIGN 7."""
import json
import torch
dev = torch.device("cuda")
n = 1 << 29   # 2 GiB of fp32
a = torch.empty(n, device=dev); b = torch.empty(n, device=dev)
def best(fn, reps=5):
    out = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out = min(out, e0.elapsed_time(e1))
    return out
a.fill_(1.0); b.copy_(a); torch.cuda.synchronize()
w = best(lambda: a.fill_(2.0)); c = best(lambda: b.copy_(a)); r = best(lambda: a.sum())
print(json.dumps({"write_TBps": 4 * n / w / 1e9, "copy_TBps_read_plus_write": 8 * n / c / 1e9, "read_TBps": 4 * n / r / 1e9,
                  "bytes": 4 * n}))
