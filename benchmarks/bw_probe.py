"""Pure write / read / copy bandwidth of this GPU at the bench tensor's size (development probe):
what a store-only or load-only stream can reach, beside the driver's copy peak."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush_w = torch.zeros(128 * 1024 * 1024, device=dev)
flush_r = torch.zeros(96 * 1024 * 1024, device=dev)


def timed(fn, flush=True, iters=12):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            flush_w.add_(1.0)
            flush_r.max()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for n in (51380224, 4 * 51380224):
    x = torch.randn(n, device=dev)
    y = torch.empty_like(x)
    dec = torch.tensor([5.0], device=dev)
    mask0 = torch.zeros(64, dtype=torch.bool, device=dev)
    lay = (n // (64 * 3136), 64, 3136)
    for name, fn, nbytes in (
            ("torch_zero_", lambda: y.zero_(), 4 * n),
            ("torch_fill_", lambda: y.fill_(1.5), 4 * n),
            ("torch_copy_", lambda: y.copy_(x), 8 * n),
            ("torch_sum(read)", lambda: x.sum(), 4 * n),
            ("reduce_stats(read)", lambda: ops.reduce_stats(x, lay, abssum=True, absmax=True), 4 * n),
            ("fq_pow2_all_pruned(write only)", lambda: ops.fq_pow2_fwd(x, dec, lay, mask=mask0, out=y), 4 * n),
            ("fq_pow2_dense", lambda: ops.fq_pow2_fwd(x, dec, lay, out=y), 8 * n)):
        for flush in (True, False):
            us = timed(fn, flush)
            print(json.dumps(dict(op=name, mb=round(4 * n / 1e6, 1), flush=flush, us=round(us, 2),
                                  gbs=round(nbytes / us / 1e3, 1))), flush=True)
    del x, y
