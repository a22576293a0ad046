"""Short-row reductions: tile / row kernels vs the column kernel on the [outer, C * inner] view (tuning key 20)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.zeros(128 * 1024 * 1024, device=dev)
flush_rd = torch.zeros(96 * 1024 * 1024, device=dev)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)
        flush_rd.max()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for lay in ((256, 256, 196), (1024, 256, 64), (256, 128, 784), (512, 512, 100), (256, 512, 256), (64, 64, 3136),
            (4096, 64, 196), (256, 2048, 64)):
    x = torch.randn(lay, device=dev)
    n = x.numel()
    row = dict(layout=list(lay), mb=round(n * 4 / 1e6, 1))
    ref = None
    for colmax in (63, 256, 1024, 4096):
        ops.set_tuning(20, colmax)
        t1 = timed(lambda: ops.reduce_stats(x, lay, abssum=True, absmax=True))
        t2 = timed(lambda: ops.reduce_stats(x, lay, minmax=True))
        r = ops.reduce_stats(x, lay, abssum=True, absmax=True, minmax=True)
        if ref is None:
            ref = r
        else:
            assert torch.equal(r["absmax"], ref["absmax"]) and torch.equal(r["min"], ref["min"])
            assert torch.allclose(r["abssum"], ref["abssum"], rtol=1e-6)
        row[f"colmax{colmax}"] = [round(t1, 1), round(t2, 1), round(4 * n / t1 / 1e3 / 6457.4, 2)]
    ops.set_tuning(20, 63)
    print(json.dumps(row), flush=True)
