"""Per-channel kernels over the activation layouts of a ResNet / transformer (development tool): which
[outer, C, inner] factorisations are far from the copy peak?    python benchmarks/layout_probe.py"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.zeros(128 * 1024 * 1024, device=dev)
flush_rd = torch.zeros(96 * 1024 * 1024, device=dev)
PEAK = 6457.4


def timed(fn, iters=8):
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


shapes = [((256, 64, 112, 112), 1), ((256, 128, 28, 28), 1), ((256, 256, 14, 14), 1), ((512, 512, 7, 7), 1),
          ((1024, 2048, 7, 7), 1), ((4096, 1000), 1), ((512, 512, 3, 3), 0), ((64, 3, 7, 7), 0), ((2048, 1024, 1, 1), 0),
          ((8, 197, 768), 2), ((8, 12, 197, 197), 1)]
for shape, ci in shapes:
    x = torch.randn(shape, device=dev)
    n = x.numel()
    if n < (1 << 22):      # make small cases big enough to time: stack copies along the batch
        rep = (1 << 24) // n
        x = x.repeat(rep, *([1] * (x.dim() - 1))) if ci != 0 else x
        n = x.numel()
    y = torch.empty_like(x)
    C = x.shape[ci]
    outer = 1
    for s_ in x.shape[:ci]:
        outer *= s_
    lay = (outer, C, n // (outer * C))
    dec = torch.full((C,), 5.0, device=dev)
    lines = torch.tensor([[-0.5, 0.5]] * C, device=dev)
    mask = (torch.arange(C, device=dev) % 4 == 0)
    r = dict(shape=list(x.shape), layout=list(lay), mb=round(n * 4 / 1e6, 1))
    for name, fn, b in (("fq_pow2_ch", lambda: ops.fq_pow2_fwd(x, dec, lay, out=y), 8),
                        ("fq_line_ch", lambda: ops.fq_line_fwd(x, lines, 8, True, lay, out=y), 8),
                        ("fq_pow2_mask75", lambda: ops.fq_pow2_fwd(x, dec[:1], lay, mask=mask, out=y), 8),
                        ("ste_bwd_ch", lambda: ops.ste_bwd(x, dec, True, 8, 0, lay, clamp_in_place=False, want_gx=True), 8),
                        ("reduce_sum_max", lambda: ops.reduce_stats(x, lay, abssum=True, absmax=True), 4),
                        ("reduce_minmax", lambda: ops.reduce_stats(x, lay, minmax=True), 4)):
        t = timed(fn)
        r[name] = (round(t, 1), round(b * n / t / 1e3 / PEAK, 2))
    print(json.dumps(r), flush=True)
    del x, y
