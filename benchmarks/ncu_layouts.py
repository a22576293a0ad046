"""ncu target for the short-feature-map layouts (development tool): one launch of each per-channel
kernel per layout, so the launch list gives kernel-only durations next to layout_probe.py's event times."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for lay in ((256, 256, 196), (512, 512, 49), (256, 128, 784), (16384, 1000, 1), (20488, 768, 1), (32, 12, 38809)):
    x = torch.randn(lay, device=dev)
    y = torch.empty_like(x)
    C = lay[1]
    dec = torch.full((C,), 5.0, device=dev)
    lines = torch.tensor([[-0.5, 0.5]] * C, device=dev)
    for _ in range(2):
        ops.fq_pow2_fwd(x, dec, lay, out=y)
        ops.fq_line_fwd(x, lines, 8, True, lay, out=y)
        ops.ste_bwd(x, dec, True, 8, 0, lay, clamp_in_place=False, want_gx=True)
        ops.reduce_stats(x, lay, abssum=True, absmax=True)
        ops.reduce_stats(x, lay, minmax=True)
    torch.cuda.synchronize()
print("done")
