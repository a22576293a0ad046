"""ncu target (development tool): the row reduction on one short-row layout.  argv: outer C inner"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

lay = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 256, 196)
x = torch.randn(lay, device="cuda:0")
for _ in range(3):
    ops.reduce_stats(x, lay, abssum=True, absmax=True)
torch.cuda.synchronize()
print("done")
