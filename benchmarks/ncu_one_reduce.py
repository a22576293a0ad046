"""ncu target (development tool): the stand-alone reduction on one layout.  argv: outer C inner [tuning-key value]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

lay = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 256, 196)
if len(sys.argv) > 5:
    ops.set_tuning(int(sys.argv[4]), int(sys.argv[5]))
x = torch.randn(lay, device="cuda:0")
for _ in range(3):
    ops.reduce_stats(x, lay, abssum=True, absmax=True)
torch.cuda.synchronize()
print("done")
