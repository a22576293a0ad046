"""ncu target: BASELINE config 3's one-launch weight access (K8) — line, scaler and decimal kinds on [4096, 4096]."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
w = torch.randn(4096, 4096, device=dev) * 0.02
for kind, width, t in ((ops.ROW_LINE, 2, 1), (ops.ROW_SCALER, 1, 0), (ops.ROW_DECIMAL, 1, 0)):
    p = torch.zeros(4096, width, device=dev)
    for i in range(3):
        ops.row_quant_fused_(w, p, kind, 4, t + i, True)
torch.cuda.synchronize()
print("done")
