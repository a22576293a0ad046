"""BASELINE config 3's one-launch weight access (K8) on [4096, 4096]: us per launch, L2 flushed between launches"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
w = torch.randn(4096, 4096, device=dev) * 0.02
flush = torch.empty(96 << 20, dtype=torch.float32, device=dev)
for name, kind, width, t in (("line", ops.ROW_LINE, 2, 1), ("scaler", ops.ROW_SCALER, 1, 0), ("decimal", ops.ROW_DECIMAL, 1, 0)):
    p = torch.zeros(4096, width, device=dev)
    ts = []
    for i in range(40):
        flush.zero_()
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.row_quant_fused_(w, p, kind, 4, t + i, True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(name, "median us", round(ts[len(ts) // 2], 2), "best", round(ts[0], 2), flush=True)
