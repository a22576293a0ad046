"""BASELINE config 3's one-launch weight access (K8) on [4096, 4096]: us per launch — cold (L2 flushed with clean lines
before every launch) and back to back (20 launches per CUDA graph).  QSPARSE_B200_LIB selects the library build."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
w = torch.randn(4096, 4096, device=dev) * 0.02
flush = torch.empty(96 << 20, dtype=torch.float32, device=dev)
out = {"lib": os.environ.get("QSPARSE_B200_LIB", "default")}
for name, kind, width, t in (("line", ops.ROW_LINE, 2, 1), ("scaler", ops.ROW_SCALER, 1, 0), ("decimal", ops.ROW_DECIMAL, 1, 0)):
    p = torch.zeros(4096, width, device=dev)
    ts = []
    for i in range(60):
        flush.zero_()
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.row_quant_fused_(w, p, kind, 4, t + i, True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.row_quant_fused_(w, p, kind, 4, t + 100, True)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(20):
            ops.row_quant_fused_(w, p, kind, 4, t + 101 + i, True)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    out[name] = {"cold_mean_us": round(sum(ts) / len(ts), 2), "back_to_back_us": round(e0.elapsed_time(e1) * 1e3 / 200, 2)}
print(out, flush=True)
