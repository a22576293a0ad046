"""Phase timing of the select kernels (needs a build with -DQSB_SELECT_TIMING):
    QSB_EXTRA_NVCC_FLAGS=-DQSB_SELECT_TIMING python -m qsparse_b200.build -f
"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops, _native as N
n = 1 << 26
torch.manual_seed(4)
v = (torch.randn(n, device="cuda") * 0.02).abs()
for _ in range(3):
    thr = ops.kth_value(v, n // 2)
torch.cuda.synchronize()
ws = N.workspace(v.device, 1)
base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
hdr = (256 + 4096 + 4096) * 8 + 3 * 2048 * 4 + 32 * 16 * 8
ticks = ws[base + hdr + 128: base + hdr + 128 + 32 * 8].view(torch.int64).cpu().tolist()
print("sampler cycles:", [ticks[i] - ticks[0] for i in range(6)])
for p in range(3):
    t = ticks[8 + p * 8: 8 + p * 8 + 5]
    print(f"pass {p} CTA0 cycles:", [x - t[0] for x in t])
