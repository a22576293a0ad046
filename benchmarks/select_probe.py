import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops
n = 1 << 26
torch.manual_seed(4)
v = (torch.randn(n, device="cuda") * 0.02).abs()
for _ in range(2):
    thr = ops.kth_value(v, n // 2)
torch.cuda.synchronize()
print(thr.item())
